"""Host-only checks of the balanced piece lists of a fused pressure pass (csrc/pass_schedule.h through the C ABI
smk_pass_schedule): every (tile, output plane) is covered exactly once, no CTA exceeds the reported cost, the shares
are balanced, and the z-step counts quoted in DESIGN.md.  No GPU needed."""
import numpy as np
import pytest


def _check(sl, W, H, lo, hi, K, nctas):
    s = sl.pass_schedule(W, H, lo, hi, K, nctas)
    tx, ty = s["tiles"]
    assert tx == -(-(W + 1) // 56) and ty == -(-(H + 1) // (32 - 2 * K))
    assert len(s["ctas"]) <= nctas
    cover = np.zeros((ty, tx, hi - lo), dtype=np.int32)
    costs = []
    for cta in s["ctas"]:
        assert cta, "no empty CTA"
        c = 0
        for bx, by, z0, z1 in cta:
            assert 0 <= bx < tx and 0 <= by < ty and lo <= z0 < z1 <= hi
            cover[by, bx, z0 - lo:z1 - lo] += 1
            c += (z1 - z0) + 2 * K - 2
        costs.append(c)
    assert (cover == 1).all(), "every output plane of every tile exactly once"
    assert max(costs) == s["cost"]
    return s, costs


@pytest.mark.parametrize("W,H,lo,hi,K,nctas", [
    (256, 256, 0, 257, 4, 148), (512, 512, 0, 513, 4, 148), (80, 80, 0, 81, 4, 148), (80, 80, 0, 81, 2, 148),
    (256, 256, 120, 393, 4, 148), (33, 20, 0, 18, 4, 148), (57, 41, 0, 10, 2, 7), (130, 100, 3, 4, 4, 148),
    (120, 50, 0, 41, 4, 1), (1024, 1024, 0, 1025, 4, 148)])
def test_pieces_cover_every_plane_once_and_are_balanced(smk, W, H, lo, hi, K, nctas):
    from smoke_simulation_b200 import slab as sl
    s, costs = _check(sl, W, H, lo, hi, K, nctas)
    tx, ty = s["tiles"]
    total = tx * ty * (hi - lo)
    if total >= 8 * nctas and nctas > 1:
        # equal shares: the busiest CTA does the ideal share of planes + the lead-ins of its pieces (+ one stub)
        ideal = total / min(nctas, len(costs))
        most = max(len(c) for c in s["ctas"])
        assert most <= -(-int(ideal) // (hi - lo)) + 2
        assert s["cost"] <= ideal + most * (2 * K - 2) + 4, (s["cost"], ideal, most)
        assert min(costs) >= 0.5 * s["cost"]


def test_balanced_needs_fewer_z_steps_than_the_chunk_grid(smk):
    """z-steps of the busiest SM: (tile, z-chunk) grid in waves vs equal shares.  (Fewer steps did NOT make the pass faster
    on the B200 -- neighbouring tiles fall out of step and their halo re-reads miss L2, DESIGN.md section 4 -- which is why
    the grid stays the default; this only pins the step counts quoted there: 104 vs 116 at 256^3, 778 vs 789 at 512^3.)"""
    from smoke_simulation_b200 import slab as sl

    def grid_cost(tiles, nz, K, sms=148):
        best = None
        for n in range(1, max(1, nz // 4) + 1):
            zc = -(-nz // n)
            ctas = tiles * -(-nz // zc)
            waves = -(-ctas // sms)
            eff = ctas / (waves * sms) * zc / (zc + 2 * K)
            if best is None or eff > best[0] + 1e-9:
                best = (eff, waves * (zc + 2 * K - 2))
        return best[1]
    for W, nz, gain in ((80, 81, None), (256, 257, 0.92), (512, 513, 0.99)):
        s = sl.pass_schedule(W, W, 0, nz, 4, 148)
        tx, ty = s["tiles"]
        gc = grid_cost(tx * ty, nz, 4)
        if W == 256:
            assert (s["cost"], gc) == (104, 116)
        if W == 512:
            assert (s["cost"], gc) == (778, 789)
        if gain is None:   # 80^3: many short chunks already fill one wave
            assert s["cost"] <= gc + 2
        else:
            assert s["cost"] <= gain * gc, (W, s["cost"], gc)

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def po():
    import pyoracle
    if not os.path.exists(pyoracle.LIB_ORACLE):
        pyoracle.build(ref=False)
    return pyoracle


@pytest.fixture(scope="session")
def smk():
    import smoke_simulation_b200 as m
    return m


def pytest_generate_tests(metafunc):
    """Every GPU test runs twice: the fused pressure passes on the TMA-staged kernel and on the round-1 kernel (the
    automatic choice depends on the grid size, and the test grids are small).  CPU tests run once."""
    if metafunc.definition.get_closest_marker("gpu") is not None and "pass_kernel" in metafunc.fixturenames:
        metafunc.parametrize("pass_kernel", ["tma", "reg"], indirect=True)


@pytest.fixture(autouse=True)
def pass_kernel(request):
    kind = getattr(request, "param", None)
    if kind is None:
        yield None
        return
    from smoke_simulation_b200 import binding
    binding.DEFAULT_PASS_KERNEL = kind
    yield kind
    binding.DEFAULT_PASS_KERNEL = None


def rel_err(a, b):
    """max|a-b| / max|b| -- the parity metric of SURVEY.md section 8(c)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    num = np.abs(a - b).max()
    return 0.0 if num == 0 else float(num / den) if den > 0 else float("inf")


def random_state(po, W, H, D, seed=1234):
    """u,v,w ~ U(-4,4), density ~ U(0,1), mask Bernoulli(0.9) + solid floor."""
    rng = np.random.default_rng(seed)
    st = {}
    for f, n in ((po.U, "u"), (po.V, "v"), (po.W, "w")):
        st[n] = rng.uniform(-4, 4, po.field_shape(f, W, H, D)).astype(np.float32)
    st["smoke"] = rng.uniform(0, 1, (D, H, W)).astype(np.float32)
    m = (rng.uniform(0, 1, (D, H, W)) < 0.9).astype(np.uint8)
    m[:, 0, :] = 0
    st["mask"] = m
    return st


def inject(po, e, st):
    e.set_field(po.U, po.BUF0, st["u"]); e.set_field(po.V, po.BUF0, st["v"]); e.set_field(po.W, po.BUF0, st["w"])
    e.set_field(po.SMOKE, po.BUF0, st["smoke"]); e.set_field(po.MASK, po.NOW, st["mask"])


ALL_FIELDS = (("smoke", 0), ("u", 1), ("v", 2), ("w", 3))


def all_fields(po, e):
    d = {"mask": e.get_field(po.MASK)}
    for n, f in ALL_FIELDS:
        d[n + "_now"] = e.get_field(f, po.NOW)
        d[n + "_past"] = e.get_field(f, po.PAST)
    return d

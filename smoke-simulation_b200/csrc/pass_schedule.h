// pass_schedule.h -- how one fused pressure pass is cut into pieces for the SMs (pure host C++, no CUDA).
//
// A pass of K half-sweeps (kernels_pressure_reg.cuh) marches (x,y) tiles along z.  A PIECE = one tile over the output
// node planes [zo0, zo1); it costs (zo1 - zo0) useful z-steps plus 2K lead-in / lead-out steps, of which the first ones
// skip their sweeps.  Any partition of (tiles x planes) into pieces gives the same bits, so the partition is purely a
// scheduling decision:
//   * a (tile, z-chunk) grid, one piece per CTA (k_pressure_reg): simple, but CTAs quantise into waves over the SMs;
//   * this header: `nctas` CTAs (one per SM), each with a LIST of pieces of equal total cost (k_pressure_reg_bal).
// STATUS: an experiment kept for ablation (smk_set_pass_ctas / SMK_PASS_CTAS), not the default.  On the B200 the equal
// shares (104 instead of 116 z-steps at 256^3) buy nothing: neighbouring tiles are no longer at the same z at the same
// time, their halo re-reads miss L2 and DRAM reads double (profiles/r1_balanced_schedule.txt).
// The reference has no counterpart (one thread per cell, cu:795-801); the tests check coverage and balance on the CPU
// through the host-only entry point smk_pass_schedule.
#pragma once
#include <algorithm>
#include <vector>

namespace sched {

struct Piece { int bx, by, zo0, zo1; };

struct PassSchedule {
    std::vector<Piece> pieces;
    std::vector<int> first; // CTA b works through pieces [first[b], first[b+1])
    int cost = 0;           // z-steps of the busiest CTA
    int nctas() const { return first.empty() ? 0 : (int)first.size() - 1; }
};

// lead = z-steps a piece spends before its first output plane, counted at the weight of a full step:
// 2K - 2 for a pass of K half-sweeps (2K lead-in / lead-out steps, the first ones without sweeps), 3 for a Jacobi piece
inline int pass_lead(int K) { return 2 * K - 2; }
constexpr int JACOBI_LEAD = 3;
inline int piece_cost(int planes, int lead) { return planes + lead; }
constexpr int MIN_PLANES = 4; // no piece shorter than this unless the whole range is (lead-in would dominate)

// greedy fill with per-CTA budget C; returns the number of CTAs used (pieces/first filled if out != nullptr)
inline int fill(int tx, int ty, int lo, int hi, int lead, int C, PassSchedule* out)
{
    const int minlen = std::min(MIN_PLANES, hi - lo);
    int ctas = 0, used = 0, maxc = 0;
    if (out) { out->pieces.clear(); out->first.assign(1, 0); }
    for (int by = 0; by < ty; by++)
        for (int bx = 0; bx < tx; bx++) {
            int z = lo;
            while (z < hi) {
                int room = C - used - lead;
                if (room < std::min(minlen, hi - z) && used > 0) { // close this CTA
                    maxc = std::max(maxc, used);
                    ctas++; used = 0;
                    if (out) out->first.push_back((int)out->pieces.size());
                    continue;
                }
                int len = std::min(hi - z, std::max(room, minlen));
                const int rem = hi - z - len;
                if (rem > 0 && rem < minlen) { // do not leave a stub behind: shorten this piece if it stays long enough
                    if (len - (minlen - rem) >= minlen) len -= minlen - rem;
                    else len = hi - z;
                }
                if (out) out->pieces.push_back(Piece{bx, by, z, z + len});
                used += piece_cost(len, lead);
                z += len;
            }
        }
    if (used > 0) {
        maxc = std::max(maxc, used);
        ctas++;
        if (out) out->first.push_back((int)out->pieces.size());
    }
    if (out) out->cost = maxc;
    return ctas;
}

// smallest budget whose greedy fill needs at most `nctas` CTAs
inline PassSchedule balance(int tx, int ty, int lo, int hi, int lead, int nctas)
{
    PassSchedule s;
    if (tx <= 0 || ty <= 0 || hi <= lo || nctas <= 0) { s.first.assign(1, 0); return s; }
    int a = piece_cost(std::min(MIN_PLANES, hi - lo), lead), b = tx * ty * piece_cost(hi - lo, lead);
    while (a < b) {
        const int m = a + (b - a) / 2;
        if (fill(tx, ty, lo, hi, lead, m, nullptr) <= nctas) b = m; else a = m + 1;
    }
    // the stub rule makes the CTA count only ALMOST monotone in the budget: look a little below for a better one
    int best = b;
    for (int c = b - 1; c >= std::max(1, b - 8); c--)
        if (fill(tx, ty, lo, hi, lead, c, nullptr) <= nctas) best = c;
    while (fill(tx, ty, lo, hi, lead, best, &s) > nctas) best++; // (never loops in practice; b is feasible)
    return s;
}

// Multi-GPU passes that read the neighbours' planes inside the kernel: a piece is a BOUNDARY piece of the lower side
// if zo0 < zl and of the upper side if zo1 > zh (kernels_pressure_reg.cuh: it reads the neighbour's planes and writes
// the planes the neighbour reads).  They wait for the neighbour's epoch and the last of them publishes this pass's, so
// every CTA runs them first: the epoch goes out early in the pass and the waits hide behind the interior pieces of the
// other CTAs.  Returns the number of boundary pieces per side in counts[2].
inline void boundary_first(PassSchedule& s, int zl, int zh, int counts[2])
{
    counts[0] = counts[1] = 0;
    for (int b = 0; b + 1 < (int)s.first.size(); b++)
        std::stable_partition(s.pieces.begin() + s.first[b], s.pieces.begin() + s.first[b + 1],
                              [&](const Piece& p) { return p.zo0 < zl || p.zo1 > zh; });
    for (const Piece& p : s.pieces) {
        if (p.zo0 < zl) counts[0]++;
        if (p.zo1 > zh) counts[1]++;
    }
}

// z-steps of the busiest SM under the plain (tile, z-chunk) grid with `slots` co-resident CTAs
inline int grid_cost(int tiles, int nz, int zchunk, int lead, int slots)
{
    const int nchunks = (nz + zchunk - 1) / zchunk;
    const long ctas = (long)tiles * nchunks;
    return (int)((ctas + slots - 1) / slots) * piece_cost(zchunk, lead);
}

} // namespace sched

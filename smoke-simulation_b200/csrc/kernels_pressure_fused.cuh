// kernels_pressure_fused.cuh -- K red/black half-sweeps per launch, temporally blocked in shared memory.
//
// The reference runs 60 half-sweeps as 60 launches that each stream u,v,w through HBM (cu:797-801,
// 25 B/cell each).  Here one launch advances an (x,y) tile K half-sweeps while marching along z
// ("2.5D" / wavefront temporal blocking):
//
//   step t:  plane t enters the shared-memory ring (prefetched into registers one step earlier)
//            sweep 1 runs on cell plane t-1, sweep 2 on plane t-2, ..., sweep K on plane t-K
//            node plane t-K is final and is written to the OUTPUT buffers
//
// Sweep j on plane c only needs sweep j-1 on planes c-1, c, c+1, which the order above guarantees; same-
// colour cells never share a face, so the per-cell arithmetic and its data dependencies are exactly those
// of K separate launches => results are bit-identical to the unfused schedule.
// A tile is loaded with a K-cell halo in x and y (and K planes of lead-in / lead-out in z); cells in the
// halo are updated with incomplete neighbourhoods and are never written back (trapezoid scheme), which is
// why the kernel is out of place: it reads the "in" buffers and writes every node of the "out" buffers.
//
// Shared-memory layout: ring of R = K+1 planes; each plane row is split by x parity, E[i] = f[x0+2i],
// O[i] = f[x0+2i+1].  A warp owns RPW rows (lane = pair of x-adjacent cells, exactly one of which has the
// active colour), so every shared-memory access of a sweep is unit-stride across the warp: no bank
// conflicts.  Global traffic is one float2 load and one float2 store per node and field (coalesced 256 B
// per warp), prefetched one z-step ahead in registers.  The RPW rows of a thread are independent and are
// processed branch-free so that their dependency chains interleave.
#pragma once
#include <cuda_runtime.h>
#include "grid.h"
#include "kernels_basic.cuh"

namespace smk {

template <int K, int NW, int RPW>
struct FusedCfg {
    static constexpr int LX = 64;           // loaded tile width  (one warp = one row of 32 cell pairs)
    static constexpr int LY = NW * RPW;     // loaded tile height (rows)
    static constexpr int R = K + 1;         // ring depth in planes
    static constexpr int PL = LX * LY;      // nodes per plane tile
    static constexpr int OX = LX - 2 * K;   // output tile width
    static constexpr int OY = LY - 2 * K;   // output tile height
    static constexpr int THREADS = NW * 32;
    static constexpr size_t SMEM = (size_t)R * PL * (3 * sizeof(float) + 1);
};

template <int K, int NW, int RPW>
__global__ void __launch_bounds__(NW * 32, 1)
k_pressure_fused(GridP g, const float* __restrict__ ui, const float* __restrict__ vi, const float* __restrict__ wi,
                 float* __restrict__ uo, float* __restrict__ vo, float* __restrict__ wo,
                 const unsigned char* __restrict__ code, int sweep0, int zchunk)
{
    using C = FusedCfg<K, NW, RPW>;
    constexpr int LX = C::LX, LY = C::LY, R = C::R, PL = C::PL;
    static_assert(K % 2 == 0, "tile origins must keep the x parity (K even)");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* su = reinterpret_cast<float*>(smem_raw);
    float* sv = su + R * PL;
    float* sw = sv + R * PL;
    unsigned short* sc = reinterpret_cast<unsigned short*>(sw + R * PL); // [R][LY][32]: code bytes of a cell pair

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int x0 = blockIdx.x * C::OX - K;
    const int y0 = blockIdx.y * C::OY - K;
    const int zo0 = g.zlo + blockIdx.z * zchunk;            // output node planes [zo0, zo1)
    const int zo1 = min(zo0 + zchunk, g.zlo + g.nzn);
    const int t0 = zo0 - K, t1 = zo1 + K - 1;               // planes that enter the ring
    const int xg = x0 + 2 * lane;                           // global x of this lane's pair (even)

    // per-thread constants: offsets inside a plane and what may be loaded / stored
    int noff[RPW], koff[RPW], srow[RPW];
    bool nok[RPW], kok[RPW], sok[RPW];
#pragma unroll
    for (int r = 0; r < RPW; r++) {
        const int yl = wid + r * NW, yg = y0 + yl;
        nok[r] = xg >= 0 && xg <= g.P - 2 && yg >= 0 && yg < g.SY;
        kok[r] = xg >= 0 && xg <= g.PC - 2 && yg >= 0 && yg < g.H;
        sok[r] = nok[r] && yl >= K && yl < LY - K && 2 * lane >= K && 2 * lane < LX - K;
        noff[r] = nok[r] ? xg + yg * g.P : 0;
        koff[r] = kok[r] ? xg + yg * g.PC : 0;
        srow[r] = yl * LX + lane;
    }

    float2 pu[RPW], pv[RPW], pw[RPW];
    unsigned short pc[RPW];

    auto prefetch = [&](int z) {
        const bool zn = z >= g.zlo && z < g.zlo + g.nzn;
        const bool zc = z >= g.zlo && z < g.zlo + g.nzc;
        const long long nb = (long long)(z - g.zlo) * g.nplane;
        const long long kb = (long long)(z - g.zlo) * g.kplane;
#pragma unroll
        for (int r = 0; r < RPW; r++) {
            pu[r] = pv[r] = pw[r] = make_float2(0.f, 0.f);
            pc[r] = 0;
            if (zn && nok[r]) {
                pu[r] = __ldg(reinterpret_cast<const float2*>(ui + nb + noff[r]));
                pv[r] = __ldg(reinterpret_cast<const float2*>(vi + nb + noff[r]));
                pw[r] = __ldg(reinterpret_cast<const float2*>(wi + nb + noff[r]));
            }
            if (zc && kok[r]) pc[r] = __ldg(reinterpret_cast<const unsigned short*>(code + kb + koff[r]));
        }
    };

    // per-thread, per-row constants of the sweep phases
    int dv1[RPW], rowpar[RPW];
#pragma unroll
    for (int r = 0; r < RPW; r++) {
        const int yl = wid + r * NW;
        dv1[r] = yl < LY - 1 ? LX : 0;                       // last tile row is never updated: keep v1's address legal
        rowpar[r] = (y0 + yl + sweep0 + 1) & 1;              // + t gives the x parity of the active colour
    }
    const float2 M1 = make_float2(-1.0f, -1.0f);

    prefetch(t0);
    int slot_t = 0; // ring slot of plane t
    for (int t = t0; t <= t1; t++) {
        // plane t: registers -> ring (same thread <-> same addresses as the store below: no extra barrier)
#pragma unroll
        for (int r = 0; r < RPW; r++) {
            const int o = slot_t * PL + srow[r];
            su[o] = pu[r].x; su[o + 32] = pu[r].y;
            sv[o] = pv[r].x; sv[o + 32] = pv[r].y;
            sw[o] = pw[r].x; sw[o + 32] = pw[r].y;
            sc[slot_t * (PL / 2) + (wid + r * NW) * 32 + lane] = pc[r];
        }
        if (t < t1) prefetch(t + 1); // in flight during the sweeps
        __syncthreads();

        // Sweep j runs on cell plane c = t-j with colour (sweep0+j-1)&1, so the x parity of the active cell of a
        // row, (y + c + sweep0 + j - 1) & 1 = (y + t + sweep0 + 1) & 1, is the same for all K sweeps of this step.
        int base[RPW], du1[RPW];
        unsigned sh[RPW], msk[RPW];
#pragma unroll
        for (int r = 0; r < RPW; r++) {
            const int par = (rowpar[r] + t) & 1;
            const bool edge = par && lane == 31;                 // cell 63 of the row: u[64] is not in the tile
            base[r] = srow[r] + par * 32;
            du1[r] = edge ? 0 : (par ? -31 : 32);                // O[i] -> E[i+1],  E[i] -> O[i]
            sh[r] = par * 8;
            msk[r] = (edge || dv1[r] == 0) ? 0u : 0xffu;
        }

#pragma unroll
        for (int j = 1; j <= K; j++) {
            if (t - j >= t0) { // cell plane c = t-j is in the ring
                int sl = slot_t - j; if (sl < 0) sl += R;          // slot of plane c
                int sl1 = sl + 1; if (sl1 == R) sl1 = 0;           // slot of plane c+1
                const int so = sl * PL, so1 = sl1 * PL, sk = sl * (PL / 2);
                int a[RPW];
                unsigned cd[RPW];
                float u0[RPW], u1[RPW], v0[RPW], v1[RPW], w0[RPW], w1[RPW], p[RPW], rr[RPW], af[RPW], dv[RPW];
                bool any_slow = false;
#pragma unroll
                for (int r = 0; r < RPW; r++) {
                    a[r] = so + base[r];
                    const unsigned cw = sc[sk + (wid + r * NW) * 32 + lane];
                    cd[r] = (cw >> sh[r]) & msk[r];
                    if (!(cd[r] & CODE_ACTIVE)) cd[r] = 0;
                    u0[r] = su[a[r]]; u1[r] = su[a[r] + du1[r]];
                    v0[r] = sv[a[r]]; v1[r] = sv[a[r] + dv1[r]];
                    w0[r] = sw[a[r]]; w1[r] = sw[so1 + base[r]];
                    const int acc = __popc(cd[r] & 63u);
                    rr[r] = c_rcp[acc];
                    af[r] = -(float)acc;
                }
                // pressure_p_fast() on pairs of rows with packed f32x2 instructions (same roundings per lane:
                // x*(-1)+y == y-x exactly rounded once, like the scalar FADDs)
#pragma unroll
                for (int r = 0; r + 1 < RPW; r += 2) {
                    float2 d = __ffma2_rn(make_float2(u0[r], u0[r + 1]), M1, make_float2(u1[r], u1[r + 1]));
                    d = __ffma2_rn(make_float2(v0[r], v0[r + 1]), M1, d);
                    d = __fadd2_rn(d, make_float2(v1[r], v1[r + 1]));
                    d = __ffma2_rn(make_float2(w0[r], w0[r + 1]), M1, d);
                    d = __fadd2_rn(d, make_float2(w1[r], w1[r + 1]));
                    const float2 R2 = make_float2(rr[r], rr[r + 1]);
                    const float2 q0 = __fmul2_rn(d, R2);
                    const float2 rem = __ffma2_rn(q0, make_float2(af[r], af[r + 1]), d);
                    const float2 q = __ffma2_rn(rem, R2, q0);
                    dv[r] = d.x; dv[r + 1] = d.y;
                    p[r] = __double2float_rn(__dmul_rn((double)q.x, -1.9));
                    p[r + 1] = __double2float_rn(__dmul_rn((double)q.y, -1.9));
                    any_slow |= (fabsf(q.x) < 1.17549435e-38f && d.x != 0.0f && cd[r] != 0) ||
                                (fabsf(q.y) < 1.17549435e-38f && d.y != 0.0f && cd[r + 1] != 0);
                }
                if (RPW & 1) {
                    constexpr int r = RPW - 1;
                    bool slow;
                    p[r] = pressure_p_fast(u0[r], u1[r], v0[r], v1[r], w0[r], w1[r], __popc(cd[r] & 63u), slow);
                    dv[r] = 1.0f;
                    any_slow |= slow && cd[r] != 0;
                }
                if (__any_sync(0xffffffffu, any_slow)) { // denormal-range divergence somewhere: exact IEEE division
#pragma unroll
                    for (int r = 0; r < RPW; r++)
                        if (cd[r]) p[r] = pressure_p(u0[r], u1[r], v0[r], v1[r], w0[r], w1[r], __popc(cd[r] & 63u));
                }
#pragma unroll
                for (int r = 0; r + 1 < RPW; r += 2) {
                    const float2 P2 = make_float2(p[r], p[r + 1]);
                    const float2 nu0 = __ffma2_rn(P2, M1, make_float2(u0[r], u0[r + 1]));
                    const float2 nu1 = __fadd2_rn(make_float2(u1[r], u1[r + 1]), P2);
                    const float2 nv0 = __ffma2_rn(P2, M1, make_float2(v0[r], v0[r + 1]));
                    const float2 nv1 = __fadd2_rn(make_float2(v1[r], v1[r + 1]), P2);
                    const float2 nw0 = __ffma2_rn(P2, M1, make_float2(w0[r], w0[r + 1]));
                    const float2 nw1 = __fadd2_rn(make_float2(w1[r], w1[r + 1]), P2);
                    u0[r] = nu0.x; u0[r + 1] = nu0.y; u1[r] = nu1.x; u1[r + 1] = nu1.y;
                    v0[r] = nv0.x; v0[r + 1] = nv0.y; v1[r] = nv1.x; v1[r + 1] = nv1.y;
                    w0[r] = nw0.x; w0[r + 1] = nw0.y; w1[r] = nw1.x; w1[r + 1] = nw1.y;
                }
                if (RPW & 1) {
                    constexpr int r = RPW - 1;
                    u0[r] = __fsub_rn(u0[r], p[r]); u1[r] = __fadd_rn(u1[r], p[r]);
                    v0[r] = __fsub_rn(v0[r], p[r]); v1[r] = __fadd_rn(v1[r], p[r]);
                    w0[r] = __fsub_rn(w0[r], p[r]); w1[r] = __fadd_rn(w1[r], p[r]);
                }
#pragma unroll
                for (int r = 0; r < RPW; r++) {
                    if (cd[r] & CODE_SX0) su[a[r]] = u0[r];
                    if (cd[r] & CODE_SX1) su[a[r] + du1[r]] = u1[r];
                    if (cd[r] & CODE_SY0) sv[a[r]] = v0[r];
                    if (cd[r] & CODE_SY1) sv[a[r] + dv1[r]] = v1[r];
                    if (cd[r] & CODE_SZ0) sw[a[r]] = w0[r];
                    if (cd[r] & CODE_SZ1) sw[so1 + base[r]] = w1[r];
                }
            }
            // no barrier between the sweeps of one step: see kernels_pressure_quad.cuh (the only shared face of
            // consecutive sweeps is w[t-j] of the thread's own column)
        }
        __syncthreads();

        // node plane t-K is final: ring -> output buffers (interior of the tile only)
        const int s = t - K;
        if (s >= zo0 && s < zo1) {
            int sl = slot_t - K; if (sl < 0) sl += R;
            const long long nb = (long long)(s - g.zlo) * g.nplane;
#pragma unroll
            for (int r = 0; r < RPW; r++) {
                if (sok[r]) {
                    const int o = sl * PL + srow[r];
                    *reinterpret_cast<float2*>(uo + nb + noff[r]) = make_float2(su[o], su[o + 32]);
                    *reinterpret_cast<float2*>(vo + nb + noff[r]) = make_float2(sv[o], sv[o + 32]);
                    *reinterpret_cast<float2*>(wo + nb + noff[r]) = make_float2(sw[o], sw[o + 32]);
                }
            }
        }
        if (++slot_t == R) slot_t = 0;
    }
}

} // namespace smk

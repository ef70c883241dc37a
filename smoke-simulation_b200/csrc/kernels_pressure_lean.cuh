// kernels_pressure_lean.cuh -- the fused pressure pass of round 2 (default): the schedule, tile geometry, lane mapping and
// arithmetic of kernels_pressure_reg.cuh (K red/black SOR half-sweeps per launch while marching a 64 x 32 tile along z,
// u and w of a lane's quad in a register ring, v in shared memory; reference: divergence cu:356-394 x 60, cu:797-801),
// rewritten around its measured limiter.  ncu on the round-1 kernel: 65 % issue-active with ~490 executed instructions
// per warp and z-step, of which ~105 in the step's head (64-bit address arithmetic and source selection for the
// prefetch), ~75 in its tail, and ~72 per sweep phase where ~40 are needed (profiles/r1_final_pressure_reg_ncu_full.txt;
// SASS of the round-1 kernel).  Changes, none of which touches a floating-point operation:
//
//   * stencil information comes in the "pcode" encoding (grid.h): "every cell of this sweep phase is ACTIVE with six
//     fluid neighbours" is one LOP3 + one vote instead of nine instructions; cells that are not updated are only
//     looked at when that test fails;
//   * global addresses are 32-bit byte offsets from warp-uniform bases (one add per step and array instead of 64-bit
//     pointer arithmetic per access); the source of a plane (local arrays or a neighbour's memory) switches the
//     uniform bases at the two plane indices where it can change instead of being re-derived at every step;
//   * the lead-in / trapezoid conditions of the K sweeps of a step collapse into one per-step sweep count;
//   * the denormal-quotient test of a pair is one min + compare on values the sweep already has.
//
// Bit-identity with K separate launches follows from the same argument as before (each piece recomputes its own
// trapezoid halo; per-cell operation order untouched) and is asserted by tests/test_parity_gpu.py and friends.
#pragma once
#include <cuda_runtime.h>
#include "grid.h"
#include "kernels_basic.cuh"
#include "kernels_pressure_reg.cuh"

namespace smk {

// One sweep phase of one lane on ring position J (plane t-J): two same-colour cells of the lane's quad.
// PAR = x parity of the active colour (warp uniform).  pv = shared address of this lane's V0 pair (parity offset included).
template <int PAR, int RS>
__device__ __forceinline__ void lean_update(float2& ue, float2& uo, float2& we, float2& wo, float2& we1, float2& wo1,
                                            float* __restrict__ pv, const unsigned cw, const bool hnz)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr unsigned SH = 8u * PAR;                 // byte 0/2 (PAR 0) or 1/3 (PAR 1) of the code word
    constexpr unsigned M_AC = 0x00C000C0u << SH;      // ACTIVE | COMPLEX of both cells
    constexpr unsigned V_A = 0x00400040u << SH;       // ... == ACTIVE, not COMPLEX
    constexpr unsigned M_C = 0x00800080u << SH;
    const float2 M1 = make_float2(-1.0f, -1.0f);
    float2 U0, U1, W0, W1;
    if (PAR == 0) { U0 = ue; U1 = uo; W0 = we; W1 = we1; }
    else {
        U0 = uo; W0 = wo; W1 = wo1;
        U1 = make_float2(ue.y, __shfl_down_sync(FULL, ue.x, 1)); // u[4h+2], u[4h+4] (next quad's first face)
    }
    float2 V0 = *reinterpret_cast<float2*>(pv), V1 = *reinterpret_cast<float2*>(pv + RS);

    float2 d = __ffma2_rn(U0, M1, U1);   // -u0 + u1           (cu:379-381, left to right, one rounding each)
    d = __ffma2_rn(V0, M1, d);           //  ... - v0
    d = __fadd2_rn(d, V1);               //  ... + v1
    d = __ffma2_rn(W0, M1, d);           //  ... - w0
    d = __fadd2_rn(d, W1);               //  ... + w1

    // tier A: both cells of every lane ACTIVE with six fluid neighbours; tier B: no COMPLEX cell (some are not updated)
    const bool allact = __all_sync(FULL, (cw & M_AC) == V_A);
    bool simple = allact;
    if (!allact) simple = __all_sync(FULL, (cw & (cw << 1) & M_C) == 0u);
    if (simple) {
        const float r6 = 0x1.555556p-3f; // RN(1/6)
        const float2 q0 = __fmul2_rn(d, make_float2(r6, r6));
        const float2 rem = __ffma2_rn(q0, make_float2(-6.0f, -6.0f), d);
        const float2 q = __ffma2_rn(rem, make_float2(r6, r6), q0);
        float2 P = p_pair_from_q(q);
        // exact for every |d| >= 2^-125 (exhaustive check); below that (and d != 0) a tie on the
        // denormal grid can round the wrong way -> exact integer quotient for those lanes
        const unsigned ax = (__float_as_uint(d.x) & 0x7fffffffu) - 1u, ay = (__float_as_uint(d.y) & 0x7fffffffu) - 1u;
        if (__any_sync(FULL, min(ax, ay) < 0x00ffffffu)) {
            if (ax < 0x00ffffffu) P.x = __double2float_rn(__dmul_rn((double)div6_tiny(d.x), M19));
            if (ay < 0x00ffffffu) P.y = __double2float_rn(__dmul_rn((double)div6_tiny(d.y), M19));
        }
        if (!allact) { // cells that are not updated: old -/+ 0 = old
            if (!(cw & (0x40u << SH))) P.x = 0.f;
            if (!(cw & (0x400000u << SH))) P.y = 0.f;
        }
        U0 = __ffma2_rn(P, M1, U0); U1 = __fadd2_rn(U1, P);
        V0 = __ffma2_rn(P, M1, V0); V1 = __fadd2_rn(V1, P);
        W0 = __ffma2_rn(P, M1, W0); W1 = __fadd2_rn(W1, P);
    } else {
        // general path: per-cell neighbour count, per-face masks (a masked face gets its old value back)
        const unsigned cs = cw >> SH;
        unsigned cA = cs & 0xffu, cB = (cs >> 16) & 0xffu;
        if (!(cA & PCODE_ACTIVE)) cA = 0;
        if (!(cB & PCODE_ACTIVE)) cB = 0;
        const int nA = __popc(cA & 63u), nB = __popc(cB & 63u);
        const float2 rr = make_float2(c_rcp[nA], c_rcp[nB]);
        const float2 q0 = __fmul2_rn(d, rr);
        const float2 rem = __ffma2_rn(q0, make_float2(-(float)nA, -(float)nB), d);
        const float2 q = __ffma2_rn(rem, rr, q0);
        float2 P = p_pair_from_q(q);
        const unsigned ax = __float_as_uint(d.x) & 0x7fffffffu, ay = __float_as_uint(d.y) & 0x7fffffffu;
        const bool sx = nA == 6 && (ax - 1u < 0x00ffffffu), sy = nB == 6 && (ay - 1u < 0x00ffffffu);
        if (__any_sync(FULL, sx || sy)) { // only acc = 6 can miss (exhaustive check, see pressure_p_fast)
            if (sx) P.x = __double2float_rn(__dmul_rn((double)div6_tiny(d.x), M19));
            if (sy) P.y = __double2float_rn(__dmul_rn((double)div6_tiny(d.y), M19));
        }
        // old - 0 and old + 0 return old (up to the sign of a zero): a masked face keeps its value
        U0 = __ffma2_rn(make_float2((cA & CODE_SX0) ? P.x : 0.f, (cB & CODE_SX0) ? P.y : 0.f), M1, U0);
        U1 = __fadd2_rn(U1, make_float2((cA & CODE_SX1) ? P.x : 0.f, (cB & CODE_SX1) ? P.y : 0.f));
        V0 = __ffma2_rn(make_float2((cA & CODE_SY0) ? P.x : 0.f, (cB & CODE_SY0) ? P.y : 0.f), M1, V0);
        V1 = __fadd2_rn(V1, make_float2((cA & CODE_SY1) ? P.x : 0.f, (cB & CODE_SY1) ? P.y : 0.f));
        W0 = __ffma2_rn(make_float2((cA & CODE_SZ0) ? P.x : 0.f, (cB & CODE_SZ0) ? P.y : 0.f), M1, W0);
        W1 = __fadd2_rn(W1, make_float2((cA & CODE_SZ1) ? P.x : 0.f, (cB & CODE_SZ1) ? P.y : 0.f));
    }
    *reinterpret_cast<float2*>(pv) = V0;
    *reinterpret_cast<float2*>(pv + RS) = V1;
    if (PAR == 0) { ue = U0; uo = U1; we = W0; we1 = W1; }
    else {
        uo = U0; wo = W0; wo1 = W1;
        ue.y = U1.x;
        const float from_left = __shfl_up_sync(FULL, U1.y, 1); // the left quad's updated u[4h]
        if (hnz) ue.x = from_left;                              // h == 0: tile edge, face stays stale (halo)
    }
}

// forcing + clamp on the way into the first pass of a step (kernels_pressure_reg.cuh: force_clamp_node), pcode flavour:
// the cell is fluid iff (p & 0xC0) != 0
__device__ __forceinline__ void force_clamp_node_p(float& u, float& v, float& w, unsigned pc, float d, bool clampable, const ForceArgs& fa)
{
    force_clamp_node(u, v, w, ((pc & 0xC0u) ? CODE_SELF : 0u) | (pc & CODE_SY0), d, clampable, fa);
}

// One PIECE of a pass: the tile (bx, by) marched over the output node planes [zo0, zo1).  MAXW: also reduce max |w| over
// the planes written (the bound of the next advection's backtrace in z) into *wmax (bit pattern, atomicMax).
template <int K, int NW, bool FORCE, bool MAXW>
__device__ __forceinline__ void lean_pass_piece(const GridP& g, float* __restrict__ uo, float* __restrict__ vo, float* __restrict__ wo,
                                                const unsigned char* __restrict__ pcode, int sweep0, const PassRange& pr,
                                                const ForceArgs& fa, float* __restrict__ sv, int bx, int by, int zo0, int zo1,
                                                unsigned* __restrict__ wmax)
{
    using C = RegCfg<K, NW>;
    constexpr int LY = C::LY, RS = C::RS, HO = C::HO, R = C::R, PLS = C::PLS;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int h = lane & 15;                                 // quad index inside the row
    // rows of a warp and the deepest sweep they need: see kernels_pressure_reg.cuh (trapezoid halo, shallow rows paired)
    constexpr bool SKIP = (K == 4 && NW >= 8 && NW % 4 == 0);
    const int hb = lane >> 4;
    const int yl = !SKIP ? wid + hb * NW : wid >= 4 ? wid + hb * ((LY - 8) / 2) : hb == 0 ? wid : wid == 3 ? LY - 1 : LY - 2 - wid;
    const int jmax = (SKIP && wid < 3) ? wid + 1 : K;
    const int x0 = bx * C::OX - C::HX;
    const int y0 = by * C::OY - K;
    const int t0 = zo0 - K, t1 = zo1 + K - 1;                // planes that enter the ring
    const int xg = x0 + 4 * h, yg = y0 + yl;

    const bool nok = xg >= 0 && xg <= g.P - 4 && yg >= 0 && yg < g.SY;
    const bool kok = xg >= 0 && xg <= g.PC - 4 && yg >= 0 && yg < g.H;
    const bool sok = nok && yl >= K && yl < LY - K && h >= C::HX / 4 && h < 16 - C::HX / 4;
    const bool dok = FORCE && xg >= 0 && xg <= g.W - 4 && yg >= 0 && yg < g.H; // W % 4 == 0 (host checks)
    const int vrow = yl * RS + 2 * h;                        // E[2h] of this lane's row inside a shared v plane
    const int rowpar = (y0 + yl + sweep0 + 1) & 1;           // (+ t) = x parity of the active colour, warp uniform
    const bool hnz = h != 0;

    for (int i = threadIdx.x; i < R * PLS; i += C::THREADS) sv[i] = 0.f; // dummy entries / dummy row: defined values
    __syncthreads();

    // register ring, index k = plane t-k
    float2 UE[R], UO[R], WE[R], WO[R];
    unsigned CW[R];
#pragma unroll
    for (int k = 0; k < R; k++) {
        UE[k] = UO[k] = WE[k] = WO[k] = make_float2(0.f, 0.f);
        CW[k] = 0;
    }

    // ---- addressing: 32-bit byte offsets of this lane's quad relative to plane t0 of a source; the source (local arrays,
    // or a neighbour's memory for planes outside the owned range) only selects warp-uniform base pointers
    const unsigned nplaneB = (unsigned)g.nplane * 4u, kplaneB = (unsigned)g.kplane, cplaneB = (unsigned)g.cplane * 4u;
    unsigned on = nok ? (unsigned)(xg + yg * g.P) * 4u : 0u;       // node arrays (u, v, w), input side
    unsigned ok = kok ? (unsigned)(xg + yg * g.PC) : 0u;           // pcode
    unsigned od = dok ? (unsigned)(xg + yg * g.W) * 4u : 0u;       // density (FORCE)
    const char *bu, *bv, *bw, *bd;                                 // bases of the CURRENT source at plane t0
    auto select = [&](const PeerPlanes& s) {
        const long long o = (long long)t0 * (long long)nplaneB;
        bu = reinterpret_cast<const char*>(s.u) + o;
        bv = reinterpret_cast<const char*>(s.v) + o;
        bw = reinterpret_cast<const char*>(s.w) + o;
        bd = reinterpret_cast<const char*>(s.smoke) + (long long)t0 * (long long)cplaneB;
    };
    const bool has_lo = pr.lower.u != nullptr, has_hi = pr.upper.u != nullptr;
    if (has_lo && t0 < pr.own_lo) select(pr.lower); else select(pr.local);
    const char* const bk = reinterpret_cast<const char*>(pcode) + (long long)(t0 - g.zlo) * (long long)kplaneB;
    // planes that exist in their source: below own_lo the lower neighbour's (from its first stored plane), above own_hi the
    // upper neighbour's (up to the top of the domain), else the local stored range
    const int vlo = has_lo ? pr.lower.zlo : g.zlo, vhi = has_hi ? g.D : g.zlo + g.nzn - 1;
    const int klo = g.zlo, khi = g.zlo + g.nzc - 1;                // cell planes with a stencil code / a density

    float4 pu, pv, pw, pd;
    unsigned pc;
    auto prefetch = [&](int z) { // called for z = t0, t0 + 1, ...: every plane exactly once, in order
        if (has_lo && z == pr.own_lo) select(pr.local);
        if (has_hi && z == pr.own_hi + 1) select(pr.upper);
        const bool zn = z >= vlo && z <= vhi, zc = z >= klo && z <= khi;
        pu = pv = pw = make_float4(0.f, 0.f, 0.f, 0.f);
        pc = 0;
        if (FORCE) {
            pd = make_float4(0.f, 0.f, 0.f, 0.f);
            if (zn && zc && dok) pd = __ldg(reinterpret_cast<const float4*>(bd + od));
            od += cplaneB;
        }
        if (zn && nok) {
            pu = __ldg(reinterpret_cast<const float4*>(bu + on));
            pv = __ldg(reinterpret_cast<const float4*>(bv + on));
            pw = __ldg(reinterpret_cast<const float4*>(bw + on));
        }
        if (zc && kok) pc = __ldg(reinterpret_cast<const unsigned*>(bk + ok));
        on += nplaneB;
        ok += kplaneB;
    };

    // output side: byte offset of this lane's quad in plane t-K relative to plane t0-K of the output arrays (u, w; v is
    // written one plane behind)
    char* const ou = reinterpret_cast<char*>(uo) + (long long)(t0 - K - g.zlo) * (long long)nplaneB;
    char* const ow = reinterpret_cast<char*>(wo) + (long long)(t0 - K - g.zlo) * (long long)nplaneB;
    char* const ovb = reinterpret_cast<char*>(vo) + (long long)(t0 - K - 1 - g.zlo) * (long long)nplaneB;
    unsigned oo = nok ? (unsigned)(xg + yg * g.P) * 4u : 0u;
    float wm = 0.f;

    prefetch(t0);
    float* slot_p = sv + vrow;                  // this lane's row in the shared-memory slot of plane t
    float* const slot_end = sv + R * PLS + vrow;
    for (int t = t0; t <= t1 + 1; t++) {
        // (a) v of plane t-K-1 became final with the previous step (behind its barrier): write it out, then reuse
        //     the slot (same thread <-> same addresses, no barrier needed in between)
        {
            const int s2 = t - K - 1;
            if (s2 >= zo0 && s2 < zo1 && sok) {
                const float2 ve = *reinterpret_cast<const float2*>(slot_p), vo2 = *reinterpret_cast<const float2*>(slot_p + HO);
                *reinterpret_cast<float4*>(ovb + oo) = make_float4(ve.x, vo2.x, ve.y, vo2.y);
            }
        }
        if (t > t1) break;
        // (b) plane t enters: u, w and the code word into ring position 0, v into shared memory
        if (FORCE) { // first pass of the step: forcing + clamp on the way in
            const bool cl = yg >= 1 && yg < g.H && t >= 1 && t < g.D; // + 1 <= x < W per node
            force_clamp_node_p(pu.x, pv.x, pw.x, pc & 255u, pd.x, cl && xg >= 1 && xg < g.W, fa);
            force_clamp_node_p(pu.y, pv.y, pw.y, (pc >> 8) & 255u, pd.y, cl && xg + 1 < g.W, fa);
            force_clamp_node_p(pu.z, pv.z, pw.z, (pc >> 16) & 255u, pd.z, cl && xg + 2 < g.W, fa);
            force_clamp_node_p(pu.w, pv.w, pw.w, pc >> 24, pd.w, cl && xg + 3 < g.W, fa);
        }
        UE[0] = make_float2(pu.x, pu.z); UO[0] = make_float2(pu.y, pu.w);
        WE[0] = make_float2(pw.x, pw.z); WO[0] = make_float2(pw.y, pw.w);
        CW[0] = pc;
        *reinterpret_cast<float2*>(slot_p) = make_float2(pv.x, pv.z);
        *reinterpret_cast<float2*>(slot_p + HO) = make_float2(pv.y, pv.w);
        if (t < t1) prefetch(t + 1); // in flight during the sweeps

        // (c) sweep j runs on cell plane t-j with colour (sweep0+j-1)&1: the active x parity of a row,
        //     (y + (t-j) + sweep0 + j - 1) & 1 = (y + t + sweep0 + 1) & 1, is the same for all K sweeps of this step.
        //     Plane t-j needs sweep j only if it lies j-1 planes above t0 (t - 2j + 1 >= t0) and the warp's rows need
        //     it (j <= jmax): one sweep count per step.  No barrier between the sweeps: they touch different planes
        //     of v, and u / w are private to the lane.
        const int nsw = min(jmax, (t - t0 + 1) >> 1);
        const int par = (rowpar + t) & 1;
        float* pj = slot_p + par * HO;
        if (par) {
#pragma unroll
            for (int j = 1; j <= K; j++) {
                pj -= PLS; if (pj < sv) pj += R * PLS;
                if (j <= nsw) lean_update<1, RS>(UE[j], UO[j], WE[j], WO[j], WE[j - 1], WO[j - 1], pj, CW[j], hnz);
            }
        } else {
#pragma unroll
            for (int j = 1; j <= K; j++) {
                pj -= PLS; if (pj < sv) pj += R * PLS;
                if (j <= nsw) lean_update<0, RS>(UE[j], UO[j], WE[j], WO[j], WE[j - 1], WO[j - 1], pj, CW[j], hnz);
            }
        }

        // (d) u and w of plane t-K are final and private to this lane: write them out now
        {
            const int s = t - K;
            if (s >= zo0 && s < zo1 && sok) {
                *reinterpret_cast<float4*>(ou + oo) = make_float4(UE[K].x, UO[K].x, UE[K].y, UO[K].y);
                *reinterpret_cast<float4*>(ow + oo) = make_float4(WE[K].x, WO[K].x, WE[K].y, WO[K].y);
                if (MAXW) wm = fmaxf(fmaxf(fmaxf(wm, fabsf(WE[K].x)), fmaxf(fabsf(WO[K].x), fabsf(WE[K].y))), fabsf(WO[K].y));
            }
            oo += nplaneB;
        }
        // (e) shift the register ring
#pragma unroll
        for (int k = K; k >= 1; k--) {
            UE[k] = UE[k - 1]; UO[k] = UO[k - 1]; WE[k] = WE[k - 1]; WO[k] = WO[k - 1]; CW[k] = CW[k - 1];
        }
        // (f) v faces written in this step are read by other rows in the next one
        __syncthreads();
        slot_p += PLS; if (slot_p == slot_end) slot_p = sv + vrow;
    }
    if (MAXW) {
        for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
        if (lane == 0 && wm > 0.f) atomicMax(wmax, __float_as_uint(wm));
    }
}

template <int K, int NW, bool FORCE, bool MAXW>
__global__ void __launch_bounds__(NW * 32, 1)
k_pressure_lean(GridP g, float* __restrict__ uo, float* __restrict__ vo, float* __restrict__ wo,
                const unsigned char* __restrict__ pcode, int sweep0, int zchunk, PassRange pr, ForceArgs fa, unsigned* __restrict__ wmax)
{
    static_assert(K % 2 == 0 && NW % 2 == 0, "a pass is whole red+black pairs; both rows of a warp share the parity");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sv = reinterpret_cast<float*>(smem_raw); // [R][LY+1][RS]

    int chunk = pr.chunk_first + (int)blockIdx.z * pr.chunk_step;
    int bside = -1; // this CTA reads / serves the neighbour on that side (PassSync, kernels_pressure_reg.cuh)
    if (pr.sync.nchunks > 0) {
        if (pr.sync.first) chunk = blockIdx.z == 0 ? 0 : blockIdx.z == 1 ? pr.sync.nchunks - 1 : (int)blockIdx.z - 1;
        else chunk = (int)blockIdx.z + 1 < pr.sync.nchunks ? (int)blockIdx.z + 1 : 0; // boundary chunks last
        if (chunk == 0 && pr.sync.wait_ctr[0]) bside = 0;
        else if (chunk == pr.sync.nchunks - 1 && pr.sync.wait_ctr[1]) bside = 1;
        if (bside >= 0) {
            if (threadIdx.x == 0) {
                const long long t0c = clock64();
                unsigned v;
                do {
                    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(pr.sync.wait_ctr[bside]) : "memory");
                    if ((int)(v - pr.sync.wait_epoch) >= 0) break;
                    if (clock64() - t0c > (long long)2e10) { pr.sync.flags[1] = 1; break; }
                    __nanosleep(100);
                } while (true);
            }
            __syncthreads();
        }
    }
    const int zo0 = pr.out_lo + chunk * zchunk; // output node planes [zo0, zo1)
    const int zo1 = min(zo0 + zchunk, pr.out_hi);
    lean_pass_piece<K, NW, FORCE, MAXW>(g, uo, vo, wo, pcode, sweep0, pr, fa, sv, (int)blockIdx.x, (int)blockIdx.y, zo0, zo1, wmax);
    if (bside >= 0) { // the last boundary CTA of this side publishes the epoch
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned done = atomicAdd(pr.sync.done_ctr[bside], 1u);
            if (done == gridDim.x * gridDim.y - 1) {
                *pr.sync.done_ctr[bside] = 0;
                __threadfence_system();
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pr.sync.sig_ctr[bside]), "r"(pr.sync.sig_epoch) : "memory");
            }
        }
    }
}

} // namespace smk

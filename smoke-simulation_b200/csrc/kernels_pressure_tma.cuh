// kernels_pressure_tma.cuh -- the fused pressure pass of round 2 (default): K red/black SOR half-sweeps per launch while
// marching a 64 x 32-cell tile along z (reference: divergence cu:356-394 x 60, schedule cu:797-801).  Same schedule, tile
// geometry, lane mapping, arithmetic and proof of bit-identity as kernels_pressure_reg.cuh (trapezoid halo, out of place;
// u and w of a lane's quad in a register ring, v in shared memory).  What changed is everything AROUND the arithmetic,
// because that is what bounded the round-1 kernel (ncu: 65 % issue-active, ~490 executed instructions per warp and
// z-step of which ~180 were address arithmetic, source selection, predicates and register moves outside the sweeps and
// ~30 per sweep phase were stencil-code handling and copies; profiles/r1_final_pressure_reg_ncu_full.txt):
//
//   * INPUT PLANES ARRIVE BY TMA.  One elected thread issues cp.async.bulk.tensor.3d loads of the tile's next u, v, w
//     (and stencil-code, and -- first pass of a step -- density) plane into a shared staging ring, NS-1 planes ahead,
//     completion on an mbarrier.  Lanes read their quad with LDS.128 at immediate offsets: no per-lane global address,
//     no prefetch registers (13 + pointers in round 1), no bounds predicates -- the TMA unit zero-fills everything
//     outside the stored planes, which is exactly the "plane does not exist" convention of the sweeps.
//     Planes beyond the slab's owned range come from the NEIGHBOUR GPU's memory through tensor maps over its
//     peer-mapped arena: the halo transfer stays fused into the pass (no ghost copies), and because the loads run
//     NS-1 z-steps ahead the NVLink round trip is off the critical path (round 1: one exposed round trip per z-step
//     in the boundary chunks, 0.89 weak-scaling efficiency at N = 8).
//   * the v ring is 8 slots of 8 KB at a 64 KB-aligned shared address: the slot of plane t-j is (tt - j*8K) & mask,
//     one add + one LOP3, instead of a compare-and-wrap chain;
//   * stencil information comes in the "pcode" encoding (grid.h): "every cell of this sweep phase is ACTIVE with six
//     fluid neighbours" is one LOP3 + one vote; the rare paths -- cells next to a solid, denormal-range quotients -- are
//     out-of-line functions, so the hot path is straight-line code the compiler does not pad with copies;
//   * the lead-in / trapezoid conditions of the K sweeps of a step collapse into one per-step sweep count;
//   * output addresses are three running per-lane pointers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <type_traits>
#include "grid.h"
#include "kernels_basic.cuh"
#include "kernels_pressure_reg.cuh"
#include "kernels_advect_tma.cuh" // mbarrier / TMA helpers

#ifndef TINY_INL
#define TINY_INL __noinline__
#endif
#ifndef GEN_INL
#define GEN_INL __noinline__
#endif
namespace smk {

template <int K, int NW, bool FORCE>
struct TmaCfg {
    static_assert(NW == 16, "the v ring below is laid out for 32-row tiles");
    static constexpr int LX = 64, LY = 2 * NW;
    static constexpr int HX = (K + 3) / 4 * 4;                 // x halo in cells (whole quads)
    static constexpr int OX = LX - 2 * HX, OY = LY - 2 * K;    // output tile
    static constexpr int NS = FORCE ? 3 : 4;                   // staging slots: planes in flight = NS - 1
    static constexpr int FB = LX * LY * 4;                     // bytes of one staged field plane (8 KB)
    static constexpr int KW = 80;                              // staged stencil-code row: 80 bytes from a 16-byte aligned x
    static constexpr int KB = (KW * LY + 127) / 128 * 128;
    static constexpr int SLOT = 3 * FB + KB + (FORCE ? FB : 0);
    static constexpr unsigned RING_ABS = 0x20000u;             // shared-window address of the v ring (64 KB aligned)
    static constexpr unsigned RING_SLOT = 0x2000u, RING_MASK = 0xE000u; // 8 slots of 32 rows x 256 B: E[0..31] | O[0..31]
    static constexpr unsigned DUMMY_ABS = 0x30000u;            // 2 rows for the lanes of tile row LY-1 (no row above them)
    static constexpr unsigned BAR_ABS = 0x30200u;              // NS mbarriers
    static constexpr unsigned SMEM_END = 0x30280u;
    static constexpr int THREADS = NW * 32;
    static_assert(NS * SLOT + 2048 <= (int)RING_ABS, "staging must fit below the v ring");
};

// tensor maps of one launch: box = one tile plane.  Kernel parameter (__grid_constant__): the TMA unit reads them from there.
struct PassMaps {
    CUtensorMap loc[3];   // u, v, w of the slab's own "in" buffers                       box 64 x 32 x 1 floats
    CUtensorMap lo[3];    // ... of the lower / upper neighbour (peer-mapped memory; unused copies of loc[] without one)
    CUtensorMap hi[3];
    CUtensorMap pcode;    // stencil codes (pcode)                                        box 80 x 32 x 1 bytes
    CUtensorMap smoke[3]; // density "now": own, lower neighbour's, upper neighbour's     box 64 x 32 x 1 floats
};

// Packed pairs live in 64-bit registers from the moment they are loaded until they are stored: the f32x2 instructions
// take .b64 operands, and every float2 <-> .b64 conversion in the source became a pair of register moves in the SASS of the
// first version of this kernel (the per-sweep "copies" of round 1's profile).  Hence a tiny vocabulary on b64 values.
typedef unsigned long long b64;
__device__ __forceinline__ b64 pk(float lo, float hi)
{
    b64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo32(b64 v)
{
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return a;
}
__device__ __forceinline__ float hi32(b64 v)
{
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return b;
}
__device__ __forceinline__ b64 ffma2(b64 a, b64 b, b64 c) // RN(a*b + c) per half
{
    b64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ b64 fadd2(b64 a, b64 b)
{
    b64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ b64 fmul2(b64 a, b64 b)
{
    b64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
template <int OFF>
__device__ __forceinline__ b64 lds64(unsigned a)
{
    b64 v;
    asm volatile("ld.shared.b64 %0, [%1+%2];" : "=l"(v) : "r"(a), "n"(OFF) : "memory");
    return v;
}
template <int OFF>
__device__ __forceinline__ void sts64(unsigned a, b64 v)
{
    asm volatile("st.shared.b64 [%0+%1], %2;" ::"r"(a), "n"(OFF), "l"(v) : "memory");
}
template <int OFF>
__device__ __forceinline__ void sts32(unsigned a, float v)
{
    asm volatile("st.shared.f32 [%0+%1], %2;" ::"r"(a), "n"(OFF), "f"(v) : "memory");
}
// p = (float)((double)q * -1.9) per half (cu:384; DMUL by the negated constant like the reference's SASS)
__device__ __forceinline__ b64 p_from_q(b64 q)
{
    return pk(__double2float_rn(__dmul_rn((double)lo32(q), M19)), __double2float_rn(__dmul_rn((double)hi32(q), M19)));
}

// bounded mbarrier wait, not unrolled (the first probe almost always succeeds: the loads run NS-1 planes ahead)
__device__ __forceinline__ bool mbar_wait1(unsigned long long* bar, unsigned parity)
{
    unsigned ok = 0;
#pragma unroll 1
    for (int it = 0; it < (1 << 26) && !ok; it++)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// ---- rare paths, out of line (they cost a call only where they are taken) ---------------------------------------------
// denormal-range quotient of an "open" cell (acc = 6): the exact integer quotient, see div6_tiny
__device__ TINY_INL b64 tiny_fix_pair(b64 d2, b64 P2)
{
    const float dx = lo32(d2), dy = hi32(d2);
    float px = lo32(P2), py = hi32(P2);
    const unsigned ax = (__float_as_uint(dx) & 0x7fffffffu) - 1u, ay = (__float_as_uint(dy) & 0x7fffffffu) - 1u;
    if (ax < 0x00ffffffu) px = __double2float_rn(__dmul_rn((double)div6_tiny(dx), M19));
    if (ay < 0x00ffffffu) py = __double2float_rn(__dmul_rn((double)div6_tiny(dy), M19));
    return pk(px, py);
}
// general update of a pair: per-cell neighbour count, per-face masks (a masked face gets its old value back).
// cs: pcode of cell A in byte 0, of cell B in byte 2.  f = {U0, U1, V0, V1, W0, W1}, updated in place.
struct PairFaces { float2 f[6]; };
__device__ GEN_INL void general_update_pair(PairFaces& io, b64 d2, unsigned cs)
{
    const float2 d = make_float2(lo32(d2), hi32(d2));
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
    if (nA == 6 && (ax - 1u < 0x00ffffffu)) P.x = __double2float_rn(__dmul_rn((double)div6_tiny(d.x), M19)); // only acc = 6 can miss
    if (nB == 6 && (ay - 1u < 0x00ffffffu)) P.y = __double2float_rn(__dmul_rn((double)div6_tiny(d.y), M19));
    const float2 M1 = make_float2(-1.0f, -1.0f);
    // old - 0 and old + 0 return old (up to the sign of a zero): a masked face keeps its value
    io.f[0] = __ffma2_rn(make_float2((cA & CODE_SX0) ? P.x : 0.f, (cB & CODE_SX0) ? P.y : 0.f), M1, io.f[0]);
    io.f[1] = __fadd2_rn(io.f[1], make_float2((cA & CODE_SX1) ? P.x : 0.f, (cB & CODE_SX1) ? P.y : 0.f));
    io.f[2] = __ffma2_rn(make_float2((cA & CODE_SY0) ? P.x : 0.f, (cB & CODE_SY0) ? P.y : 0.f), M1, io.f[2]);
    io.f[3] = __fadd2_rn(io.f[3], make_float2((cA & CODE_SY1) ? P.x : 0.f, (cB & CODE_SY1) ? P.y : 0.f));
    io.f[4] = __ffma2_rn(make_float2((cA & CODE_SZ0) ? P.x : 0.f, (cB & CODE_SZ0) ? P.y : 0.f), M1, io.f[4]);
    io.f[5] = __fadd2_rn(io.f[5], make_float2((cA & CODE_SZ1) ? P.x : 0.f, (cB & CODE_SZ1) ? P.y : 0.f));
}
__device__ __forceinline__ float2 f2(b64 v) { return make_float2(lo32(v), hi32(v)); }
__device__ __forceinline__ b64 pk(float2 v) { return pk(v.x, v.y); }

// rare fix-ups of the fast path, out of line: denormal-range quotients (exact integer quotient) and cells that are not
// updated (P = 0: old -/+ 0 = old).  act_lo / act_hi: the ACTIVE bits of the two cells.
__device__ __noinline__ b64 rare_fix_pair(b64 d2, b64 P2, unsigned act_lo, unsigned act_hi)
{
    const float dx = lo32(d2), dy = hi32(d2);
    float px = lo32(P2), py = hi32(P2);
    const unsigned ax = (__float_as_uint(dx) & 0x7fffffffu) - 1u, ay = (__float_as_uint(dy) & 0x7fffffffu) - 1u;
    if (ax < 0x00ffffffu) px = __double2float_rn(__dmul_rn((double)div6_tiny(dx), M19));
    if (ay < 0x00ffffffu) py = __double2float_rn(__dmul_rn((double)div6_tiny(dy), M19));
    return pk(act_lo ? px : 0.f, act_hi ? py : 0.f);
}

// One sweep phase of one lane on ring position J (plane t-J): two same-colour cells of the lane's quad, in two parts.
// PAR = x parity of the active colour (warp uniform).  a = shared address of this lane's E pair of the plane's row
// (the parity offset is an immediate).  Pairs: E = (f[4h], f[4h+2]), O = (f[4h+1], f[4h+3]).
//
// Part A -- everything that does NOT depend on the previous sweep of the same z-step: the v faces, the u face of the
// next quad, the stencil tests and four of the five additions of the divergence (cu:379-381 adds w1 LAST, and w1 = the
// w0 of the sweep before is the only value two consecutive sweeps of a step share).  The step issues part A of sweep j+1
// in front of part B of sweep j, in straight-line code: the independent loads, shuffles and adds fill the latency of
// the serial chain (add, 3 x f32x2, F2F, DMUL, F2F, update w0) instead of queueing behind it.
struct SweepA {
    b64 V0, V1, U1, dA;
    bool allact, simple;
};
template <int PAR, bool GENERAL>
__device__ __forceinline__ SweepA sweep_a(const b64 ue, const b64 uo, const b64 we, const b64 wo, const unsigned a, const unsigned cw)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr unsigned SH = 8u * PAR;                 // byte 0/2 (PAR 0) or 1/3 (PAR 1) of the code word
    constexpr unsigned M_AC = 0x00C000C0u << SH;      // ACTIVE | COMPLEX of both cells
    constexpr unsigned V_A = 0x00400040u << SH;       // ... == ACTIVE, not COMPLEX
    constexpr unsigned M_C = 0x00800080u << SH;
    const b64 M1 = pk(-1.0f, -1.0f);
    SweepA r;
    r.V0 = lds64<PAR * 128>(a);
    r.V1 = lds64<PAR * 128 + 256>(a);
    const b64 U0 = PAR == 0 ? ue : uo, W0 = PAR == 0 ? we : wo;
    r.U1 = PAR == 0 ? uo : pk(hi32(ue), __shfl_down_sync(FULL, lo32(ue), 1)); // PAR 1: u[4h+2], u[4h+4] (next quad's first face)
    b64 d = ffma2(U0, M1, r.U1);         // -u0 + u1           (cu:379-381, left to right, one rounding each)
    d = ffma2(r.V0, M1, d);              //  ... - v0
    d = fadd2(d, r.V1);                  //  ... + v1
    r.dA = ffma2(W0, M1, d);             //  ... - w0
    // tier A: both cells of every lane ACTIVE with six fluid neighbours; tier B: no COMPLEX cell (some are not updated)
    r.allact = __all_sync(FULL, (cw & M_AC) == V_A);
    r.simple = !GENERAL || r.allact || __all_sync(FULL, (cw & (cw << 1) & M_C) == 0u);
    return r;
}
template <int PAR, bool GENERAL>
__device__ __forceinline__ void sweep_b(const SweepA& A, b64& ue, b64& uo, b64& we, b64& wo, b64& we1, b64& wo1, const unsigned a,
                                        const unsigned cw, const bool hnz)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr unsigned SH = 8u * PAR;
    const b64 M1 = pk(-1.0f, -1.0f), R6 = pk(0x1.555556p-3f, 0x1.555556p-3f) /* RN(1/6) */, M6 = pk(-6.0f, -6.0f);
    b64& U0r = PAR == 0 ? ue : uo;
    b64& W0r = PAR == 0 ? we : wo;
    b64& W1r = PAR == 0 ? we1 : wo1;
    const b64 d = fadd2(A.dA, W1r);      //  ... + w1
    b64 U1 = A.U1, V0 = A.V0, V1 = A.V1;
    if (A.simple) {
        // q = d / 6 by reciprocal + one correction (exact for |d| >= 2^-125: exhaustive check)
        const b64 q0 = fmul2(d, R6);
        const b64 rem = ffma2(q0, M6, d);
        const b64 q = ffma2(rem, R6, q0);
        b64 P = p_from_q(q);
        // below 2^-125 (and d != 0) a tie on the denormal grid can round the wrong way; cells that are not ACTIVE get
        // P = 0: one branch for both rare cases
        const unsigned ax = (__float_as_uint(lo32(d)) & 0x7fffffffu) - 1u, ay = (__float_as_uint(hi32(d)) & 0x7fffffffu) - 1u;
        if (__any_sync(FULL, min(ax, ay) < 0x00ffffffu) || !A.allact) P = rare_fix_pair(d, P, cw & (0x40u << SH), cw & (0x400000u << SH));
        W0r = ffma2(P, M1, W0r);         // first: the next sweep of this step waits for it
        W1r = fadd2(W1r, P);
        V0 = ffma2(P, M1, V0); V1 = fadd2(V1, P);
        U0r = ffma2(P, M1, U0r); U1 = fadd2(U1, P);
    } else {
        PairFaces io;
        io.f[0] = f2(U0r); io.f[1] = f2(U1); io.f[2] = f2(V0); io.f[3] = f2(V1); io.f[4] = f2(W0r); io.f[5] = f2(W1r);
        general_update_pair(io, d, cw >> SH);
        U0r = pk(io.f[0]); U1 = pk(io.f[1]); V0 = pk(io.f[2]); V1 = pk(io.f[3]); W0r = pk(io.f[4]); W1r = pk(io.f[5]);
    }
    sts64<PAR * 128>(a, V0);
    sts64<PAR * 128 + 256>(a, V1);
    if (PAR == 0) uo = U1;
    else {
        const float from_left = __shfl_up_sync(FULL, hi32(U1), 1); // the left quad's updated u[4h]
        ue = pk(hnz ? from_left : lo32(ue), lo32(U1));              // h == 0: tile edge, face stays stale (halo)
    }
}

__device__ __forceinline__ void force_clamp_node_pc(float& u, float& v, float& w, unsigned pc, float d, bool clampable, const ForceArgs& fa)
{
    // the cell is fluid iff (pcode & 0xC0) != 0 (grid.h)
    force_clamp_node(u, v, w, ((pc & 0xC0u) ? CODE_SELF : 0u) | (pc & CODE_SY0), d, clampable, fa);
}

// One PIECE of a pass: the tile (bx, by) marched over the output node planes [zo0, zo1) (K lead-in planes below, K - 1
// above).  MAXW: also reduce max |w| over the planes written (bound of the next advection's backtrace in z).
template <int K, int NW, bool FORCE, bool MAXW, bool GENERAL>
__device__ __forceinline__ void tma_pass_piece(const GridP& g, const PassMaps& maps, float* __restrict__ uo, float* __restrict__ vo,
                                               float* __restrict__ wo, int sweep0, const PassRange& pr, const ForceArgs& fa,
                                               unsigned char* __restrict__ smem, int bx, int by, int zo0, int zo1,
                                               unsigned* __restrict__ wmax, int* __restrict__ flags)
{
    using C = TmaCfg<K, NW, FORCE>;
    constexpr int LY = C::LY, R = K + 1, NS = C::NS;
    // (the warp index through a shuffle: the compiler then knows that everything derived from it is warp-uniform and
    // emits plain uniform branches around the sweeps instead of divergence bookkeeping)
    const int lane = threadIdx.x & 31, wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int h = lane & 15;                                 // quad index inside the row
    // rows of a warp and the deepest sweep they need: kernels_pressure_reg.cuh (trapezoid halo, shallow rows paired)
    constexpr bool SKIP = (K == 4 && NW >= 8 && NW % 4 == 0);
    const int hb = lane >> 4;
    const int yl = !SKIP ? wid + hb * NW : wid >= 4 ? wid + hb * ((LY - 8) / 2) : hb == 0 ? wid : wid == 3 ? LY - 1 : LY - 2 - wid;
    const int jmax = (SKIP && wid < 3) ? wid + 1 : K;
    const int x0 = bx * C::OX - C::HX;
    const int y0 = by * C::OY - K;
    const int t0 = zo0 - K, t1 = zo1 + K - 1;                // planes that enter the ring
    const int xg = x0 + 4 * h, yg = y0 + yl;
    const int x0k = x0 & ~15;                                // first staged code byte of a row (16-byte aligned, <= x0)

    const bool nok = xg >= 0 && xg <= g.P - 4 && yg >= 0 && yg < g.SY;
    const bool sok = nok && yl >= K && yl < LY - K && h >= C::HX / 4 && h < 16 - C::HX / 4;
    // (+ t) = x parity of the active colour; both rows of a warp share it (broadcast: warp uniform for the compiler too)
    const int rowpar = __shfl_sync(0xffffffffu, (y0 + yl + sweep0 + 1) & 1, 0);
    const bool hnz = h != 0;

    // ---- shared memory: [staging ring | ... | v ring at RING_ABS | dummy rows | mbarriers]
    const unsigned dyn0 = smem_u32(smem);
    unsigned char* const stage = smem;                                                  // NS slots of C::SLOT bytes
    unsigned long long* const bars = reinterpret_cast<unsigned long long*>(smem + (C::BAR_ABS - dyn0));
    {   // defined values everywhere in the v ring (dummy rows, halo rows of planes that never enter)
        float4* z4 = reinterpret_cast<float4*>(smem + (C::RING_ABS - dyn0));
        for (int i = threadIdx.x; i < (int)(C::BAR_ABS - C::RING_ABS) / 16; i += C::THREADS) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; i++) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // TMA producer (thread 0): plane z of the current source into staging slot `slot`
    const bool has_lo = pr.lower.u != nullptr, has_hi = pr.upper.u != nullptr;
    auto issue = [&](int z, int slot) {
        const CUtensorMap* m = maps.loc;
        const CUtensorMap* ms = &maps.smoke[0];
        int zr = z - g.zlo;
        if (has_lo && z < pr.own_lo) { m = maps.lo; ms = &maps.smoke[1]; zr = z - pr.lower.zlo; }
        else if (has_hi && z > pr.own_hi) { m = maps.hi; ms = &maps.smoke[2]; zr = z - pr.upper.zlo; }
        unsigned char* dst = stage + slot * C::SLOT;
        mbar_expect_tx(&bars[slot], 3u * C::FB + (unsigned)(C::KW * LY) + (FORCE ? (unsigned)C::FB : 0u));
        tma_load_3d(dst, &m[0], x0, y0, zr, &bars[slot]);
        tma_load_3d(dst + C::FB, &m[1], x0, y0, zr, &bars[slot]);
        tma_load_3d(dst + 2 * C::FB, &m[2], x0, y0, zr, &bars[slot]);
        tma_load_3d(dst + 3 * C::FB, &maps.pcode, x0k, y0, z - g.zlo, &bars[slot]);
        if (FORCE) tma_load_3d(dst + 3 * C::FB + C::KB, ms, x0, y0, zr, &bars[slot]);
    };
    if (threadIdx.x == 0)
        for (int i = 0; i < NS - 1 && t0 + i <= t1; i++) issue(t0 + i, i);

    // register ring, index k = plane t-k
    b64 UE[R], UO[R], WE[R], WO[R];
    unsigned CW[R];
#pragma unroll
    for (int k = 0; k < R; k++) {
        UE[k] = UO[k] = WE[k] = WO[k] = 0ull;
        CW[k] = 0;
    }

    // per-lane constants of the shared-memory accesses
    const unsigned lq = (unsigned)(yl * 256 + 8 * h);                  // this lane's E pair inside a v-ring slot
    const bool top = yl == LY - 1;                                      // no row above: sweeps go to the dummy rows
    const unsigned sw_base = top ? C::DUMMY_ABS + 8u * h : C::RING_ABS + lq;
    const unsigned sw_mask = top ? 0u : C::RING_MASK;
    const unsigned char* const lst = stage + (yl * 64 + 4 * h) * 4;     // this lane's quad inside a staged field plane
    const unsigned char* const lkc = stage + 3 * C::FB + yl * C::KW + (x0 - x0k) + 4 * h; // ... its four code bytes

    // output: running pointers to this lane's quad in plane t-K (u, w) and t-K-1 (v)
    const long long o0 = (long long)(t0 - K - g.zlo) * g.nplane + (nok ? xg + yg * g.P : 0);
    float* pou = uo + o0;
    float* pow_ = wo + o0;
    float* pov = vo + (o0 - g.nplane);
    float wm = 0.f;

    unsigned tt = 0;            // (t - t0) << 13: slot bits of plane t in the v ring (masked)
    int ss = 0;                 // staging slot of plane t
    unsigned ph = 0;            // parity bits of the NS mbarriers (bit i flips every time slot i is consumed)
    int t = t0;
    bool failed = false;

    // One z-step.  ROT = (t - t0) mod R and PAR = x parity of the active colour are COMPILE-TIME: plane p lives in the
    // physical ring entry p mod R for its whole life (position k of step t is entry (ROT - k) mod R), so the ring is never
    // shifted -- the loop below is unrolled over the 2R combinations instead (round 1 moved 36 registers per step).
    // Returns false when the piece is finished (or a TMA transaction was lost).
    auto step = [&](auto rot_c, auto par_c) -> bool {
        constexpr int ROT = decltype(rot_c)::value, PAR = decltype(par_c)::value;
        auto ph_of = [](int k) { return ((ROT - k) % R + R) % R; };
        // (a) v of plane t-K-1 became final with the previous step (behind its barrier): write it out
        {
            const int s2 = t - K - 1;
            if (s2 >= zo0 && s2 < zo1 && sok) {
                const unsigned a = C::RING_ABS + ((tt - (unsigned)(K + 1) * C::RING_SLOT) & C::RING_MASK) + lq;
                const b64 ve = lds64<0>(a), vq = lds64<128>(a);
                *reinterpret_cast<float4*>(pov) = make_float4(lo32(ve), lo32(vq), hi32(ve), hi32(vq));
            }
        }
        if (t > t1) return false;
        // (b) plane t enters: wait for its staged copy; u, w and the code word go to ring position 0, v to the v ring.
        //     The producer is a lane of warp 0, whose rows need one sweep per step instead of K: it has the time.
        if (threadIdx.x == 0 && t + NS - 1 <= t1) issue(t + NS - 1, ss == 0 ? NS - 1 : ss - 1); // the slot plane t-1 just left
        if (!mbar_wait1(&bars[ss], (ph >> ss) & 1u)) { flags[2] = 1; failed = true; return false; }
        ph ^= 1u << ss;
        float4 pu = *reinterpret_cast<const float4*>(lst + ss * C::SLOT);
        float4 pv = *reinterpret_cast<const float4*>(lst + ss * C::SLOT + C::FB);
        float4 pw = *reinterpret_cast<const float4*>(lst + ss * C::SLOT + 2 * C::FB);
        unsigned pc = *reinterpret_cast<const unsigned*>(lkc + ss * C::SLOT);
        if (FORCE) { // first pass of the step: forcing + clamp on the way in
            const float4 pd = *reinterpret_cast<const float4*>(lst + ss * C::SLOT + 3 * C::FB + C::KB);
            const bool cl = yg >= 1 && yg < g.H && t >= 1 && t < g.D; // + 1 <= x < W per node
            force_clamp_node_pc(pu.x, pv.x, pw.x, pc & 255u, pd.x, cl && xg >= 1 && xg < g.W, fa);
            force_clamp_node_pc(pu.y, pv.y, pw.y, (pc >> 8) & 255u, pd.y, cl && xg + 1 < g.W, fa);
            force_clamp_node_pc(pu.z, pv.z, pw.z, (pc >> 16) & 255u, pd.z, cl && xg + 2 < g.W, fa);
            force_clamp_node_pc(pu.w, pv.w, pw.w, pc >> 24, pd.w, cl && xg + 3 < g.W, fa);
        }
        UE[ph_of(0)] = pk(pu.x, pu.z); UO[ph_of(0)] = pk(pu.y, pu.w);
        WE[ph_of(0)] = pk(pw.x, pw.z); WO[ph_of(0)] = pk(pw.y, pw.w);
        CW[ph_of(0)] = pc;
        {
            const unsigned a = C::RING_ABS + (tt & C::RING_MASK) + lq;
            sts32<0>(a, pv.x); sts32<4>(a, pv.z); sts32<128>(a, pv.y); sts32<132>(a, pv.w);
        }
        // (c) sweep j runs on cell plane t-j with colour (sweep0+j-1)&1: the active x parity of a row,
        //     (y + (t-j) + sweep0 + j - 1) & 1 = (y + t + sweep0 + 1) & 1, is the same for all K sweeps of this step.
        //     No barrier between the sweeps: they touch different planes of v, and u / w are private to the lane.
        //     ALL K sweeps run in every step and on every row: the ones the trapezoid halo does not need (sweep j of a
        //     plane less than j-1 above t0, or of a row closer than j to the tile edge) only ever feed values that are
        //     not needed either -- which is why round 1 could skip them; here they buy straight-line code, and the warps
        //     they would have spared wait at the step's barrier anyway.
        {
            auto adr = [&](int j) { return ((tt - (unsigned)j * C::RING_SLOT) & sw_mask) | sw_base; };
            SweepA A = sweep_a<PAR, GENERAL>(UE[ph_of(1)], UO[ph_of(1)], WE[ph_of(1)], WO[ph_of(1)], adr(1), CW[ph_of(1)]);
#pragma unroll
            for (int j = 1; j <= K; j++) {
                SweepA An = A;
                if (j < K) An = sweep_a<PAR, GENERAL>(UE[ph_of(j + 1)], UO[ph_of(j + 1)], WE[ph_of(j + 1)], WO[ph_of(j + 1)], adr(j + 1), CW[ph_of(j + 1)]);
                sweep_b<PAR, GENERAL>(A, UE[ph_of(j)], UO[ph_of(j)], WE[ph_of(j)], WO[ph_of(j)], WE[ph_of(j - 1)], WO[ph_of(j - 1)], adr(j),
                                      CW[ph_of(j)], hnz);
                A = An;
            }
        }
        // (d) u and w of plane t-K are final and private to this lane: write them out now
        {
            const int s = t - K;
            if (s >= zo0 && s < zo1 && sok) {
                const b64 ue = UE[ph_of(K)], uq = UO[ph_of(K)], we = WE[ph_of(K)], wq = WO[ph_of(K)];
                *reinterpret_cast<float4*>(pou) = make_float4(lo32(ue), lo32(uq), hi32(ue), hi32(uq));
                *reinterpret_cast<float4*>(pow_) = make_float4(lo32(we), lo32(wq), hi32(we), hi32(wq));
                if (MAXW) wm = fmaxf(fmaxf(fmaxf(wm, fabsf(lo32(we))), fmaxf(fabsf(lo32(wq)), fabsf(hi32(we)))), fabsf(hi32(wq)));
            }
            pou += g.nplane; pow_ += g.nplane; pov += g.nplane;
        }
        // (f) v faces written in this step are read by other rows in the next one; the staged copy of plane t is free
        __syncthreads();
        tt += C::RING_SLOT;
        if (++ss == NS) ss = 0;
        t++;
        return true;
    };
    // unrolled over ring rotation x parity: step s of the pattern has ROT = s mod R, PAR = (s + c) & 1 with
    // c = parity of this warp's first step.  A warp with c = 1 enters the pattern at s = R (same rotation, other parity).
    static_assert(R % 2 == 1, "the 2R pattern needs an odd ring depth");
    bool second_half_only = ((rowpar + t0) & 1) != 0;
    if (GENERAL) { // the rare CTAs with COMPLEX cells: compact code -- one step per iteration, ring shifted by register moves
        for (;;) {
            const bool go = ((rowpar + t) & 1) ? step(std::integral_constant<int, 0>{}, std::integral_constant<int, 1>{})
                                               : step(std::integral_constant<int, 0>{}, std::integral_constant<int, 0>{});
            if (!go) break;
            // position k of the next step = position k-1 of this one; with ROT = 0 position k is entry (R - k) % R
#pragma unroll
            for (int k = K; k >= 1; k--) {
                const int to = (R - k) % R, from = (R - (k - 1)) % R;
                UE[to] = UE[from]; UO[to] = UO[from]; WE[to] = WE[from]; WO[to] = WO[from]; CW[to] = CW[from];
            }
        }
    } else
    for (;;) {
        if (!second_half_only) {
            if (!step(std::integral_constant<int, 0 % R>{}, std::integral_constant<int, 0>{})) break;
            if (!step(std::integral_constant<int, 1 % R>{}, std::integral_constant<int, 1>{})) break;
            if (!step(std::integral_constant<int, 2 % R>{}, std::integral_constant<int, 0>{})) break;
            if (R > 3) {
                if (!step(std::integral_constant<int, 3 % R>{}, std::integral_constant<int, 1>{})) break;
                if (!step(std::integral_constant<int, 4 % R>{}, std::integral_constant<int, 0>{})) break;
            }
        }
        second_half_only = false;
        if (!step(std::integral_constant<int, 0 % R>{}, std::integral_constant<int, 1>{})) break;
        if (!step(std::integral_constant<int, 1 % R>{}, std::integral_constant<int, 0>{})) break;
        if (!step(std::integral_constant<int, 2 % R>{}, std::integral_constant<int, 1>{})) break;
        if (R > 3) {
            if (!step(std::integral_constant<int, 3 % R>{}, std::integral_constant<int, 0>{})) break;
            if (!step(std::integral_constant<int, 4 % R>{}, std::integral_constant<int, 1>{})) break;
        }
    }
    if (failed) return;
    if (MAXW) {
        for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
        if (lane == 0 && wm > 0.f) atomicMax(wmax, __float_as_uint(wm));
    }
}

template <int K, int NW, bool FORCE, bool MAXW>
__global__ void __launch_bounds__(NW * 32, 1)
k_pressure_tma(GridP g, const __grid_constant__ PassMaps maps, float* __restrict__ uo, float* __restrict__ vo, float* __restrict__ wo,
               int sweep0, int zchunk, PassRange pr, ForceArgs fa, unsigned* __restrict__ wmax, int* __restrict__ flags,
               const unsigned char* __restrict__ cflag)
{
    static_assert(K % 2 == 0 && NW % 2 == 0, "a pass is whole red+black pairs; both rows of a warp share the parity");
    using C = TmaCfg<K, NW, FORCE>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (smem_u32(smem_raw) > 2048u) { // the layout is in absolute shared-window addresses: the dynamic part must start low
        if (threadIdx.x == 0) flags[2] = 2;
        return;
    }

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
                // the neighbour's planes are read by the TMA unit (async proxy): order its reads behind the acquire
                asm volatile("fence.proxy.async;" ::: "memory");
            }
            __syncthreads();
        }
    }
    const int zo0 = pr.out_lo + chunk * zchunk; // output node planes [zo0, zo1)
    const int zo1 = min(zo0 + zchunk, pr.out_hi);
    if (zo0 < zo1) {
        // Does any plane this piece loads hold a COMPLEX cell inside the tile (halo included)?  cflag[z][by][bx] is written
        // with the stencil codes (k_codes*, K = 4 tile geometry); without it the general update is always compiled in.
        int cx = 1;
        if (K == 4 && cflag) {
            cx = 0;
            const int za = max(zo0 - K, g.zlo), zb = min(zo1 + K - 1, g.zlo + g.nzc - 1);
            const int per = (int)(gridDim.x * gridDim.y), me = (int)(blockIdx.y * gridDim.x + blockIdx.x);
            for (int z = za + (int)threadIdx.x; z <= zb; z += (int)blockDim.x) cx |= cflag[(long long)(z - g.zlo) * per + me];
            cx = __syncthreads_or(cx);
        }
        if (cx) tma_pass_piece<K, NW, FORCE, MAXW, true>(g, maps, uo, vo, wo, sweep0, pr, fa, smem_raw, (int)blockIdx.x, (int)blockIdx.y, zo0, zo1, wmax, flags);
        else tma_pass_piece<K, NW, FORCE, MAXW, false>(g, maps, uo, vo, wo, sweep0, pr, fa, smem_raw, (int)blockIdx.x, (int)blockIdx.y, zo0, zo1, wmax, flags);
    }
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

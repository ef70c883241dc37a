// kernels_pressure_tma.cuh -- the fused pressure pass of round 2 (default): K red/black SOR half-sweeps per launch while
// marching a 64 x 32-cell tile along z (reference: divergence cu:356-394 x 60, schedule cu:797-801).  Same schedule, tile
// geometry, lane mapping, arithmetic and proof of bit-identity as kernels_pressure_reg.cuh (trapezoid halo, out of place;
// u and w of a lane's quad in a register ring, v in shared memory).  What changed is everything AROUND the arithmetic,
// because that is what bounded the round-1 kernel (ncu: 65 % issue-active, ~490 executed instructions per warp and
// z-step of which ~180 were address arithmetic, source selection, predicates and register moves outside the sweeps and
// ~30 per sweep phase were stencil-code handling and copies; profiles/r1_final_pressure_reg_ncu_full.txt):
//
//   * INPUT PLANES ARRIVE BY TMA.  One elected thread issues ONE cp.async.bulk.tensor.4d load (x, y, z, field) for the
//     tile's next u, v, w plane (plus one for the stencil codes and -- first pass of a step -- the density) into a
//     shared staging ring, NS-1 planes ahead, completion on an mbarrier.  Lanes read their quad with LDS.128 at immediate offsets: no per-lane global address,
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
//   * the register ring is never shifted in pieces without COMPLEX cells (loop unrolled over the K+2 rotations); pieces
//     with COMPLEX cells (the floor row of every scene) run one compact step body per colour parity and shift the ring by
//     moves -- unrolled, their code (every tier inline) overflowed the instruction cache;
//   * p = (float)((double)q * -1.9) is evaluated exactly in binary32 (p_from_q: two fused roundings that bracket the doubly
//     rounded product + a tie-to-even select; all 2^32 inputs checked on CPU and GPU), the double-precision sequence
//     remains for denormal-range divergences only;
//   * pieces of the bottom tile row are dispatched first, and on one GPU the z-chunks have unequal lengths (long first, one
//     short last) chosen by a simulation of the hardware's dispatch order (smk_api.cu, pick_chunks_tma).
// Measurements, the experiments that did not work and the instruction budget of a z-step: profiles/r2_tma_pass_final.txt.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <type_traits>
#include "grid.h"
#include "kernels_basic.cuh"
#include "kernels_pressure_reg.cuh"
#include "kernels_advect_tma.cuh" // mbarrier / TMA helpers

#ifndef GEN_INL
#define GEN_INL __forceinline__
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
    static constexpr unsigned BAR_ABS = 0x30200u;              // NS mbarriers of the staging slots + the step barrier
    static constexpr unsigned SMEM_END = 0x30280u;
    static constexpr int THREADS = NW * 32;
    static_assert(NS * SLOT + 2048 <= (int)RING_ABS, "staging must fit below the v ring");
};

// tensor maps of one launch: box = one tile plane.  Kernel parameter (__grid_constant__): the TMA unit reads them from there.
struct PassMaps {
    // u, v, w of one source as ONE 4-D tensor (x, y, z, field): the three arrays of a buffer set lie a constant stride
    // apart in the arena, so a single TMA instruction brings the plane of all three (box 64 x 32 x 1 x 3 floats).
    // Measured with one instruction per array: the producer lane needed ~1550 cycles per z-step and the whole CTA
    // waited for it.  src 0 = the slab's own "in" buffers, 1 / 2 = the lower / upper neighbour's (peer-mapped memory).
    CUtensorMap uvw[3];
    CUtensorMap pcode;    // stencil codes (pcode)                                        box 80 x 32 x 1 bytes
    CUtensorMap smoke[3]; // density "now": own, lower neighbour's, upper neighbour's     box 64 x 32 x 1 floats
};

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int x, int y, int z, int f, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(f), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

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
// p = (float)((double)q * -1.9) per half (cu:384; DMUL by the negated constant like the reference's SASS): the exact
// sequence, 4 F2F + 2 DMUL.  F2F runs at 16 lanes/clk/SM and the three instructions sit in the middle of the dependent
// chain of a sweep (~110-150 cycles measured): the sweeps use p_from_q below and keep this one for denormal-range inputs.
__device__ __forceinline__ b64 p_from_q_f64(b64 q)
{
    return pk(__double2float_rn(__dmul_rn((double)lo32(q), M19)), __double2float_rn(__dmul_rn((double)hi32(q), M19)));
}
// The same value in binary32 arithmetic, for q = 0 and |q| >= 2^-99 (tools/experiments/omega_fp32_exhaustive.c checks
// all 2^32 inputs against the double product, zeros' signs included).  -1.9 (binary64) = ch + cl + ...;
//     P+ = fma(q, ch, q * (cl + 2^-39)),   P- = fma(q, ch, q * (cl - 2^-39))
// bracket the doubly rounded product: if they agree that is the answer.  If not, a binary32 rounding boundary lies within
// 2^-40 |p| of the product; 19 q / 10 lives on a lattice of tenths of an ulp, so it is an exact tie of 19 q / 10, the
// binary64 product (off by 8.9e-17 relative, less than half an ulp of binary64) rounds ONTO the midpoint and the
// conversion breaks the tie to even: take the even one of the two adjacent values.  No branch, no F2F, no DMUL.
__device__ __forceinline__ unsigned even_of(unsigned bp, unsigned bm)
{
    return (bp & 1u) == 0u ? bp : bm; // equal: that value; adjacent bit patterns: exactly one of them is even
}
__device__ __forceinline__ b64 p_from_q(b64 q)
{
    const b64 CH = pk(-0x1.e66666p+0f, -0x1.e66666p+0f), CLP = pk(-0x1.99919ap-26f, -0x1.99919ap-26f), CLM = pk(-0x1.99a19ap-26f, -0x1.99a19ap-26f);
    const b64 pp = ffma2(q, CH, fmul2(q, CLP)), pm = ffma2(q, CH, fmul2(q, CLM));
    return pk(__uint_as_float(even_of(__float_as_uint(lo32(pp)), __float_as_uint(lo32(pm)))),
              __uint_as_float(even_of(__float_as_uint(hi32(pp)), __float_as_uint(hi32(pm)))));
}
// self-check (smk_selfcheck_omega): every binary32 q of [first, first + count) through both evaluations on the device
__global__ void k_omega_check(unsigned first, unsigned long long count, unsigned long long* __restrict__ out)
{
    unsigned long long bad = 0, ties = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < count; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned b = first + (unsigned)i, a = b & 0x7fffffffu;
        if (a >= 0x7f800000u || (a != 0u && a < 0x0E000000u)) continue; // inf / nan; 0 < |q| < 2^-99 never reaches p_from_q
        const float q = __uint_as_float(b);
        const b64 q2 = pk(q, -q), fast = p_from_q(q2), ref = p_from_q_f64(q2);
        bad += (fast != ref);
        const b64 CH = pk(-0x1.e66666p+0f, -0x1.e66666p+0f), CLP = pk(-0x1.99919ap-26f, -0x1.99919ap-26f), CLM = pk(-0x1.99a19ap-26f, -0x1.99a19ap-26f);
        ties += (ffma2(q2, CH, fmul2(q2, CLP)) != ffma2(q2, CH, fmul2(q2, CLM)));
    }
    if (bad) atomicAdd(out, bad);
    if (ties) atomicAdd(out + 1, ties);
}

// 0 < |d| < 2^-96 takes the slow path: exact integer quotient below 2^-125 (div6_tiny), F2F/DMUL product.  With n <= 6
// neighbours |q| = |d / n| >= 2^-99 above it.  Key of a value: 2 * bits - 2 (mod 2^32: the sign falls out, zero wraps to
// the top) -- one multiply-add per value; tiny  <=>  key < TINY_D.
__device__ __forceinline__ unsigned tiny_key(float d) { return __float_as_uint(d) * 2u - 2u; }
constexpr unsigned TINY_D = 2u * 0x0F800000u - 2u;

// bounded mbarrier wait, not unrolled (the first probe almost always succeeds: the loads run NS-1 planes ahead)
__device__ __forceinline__ bool mbar_wait1(unsigned long long* bar, unsigned parity)
{
    unsigned ok = 0;
#pragma unroll 1
    for (int it = 0; it < (1 << 21) && !ok; it++)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(200000u) : "memory"); // suspend-time hint (ns): sleep, do not spin
    return ok != 0;
}

// Denormal-range quotients of "open" cells (acc = 6): the reciprocal + correction sequence can break a tie on the denormal
// grid the wrong way for 0 < |d| < 2^-125; those lanes take the exact integer quotient (div6_tiny).  Inline and on the
// QUOTIENT (one conversion to p afterwards): the decaying front of the SOR iteration is full of such values for the
// first ticks of a scene, so this is not a cold path there.
__device__ __forceinline__ b64 fix_tiny_q(b64 q, b64 d, unsigned ax, unsigned ay, bool six_lo = true, bool six_hi = true)
{
    float qx = lo32(q), qy = hi32(q);
    if (six_lo && ax < 0x00ffffffu) qx = div6_tiny(lo32(d));
    if (six_hi && ay < 0x00ffffffu) qy = div6_tiny(hi32(d));
    return pk(qx, qy);
}
// the denormal-range path of a sweep: exact quotient where the correction sequence can miss, double-precision product
#ifndef SMK_SLOW_P_INLINE
__device__ __noinline__
#else
__device__ __forceinline__
#endif
b64 slow_p(b64 q, b64 d, bool six_lo, bool six_hi)
{
    const unsigned ax = (__float_as_uint(lo32(d)) & 0x7fffffffu) - 1u, ay = (__float_as_uint(hi32(d)) & 0x7fffffffu) - 1u;
    return p_from_q_f64(fix_tiny_q(q, d, ax, ay, six_lo, six_hi));
}

// ---- rare paths, out of line (they cost a call only where they are taken) ---------------------------------------------
// denormal-range quotient of an "open" cell (acc = 6): the exact integer quotient, see div6_tiny
// ... and cells that are not updated (P = 0: old -/+ 0 = old).  act_lo / act_hi: the ACTIVE bits of the two cells.
// six_lo / six_hi: the cell has six fluid neighbours (only the quotient by 6 can miss; see pressure_p_fast).
__device__ __noinline__ b64 rare_fix_pair(b64 d2, b64 P2, unsigned act_lo, unsigned act_hi, unsigned six_lo, unsigned six_hi)
{
    const float dx = lo32(d2), dy = hi32(d2);
    float px = lo32(P2), py = hi32(P2);
    const unsigned ax = (__float_as_uint(dx) & 0x7fffffffu) - 1u, ay = (__float_as_uint(dy) & 0x7fffffffu) - 1u;
    if (six_lo && ax < 0x00ffffffu) px = __double2float_rn(__dmul_rn((double)div6_tiny(dx), M19));
    if (six_hi && ay < 0x00ffffffu) py = __double2float_rn(__dmul_rn((double)div6_tiny(dy), M19));
    return pk(act_lo ? px : 0.f, act_hi ? py : 0.f);
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

// One sweep phase of one lane on ring position J (plane t-J): two same-colour cells of the lane's quad.
// PAR = x parity of the active colour (warp uniform).  a = shared address of this lane's E pair of the plane's row
// (the parity offset is an immediate).  Pairs: E = (f[4h], f[4h+2]), O = (f[4h+1], f[4h+3]).
// GENERAL = false: the CTA's planes hold no COMPLEX cell (flags written with the stencil codes, grid.h) -- the general
// update is not even compiled in, so the hot path has no control-flow merge for the compiler to resolve with copies.
#ifdef SMK_TRACE_FINE
#define FT(k) do { if (ft) ft[k] = clock64(); } while (0)
#else
#define FT(k) do { } while (0)
#endif
template <int PAR, bool GENERAL>
__device__ __forceinline__ void tma_update(b64& ue, b64& uo, b64& we, b64& wo, b64& we1, b64& wo1, const unsigned a, const unsigned cw,
                                           const bool hnz, long long* ft = nullptr)
{
    FT(0);
    constexpr unsigned FULL = 0xffffffffu;
    constexpr unsigned SH = 8u * PAR;                 // byte 0/2 (PAR 0) or 1/3 (PAR 1) of the code word
    constexpr unsigned M_AC = 0x00C000C0u << SH;      // ACTIVE | COMPLEX of both cells
    constexpr unsigned V_A = 0x00400040u << SH;       // ... == ACTIVE, not COMPLEX
    constexpr unsigned M_C = 0x00800080u << SH;
    const b64 M1 = pk(-1.0f, -1.0f), R6 = pk(0x1.555556p-3f, 0x1.555556p-3f) /* RN(1/6) */, M6 = pk(-6.0f, -6.0f);
    b64 V0 = lds64<PAR * 128>(a), V1 = lds64<PAR * 128 + 256>(a);
    b64 U0, U1, W0, W1;
    if (PAR == 0) { U0 = ue; U1 = uo; W0 = we; W1 = we1; }
    else {
        U0 = uo; W0 = wo; W1 = wo1;
        U1 = pk(hi32(ue), __shfl_down_sync(FULL, lo32(ue), 1)); // u[4h+2], u[4h+4] (next quad's first face)
    }
    b64 d = ffma2(U0, M1, U1);           // -u0 + u1           (cu:379-381, left to right, one rounding each)
    d = ffma2(V0, M1, d);                //  ... - v0
    d = fadd2(d, V1);                    //  ... + v1
    d = ffma2(W0, M1, d);                //  ... - w0
    d = fadd2(d, W1);                    //  ... + w1
    FT(1);

    // tier A: both cells of every lane ACTIVE with six fluid neighbours; tier B: no COMPLEX cell (some are not updated)
    const bool allact = __all_sync(FULL, (cw & M_AC) == V_A);
    if (!GENERAL || allact || __all_sync(FULL, (cw & (cw << 1) & M_C) == 0u)) {
        // q = d / 6 by reciprocal + one correction (exact for |d| >= 2^-125: exhaustive check)
        const b64 q0 = fmul2(d, R6);
        const b64 rem = ffma2(q0, M6, d);
        b64 q = ffma2(rem, R6, q0);
        FT(2);
        // below 2^-125 (and d != 0) a tie on the denormal grid can round the wrong way -> exact integer quotient
        const unsigned ax = tiny_key(lo32(d)), ay = tiny_key(hi32(d));
        b64 P;
        if (__any_sync(FULL, min(ax, ay) < TINY_D)) { P = slow_p(q, d, true, true); FT(5); }
        else P = p_from_q(q);
        FT(6);
        FT(3);
        if (!allact) // cells that are not ACTIVE get P = 0 (old -/+ 0 = old)
            P = pk((cw & (0x40u << SH)) ? lo32(P) : 0.f, (cw & (0x400000u << SH)) ? hi32(P) : 0.f);
        U0 = ffma2(P, M1, U0); U1 = fadd2(U1, P);
        V0 = ffma2(P, M1, V0); V1 = fadd2(V1, P);
        W0 = ffma2(P, M1, W0); W1 = fadd2(W1, P);
    } else if (__all_sync(FULL, (((cw >> SH) & 0xC0u) != 0xC0u || ((cw >> SH) & 0x3Fu) == 0x3Bu) &&
                                (((cw >> (SH + 16)) & 0xC0u) != 0xC0u || ((cw >> (SH + 16)) & 0x3Fu) == 0x3Bu))) {
        // tier F: the only COMPLEX cells are cells standing on a solid one (every neighbour fluid but y-1) -- the row
        // above the floor that every scene of the reference has (cu:200-207).  Same sequence with acc = 5 for those
        // cells (exact for every input, see pressure_p_fast) and their lower v face left alone.
        const bool fa_ = ((cw >> SH) & 0xC0u) == 0xC0u, fb_ = ((cw >> (SH + 16)) & 0xC0u) == 0xC0u;
        const float r5 = 0x1.99999ap-3f, r6 = 0x1.555556p-3f;
        const b64 RR2 = pk(fa_ ? r5 : r6, fb_ ? r5 : r6), MN = pk(fa_ ? -5.0f : -6.0f, fb_ ? -5.0f : -6.0f);
        const b64 q0 = fmul2(d, RR2);
        const b64 rem = ffma2(q0, MN, d);
        b64 q = ffma2(rem, RR2, q0);
        const unsigned ax = tiny_key(lo32(d)), ay = tiny_key(hi32(d));
        b64 P;
        if (__any_sync(FULL, min(ax, ay) < TINY_D)) P = slow_p(q, d, !fa_, !fb_);
        else P = p_from_q(q);
        if (!allact) P = pk((cw & (0x40u << SH)) ? lo32(P) : 0.f, (cw & (0x400000u << SH)) ? hi32(P) : 0.f);
        const b64 Pv0 = pk(fa_ ? 0.f : lo32(P), fb_ ? 0.f : hi32(P));
        U0 = ffma2(P, M1, U0); U1 = fadd2(U1, P);
        V0 = ffma2(Pv0, M1, V0); V1 = fadd2(V1, P);
        W0 = ffma2(P, M1, W0); W1 = fadd2(W1, P);
    } else {
        PairFaces io;
        io.f[0] = f2(U0); io.f[1] = f2(U1); io.f[2] = f2(V0); io.f[3] = f2(V1); io.f[4] = f2(W0); io.f[5] = f2(W1);
        general_update_pair(io, d, cw >> SH);
        U0 = pk(io.f[0]); U1 = pk(io.f[1]); V0 = pk(io.f[2]); V1 = pk(io.f[3]); W0 = pk(io.f[4]); W1 = pk(io.f[5]);
    }
    sts64<PAR * 128>(a, V0);
    sts64<PAR * 128 + 256>(a, V1);
    if (PAR == 0) { ue = U0; uo = U1; we = W0; we1 = W1; }
    else {
        uo = U0; wo = W0; wo1 = W1;
        const float from_left = __shfl_up_sync(FULL, hi32(U1), 1); // the left quad's updated u[4h]
        ue = pk(hnz ? from_left : lo32(ue), lo32(U1));              // h == 0: tile edge, face stays stale (halo)
    }
    FT(4);
}

__device__ __forceinline__ void force_clamp_node_pc(float& u, float& v, float& w, unsigned pc, float d, bool clampable, const ForceArgs& fa)
{
    // the cell is fluid iff (pcode & 0xC0) != 0 (grid.h)
    force_clamp_node(u, v, w, ((pc & 0xC0u) ? CODE_SELF : 0u) | (pc & CODE_SY0), d, clampable, fa);
}

// TMA producer, out of line (one lane of warp 0 calls it once per z-step): plane z of its source into a staging slot.
struct IssueArgs {
    const PassMaps* maps;
    unsigned long long* bars;
    unsigned char* stage;
    int x0, y0, x0k, zlo, own_lo, own_hi, lo_zlo, hi_zlo;
    bool has_lo, has_hi;
    long long* trace; // development aid
    int t0;
};
template <int K, int NW, bool FORCE>
__device__ __forceinline__ void tma_issue_plane(const IssueArgs& q, int z, int slot)
{
    using C = TmaCfg<K, NW, FORCE>;
    int src = 0, zr = z - q.zlo;
    if (q.has_lo && z < q.own_lo) { src = 1; zr = z - q.lo_zlo; }
    else if (q.has_hi && z > q.own_hi) { src = 2; zr = z - q.hi_zlo; }
    const CUtensorMap* ms = &q.maps->smoke[src];
    unsigned char* dst = q.stage + slot * C::SLOT;
#ifdef SMK_PASS_TRACE
    long long* tr = (q.trace && z - q.t0 < 80 && z >= q.t0) ? q.trace + 16 * 80 * 8 + (z - q.t0) * 8 : nullptr;
#else
    long long* const tr = nullptr;
#endif
    if (tr) tr[0] = clock64();
    mbar_expect_tx(&q.bars[slot], 3u * C::FB + (unsigned)(C::KW * C::LY) + (FORCE ? (unsigned)C::FB : 0u));
    if (tr) tr[1] = clock64();
    tma_load_4d(dst, &q.maps->uvw[src], q.x0, q.y0, zr, 0, &q.bars[slot]);
    if (tr) tr[2] = clock64();
    tma_load_3d(dst + 3 * C::FB, &q.maps->pcode, q.x0k, q.y0, z - q.zlo, &q.bars[slot]);
    if (tr) tr[3] = clock64();
    if (FORCE) tma_load_3d(dst + 3 * C::FB + C::KB, ms, q.x0, q.y0, zr, &q.bars[slot]);
}

// One PIECE of a pass: the tile (bx, by) marched over the output node planes [zo0, zo1) (K lead-in planes below, K - 1
// above).
//
// Step t:  (a) write out v of plane t-K-1;  (b) ENTER plane t+1 -- wait for its staged copy, registers + v ring;
//          (c) K sweeps on planes t-1 .. t-K;  (d) write out u, w of plane t-K;  (e) barrier.
// (b) sits in front of (c) in program order but nothing in (c) depends on it (plane t+1 is first touched by sweep 1 of
// step t+1, as the w above plane t): its mbarrier probe, its four LDS and its v stores drain while the sweeps run, and a
// warp that leaves the barrier finds everything the next sweeps need in registers and shared memory.
// The register ring has K+2 entries; plane p lives in entry (p - t0 + c) mod (K+2) for its whole life, and the loop is
// unrolled over the K+2 rotations (no ring shift; round 1 moved 36 registers per step).  K+2 is even, so the rotation
// also fixes the x parity of the active colour (c = parity of the warp's first step selects where it enters the pattern).
template <int K, int NW, bool FORCE, bool GENERAL>
__device__ __forceinline__ void tma_pass_piece(const GridP& g, const PassMaps& maps, float* __restrict__ uo, float* __restrict__ vo,
                                               float* __restrict__ wo, int sweep0, const PassRange& pr, const ForceArgs& fa,
                                               unsigned char* __restrict__ smem, int bx, int by, int zo0, int zo1,
                                               int* __restrict__ flags, long long* __restrict__ trace = nullptr)
{
    using C = TmaCfg<K, NW, FORCE>;
    constexpr int LY = C::LY, RR = K + 2, NS = C::NS;
    static_assert(RR % 2 == 0, "the unrolled pattern fixes the colour parity per rotation");
    // (the warp index through a shuffle: the compiler then knows that everything derived from it is warp-uniform and
    // emits plain uniform branches around the sweeps instead of divergence bookkeeping)
    const int lane = threadIdx.x & 31, wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int h = lane & 15;                                 // quad index inside the row
    // rows of a warp and the deepest sweep they need: kernels_pressure_reg.cuh (trapezoid halo, shallow rows paired)
    constexpr bool SKIP = (K == 4 && NW >= 8 && NW % 4 == 0);
    const int hb = lane >> 4;
    const int widr = wid;
    const int yl = !SKIP ? widr + hb * NW : widr >= 4 ? widr + hb * ((LY - 8) / 2) : hb == 0 ? widr : widr == 3 ? LY - 1 : LY - 2 - widr;
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
        mbar_init(&bars[NS], C::THREADS); // the step barrier
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const IssueArgs iq{&maps, bars, stage, x0, y0, x0k, g.zlo, pr.own_lo, pr.own_hi, pr.lower.zlo, pr.upper.zlo,
                       pr.lower.u != nullptr, pr.upper.u != nullptr, trace, t0};
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&maps.uvw[0]); tma_prefetch_desc(&maps.pcode);
        if (iq.has_lo) tma_prefetch_desc(&maps.uvw[1]);
        if (iq.has_hi) tma_prefetch_desc(&maps.uvw[2]);
        if (FORCE) tma_prefetch_desc(&maps.smoke[0]);
        for (int i = 0; i < NS && t0 + i <= t1; i++) tma_issue_plane<K, NW, FORCE>(iq, t0 + i, i);
    }

    // register ring
    b64 UE[RR], UO[RR], WE[RR], WO[RR];
    unsigned CW[RR];
#pragma unroll
    for (int k = 0; k < RR; k++) {
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
    long long oo = o0;           // (one running offset, not three running pointers: registers)
#define pou (uo + oo)
#define pow_ (wo + oo)
#define pov (vo + (oo - g.nplane))

    unsigned tt = 0;            // (t - t0) << 13: slot bits of plane t in the v ring (masked)
    int ss = 0;                 // staging slot of plane t+1 (the plane that enters during step t)
    unsigned ph = 0;            // parity bits of the NS mbarriers (bit i flips every time slot i is consumed)
    int t = t0;
    bool failed = false;

    // enter plane z (staged in slot ss, v-ring slot bits vt) into the ring entry E
    auto enter = [&](b64& ue, b64& uq, b64& we, b64& wq, unsigned& cw, int z, unsigned vt) -> bool {
        if (!mbar_wait1(&bars[ss], (ph >> ss) & 1u)) { flags[2] = 1; failed = true; return false; }
        ph ^= 1u << ss;
        float4 pu = *reinterpret_cast<const float4*>(lst + ss * C::SLOT);
        float4 pv = *reinterpret_cast<const float4*>(lst + ss * C::SLOT + C::FB);
        float4 pw = *reinterpret_cast<const float4*>(lst + ss * C::SLOT + 2 * C::FB);
        unsigned pc = *reinterpret_cast<const unsigned*>(lkc + ss * C::SLOT);
        if (FORCE) { // first pass of the step: forcing + clamp on the way in
            const float4 pd = *reinterpret_cast<const float4*>(lst + ss * C::SLOT + 3 * C::FB + C::KB);
            const bool cl = yg >= 1 && yg < g.H && z >= 1 && z < g.D; // + 1 <= x < W per node
            force_clamp_node_pc(pu.x, pv.x, pw.x, pc & 255u, pd.x, cl && xg >= 1 && xg < g.W, fa);
            force_clamp_node_pc(pu.y, pv.y, pw.y, (pc >> 8) & 255u, pd.y, cl && xg + 1 < g.W, fa);
            force_clamp_node_pc(pu.z, pv.z, pw.z, (pc >> 16) & 255u, pd.z, cl && xg + 2 < g.W, fa);
            force_clamp_node_pc(pu.w, pv.w, pw.w, pc >> 24, pd.w, cl && xg + 3 < g.W, fa);
        }
        ue = pk(pu.x, pu.z); uq = pk(pu.y, pu.w);
        we = pk(pw.x, pw.z); wq = pk(pw.y, pw.w);
        cw = pc;
        const unsigned a = C::RING_ABS + (vt & C::RING_MASK) + lq;
        sts32<0>(a, pv.x); sts32<4>(a, pv.z); sts32<128>(a, pv.y); sts32<132>(a, pv.w);
        if (++ss == NS) ss = 0;
        return true;
    };

    // Split-phase step barrier (an mbarrier every thread arrives on): a thread ARRIVES right after its last v store of a
    // step and WAITS just before the first sweep of the next one.  What lies in between -- writing out u and w, the
    // producer's TMA issue, entering the next plane -- overlaps the skew between the warps instead of following it.
    // development aid: timestamps of three warps of one CTA (tools/cta_times.py trace)
#ifdef SMK_PASS_TRACE // (make EXTRA=-DSMK_PASS_TRACE: the timestamps cost code size in every unrolled step)
#define TRACE(k)                                                                                              \
    do {                                                                                                      \
        if (trace && lane == 0 && (t - t0) < 80) trace[((wid * 80) + (t - t0)) * 8 + (k)] = clock64();         \
    } while (0)
#else
#define TRACE(k) do { } while (0)
#endif
    unsigned long long* const sbar = bars + NS;
    unsigned sph = 0;
    auto step_arrive = [&]() { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(sbar)) : "memory"); };
    auto step_wait = [&]() -> bool {
        if (!mbar_wait1(sbar, sph)) { flags[2] = 1; failed = true; return false; }
        sph ^= 1u;
        return true;
    };

    // One z-step at rotation ROT (compile time): position k (plane t-k, k = -1 .. K) is ring entry (ROT - k) mod RR, and
    // the colour parity of the step is ROT & 1.  Returns false when the piece is finished (or a TMA transaction was lost).
    auto step = [&](auto rot_c, auto par_c) -> bool {
        constexpr int ROT = decltype(rot_c)::value, PAR = decltype(par_c)::value;
        auto ph_of = [](int k) { return ((ROT - k) % RR + RR) % RR; };
        // every row has finished the sweeps of step t-1 (and entered plane t)
        if (!step_wait()) return false;
        TRACE(3);
        if (t > t1) { // only the last plane's v is left to write
            const int s2 = t - K - 1;
            if (s2 >= zo0 && s2 < zo1 && sok) {
                const unsigned a = C::RING_ABS + ((tt - (unsigned)(K + 1) * C::RING_SLOT) & C::RING_MASK) + lq;
                const b64 ve = lds64<0>(a), vq = lds64<128>(a);
                *reinterpret_cast<float4*>(pov) = make_float4(lo32(ve), lo32(vq), hi32(ve), hi32(vq));
            }
            return false;
        }
        // Plane t-1 was entered during step t-2, in front of every thread's arrive of step t-1: its staging slot has been
        // read by everyone and is refilled now.  The producer is a lane of warp 0, whose rows need one sweep per step
        // instead of K: it has the time.
        if (threadIdx.x == 0 && t > t0 && t - 1 + NS <= t1) tma_issue_plane<K, NW, FORCE>(iq, t - 1 + NS, (ss + NS - 2) % NS);
        TRACE(0);
        // (c) sweep j runs on cell plane t-j with colour (sweep0+j-1)&1: the active x parity of a row,
        //     (y + (t-j) + sweep0 + j - 1) & 1 = (y + t + sweep0 + 1) & 1, is the same for all K sweeps of this step.
        //     Plane t-j needs sweep j only if it lies j-1 planes above t0 (t - 2j + 1 >= t0) and the warp's rows need
        //     it (j <= jmax): one sweep count per step.  No barrier between the sweeps: they touch different planes
        //     of v, and u / w are private to the lane.
        const int nsw = min(jmax, (t - t0 + 1) >> 1);
#pragma unroll
        for (int j = 1; j <= K; j++) {
            if (j <= nsw)
                tma_update<PAR, GENERAL>(UE[ph_of(j)], UO[ph_of(j)], WE[ph_of(j)], WO[ph_of(j)], WE[ph_of(j - 1)], WO[ph_of(j - 1)],
                                         ((tt - (unsigned)j * C::RING_SLOT) & sw_mask) | sw_base, CW[ph_of(j)], hnz,
#ifdef SMK_TRACE_FINE
                                         (j == 2 && trace && lane == 0 && (t - t0) < 80) ? trace + 17 * 80 * 8 + ((wid * 80) + (t - t0)) * 8 : nullptr
#else
                                         nullptr
#endif
                );
            TRACE(3 + j);
            if (j == 1) { // (a) v of plane t-K-1 became final with the previous step: write it out (off the critical path)
                const int s2 = t - K - 1;
                if (s2 >= zo0 && s2 < zo1 && sok) {
                    const unsigned a = C::RING_ABS + ((tt - (unsigned)(K + 1) * C::RING_SLOT) & C::RING_MASK) + lq;
                    const b64 ve = lds64<0>(a), vq = lds64<128>(a);
                    *reinterpret_cast<float4*>(pov) = make_float4(lo32(ve), lo32(vq), hi32(ve), hi32(vq));
                }
            }
        }
    
        // (e) this thread's v faces of step t are written
        TRACE(1);
        step_arrive();
        // (d) u and w of plane t-K are final and private to this lane: write them out now
        {
            const int s = t - K;
            if (s >= zo0 && s < zo1 && sok) {
                const b64 ue = UE[ph_of(K)], uq = UO[ph_of(K)], we = WE[ph_of(K)], wq = WO[ph_of(K)];
                *reinterpret_cast<float4*>(pou) = make_float4(lo32(ue), lo32(uq), hi32(ue), hi32(uq));
                *reinterpret_cast<float4*>(pow_) = make_float4(lo32(we), lo32(wq), hi32(we), hi32(wq));
            }
            oo += g.nplane;
        }
        // (b) plane t+1 enters (its staging slot was filled NS-1 steps ago): registers + v ring.  Its v stores are
        //     ordered before this thread's NEXT arrive, which is what the rows that read them (step t+2) wait for.
        if (t + 1 <= t1) {
            if (!enter(UE[ph_of(-1)], UO[ph_of(-1)], WE[ph_of(-1)], WO[ph_of(-1)], CW[ph_of(-1)], t + 1, tt + C::RING_SLOT)) return false;
        }
        TRACE(2);
        tt += C::RING_SLOT;
        t++;
        return true;
    };

    // prologue: plane t0 enters position 0 of the first step, i.e. ring entry c
    const bool c1 = ((rowpar + t0) & 1) != 0;
    {
        bool ok;
        if (c1) ok = enter(UE[1], UO[1], WE[1], WO[1], CW[1], t0, 0u);
        else ok = enter(UE[0], UO[0], WE[0], WO[0], CW[0], t0, 0u);
        if (!ok) return;
        step_arrive();   // plane t0 is in: counts as "step t0 - 1 done" for the first wait
    }
    using I0 = std::integral_constant<int, 0>;
    using I1 = std::integral_constant<int, 1>;
    if (GENERAL) {
        // Tiles with COMPLEX cells (the floor row of every scene, obstacle surfaces): every tier of the update is compiled in,
        // so the step body is large -- unrolled six times it no longer fits the instruction cache, and the traces showed every
        // vote / branch of the hot tier paying an instruction fetch (2.2x per sweep).  Compact code instead: one step body per
        // colour parity, plane t0 in ring entry 0, and the ring shifted by register moves after every step.
        if (c1) { // the prologue put plane t0 into entry 1
            UE[0] = UE[1]; UO[0] = UO[1]; WE[0] = WE[1]; WO[0] = WO[1]; CW[0] = CW[1];
        }
        for (;;) {
            t = __shfl_sync(0xffffffffu, t, 0); // (tells the compiler the step counter is warp-uniform: no divergence bookkeeping)
            const bool go = ((rowpar + t) & 1) ? step(I0{}, I1{}) : step(I0{}, I0{});
            if (!go) break;
            // position k of the next step = position k-1 of this one; with ROT = 0 position k is entry (RR - k) % RR
#pragma unroll
            for (int k = K; k >= 0; k--) {
                const int to = (RR - k) % RR, from = (RR - (k - 1)) % RR;
                UE[to] = UE[from]; UO[to] = UO[from]; WE[to] = WE[from]; WO[to] = WO[from]; CW[to] = CW[from];
            }
        }
    } else {
        bool skip0 = c1;
        for (;;) {
            if (!skip0) { if (!step(I0{}, I0{})) break; }
            skip0 = false;
            if (!step(I1{}, I1{})) break;
            if (!step(std::integral_constant<int, 2 % RR>{}, I0{})) break;
            if (!step(std::integral_constant<int, 3 % RR>{}, I1{})) break;
            if (RR > 4) {
                if (!step(std::integral_constant<int, 4 % RR>{}, I0{})) break;
                if (!step(std::integral_constant<int, 5 % RR>{}, I1{})) break;
            }
        }
    }
    if (failed) return;
}
#undef pou
#undef pow_
#undef pov

template <int K, int NW, bool FORCE>
__global__ void __launch_bounds__(NW * 32, 1)
k_pressure_tma(GridP g, const __grid_constant__ PassMaps maps, float* __restrict__ uo, float* __restrict__ vo, float* __restrict__ wo,
               int sweep0, int zchunk, PassRange pr, ForceArgs fa, int* __restrict__ flags,
               const unsigned char* __restrict__ cflag, long long* __restrict__ dbg)
{
    const long long dbg_t0 = dbg ? clock64() : 0;
    static_assert(K % 2 == 0 && NW % 2 == 0, "a pass is whole red+black pairs; both rows of a warp share the parity");
    using C = TmaCfg<K, NW, FORCE>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if (smem_u32(smem_raw) > 2048u) { // the layout is in absolute shared-window addresses: the dynamic part must start low
        if (threadIdx.x == 0) flags[2] = 2;
        return;
    }

    // CTAs are dispatched in linear block order and a piece of the bottom tile row (the floor: COMPLEX cells, the
    // slower variant) takes ~1.5x as long as the others: hand those out first, so that they never form the tail.
    // With the in-kernel handshake (PassSync) the boundary chunks sit at the end of the z order unless this is the
    // first pass of a step (they wait for the neighbours, so they start as late as possible): their floor pieces stay
    // with them -- interior floor, interior rest, boundary floor, boundary rest.
    int bx = (int)blockIdx.x, by = (int)blockIdx.y, bz = (int)blockIdx.z;
    if (gridDim.y > 1) {
        const int gx = (int)gridDim.x, gy = (int)gridDim.y, gz = (int)gridDim.z;
        const int zb = (pr.sync.nchunks > 0 && !pr.sync.first) ? min(gz, 2) : 0; // trailing boundary z's
        const int za = gz - zb;
        int L = bx + gx * (by + gy * bz), z0 = 0, nz = za;
        if (L >= gx * gy * za) { L -= gx * gy * za; z0 = za; nz = zb; }
        if (L < gx * nz) { by = 0; bx = L % gx; bz = z0 + L / gx; }
        else { L -= gx * nz; bx = L % gx; L /= gx; by = 1 + L % (gy - 1); bz = z0 + L / (gy - 1); }
    }
    int chunk = pr.chunk_first + bz * pr.chunk_step;
    int bside = -1; // this CTA reads / serves the neighbour on that side (PassSync, kernels_pressure_reg.cuh)
    if (pr.sync.nchunks > 0) {
        if (pr.sync.first) chunk = bz == 0 ? 0 : bz == 1 ? pr.sync.nchunks - 1 : bz - 1;
        else chunk = bz + 1 < pr.sync.nchunks ? bz + 1 : 0; // boundary chunks last
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
    int zo0 = pr.out_lo + chunk * zchunk; // output node planes [zo0, zo1)
    int zo1 = min(zo0 + zchunk, pr.out_hi);
    if (pr.nzcut > 0) { zo0 = pr.zcut[chunk]; zo1 = pr.zcut[chunk + 1]; }
    if (zo0 < zo1) {
        // Does any plane this piece loads hold a COMPLEX cell inside the tile (halo included)?  cflag[z][by][bx] is written
        // with the stencil codes (k_codes*, K = 4 tile geometry); without it the general update is always compiled in.
        int cx = 1;
        if (K == 4 && cflag) {
            cx = 0;
            const int za = max(zo0 - K, g.zlo), zb = min(zo1 + K - 1, g.zlo + g.nzc - 1);
            const int per = (int)(gridDim.x * gridDim.y), me = by * (int)gridDim.x + bx;
            for (int z = za + (int)threadIdx.x; z <= zb; z += (int)blockDim.x) cx |= cflag[(long long)(z - g.zlo) * per + me];
            cx = __syncthreads_or(cx);
        }
        if (cx) tma_pass_piece<K, NW, FORCE, true>(g, maps, uo, vo, wo, sweep0, pr, fa, smem_raw, bx, by, zo0, zo1, flags,
                                                         (dbg && bx == dbg[(1 << 17) - 3] && by == dbg[(1 << 17) - 2] && bz == dbg[(1 << 17) - 1]) ? dbg + (1 << 17) : nullptr);
        else tma_pass_piece<K, NW, FORCE, false>(g, maps, uo, vo, wo, sweep0, pr, fa, smem_raw, bx, by, zo0, zo1, flags,
                                                       (dbg && bx == dbg[(1 << 17) - 3] && by == dbg[(1 << 17) - 2] && bz == dbg[(1 << 17) - 1]) ? dbg + (1 << 17) : nullptr);
        if (dbg && threadIdx.x == 0) { // development aid (SMK_PASS_DEBUG): per-CTA start, duration, SM and variant
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            long long* o = dbg + 4 * ((long long)(bz * (int)gridDim.y + by) * (int)gridDim.x + bx);
            o[0] = dbg_t0; o[1] = clock64() - dbg_t0; o[2] = smid; o[3] = cx * 1000 + (zo1 - zo0);
        }
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

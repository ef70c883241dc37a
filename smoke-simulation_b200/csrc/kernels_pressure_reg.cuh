// kernels_pressure_reg.cuh -- temporally blocked pressure kernel, register-resident columns (the default).
//
// Same schedule and the same proof of bit-identity as kernels_pressure_fused.cuh (K half-sweeps per launch while
// marching an (x,y) tile along z; trapezoid halo; out of place).  What changed is WHERE the ring of K+1 planes
// lives.  Measurements of the all-shared-memory versions (profiles/r1_fused_smem_ablation.txt) showed ~116 B of
// shared-memory traffic per cell update and the block barriers of every z-step dominating the time, far from both
// the HBM and the issue limits.  Observation: with a lane owning a QUAD of x-consecutive cells of one row,
//
//   * the w faces of a cell column are touched by that column only            -> registers, never shared
//   * the u faces u[4h+1..4h+3] are interior to the quad, u[4h] is written only by the lane to the left
//     (and only in odd-parity steps)                                          -> registers + one shuffle pair
//   * only the v faces are shared between rows (different half-warps / warps) -> shared memory
//
// so a sweep phase costs four 64-bit shared-memory accesses instead of fifteen, and -- because consecutive sweeps
// of one z-step touch different planes except for the w face of the thread's own column -- the K sweeps of a step
// need no barrier between them.  One __syncthreads per z-step remains (v faces cross rows between steps).
//
// Register ring: UE/UO/WE/WO[k] hold plane t-k (k = 0..K) as float2 pairs split by x parity, E = (f[4h], f[4h+2]),
// O = (f[4h+1], f[4h+3]); the pairs feed the packed f32x2 arithmetic directly.  The ring is shifted by register moves
// once per step (static indices only, no local memory).
#pragma once
#include <cuda_runtime.h>
#include "grid.h"
#include "kernels_basic.cuh"

namespace smk {

template <int K, int NW>
struct RegCfg {
    static constexpr int LX = 64;              // loaded tile width in cells (one half-warp = one row of 16 quads)
    static constexpr int LY = 2 * NW;          // loaded tile height
    static constexpr int RS = 68;              // floats per shared v row: E[0..33] | O[0..33]
    static constexpr int HO = 34;
    static constexpr int R = K + 1;            // ring depth
    static constexpr int PLS = (LY + 1) * RS;  // floats per shared v plane (one dummy row)
    static constexpr int HX = (K + 3) / 4 * 4; // x halo: K rounded up to whole quads (the output region starts on a quad)
    static constexpr int OX = LX - 2 * HX, OY = LY - 2 * K;
    static constexpr int THREADS = NW * 32;
    static constexpr size_t SMEM = (size_t)R * PLS * sizeof(float);
};

// exact q = d / acc for the neighbour counts 1..6 (see pressure_p_fast); returns true if any lane needs the
// IEEE fallback (denormal-range d with acc = 6 is the only inexact case)
// -OVER_RELAXATION as a double (cu:38, 384) in constant memory: DMUL takes it as a constant-bank operand instead of two
// register moves per sweep
__constant__ double c_m19 = -1.9;
#define M19 c_m19
__device__ __forceinline__ float2 p_pair_from_q(float2 q)
{
    return make_float2(__double2float_rn(__dmul_rn((double)q.x, M19)), __double2float_rn(__dmul_rn((double)q.y, M19)));
}

// One sweep phase of one lane on ring position J (plane t-J): two same-colour cells of the lane's quad.
// PAR = x parity of the active colour (warp uniform).  ue/uo/we/wo = the register ring, sv = shared v plane.
template <int PAR, int RS, int HO>
__device__ __forceinline__ void reg_update(float2& ue, float2& uo, float2& we, float2& wo, float2& we1, float2& wo1,
                                           float* __restrict__ sv, int vrow, unsigned cw, int h)
{
    const float2 M1 = make_float2(-1.0f, -1.0f);
    float2 U0, U1, W0, W1;
    if (PAR == 0) { U0 = ue; U1 = uo; W0 = we; W1 = we1; }
    else {
        U0 = uo; W0 = wo; W1 = wo1;
        U1 = make_float2(ue.y, __shfl_down_sync(0xffffffffu, ue.x, 1)); // u[4h+2], u[4h+4] (next quad's first face)
    }
    float* pv = sv + vrow + PAR * HO;
    float2 V0 = *reinterpret_cast<float2*>(pv), V1 = *reinterpret_cast<float2*>(pv + RS);
    const unsigned cs = cw >> (PAR * 8); // byte 0 = code of cell A, byte 2 = code of cell B

    float2 d = __ffma2_rn(U0, M1, U1);   // -u0 + u1           (cu:379-381, left to right, one rounding each)
    d = __ffma2_rn(V0, M1, d);           //  ... - v0
    d = __fadd2_rn(d, V1);               //  ... + v1
    d = __ffma2_rn(W0, M1, d);           //  ... - w0
    d = __fadd2_rn(d, W1);               //  ... + w1

    // Fast path: every cell of the warp is either "open" (code 0xff: fluid, six fluid neighbours, acc = 6, every
    // face updated) or not updated at all (no ACTIVE bit: solid, domain boundary, outside the domain / slab).
    // Only cells next to a solid cell need the general path below.
    const unsigned actm = cs & 0x00400040u;                          // ACTIVE bits of cells A and B
    const unsigned want = (actm << 2) - (actm >> 6);                 // 0xff in the byte of every ACTIVE cell
    const bool simple = ((cs & 0x00ff00ffu) & want) == want;         // ACTIVE => code == 0xff
    if (__all_sync(0xffffffffu, simple)) {
        const float r6 = 0x1.555556p-3f; // RN(1/6)
        const float2 q0 = __fmul2_rn(d, make_float2(r6, r6));
        const float2 rem = __ffma2_rn(q0, make_float2(-6.0f, -6.0f), d);
        const float2 q = __ffma2_rn(rem, make_float2(r6, r6), q0);
        float2 P = p_pair_from_q(q);
        // the correction sequence is exact for every |d| >= 2^-125 (exhaustive check); below that (and d != 0)
        // a tie on the denormal grid can round the wrong way -> IEEE division for those lanes
        // (tested on every cell, active or not: a cell that is not updated gets P = 0 below anyway)
        const unsigned ax = (__float_as_uint(d.x) & 0x7fffffffu) - 1u, ay = (__float_as_uint(d.y) & 0x7fffffffu) - 1u;
        if (__any_sync(0xffffffffu, min(ax, ay) < 0x00ffffffu)) {
            if (ax < 0x00ffffffu) P.x = __double2float_rn(__dmul_rn((double)div6_tiny(d.x), M19));
            if (ay < 0x00ffffffu) P.y = __double2float_rn(__dmul_rn((double)div6_tiny(d.y), M19));
        }
        if (!(actm & 0x40u)) P.x = 0.f;      // cells that are not updated: old -/+ 0 = old
        if (!(actm & 0x400000u)) P.y = 0.f;
        U0 = __ffma2_rn(P, M1, U0); U1 = __fadd2_rn(U1, P);
        V0 = __ffma2_rn(P, M1, V0); V1 = __fadd2_rn(V1, P);
        W0 = __ffma2_rn(P, M1, W0); W1 = __fadd2_rn(W1, P);
    } else {
        // general path: per-cell neighbour count, per-face masks (a masked face gets its old value back)
        unsigned cA = cs & 0xffu, cB = (cs >> 16) & 0xffu;
        if (!(cA & CODE_ACTIVE)) cA = 0;
        if (!(cB & CODE_ACTIVE)) cB = 0;
        const int nA = __popc(cA & 63u), nB = __popc(cB & 63u);
        const float2 rr = make_float2(c_rcp[nA], c_rcp[nB]);
        const float2 q0 = __fmul2_rn(d, rr);
        const float2 rem = __ffma2_rn(q0, make_float2(-(float)nA, -(float)nB), d);
        const float2 q = __ffma2_rn(rem, rr, q0);
        float2 P = p_pair_from_q(q);
        const unsigned ax = __float_as_uint(d.x) & 0x7fffffffu, ay = __float_as_uint(d.y) & 0x7fffffffu;
        const bool sx = nA == 6 && (ax - 1u < 0x00ffffffu), sy = nB == 6 && (ay - 1u < 0x00ffffffu);
        if (__any_sync(0xffffffffu, sx || sy)) { // only acc = 6 can miss (exhaustive check, see pressure_p_fast)
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
        const float from_left = __shfl_up_sync(0xffffffffu, U1.y, 1); // the left quad's updated u[4h]
        if (h != 0) ue.x = from_left;                                  // h == 0: tile edge, face stays stale (halo)
    }
}

// Where the planes outside the slab's OWNED range come from in a multi-GPU run: directly from the neighbour's memory
// (peer-mapped over NVLink), i.e. the halo transfer is fused into the pressure pass plane by plane -- no ghost copy,
// no separate exchange.  u == nullptr: no neighbour on that side / single GPU (planes come from the local arrays).
// All four pointers are VIRTUAL PLANE-0 bases (pointer to the first stored plane minus zlo planes, set by the host), so
// that a lane addresses plane z of any source with the same running offset.
struct PeerPlanes {
    const float* u; const float* v; const float* w; // the neighbour's current "in" buffers
    int zlo;                                         // global index of the neighbour's first stored plane
    const float* smoke;                              // the neighbour's density "now" (first pass with fused forcing)
};
// Forcing + max-velocity clamp (integrate cu:315-329, velocityConfinement cu:331-352; arithmetic of k_force_clamp)
// applied to every node as it is loaded by the FIRST pass of a step: both are pointwise in the node index, so
// "force everything, then sweep" and "force each node on its way into the sweeps" give the same bits, and the
// separate read-modify-write of u, v, w disappears.
struct ForceArgs {
    const float* smoke; // density "now" (cell layout)
    float dt, gravity, alpha;
};
// In-kernel handshake of a pass that reads neighbour planes (one launch per pass, no helper kernels or streams):
// the CTAs of the first / last z-chunk -- the only ones that touch a neighbour's planes -- are scheduled FIRST, spin
// (one thread, acquire loads, clock timeout) until that neighbour has published `wait_epoch`, and the last of them to
// finish publishes `sig_epoch` into the neighbour's counter: "my boundary planes of this pass are written and I no
// longer read yours".  Interior CTAs never wait.  (A kernel only ever waits for kernels enqueued before it.)
struct PassSync {
    const unsigned* wait_ctr[2]; // my counters, written by the lower / upper neighbour (nullptr: no neighbour)
    unsigned* sig_ctr[2];        // the neighbours' counters I publish to
    unsigned* done_ctr[2];       // local: boundary CTAs of this pass that have finished, per side
    unsigned wait_epoch, sig_epoch;
    int* flags;                  // [1] raised on timeout
    int nchunks;                 // > 0 enables the scheme
    int first;                   // 1: block z = 0, 1 are the boundary chunks (default); 0: they come last
    int npieces[2];              // k_pressure_reg_bal: boundary pieces per side (the last one to finish publishes)
};
struct PassRange {
    int out_lo, out_hi;  // node planes this launch writes: [out_lo, out_hi)
    int own_lo, own_hi;  // node planes owned by this slab (inclusive); outside them a peer is the source if present
    int chunk_first, chunk_step; // z-chunk of block z = chunk_first + blockIdx.z * chunk_step (boundary / interior launches)
    PeerPlanes lower, upper;
    PeerPlanes local;    // this slab's own input set and density, as plane-0 bases like the neighbours'
    PassSync sync;
    // z-chunks of unequal length (TMA-staged pass on one GPU): chunk c writes the planes [zcut[c], zcut[c+1]); nzcut = 0:
    // equal chunks of `zchunk` planes.  Long chunks first, one short chunk last: it fills the tail of the last wave.
    int nzcut;
    int zcut[18];
};

__device__ __forceinline__ void force_clamp_node(float& u, float& v, float& w, unsigned cd, float d, bool clampable,
                                                 const ForceArgs& fa)
{
    if ((cd & CODE_SELF) && (cd & CODE_SY0)) {
        const float t = __fmaf_rn(__fmul_rn(d, fa.gravity), fa.dt, __fmul_rn(__fmul_rn(d, fa.alpha), fa.dt));
        v = __fadd_rn(t, v);
    }
    if (clampable) {
        const float L = __fmaf_rn(w, w, __fmaf_rn(u, u, __fmul_rn(v, v)));
        const float t = __fmul_rn(L, fa.dt);
        if (t > 9.0f) {
            const float k = __fdiv_rn(9.0f, t);
            u = __fmul_rn(u, k);
            w = __fmul_rn(w, k);
            v = __fmul_rn(v, k);
        }
    }
}

// One PIECE of a pass: the tile (bx, by) marched over the output node planes [zo0, zo1) (K lead-in planes below, K - 1
// above).  Any split of a pass into pieces gives the same bits (each piece recomputes its own trapezoid halo from the
// input set), so the callers are free to schedule pieces: k_pressure_reg = one piece per CTA on a (tile, z-chunk) grid,
// k_pressure_reg_bal = one CTA per SM working through a list of pieces of equal total cost.
template <int K, int NW, bool FORCE>
__device__ __forceinline__ void reg_pass_piece(const GridP& g, const float* __restrict__ ui, const float* __restrict__ vi,
                                               const float* __restrict__ wi, float* __restrict__ uo, float* __restrict__ vo,
                                               float* __restrict__ wo, const unsigned char* __restrict__ code, int sweep0,
                                               const PassRange& pr, const ForceArgs& fa, float* __restrict__ sv, int bx, int by,
                                               int zo0, int zo1)
{
    using C = RegCfg<K, NW>;
    constexpr int LY = C::LY, RS = C::RS, HO = C::HO, R = C::R, PLS = C::PLS;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int h = lane & 15;                                 // quad index inside the row
    // This half-warp's row.  Trapezoid halo: after sweep s the face rows [s, LY-1-s] of the tile are right, so sweep j is
    // only needed on the cell rows [j-1, LY-1-j]: row r needs the sweeps j <= min(r+1, LY-1-r).  Both rows of a warp must
    // share the x parity of the active colour, so the shallow rows are paired with each other -- (0,LY-2) (1,LY-3) (2,LY-4),
    // needing 1, 2, 3 sweeps -- and those warps skip the sweeps their rows do not need (6 of the 64 warp-sweeps of a
    // step at LY = 32); (3,LY-1) and the remaining pairs (w, w+(LY-8)/2) run all K.  K = 2 keeps the plain (w, w+NW) pairing.
    constexpr bool SKIP = (K == 4 && NW >= 8 && NW % 4 == 0); // ((LY - 8) / 2 even: the pairs (w, w + (LY-8)/2) share the parity)
    const int hb = lane >> 4;
    const int yl = !SKIP ? wid + hb * NW : wid >= 4 ? wid + hb * ((LY - 8) / 2) : hb == 0 ? wid : wid == 3 ? LY - 1 : LY - 2 - wid;
    const int jmax = (SKIP && wid < 3) ? wid + 1 : K; // deepest sweep this warp's rows need (warp uniform)
    const int x0 = bx * C::OX - C::HX;
    const int y0 = by * C::OY - K;
    const int t0 = zo0 - K, t1 = zo1 + K - 1;                // planes that enter the ring
    const int xg = x0 + 4 * h, yg = y0 + yl;

    const bool nok = xg >= 0 && xg <= g.P - 4 && yg >= 0 && yg < g.SY;
    const bool kok = xg >= 0 && xg <= g.PC - 4 && yg >= 0 && yg < g.H;
    const bool sok = nok && yl >= K && yl < LY - K && h >= C::HX / 4 && h < 16 - C::HX / 4;
    const int noff = nok ? xg + yg * g.P : 0;
    const int koff = kok ? xg + yg * g.PC : 0;
    const int vrow = yl * RS + 2 * h;                        // E[2h] of this lane's row inside a shared v plane
    const int rowpar = (y0 + yl + sweep0 + 1) & 1;           // (+ t) = x parity of the active colour, warp uniform

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

    float4 pu, pv, pw, pd;
    unsigned pc;
    const bool dok = FORCE && xg >= 0 && xg <= g.W - 4 && yg >= 0 && yg < g.H; // W % 4 == 0 (host checks)
    const int doff = dok ? xg + yg * g.W : 0;
    // Running element offsets of this lane's quad in plane z relative to the (virtual) plane 0 of a source: one 64-bit
    // add per step instead of a 64-bit multiply per pointer.  The source -- the local arrays, or a neighbour's memory for
    // planes outside the owned range -- only selects the warp-uniform base pointers (kernel parameters).
    long long bn = (long long)t0 * g.nplane + noff, bk = (long long)(t0 - g.zlo) * g.kplane + koff, bd = (long long)t0 * g.cplane + doff;
    auto prefetch = [&](int z) { // called for z = t0, t0 + 1, ...: every plane exactly once, in order
        const float *su = pr.local.u, *sv_ = pr.local.v, *sw = pr.local.w, *sd = pr.local.smoke;
        bool zn = z >= g.zlo && z < g.zlo + g.nzn;
        if (z < pr.own_lo && pr.lower.u) { su = pr.lower.u; sv_ = pr.lower.v; sw = pr.lower.w; sd = pr.lower.smoke; zn = z >= pr.lower.zlo; }
        else if (z > pr.own_hi && pr.upper.u) { su = pr.upper.u; sv_ = pr.upper.v; sw = pr.upper.w; sd = pr.upper.smoke; zn = z <= g.D; }
        const bool zc = z >= g.zlo && z < g.zlo + g.nzc;
        if (FORCE) {
            pd = make_float4(0.f, 0.f, 0.f, 0.f);
            if (zn && zc && dok) pd = __ldg(reinterpret_cast<const float4*>(sd + bd));
            bd += g.cplane;
        }
        pu = pv = pw = make_float4(0.f, 0.f, 0.f, 0.f);
        pc = 0;
        if (zn && nok) {
            pu = __ldg(reinterpret_cast<const float4*>(su + bn));
            pv = __ldg(reinterpret_cast<const float4*>(sv_ + bn));
            pw = __ldg(reinterpret_cast<const float4*>(sw + bn));
        }
        if (zc && kok) pc = __ldg(reinterpret_cast<const unsigned*>(code + bk));
        bn += g.nplane;
        bk += g.kplane;
    };

    // running offset of this lane's quad in the OUTPUT plane t-K (u, w; v is written one plane behind)
    long long bo = (long long)(t0 - K - g.zlo) * g.nplane + noff;
    float* const vo_below = vo - g.nplane;
    prefetch(t0);
    int slot_t = 0; // shared-memory slot of plane t (the same slot held plane t-K-1)
    for (int t = t0; t <= t1 + 1; t++) {
        // (a) v of plane t-K-1 became final with the previous step (behind its barrier): write it out, then reuse
        //     the slot (same thread <-> same addresses, no barrier needed in between)
        {
            const int s2 = t - K - 1;
            if (s2 >= zo0 && s2 < zo1 && sok) {
                const float* p = sv + slot_t * PLS + vrow;
                const float2 ve = *reinterpret_cast<const float2*>(p), vo2 = *reinterpret_cast<const float2*>(p + HO);
                *reinterpret_cast<float4*>(vo_below + bo) = make_float4(ve.x, vo2.x, ve.y, vo2.y);
            }
        }
        if (t > t1) break;
        // (b) plane t enters: u, w and the code word into ring position 0, v into shared memory
        if (FORCE) { // first pass of the step: forcing + clamp on the way in
            const bool cl = yg >= 1 && yg < g.H && t >= 1 && t < g.D; // + 1 <= x < W per node
            force_clamp_node(pu.x, pv.x, pw.x, pc & 255u, pd.x, cl && xg >= 1 && xg < g.W, fa);
            force_clamp_node(pu.y, pv.y, pw.y, (pc >> 8) & 255u, pd.y, cl && xg + 1 < g.W, fa);
            force_clamp_node(pu.z, pv.z, pw.z, (pc >> 16) & 255u, pd.z, cl && xg + 2 < g.W, fa);
            force_clamp_node(pu.w, pv.w, pw.w, pc >> 24, pd.w, cl && xg + 3 < g.W, fa);
        }
        UE[0] = make_float2(pu.x, pu.z); UO[0] = make_float2(pu.y, pu.w);
        WE[0] = make_float2(pw.x, pw.z); WO[0] = make_float2(pw.y, pw.w);
        CW[0] = pc;
        {
            float* p = sv + slot_t * PLS + vrow;
            *reinterpret_cast<float2*>(p) = make_float2(pv.x, pv.z);
            *reinterpret_cast<float2*>(p + HO) = make_float2(pv.y, pv.w);
        }
        if (t < t1) prefetch(t + 1); // in flight during the sweeps

        // (c) sweep j runs on cell plane t-j with colour (sweep0+j-1)&1: the active x parity of a row,
        //     (y + (t-j) + sweep0 + j - 1) & 1 = (y + t + sweep0 + 1) & 1, is the same for all K sweeps of this step.
        //     No barrier between the sweeps: they touch different planes of v, and u / w are private to the lane.
        const int par = (rowpar + t) & 1;
        if (par) {
#pragma unroll
            for (int j = 1; j <= K; j++) {
                if (t - 2 * j + 1 >= t0 && j <= jmax) { // plane t-j needs sweep j only if it lies j-1 planes above t0
                    int sl = slot_t - j; if (sl < 0) sl += R;
                    reg_update<1, RS, HO>(UE[j], UO[j], WE[j], WO[j], WE[j - 1], WO[j - 1], sv, sl * PLS + vrow, CW[j], h);
                }
            }
        } else {
#pragma unroll
            for (int j = 1; j <= K; j++) {
                if (t - 2 * j + 1 >= t0 && j <= jmax) { // plane t-j needs sweep j only if it lies j-1 planes above t0
                    int sl = slot_t - j; if (sl < 0) sl += R;
                    reg_update<0, RS, HO>(UE[j], UO[j], WE[j], WO[j], WE[j - 1], WO[j - 1], sv, sl * PLS + vrow, CW[j], h);
                }
            }
        }

        // (d) u and w of plane t-K are final and private to this lane: write them out now
        {
            const int s = t - K;
            if (s >= zo0 && s < zo1 && sok) {
                *reinterpret_cast<float4*>(uo + bo) = make_float4(UE[K].x, UO[K].x, UE[K].y, UO[K].y);
                *reinterpret_cast<float4*>(wo + bo) = make_float4(WE[K].x, WO[K].x, WE[K].y, WO[K].y);
            }
            bo += g.nplane;
        }
        // (e) shift the register ring
#pragma unroll
        for (int k = K; k >= 1; k--) {
            UE[k] = UE[k - 1]; UO[k] = UO[k - 1]; WE[k] = WE[k - 1]; WO[k] = WO[k - 1]; CW[k] = CW[k - 1];
        }
        // (f) v faces written in this step are read by other rows in the next one
        __syncthreads();
        if (++slot_t == R) slot_t = 0;
    }
}

template <int K, int NW, bool FORCE>
__global__ void __launch_bounds__(NW * 32, 1)
k_pressure_reg(GridP g, const float* __restrict__ ui, const float* __restrict__ vi, const float* __restrict__ wi,
               float* __restrict__ uo, float* __restrict__ vo, float* __restrict__ wo,
               const unsigned char* __restrict__ code, int sweep0, int zchunk, PassRange pr, ForceArgs fa)
{
    using C = RegCfg<K, NW>;
    static_assert(K % 2 == 0 && NW % 2 == 0, "a pass is whole red+black pairs; both rows of a warp share the parity");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sv = reinterpret_cast<float*>(smem_raw); // [R][LY+1][RS]

    int chunk = pr.chunk_first + (int)blockIdx.z * pr.chunk_step;
    int bside = -1; // this CTA reads / serves the neighbour on that side
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
    reg_pass_piece<K, NW, FORCE>(g, ui, vi, wi, uo, vo, wo, code, sweep0, pr, fa, sv, (int)blockIdx.x, (int)blockIdx.y, zo0, zo1);
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

// Balanced schedule (opt-in experiment, see pass_schedule.h): ONE CTA per SM, each working through its own list of pieces -- pieces[first[b] .. first[b+1]) =
// (tile x, tile y, zo0, zo1) -- which the host cuts so that every CTA marches the same number of z-steps (lead-in planes
// included).  A (tile, z-chunk) grid quantises into waves (275 CTAs on 148 SMs = 2 rounds of 58 steps at 256^3); equal
// shares need 104.
// Multi-GPU (pr.sync.nchunks > 0): a BOUNDARY piece of a side is one that reads the neighbour's planes, which is exactly
// when it writes planes the neighbour reads (zo0 - K < own_lo  <=>  zo0 < own_lo + K; likewise above).  The host puts
// the boundary pieces first in every list; a CTA waits for the neighbour's epoch before its first boundary piece of a
// side, and the last boundary piece per side to finish publishes this pass's epoch (same protocol as k_pressure_reg).
template <int K, int NW, bool FORCE>
__global__ void __launch_bounds__(NW * 32, 1)
k_pressure_reg_bal(GridP g, const float* __restrict__ ui, const float* __restrict__ vi, const float* __restrict__ wi,
                   float* __restrict__ uo, float* __restrict__ vo, float* __restrict__ wo,
                   const unsigned char* __restrict__ code, int sweep0, PassRange pr, ForceArgs fa,
                   const int4* __restrict__ pieces, const int* __restrict__ first)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sv = reinterpret_cast<float*>(smem_raw); // [R][LY+1][RS]
    const int p0 = first[blockIdx.x], p1 = first[blockIdx.x + 1];
    unsigned waited = 0;
    for (int p = p0; p < p1; p++) {
        const int4 pc = pieces[p];
        if (p > p0) __syncthreads(); // the previous piece's last v write-out reads shared memory
        unsigned side = 0;
        if (pr.sync.nchunks > 0) {
            if (pr.sync.wait_ctr[0] && pc.z - K < pr.own_lo) side |= 1u;
            if (pr.sync.wait_ctr[1] && pc.w + K - 1 > pr.own_hi) side |= 2u;
            const unsigned need = side & ~waited;
            if (need) {
                if (threadIdx.x == 0) {
                    for (int sd = 0; sd < 2; sd++) {
                        if (!((need >> sd) & 1u)) continue;
                        const long long t0c = clock64();
                        unsigned v;
                        do {
                            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(pr.sync.wait_ctr[sd]) : "memory");
                            if ((int)(v - pr.sync.wait_epoch) >= 0) break;
                            if (clock64() - t0c > (long long)2e10) { pr.sync.flags[1] = 1; break; }
                            __nanosleep(100);
                        } while (true);
                    }
                }
                __syncthreads();
                waited |= need;
            }
        }
        reg_pass_piece<K, NW, FORCE>(g, ui, vi, wi, uo, vo, wo, code, sweep0, pr, fa, sv, pc.x, pc.y, pc.z, pc.w);
        if (side) {
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int sd = 0; sd < 2; sd++) {
                    if (!((side >> sd) & 1u)) continue;
                    const unsigned done = atomicAdd(pr.sync.done_ctr[sd], 1u);
                    if (done == (unsigned)pr.sync.npieces[sd] - 1u) {
                        *pr.sync.done_ctr[sd] = 0;
                        __threadfence_system();
                        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pr.sync.sig_ctr[sd]), "r"(pr.sync.sig_epoch) : "memory");
                    }
                }
            }
        }
    }
}

} // namespace smk

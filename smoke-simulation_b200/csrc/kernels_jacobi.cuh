// kernels_jacobi.cuh -- damped-Jacobi pressure iteration on the face velocities.
//
// EXTENSION: the reference has one pressure solver, red-black SOR (cu:356-394).  BASELINE configs[1] and SURVEY H9
// name a Jacobi variant as the GPU-friendly alternative; its arithmetic is specified in DESIGN.md (section "Jacobi extension")
// and checked bit for bit by the parity tests:
//   p_c  = (float)((double)(-div_c / (float)acc_c) * (2.0/3.0))   for ACTIVE cells (div, acc as in the half-sweep), else 0
//   u[x] = fma(p_(x-1), s_c, fma(-p_c, s_(x-1), u[x]))            c = cell x; likewise v with y-1, w with z-1
// Every cell reads OLD velocities only, so the iteration is out of place (in -> out, the caller swaps the sets).
// One CTA = a 128 x 8 column of nodes marching in z; a lane owns a QUAD of four x-consecutive nodes (16-byte loads
// and stores, one address computation per four cells), a warp one row.  The p of a lane's own cells stays in
// registers for the next plane (it is that plane's p_(z-1)) and goes to the x-neighbour by shuffle; the rows pass
// p to the row above through a double-buffered shared tile (one barrier per plane).  A ninth warp computes p of the
// halo row y0-1 and the halo column x0-1.  Inputs are loaded one plane ahead.  Algorithmic traffic per cell and iteration: 12 B read + 12 B written + 1 B code = 25 B, the same as one
// reference half-sweep.
#pragma once
#include "kernels_basic.cuh"

namespace smk {

__device__ __forceinline__ float jacobi_p(float u0, float u1, float v0, float v1, float w0, float w1, int acc)
{
    float div = __fadd_rn(-u0, u1);
    div = __fadd_rn(div, -v0);
    div = __fadd_rn(div, v1);
    div = __fadd_rn(div, -w0);
    div = __fadd_rn(div, w1);
    const float q = __fdiv_rn(-div, (float)acc);
    return __double2float_rn(__dmul_rn((double)q, 2.0 / 3.0));
}

// p of an arbitrary cell, read from global memory (halo cells and the plane below a z chunk)
__device__ __forceinline__ float jacobi_cell_p(const GridP& g, const float* __restrict__ u, const float* __restrict__ v,
                                               const float* __restrict__ w, const unsigned char* __restrict__ code, int x,
                                               int y, int z)
{
    if (x < 0 || y < 0 || z < 0 || x >= g.W || y >= g.H || z >= g.D) return 0.0f;
    const unsigned cd = code[code_index(g, x, y, z)];
    if (!(cd & CODE_ACTIVE)) return 0.0f;
    const long long n = node_index(g, x, y, z);
    return jacobi_p(u[n], u[n + 1], v[n], v[n + g.P], w[n], w[n + g.nplane], __popc(cd & 63u));
}

constexpr int JQ = 4, JTX = 32 * JQ, JTY = 8, JTHREADS = 32 * (JTY + 1);

__device__ __forceinline__ float jacobi_p_fast(float u0, float u1, float v0, float v1, float w0, float w1, unsigned cd)
{
    if (!(cd & CODE_ACTIVE)) return 0.0f;
    float div = __fadd_rn(-u0, u1);
    div = __fadd_rn(div, -v0);
    div = __fadd_rn(div, v1);
    div = __fadd_rn(div, -w0);
    div = __fadd_rn(div, w1);
    // -div / acc by reciprocal + exact remainder correction (see pressure_p_fast)
    const int acc = __popc(cd & 63u);
    const float nd = -div, r = c_rcp[acc];
    const float q0 = __fmul_rn(nd, r);
    float q = __fmaf_rn(__fmaf_rn(-q0, (float)acc, nd), r, q0);
    // exact for every finite input except acc = 6 with 0 < |nd| < 2^-125 (exhaustive check): those take the integer
    // quotient.  (The decaying front of a Jacobi iteration is full of such values: no IEEE-division slow path here.)
    if (acc == 6 && ((__float_as_uint(nd) & 0x7fffffffu) - 1u < 0x00ffffffu)) q = div6_tiny(nd);
    return __double2float_rn(__dmul_rn((double)q, 2.0 / 3.0));
}

// Inputs of one quad at one plane (w of the plane itself is carried over from the previous plane's w1).
struct JQuad {
    float4 u0, v0, v1, w1;
    float ux; // u[x + 4]
    unsigned cd; // four stencil codes
};

__device__ __forceinline__ JQuad jacobi_load(const GridP& g, const float* __restrict__ u, const float* __restrict__ v,
                                             const float* __restrict__ w, const unsigned char* __restrict__ code,
                                             long long n, long long kc, bool node, bool cells, bool has_ux, bool top)
{
    JQuad r;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    r.u0 = r.v0 = r.v1 = r.w1 = z4;
    r.ux = 0.f;
    r.cd = 0u;
    if (node) {
        r.u0 = *reinterpret_cast<const float4*>(u + n);
        r.v0 = *reinterpret_cast<const float4*>(v + n);
        if (!top) r.w1 = *reinterpret_cast<const float4*>(w + n + g.nplane);
        if (cells && !top) {
            r.cd = *reinterpret_cast<const unsigned*>(code + kc);
            r.v1 = *reinterpret_cast<const float4*>(v + n + g.P);
            if (has_ux) r.ux = u[n + 4];
        }
    }
    return r;
}

// one cell of the halo column x0-1 (scalar), same one-plane-ahead scheme
struct JCell {
    float u0, u1, v0, v1, w1;
    unsigned cd;
};
__device__ __forceinline__ JCell jacobi_load_cell(const GridP& g, const float* __restrict__ u, const float* __restrict__ v,
                                                  const float* __restrict__ w, const unsigned char* __restrict__ code,
                                                  long long n, long long kc, bool on)
{
    JCell r{0.f, 0.f, 0.f, 0.f, 0.f, 0u};
    if (on) {
        r.cd = code[kc];
        r.u0 = u[n];
        r.u1 = u[n + 1];
        r.v0 = v[n];
        r.v1 = v[n + g.P];
        r.w1 = w[n + g.nplane];
    }
    return r;
}

__device__ __forceinline__ float4 jacobi_quad_p(const JQuad& q, const float4& w0)
{
    float4 p;
    p.x = jacobi_p_fast(q.u0.x, q.u0.y, q.v0.x, q.v1.x, w0.x, q.w1.x, q.cd & 255u);
    p.y = jacobi_p_fast(q.u0.y, q.u0.z, q.v0.y, q.v1.y, w0.y, q.w1.y, (q.cd >> 8) & 255u);
    p.z = jacobi_p_fast(q.u0.z, q.u0.w, q.v0.z, q.v1.z, w0.z, q.w1.z, (q.cd >> 16) & 255u);
    p.w = jacobi_p_fast(q.u0.w, q.ux, q.v0.w, q.v1.w, w0.w, q.w1.w, q.cd >> 24);
    return p;
}

// new face value: low-side cell first (Jacobi extension); `on` = the cell exists and has this low face
__device__ __forceinline__ float jacobi_face(float f, float pc, float pn, unsigned cd, unsigned sbit, bool on)
{
    const float t = __fmaf_rn(-pc, (cd & sbit) ? 1.0f : 0.0f, f);
    const float r = __fmaf_rn(pn, (cd & CODE_SELF) ? 1.0f : 0.0f, t);
    return on ? r : f;
}

// One piece of an iteration: the tile (bx, by) over the node planes [za, zb).
__device__ __forceinline__ void jacobi_piece(const GridP& g, const float* __restrict__ ui, const float* __restrict__ vi,
                                             const float* __restrict__ wi, float* __restrict__ uo,
                                             float* __restrict__ vo, float* __restrict__ wo,
                                             const unsigned char* __restrict__ code, int xtiles, float (*ps)[JTY + 1][JTX],
                                             float (*pcol)[JTY + 1], int bx, int by, int za, int zb)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int x0 = bx * JTX, y0 = by * JTY;
    const bool halo = wid == JTY;
    const int sy = halo ? 0 : wid + 1;
    const int x = x0 + JQ * lane, y = y0 + sy - 1;
    const bool node = !halo && x < g.P && y <= g.H;          // this lane stores the quad
    const bool rd = (node || (halo && y >= 0 && x < g.P));   // this lane loads the quad
    const bool cells = x < g.W && y >= 0 && y < g.H;         // at least the first cell of the quad exists
    const bool has_ux = x + JQ < g.P;
    // second duty of the halo warp: one cell of the column x0-1 per lane
    const int xB = x0 - 1, yB = y0 + lane;
    const bool colB = halo && lane < JTY && xB >= 0 && yB < g.H;
    // third duty, only when W is a multiple of the tile width: the two pad quads x in [W, P) of every row are copied
    // through by the halo warp of the last x tile (they never change, but the output set needs them)
    const int xC = xtiles * JTX + JQ * (lane & 1), yC = y0 + (lane >> 1);
    const bool copyC = halo && bx == xtiles - 1 && xC < g.P && lane < 2 * JTY && yC <= g.H;

    long long n = node_index(g, x, max(y, 0), za), kc = code_index(g, min(x, g.W - 1), min(max(y, 0), g.H - 1), min(za, g.D - 1));
    long long nB = 0, kB = 0, nC = 0;
    if (colB) { nB = node_index(g, xB, yB, za); kB = code_index(g, xB, yB, min(za, g.D - 1)); }
    if (copyC) nC = node_index(g, xC, yC, za);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 pb = z4; // p of the cells below
    if (!halo && cells && za > 0) {
        pb.x = jacobi_cell_p(g, ui, vi, wi, code, x, y, za - 1);
        pb.y = jacobi_cell_p(g, ui, vi, wi, code, x + 1, y, za - 1);
        pb.z = jacobi_cell_p(g, ui, vi, wi, code, x + 2, y, za - 1);
        pb.w = jacobi_cell_p(g, ui, vi, wi, code, x + 3, y, za - 1);
    }
    float4 w0 = rd ? *reinterpret_cast<const float4*>(wi + n) : z4;
    float w0B = colB ? wi[nB] : 0.0f;
    JQuad cur = jacobi_load(g, ui, vi, wi, code, n, kc, rd, cells, has_ux, za >= g.D);
    JCell curB = jacobi_load_cell(g, ui, vi, wi, code, nB, kB, colB && za < g.D);
    const unsigned first = (x == 0) ? 0u : 1u;
    for (int z = za; z < zb; z++) {
        const bool top = z >= g.D; // node plane D: no cells, copied through
        JQuad nx1 = cur;
        JCell nx1B = curB;
        if (z + 1 < zb) {
            nx1 = jacobi_load(g, ui, vi, wi, code, n + g.nplane, kc + g.kplane, rd, cells, has_ux, z + 1 >= g.D);
            if (halo) nx1B = jacobi_load_cell(g, ui, vi, wi, code, nB + g.nplane, kB + g.kplane, colB && z + 1 < g.D);
        }
        const float4 p = jacobi_quad_p(cur, w0);
        float (*pz)[JTX] = ps[z & 1];
        *reinterpret_cast<float4*>(&pz[sy][JQ * lane]) = p;
        if (halo && lane < JTY)
            pcol[z & 1][lane + 1] = jacobi_p_fast(curB.u0, curB.u1, curB.v0, curB.v1, w0B, curB.w1, curB.cd);
        float pl = __shfl_up_sync(0xffffffffu, p.w, 1);
        __syncthreads();
        if (node) {
            if (lane == 0) pl = pcol[z & 1][sy];
            const float4 py = *reinterpret_cast<const float4*>(&pz[sy - 1][JQ * lane]);
            const bool row = !top && y < g.H;
            const bool c0 = row && x < g.W, c1 = row && x + 1 < g.W, c2 = row && x + 2 < g.W, c3 = row && x + 3 < g.W;
            const unsigned d0 = cur.cd & 255u, d1 = (cur.cd >> 8) & 255u, d2 = (cur.cd >> 16) & 255u, d3 = cur.cd >> 24;
            float4 un, vn, wn;
            un.x = jacobi_face(cur.u0.x, p.x, pl, d0, CODE_SX0, c0 && first);
            un.y = jacobi_face(cur.u0.y, p.y, p.x, d1, CODE_SX0, c1);
            un.z = jacobi_face(cur.u0.z, p.z, p.y, d2, CODE_SX0, c2);
            un.w = jacobi_face(cur.u0.w, p.w, p.z, d3, CODE_SX0, c3);
            const bool yy = y > 0;
            vn.x = jacobi_face(cur.v0.x, p.x, py.x, d0, CODE_SY0, c0 && yy);
            vn.y = jacobi_face(cur.v0.y, p.y, py.y, d1, CODE_SY0, c1 && yy);
            vn.z = jacobi_face(cur.v0.z, p.z, py.z, d2, CODE_SY0, c2 && yy);
            vn.w = jacobi_face(cur.v0.w, p.w, py.w, d3, CODE_SY0, c3 && yy);
            const bool zz = z > 0;
            wn.x = jacobi_face(w0.x, p.x, pb.x, d0, CODE_SZ0, c0 && zz);
            wn.y = jacobi_face(w0.y, p.y, pb.y, d1, CODE_SZ0, c1 && zz);
            wn.z = jacobi_face(w0.z, p.z, pb.z, d2, CODE_SZ0, c2 && zz);
            wn.w = jacobi_face(w0.w, p.w, pb.w, d3, CODE_SZ0, c3 && zz);
            *reinterpret_cast<float4*>(uo + n) = un;
            *reinterpret_cast<float4*>(vo + n) = vn;
            *reinterpret_cast<float4*>(wo + n) = wn;
        }
        if (copyC) {
            *reinterpret_cast<float4*>(uo + nC) = *reinterpret_cast<const float4*>(ui + nC);
            *reinterpret_cast<float4*>(vo + nC) = *reinterpret_cast<const float4*>(vi + nC);
            *reinterpret_cast<float4*>(wo + nC) = *reinterpret_cast<const float4*>(wi + nC);
        }
        pb = p;
        w0 = cur.w1;
        w0B = curB.w1;
        cur = nx1;
        curB = nx1B;
        n += g.nplane;
        kc += g.kplane;
        nB += g.nplane;
        kB += g.kplane;
        nC += g.nplane;
    }
}

// (tile, z-chunk) grid: one piece per CTA
__global__ void __launch_bounds__(JTHREADS) k_jacobi(GridP g, const float* __restrict__ ui, const float* __restrict__ vi,
                                                     const float* __restrict__ wi, float* __restrict__ uo,
                                                     float* __restrict__ vo, float* __restrict__ wo,
                                                     const unsigned char* __restrict__ code, int zchunk, int xtiles)
{
    __shared__ __align__(16) float ps[2][JTY + 1][JTX]; // row 0 = y0-1, row r = y0 + r - 1
    __shared__ float pcol[2][JTY + 1];                  // p of the cells (x0-1, y0 + r - 1)
    const int za = blockIdx.z * zchunk, zb = min(za + zchunk, g.D + 1); // node planes [za, zb)
    jacobi_piece(g, ui, vi, wi, uo, vo, wo, code, xtiles, ps, pcol, (int)blockIdx.x, (int)blockIdx.y, za, zb);
}

// balanced piece lists (csrc/pass_schedule.h): as many CTAs as fit the GPU at once, equal z-step shares
__global__ void __launch_bounds__(JTHREADS) k_jacobi_bal(GridP g, const float* __restrict__ ui, const float* __restrict__ vi,
                                                         const float* __restrict__ wi, float* __restrict__ uo,
                                                         float* __restrict__ vo, float* __restrict__ wo,
                                                         const unsigned char* __restrict__ code, int xtiles,
                                                         const int4* __restrict__ pieces, const int* __restrict__ first)
{
    __shared__ __align__(16) float ps[2][JTY + 1][JTX];
    __shared__ float pcol[2][JTY + 1];
    const int p0 = first[blockIdx.x], p1 = first[blockIdx.x + 1];
    for (int p = p0; p < p1; p++) {
        const int4 pc = pieces[p];
        if (p > p0) __syncthreads(); // the shared p tiles of the previous piece are still being read
        jacobi_piece(g, ui, vi, wi, uo, vo, wo, code, xtiles, ps, pcol, pc.x, pc.y, pc.z, pc.w);
    }
}

} // namespace smk

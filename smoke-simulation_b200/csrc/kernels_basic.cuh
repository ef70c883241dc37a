// kernels_basic.cuh -- one-stage-per-launch sm_100a kernels of the smoke step.
//
// These are the straightforward streaming kernels: every thread owns one cell / node of an x-fastest
// row, all global accesses of a warp are unit-stride.  They are the parity baseline for the fused /
// temporally blocked kernels and the path used for grids those kernels do not cover.
//
// Arithmetic contract: every floating-point operation below is spelled with an explicit rounding
// intrinsic (__fmul_rn / __fadd_rn / __fmaf_rn / __fdiv_rn / __dmul_rn ...) in exactly the order and
// fusion pattern that nvcc 12.9 generates for the reference kernels on sm_100a (read off the SASS of
// project/smokeSimulation.cu; the pattern is restated at each function).  The compiler can therefore
// neither contract nor re-associate anything, and results are bit-identical to the reference step.
#pragma once
#include <climits>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "grid.h"

namespace smk {

__device__ __forceinline__ long long node_index(const GridP& g, int x, int y, int z)
{
    return (long long)x + (long long)y * g.P + (long long)(z - g.zlo) * g.nplane;
}
__device__ __forceinline__ long long cell_index(const GridP& g, int x, int y, int z)
{
    return (long long)x + (long long)y * g.W + (long long)(z - g.zlo) * g.cplane;
}
__device__ __forceinline__ long long mask_index(const GridP& g, int x, int y, int z)
{
    return (long long)x + (long long)y * g.W + (long long)(z - g.mzlo) * g.cplane;
}
__device__ __forceinline__ long long code_index(const GridP& g, int x, int y, int z)
{
    return (long long)x + (long long)y * g.PC + (long long)(z - g.zlo) * g.kplane;
}

// ---------------------------------------------------------------------------------------------------
// Source + obstacle fill.  Reference: fillSmoke cu:251-273, fillObstacle cu:289-313, drawObjects
// cu:714-771.  Interior cells only.  dist = powf(x-sx,2)+powf(y-sy,2)+powf(z-sz,2) with the cell
// coordinate converted int->float first; the same libdevice powf as the reference (mask ties, H4).
// Sources: dist < r*r  =>  density 1.0 in BOTH buffers.  Obstacles: the cell is rewritten for every
// obstacle in order, so the last obstacle decides; with no obstacle the mask is left alone.
// One thread per cell of an x-row; grid.y walks the z planes [za, zb).
// A cell can only be inside a sphere if it is inside the sphere's bounding box grown by one cell: for |d| >= r + 1,
// powf(d,2) >= (r+1)^2 (1 - 2^-21) > r^2 whatever libdevice's rounding (powf is accurate to a few ulp).  Outside that
// box the comparison `dist < r*r` is decided without evaluating powf; inside it the reference expression is evaluated
// unchanged, so ties on the sphere surface come out exactly as in the reference.
__device__ __forceinline__ bool in_sphere(int x, int y, int z, const float* sp)
{
    const float r1 = fabsf(sp[3]) + 1.0f;
    if (fabsf(x - sp[0]) >= r1 || fabsf(y - sp[1]) >= r1 || fabsf(z - sp[2]) >= r1) return false;
    const float dist = powf(x - sp[0], 2) + powf(y - sp[1], 2) + powf(z - sp[2], 2);
    return dist < sp[3] * sp[3];
}

__global__ void __launch_bounds__(256) k_fill(GridP g, float* __restrict__ smoke0, float* __restrict__ smoke1,
                                              unsigned char* __restrict__ mask, ObjP o, int za)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.W * g.H) return;
    const int y = i / g.W, x = i - y * g.W;
    const int z = za + blockIdx.y; // walks the stored MASK planes (one more than the cell planes on slab-interior sides)
    if (x < 1 || y < 1 || z < 1 || x >= g.W - 1 || y >= g.H - 1 || z >= g.D - 1) return;
    if (z >= g.zlo && z < g.zlo + g.nzc) {
        for (int k = 0; k < o.nsrc; k++) {
            if (in_sphere(x, y, z, o.src[k])) {
                const long long c = cell_index(g, x, y, z);
                smoke0[c] = 1.0f;
                smoke1[c] = 1.0f;
            }
        }
    }
    if (o.nobs > 0) {
        unsigned char sv = 1;
        for (int k = 0; k < o.nobs; k++) // the last obstacle decides (cu:304-310) unless the union extension is on
            sv = in_sphere(x, y, z, o.obs[k]) ? 0 : (o.obstacle_union ? sv : 1);
        mask[mask_index(g, x, y, z)] = sv;
    }
}

// Sources only, over the sources' bounding boxes (the steady-state fill: the mask and the stencil codes are a pure
// function of the obstacle list and are reused until that list -- or the mask itself -- changes).  A cell outside the
// box |d| < |r| + 1 is never inside the sphere (see in_sphere), so scanning the boxes writes exactly the cells the
// whole-grid kernel writes.  grid.z = source index x box planes.
struct SrcBoxes {
    int lo[SMK_MAX_OBJ][3]; // first cell of every source's box
    int n[3];               // box extent (max over sources)
};
__global__ void __launch_bounds__(256) k_fill_sources(GridP g, float* __restrict__ smoke0, float* __restrict__ smoke1,
                                                      ObjP o, SrcBoxes b)
{
    const int k = blockIdx.z / b.n[2], dz = blockIdx.z - k * b.n[2];
    const int dx = blockIdx.x * 32 + threadIdx.x, dy = blockIdx.y * 8 + threadIdx.y;
    if (dx >= b.n[0] || dy >= b.n[1]) return;
    const int x = b.lo[k][0] + dx, y = b.lo[k][1] + dy, z = b.lo[k][2] + dz;
    if (x < 1 || y < 1 || z < 1 || x >= g.W - 1 || y >= g.H - 1 || z >= g.D - 1) return;
    if (z < g.zlo || z >= g.zlo + g.nzc) return;
    if (in_sphere(x, y, z, o.src[k])) {
        const long long c = cell_index(g, x, y, z);
        smoke0[c] = 1.0f;
        smoke1[c] = 1.0f;
    }
}

// ---------------------------------------------------------------------------------------------------
// Stencil codes.  Not a reference kernel: it folds the seven mask reads of divergence (cu:365-376)
// and the "both cells fluid" tests of integrate / advection (cu:321, 534, 564, 594, 623) into one byte
// per cell so that the hot kernels read 1 B/cell of mask information.  Neighbours outside the domain (or
// outside the stored slab) count as solid; such cells are never ACTIVE.
// "This plane holds a COMPLEX cell inside tile (bx, by)" for the tile geometry of the K = 4 fused pass (56 x 24 output
// cells, 4-cell halo: tile b loads cells [56 bx - 4, 56 bx + 60) x [24 by - 4, 24 by + 28)): cflag[z - zlo][by][bx] = 1.
// A CTA of the pass whose planes carry no flag runs the variant without the general update (kernels_pressure_tma.cuh).
struct TileFlags { unsigned char* f; int tx, ty; };
__device__ __forceinline__ void mark_complex(const TileFlags& tf, int x, int y, int zrel)
{
    if (!tf.f) return;
    const int bx1 = min((x + 4) / 56, tf.tx - 1), bx0 = max(0, (x - 59 + 55) / 56);
    const int by1 = min((y + 4) / 24, tf.ty - 1), by0 = max(0, (y - 27 + 23) / 24);
    for (int by = by0; by <= by1; by++)
        for (int bx = bx0; bx <= bx1; bx++) tf.f[((long long)zrel * tf.ty + by) * tf.tx + bx] = 1;
}

__global__ void __launch_bounds__(256) k_codes(GridP g, const unsigned char* __restrict__ mask,
                                               unsigned char* __restrict__ code, unsigned char* __restrict__ pcode, int za, TileFlags tf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.W * g.H) return;
    const int y = i / g.W, x = i - y * g.W;
    const int z = za + blockIdx.y;
    const long long c = mask_index(g, x, y, z);
    const int mhi = g.mzlo + g.nzm; // first mask plane not stored
    unsigned v = 0;
    if (x > 0 && mask[c - 1]) v |= CODE_SX0;
    if (x < g.W - 1 && mask[c + 1]) v |= CODE_SX1;
    if (y > 0 && mask[c - g.W]) v |= CODE_SY0;
    if (y < g.H - 1 && mask[c + g.W]) v |= CODE_SY1;
    if (z > g.mzlo && mask[c - g.cplane]) v |= CODE_SZ0;
    if (z < mhi - 1 && mask[c + g.cplane]) v |= CODE_SZ1;
    const bool interior = x >= 1 && y >= 1 && z >= 1 && x < g.W - 1 && y < g.H - 1 && z < g.D - 1;
    const bool self = mask[c] != 0;
    if (self) v |= CODE_SELF;
    const bool active = interior && self && (v & 63u);
    if (active) v |= CODE_ACTIVE;
    code[code_index(g, x, y, z)] = (unsigned char)v;
    const unsigned m6 = v & 63u;
    pcode[code_index(g, x, y, z)] = (unsigned char)(m6 | (active ? PCODE_ACTIVE : 0u) | ((active ? m6 != 63u : self) ? PCODE_COMPLEX : 0u));
    if (active && m6 != 63u) mark_complex(tf, x, y, z - g.zlo);
}

// Four cells per thread (W % 4 == 0): the mask rows are read as 32-bit words (bytes are 0/1), the six neighbour
// bits of the four cells are assembled with whole-word arithmetic, one 32-bit store.  7 loads per 4 cells instead of 28.
__global__ void __launch_bounds__(256) k_codes4(GridP g, const unsigned char* __restrict__ mask,
                                                unsigned char* __restrict__ code, unsigned char* __restrict__ pcode, int za, TileFlags tf)
{
    const int W4 = g.W >> 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W4 * g.H) return;
    const int y = i / W4, x = (i - y * W4) << 2;
    const int z = za + blockIdx.y;
    const unsigned char* m = mask + mask_index(g, x, y, z);
    const int mhi = g.mzlo + g.nzm;
    const unsigned c = *reinterpret_cast<const unsigned*>(m);
    const unsigned left = x > 0 ? m[-1] : 0u, right = x + 4 < g.W ? m[4] : 0u;
    const unsigned ym = y > 0 ? *reinterpret_cast<const unsigned*>(m - g.W) : 0u;
    const unsigned yp = y < g.H - 1 ? *reinterpret_cast<const unsigned*>(m + g.W) : 0u;
    const unsigned zm = z > g.mzlo ? *reinterpret_cast<const unsigned*>(m - g.cplane) : 0u;
    const unsigned zp = z < mhi - 1 ? *reinterpret_cast<const unsigned*>(m + g.cplane) : 0u;
    const unsigned b = 0x01010101u;
    // a mask byte is fluid when non-zero (the reference tests `!= 0`): normalise every byte to 0/1
    auto norm = [b](unsigned w) { return (w | (w >> 1) | (w >> 2) | (w >> 3) | (w >> 4) | (w >> 5) | (w >> 6) | (w >> 7)) & b; };
    const unsigned cs = norm(c), sx0 = (cs << 8) | (left ? 1u : 0u), sx1 = (cs >> 8) | (right ? 0x01000000u : 0u);
    unsigned v = sx0 * CODE_SX0 + sx1 * CODE_SX1 + norm(ym) * CODE_SY0 + norm(yp) * CODE_SY1 + norm(zm) * CODE_SZ0 + norm(zp) * CODE_SZ1;
    const unsigned any_nb = ((v + 0x3f3f3f3fu) >> 6) & b;      // per byte: at least one fluid neighbour
    unsigned interior = 0u;
    if (y >= 1 && z >= 1 && y < g.H - 1 && z < g.D - 1) {
        interior = b;
        if (x == 0) interior &= ~0xffu;
        if (x + 4 == g.W) interior &= ~0xff000000u;
    }
    const unsigned act = any_nb & cs & interior;                // per byte 0/1: ACTIVE
    const unsigned full = ((v + b) >> 6) & b;                   // per byte: all six neighbours fluid (0x3f + 1 = 0x40)
    const unsigned hi = (act & ~full) | (~act & cs & b);        // pcode bit 7: COMPLEX if ACTIVE, else SELF
    *reinterpret_cast<unsigned*>(pcode + code_index(g, x, y, z)) = v | act * PCODE_ACTIVE | hi * PCODE_COMPLEX;
    if (act & hi) mark_complex(tf, x, y, z - g.zlo); // (tile borders are multiples of 4 in x: the quad lies in the same tiles)
    v |= cs * CODE_SELF | act * CODE_ACTIVE;
    *reinterpret_cast<unsigned*>(code + code_index(g, x, y, z)) = v;
}

// ---------------------------------------------------------------------------------------------------
// Forcing + clamp, fused (both are pointwise on the same staggered index).
// integrate cu:315-329: v-face (x,y,z), x<W, 1<=y<H, z<D, cell and the cell below fluid:
//     v += smoke*gravity*dt + (alpha*smoke)*dt
//   SASS: t = fma(smoke*gravity, dt, (smoke*alpha)*dt);  v = t + v
// velocityConfinement cu:331-352 (max-velocity clamp): 1<=x<W, 1<=y<H, 1<=z<D, same index for u,v,w:
//     L = u*u+v*v+w*w;  if (L*dt > 9)  u,v,w *= 9/(L*dt)
//   SASS: L = fma(w,w, fma(u,u, v*v));  t = L*dt;  k = 9/t (IEEE);  three FMULs
// One thread per node of a row; grid.y walks node planes [za, zb).
__global__ void __launch_bounds__(256) k_force_clamp(GridP g, float* __restrict__ u, float* __restrict__ v,
                                                     float* __restrict__ w, const float* __restrict__ smoke,
                                                     const unsigned char* __restrict__ code, float dt, float gravity,
                                                     float alpha, int za)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.P * g.SY) return;
    const int y = i / g.P, x = i - y * g.P;
    const int z = za + blockIdx.y;
    if (x >= g.W || y < 1 || y >= g.H || z >= g.D) return;
    const long long n = node_index(g, x, y, z);
    float vv = v[n];
    bool dirty = false;
    {
        const unsigned cd = code[code_index(g, x, y, z)];
        if ((cd & CODE_SELF) && (cd & CODE_SY0)) {
            const float d = smoke[cell_index(g, x, y, z)];
            const float t = __fmaf_rn(__fmul_rn(d, gravity), dt, __fmul_rn(__fmul_rn(d, alpha), dt));
            vv = __fadd_rn(t, vv);
            dirty = true;
        }
    }
    if (x >= 1 && z >= 1) {
        const float uu = u[n], ww = w[n];
        const float L = __fmaf_rn(ww, ww, __fmaf_rn(uu, uu, __fmul_rn(vv, vv)));
        const float t = __fmul_rn(L, dt);
        if (t > 9.0f) {
            const float k = __fdiv_rn(9.0f, t);
            u[n] = __fmul_rn(uu, k);
            w[n] = __fmul_rn(ww, k);
            vv = __fmul_rn(vv, k);
            dirty = true;
        }
    }
    if (dirty) v[n] = vv;
}

// ---------------------------------------------------------------------------------------------------
// The pressure update of ONE cell (divergence cu:356-394), shared by every pressure kernel.
//   div = ((((-u0 + u1) - v0) + v1) - w0) + w1                       (cu:379-381, left to right)
//   p   = (float)((double)(div / acc) * -1.9)   == (float)((double)(-div/acc) * 1.9)   (cu:384; the
//         reference SASS folds the negation into the constant: DMUL by -1.9)
//   u0 -= p*sx0  ...  w1 += p*sz1;   p*s is exact (s in {0,1}) so each update is one rounding.
__device__ __forceinline__ float pressure_p(float u0, float u1, float v0, float v1, float w0, float w1, int acc)
{
    float div = __fadd_rn(-u0, u1);
    div = __fadd_rn(div, -v0);
    div = __fadd_rn(div, v1);
    div = __fadd_rn(div, -w0);
    div = __fadd_rn(div, w1);
    const float q = __fdiv_rn(div, (float)acc);
    return __double2float_rn(__dmul_rn((double)q, -1.9));
}

// Branch-free variant used by the hot kernels.  div/acc with acc in {1..6} is computed as
//   r = RN(1/acc) (table), q0 = RN(div*r), rem = fma(-q0, acc, div) (exact), q = fma(rem, r, q0)
// which equals the IEEE quotient for EVERY finite div with biased exponent >= 2 (checked exhaustively on all
// 2^32 inputs for acc = 1..6; the only misses are denormal-range inputs with acc = 6).  `slow` is set for the
// inputs outside that proven range (tiny non-zero, huge, inf, nan); the caller then recomputes that cell with
// pressure_p().  (Non-finite divergence is not flagged: inf/nan fields have no defined parity.)  This removes
// MUFU.RCP and the FCHK/CALL slow-path structure from the inner loop so that the compiler can interleave
// independent cell updates.
__constant__ float c_rcp[8] = {0.0f, 1.0f, 0.5f, 0x1.555556p-2f, 0.25f, 0x1.99999ap-3f, 0x1.555556p-3f, 0.0f};

__device__ __forceinline__ float pressure_p_fast(float u0, float u1, float v0, float v1, float w0, float w1, int acc,
                                                 bool& slow)
{
    float div = __fadd_rn(-u0, u1);
    div = __fadd_rn(div, -v0);
    div = __fadd_rn(div, v1);
    div = __fadd_rn(div, -w0);
    div = __fadd_rn(div, w1);
    const float r = c_rcp[acc];
    const float q0 = __fmul_rn(div, r);
    const float rem = __fmaf_rn(-q0, (float)acc, div);
    const float q = __fmaf_rn(rem, r, q0);
    // the proven-exact range is |div| >= 2^-125; a quotient below FLT_MIN can only come from below that range
    slow = (fabsf(q) < 1.17549435e-38f) && (div != 0.0f);
    return __double2float_rn(__dmul_rn((double)q, -1.9));
}

// Exact d / 6 for 0 < |d| < 2^-125 (biased exponent 0 or 1), the only inputs for which the reciprocal-plus-
// correction sequence can miss (a tie on the denormal grid).  In that range the bit pattern of |d| IS the value in
// units of 2^-149 (the denormal encoding continues into the first normal binade), the quotient is below 2^23 units,
// so IEEE round-to-nearest-even is plain integer arithmetic: m = 6k + r  ->  k + (r > 3 || (r == 3 && k odd)).
__device__ __forceinline__ float div6_tiny(float d)
{
    const unsigned b = __float_as_uint(d), m = b & 0x7fffffffu;
    const unsigned k = __umulhi(m, 0xAAAAAAABu) >> 2; // m / 6
    const unsigned r = m - 6u * k;
    const unsigned q = k + ((r > 3u || (r == 3u && (k & 1u))) ? 1u : 0u);
    return __uint_as_float(q | (b & 0x80000000u));
}

// One red or black half-sweep, in place (cu:356-394; schedule cu:797-801).
// offset 0 <-> (x+y+z) even, 1 <-> odd.  One thread per PAIR of x-adjacent cells: exactly one cell of
// the pair has the active colour, so a warp covers 64 consecutive cells of a row and every lane works.
// Same-colour cells share no face => race-free in place.  Block = 64 pairs x 4 rows; grid.z walks planes.
// The stencil code and the six faces are loaded together (independent loads, one DRAM latency), the
// activity test comes after.
__global__ void __launch_bounds__(256) k_pressure_half(GridP g, float* __restrict__ u, float* __restrict__ v,
                                                       float* __restrict__ w, const unsigned char* __restrict__ code,
                                                       int offset, int za)
{
    const int xp = blockIdx.x * 64 + threadIdx.x;
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int z = za + blockIdx.z;
    const int x = 2 * xp + ((y + z + offset) & 1);
    if (x < 1 || y < 1 || x >= g.W - 1 || y >= g.H - 1) return; // interior cells only (z range set by the launch)
    const long long n = node_index(g, x, y, z);
    const long long nx = n + 1, ny = n + g.P, nz = n + g.nplane;
    const unsigned cd = code[code_index(g, x, y, z)];
    const float u0 = u[n], u1 = u[nx], v0 = v[n], v1 = v[ny], w0 = w[n], w1 = w[nz];
    if (!(cd & CODE_ACTIVE)) return;
    const int acc = __popc(cd & 63u);
    bool slow;
    float p = pressure_p_fast(u0, u1, v0, v1, w0, w1, acc, slow);
    if (slow) p = pressure_p(u0, u1, v0, v1, w0, w1, acc);
    if (cd & CODE_SX0) u[n] = __fsub_rn(u0, p);
    if (cd & CODE_SX1) u[nx] = __fadd_rn(u1, p);
    if (cd & CODE_SY0) v[n] = __fsub_rn(v0, p);
    if (cd & CODE_SY1) v[ny] = __fadd_rn(v1, p);
    if (cd & CODE_SZ0) w[n] = __fsub_rn(w0, p);
    if (cd & CODE_SZ1) w[nz] = __fadd_rn(w1, p);
}

// ---------------------------------------------------------------------------------------------------
// Clamped trilinear sample (sampleSmoke cu:451-484).  Clamp bounds (bx,by,bz) = cell dims - 1 as floats,
// also for staggered fields; (sy, sz) are the strides of the sampled array; zlo its first stored plane.
//   p = max(min(pos, b), 1);  q = p - delta;  i0 = (int)min(floor(q), b);  w1 = q - i0;  w0 = 1 - w1;
//   i1 = (int)min(i0 + 1, b)
// SASS: weights ((xw*yw)*zw) by FMULs; acc = RN(w100*f100); then FFMA for 000,010,110,001,101,011,111.
struct Tri {
    int x0, x1, y0, y1, z0, z1;
    float xw0, xw1, yw0, yw1, zw0, zw1;
};
__device__ __forceinline__ void tri_axis(float pos, float delta, float b, int& i0, int& i1, float& w0, float& w1)
{
    const float p = fmaxf(fminf(pos, b), 1.0f);
    const float q = __fadd_rn(p, -delta);
    // f0 = (float)i0 and f0 + 1 = (float)(i0 + 1) exactly (integer-valued floats in [0, 2^24)): the int -> float
    // conversions of the reference expression are skipped, not changed
    const float f0 = fminf(floorf(q), b);
    i0 = (int)f0;
    w1 = __fadd_rn(q, -f0);
    w0 = __fadd_rn(1.0f, -w1);
    // (int)min(f0 + 1, b): f0 and b are integer-valued and f0 <= b, so that is i0 + 1 unless f0 == b -- one compare and
    // one add instead of FADD / FMNMX / F2I (the conversion pipe runs at a quarter of the rate)
    i1 = i0 + (f0 < b ? 1 : 0);
}
__device__ __forceinline__ float tri_combine(const Tri& t, float f000, float f100, float f010, float f110,
                                             float f001, float f101, float f011, float f111)
{
    const float a00 = __fmul_rn(t.xw0, t.yw0), a10 = __fmul_rn(t.xw1, t.yw0);
    const float a01 = __fmul_rn(t.xw0, t.yw1), a11 = __fmul_rn(t.xw1, t.yw1);
    float acc = __fmul_rn(__fmul_rn(a10, t.zw0), f100);
    acc = __fmaf_rn(__fmul_rn(a00, t.zw0), f000, acc);
    acc = __fmaf_rn(__fmul_rn(a01, t.zw0), f010, acc);
    acc = __fmaf_rn(__fmul_rn(a11, t.zw0), f110, acc);
    acc = __fmaf_rn(__fmul_rn(a00, t.zw1), f001, acc);
    acc = __fmaf_rn(__fmul_rn(a10, t.zw1), f101, acc);
    acc = __fmaf_rn(__fmul_rn(a01, t.zw1), f011, acc);
    acc = __fmaf_rn(__fmul_rn(a11, t.zw1), f111, acc);
    return acc;
}
// `zv` = [lo, hi] planes of f that hold valid data (the whole domain on one GPU; the slab's current valid interval in
// a multi-GPU run): a backtrace that leaves it raises *flag (-> SMK_ERR_REACH) instead of silently reading a stale ghost.
__device__ __forceinline__ float sample_global(const float* __restrict__ f, long long sy, long long sz, int zlo,
                                               float px, float py, float pz, float dx, float dy, float dz,
                                               float bx, float by, float bz, int2 zv, int* __restrict__ flag)
{
    Tri t;
    tri_axis(px, dx, bx, t.x0, t.x1, t.xw0, t.xw1);
    tri_axis(py, dy, by, t.y0, t.y1, t.yw0, t.yw1);
    tri_axis(pz, dz, bz, t.z0, t.z1, t.zw0, t.zw1);
    if (t.z0 < zv.x || t.z1 > zv.y) { // slab runs only; clamp so the load itself stays inside the stored planes
        *flag = 1;
        t.z0 = max(zv.x, min(zv.y, t.z0)); t.z1 = max(zv.x, min(zv.y, t.z1));
    }
    const float* r00 = f + t.y0 * sy + (long long)(t.z0 - zlo) * sz;
    const float* r10 = f + t.y1 * sy + (long long)(t.z0 - zlo) * sz;
    const float* r01 = f + t.y0 * sy + (long long)(t.z1 - zlo) * sz;
    const float* r11 = f + t.y1 * sy + (long long)(t.z1 - zlo) * sz;
    return tri_combine(t, r00[t.x0], r00[t.x1], r10[t.x0], r10[t.x1], r01[t.x0], r01[t.x1], r11[t.x0], r11[t.x1]);
}

// The same sample with 32-bit element indices (fields of fewer than 2^31 stored elements -- every configuration of
// SURVEY 8(d); the host checks): one 32-bit multiply-add per row and one address per corner instead of 64-bit
// arithmetic throughout.  sy, sz = row and plane stride in elements.
__device__ __forceinline__ float sample_global32(const float* __restrict__ f, int sy, int sz, int zlo,
                                                 float px, float py, float pz, float dx, float dy, float dz,
                                                 float bx, float by, float bz, int2 zv, int* __restrict__ flag)
{
    Tri t;
    tri_axis(px, dx, bx, t.x0, t.x1, t.xw0, t.xw1);
    tri_axis(py, dy, by, t.y0, t.y1, t.yw0, t.yw1);
    tri_axis(pz, dz, bz, t.z0, t.z1, t.zw0, t.zw1);
    if (t.z0 < zv.x || t.z1 > zv.y) { // slab runs only; clamp so the load itself stays inside the stored planes
        *flag = 1;
        t.z0 = max(zv.x, min(zv.y, t.z0)); t.z1 = max(zv.x, min(zv.y, t.z1));
    }
    const int r0 = (t.z0 - zlo) * sz, r1 = (t.z1 - zlo) * sz, q0 = t.y0 * sy, q1 = t.y1 * sy;
    const int r00 = r0 + q0, r10 = r0 + q1, r01 = r1 + q0, r11 = r1 + q1;
    return tri_combine(t, f[r00 + t.x0], f[r00 + t.x1], f[r10 + t.x0], f[r10 + t.x1], f[r01 + t.x0], f[r01 + t.x1],
                       f[r11 + t.x0], f[r11 + t.x1]);
}

// 8-point face sums (avgU/avgV/avgW cu:409-447) in source order, then *0.125 (= /8 exactly).
// f points at node (x,y,z); P = row pitch, S = plane stride.
__device__ __forceinline__ float avg_u8(const float* __restrict__ f, long long P, long long S)
{
    float a = f[-S];
    a = __fadd_rn(a, f[1 - S]);
    a = __fadd_rn(a, f[-P - S]);
    a = __fadd_rn(a, f[1 - P - S]);
    a = __fadd_rn(a, f[0]);
    a = __fadd_rn(a, f[1]);
    a = __fadd_rn(a, f[-P]);
    a = __fadd_rn(a, f[1 - P]);
    return __fmul_rn(a, 0.125f);
}
__device__ __forceinline__ float avg_v8(const float* __restrict__ f, long long P, long long S)
{
    float a = f[-S];
    a = __fadd_rn(a, f[-1 - S]);
    a = __fadd_rn(a, f[P - S]);
    a = __fadd_rn(a, f[P - 1 - S]);
    a = __fadd_rn(a, f[0]);
    a = __fadd_rn(a, f[-1]);
    a = __fadd_rn(a, f[P]);
    a = __fadd_rn(a, f[P - 1]);
    return __fmul_rn(a, 0.125f);
}
__device__ __forceinline__ float avg_w8(const float* __restrict__ f, long long P, long long S)
{
    float a = f[0];
    a = __fadd_rn(a, f[-1]);
    a = __fadd_rn(a, f[-P]);
    a = __fadd_rn(a, f[-P - 1]);
    a = __fadd_rn(a, f[-S]);
    a = __fadd_rn(a, f[-1 - S]);
    a = __fadd_rn(a, f[-P - S]);
    a = __fadd_rn(a, f[-P - 1 - S]);
    return __fmul_rn(a, 0.125f);
}

// ---------------------------------------------------------------------------------------------------
// Adaptive advection margin of a slab run: the reference passes wall-clock time as dt (main.cpp:891-895),
// so a backtrace may reach any number of planes.  After the last pressure pass k_absmax_w reduces max |w| over the owned
// planes (warp shuffles + one atomicMax per block) and k_margin turns it into M = planes of "now" data a node needs
// beyond its own plane: ceil(dt * max|w|) for the backtrace + 1 for the trilinear corner (the 8-point averages need 1).
// The advection launches read M from device memory: the planes at least M inside the slab are advected while the
// ghost pull is in flight, the M-wide strips at the slab ends after it.  No host round trip, no fixed assumption on dt.
__global__ void __launch_bounds__(256) k_absmax_w(const float4* __restrict__ w4, size_t n4, unsigned* __restrict__ out)
{
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(w4 + i);
        m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
        for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
    }
}
// dyn[0] = max |w| bits (reset here), dyn[1] = margin M in planes (clamped to the ghost depth: beyond that the device-side
// guard of the samplers raises SMK_ERR_REACH -- the ghost allocation itself is then too small), dyn[2] = unclamped M
__global__ void k_margin(unsigned* __restrict__ dyn, float dt, int ghost)
{
    const float r = __uint_as_float(dyn[0]) * fabsf(dt) * 1.000001f; // (the 8-point average may exceed max |w| by a few ulp)
    int M = (r < 1.0e6f ? (int)ceilf(r) : 1000000) + 1;
    if (M < 2) M = 2;
    dyn[2] = (unsigned)M;
    dyn[1] = (unsigned)min(M, ghost);
    dyn[0] = 0u;
}
// plane range of an advection launch as a function of the margin: mode 0 = [za, zb) as given; 1 = the planes at least M
// inside the slab; 2 / 3 = the strip below / above them.  own_lo / own_hi = first / last owned node plane, or far
// outside the grid on a side without a neighbour.
struct DynRange {
    const unsigned* dyn;
    int mode, own_lo, own_hi;
};
__device__ __forceinline__ void dyn_range(const DynRange& d, int& za, int& zb)
{
    if (d.mode == 0) return;
    const int M = (int)d.dyn[1];
    const int lo = d.own_lo + M, hi = max(lo, d.own_hi - M + 1); // interior [lo, hi); empty when the slab is thinner than 2M
    if (d.mode == 1) { za = max(za, lo); zb = min(zb, hi); }
    else if (d.mode == 2) zb = min(zb, lo);
    else za = max(za, hi);
}

// (float)((double)i + 0.5): the reference evaluates "y + 0.5" in double and narrows (cu:536-537)
__device__ __forceinline__ float half_up(int i) { return __double2float_rn(__dadd_rn((double)i, 0.5)); }

// ---------------------------------------------------------------------------------------------------
// Self-advection of u, v, w in ONE kernel (velocityAdvectionU/V/W cu:527-615), now -> past.
//   U: 1<=x<W,  1<=y<H-1, 1<=z<D-1, s[x]&&s[x-1];  pos0 = (x, y+.5, z+.5);  vel = (u, avgV, avgW)
//   V: 1<=x<W-1, 1<=y<H,  1<=z<D-1, s[y]&&s[y-1];  pos0 = (x+.5, y, z+.5);  vel = (avgU, v, avgW)
//   W: 1<=x<W-1, 1<=y<H-1, 1<=z<D,  s[z]&&s[z-1];  pos0 = (x+.5, y+.5, z);  vel = (avgU, avgV, w)
//   pos = pos0 - dt*vel   SASS: fma(-vel, dt, pos0)
// Faces that fail the test are NOT written (the destination keeps its old value).
// The three 8-point sums are shared between the components that use them (identical values).
__global__ void __launch_bounds__(256) k_advect_velocity(GridP g, const float* __restrict__ u0,
                                                         const float* __restrict__ v0, const float* __restrict__ w0,
                                                         float* __restrict__ u1, float* __restrict__ v1,
                                                         float* __restrict__ w1, const unsigned char* __restrict__ code,
                                                         float dt, int za, int zb, int2 zv, int* __restrict__ flag, DynRange dr)
{
    dyn_range(dr, za, zb);
    // 3-D thread blocks (32 x BY x BZ nodes): the 3x3x3 neighbourhoods of a block overlap in L1
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int z = za + blockIdx.z * blockDim.z + threadIdx.z;
    if (x < 1 || y < 1 || z < 1 || x >= g.W || y >= g.H || z >= zb || z >= g.D) return;
    const unsigned cd = code[code_index(g, x, y, z)];
    if (!(cd & CODE_SELF)) return;
    const bool doU = (cd & CODE_SX0) && y < g.H - 1 && z < g.D - 1;
    const bool doV = (cd & CODE_SY0) && x < g.W - 1 && z < g.D - 1;
    const bool doW = (cd & CODE_SZ0) && x < g.W - 1 && y < g.H - 1;
    if (!(doU || doV || doW)) return;
    const long long n = node_index(g, x, y, z);
    const long long P = g.P, S = g.nplane;
    const float bx = (float)(unsigned)(g.W - 1), by = (float)(unsigned)(g.H - 1), bz = (float)(unsigned)(g.D - 1);
    float au = 0.f, av = 0.f, aw = 0.f;
    if (doV || doW) au = avg_u8(u0 + n, P, S);
    if (doU || doW) av = avg_v8(v0 + n, P, S);
    if (doU || doV) aw = avg_w8(w0 + n, P, S);
    const float xh = half_up(x), yh = half_up(y), zh = half_up(z);
    if (doU) {
        const float px = __fmaf_rn(-u0[n], dt, (float)x);
        const float py = __fmaf_rn(-av, dt, yh);
        const float pz = __fmaf_rn(-aw, dt, zh);
        u1[n] = sample_global(u0, P, S, g.zlo, px, py, pz, 0.f, .5f, .5f, bx, by, bz, zv, flag);
    }
    if (doV) {
        const float px = __fmaf_rn(-au, dt, xh);
        const float py = __fmaf_rn(-v0[n], dt, (float)y);
        const float pz = __fmaf_rn(-aw, dt, zh);
        v1[n] = sample_global(v0, P, S, g.zlo, px, py, pz, .5f, 0.f, .5f, bx, by, bz, zv, flag);
    }
    if (doW) {
        const float px = __fmaf_rn(-au, dt, xh);
        const float py = __fmaf_rn(-av, dt, yh);
        const float pz = __fmaf_rn(-w0[n], dt, (float)z);
        w1[n] = sample_global(w0, P, S, g.zlo, px, py, pz, .5f, .5f, 0.f, bx, by, bz, zv, flag);
    }
}

// ---------------------------------------------------------------------------------------------------
// Density advection (advectSmoke cu:617-638) with the NEW velocities (cu:810).  Interior fluid cells.
//   u_t = (u[x]+u[x+1])/2;  pos = (float)(((double)x + 0.5) - (double)(u_t*dt))
//   SASS: h = (a+b) * -0.5;  m = h*dt;  pos = (float)(((double)x + 0.5) + (double)m)
__global__ void __launch_bounds__(256) k_advect_smoke(GridP g, const float* __restrict__ s0, float* __restrict__ s1,
                                                      const float* __restrict__ u, const float* __restrict__ v,
                                                      const float* __restrict__ w, const unsigned char* __restrict__ code,
                                                      float dt, int za, int zb, int2 zv, int* __restrict__ flag)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int z = za + blockIdx.z * blockDim.z + threadIdx.z;
    if (x < 1 || y < 1 || z < 1 || x >= g.W - 1 || y >= g.H - 1 || z >= zb || z >= g.D - 1) return;
    const long long c = cell_index(g, x, y, z);
    if (!(code[code_index(g, x, y, z)] & CODE_SELF)) return;
    const long long n = node_index(g, x, y, z);
    const float mu = __fmul_rn(__fmul_rn(__fadd_rn(u[n], u[n + 1]), -0.5f), dt);
    const float mv = __fmul_rn(__fmul_rn(__fadd_rn(v[n], v[n + g.P]), -0.5f), dt);
    const float mw = __fmul_rn(__fmul_rn(__fadd_rn(w[n], w[n + g.nplane]), -0.5f), dt);
    const float px = __double2float_rn(__dadd_rn(__dadd_rn((double)x, 0.5), (double)mu));
    const float py = __double2float_rn(__dadd_rn(__dadd_rn((double)y, 0.5), (double)mv));
    const float pz = __double2float_rn(__dadd_rn(__dadd_rn((double)z, 0.5), (double)mw));
    const float bx = (float)(unsigned)(g.W - 1), by = (float)(unsigned)(g.H - 1), bz = (float)(unsigned)(g.D - 1);
    s1[c] = sample_global(s0, g.W, g.cplane, g.zlo, px, py, pz, .5f, .5f, .5f, bx, by, bz, zv, flag);
}

// The same kernel with 32-bit element indices (host: every stored field has fewer than 2^31 elements).
__global__ void __launch_bounds__(256) k_advect_smoke32(GridP g, const float* __restrict__ s0, float* __restrict__ s1,
                                                        const float* __restrict__ u, const float* __restrict__ v,
                                                        const float* __restrict__ w, const unsigned char* __restrict__ code,
                                                        float dt, int za, int zb, int2 zv, int* __restrict__ flag)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int z = za + blockIdx.z * blockDim.z + threadIdx.z;
    if (x < 1 || y < 1 || z < 1 || x >= g.W - 1 || y >= g.H - 1 || z >= zb || z >= g.D - 1) return;
    const int zr = z - g.zlo, cpl = (int)g.cplane, npl = (int)g.nplane;
    if (!(code[x + y * g.PC + zr * (int)g.kplane] & CODE_SELF)) return;
    const int n = x + y * g.P + zr * npl;
    const float mu = __fmul_rn(__fmul_rn(__fadd_rn(u[n], u[n + 1]), -0.5f), dt);
    const float mv = __fmul_rn(__fmul_rn(__fadd_rn(v[n], v[n + g.P]), -0.5f), dt);
    const float mw = __fmul_rn(__fmul_rn(__fadd_rn(w[n], w[n + npl]), -0.5f), dt);
    const float px = __double2float_rn(__dadd_rn(__dadd_rn((double)x, 0.5), (double)mu));
    const float py = __double2float_rn(__dadd_rn(__dadd_rn((double)y, 0.5), (double)mv));
    const float pz = __double2float_rn(__dadd_rn(__dadd_rn((double)z, 0.5), (double)mw));
    const float bx = (float)(unsigned)(g.W - 1), by = (float)(unsigned)(g.H - 1), bz = (float)(unsigned)(g.D - 1);
    s1[x + y * g.W + zr * cpl] = sample_global32(s0, g.W, cpl, g.zlo, px, py, pz, .5f, .5f, .5f, bx, by, bz, zv, flag);
}

// ---------------------------------------------------------------------------------------------------
// Bounding box (in y and z; whole x rows) of the non-zero cells of a cell field: box4 = {y0, y1, z0, z1}, half-open,
// initialised by k_box_init.  One warp per row (W % 4 == 0), grid-stride over the H * nz rows; a CTA folds its rows in
// shared memory and issues four atomics.  Used by the sparse blocking readback (smk_api.cu, readback_box).
__global__ void k_box_init(int* box4)
{
    box4[0] = INT_MAX; box4[1] = INT_MIN; box4[2] = INT_MAX; box4[3] = INT_MIN;
}
__global__ void __launch_bounds__(256) k_density_rows_box(const float* __restrict__ s, int W, int H, int nz, int z_first,
                                                          int* __restrict__ box4)
{
    __shared__ int sb[4];
    if (threadIdx.x == 0) { sb[0] = INT_MAX; sb[1] = INT_MIN; sb[2] = INT_MAX; sb[3] = INT_MIN; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const long long rows = (long long)H * nz;
    for (long long r = warp; r < rows; r += nwarps) {
        const float4* row = reinterpret_cast<const float4*>(s + r * W);
        unsigned any = 0;
        for (int i = lane; i < W / 4; i += 32) {
            const float4 q = row[i];
            any |= (__float_as_uint(q.x) | __float_as_uint(q.y) | __float_as_uint(q.z) | __float_as_uint(q.w)) << 1; // -0.0 counts as zero
        }
        if (__any_sync(0xffffffffu, any != 0) && lane == 0) {
            const int z = (int)(r / H), y = (int)(r - (long long)z * H);
            atomicMin(&sb[0], y); atomicMax(&sb[1], y + 1);
            atomicMin(&sb[2], z + z_first); atomicMax(&sb[3], z + z_first + 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && sb[1] != INT_MIN) {
        atomicMin(&box4[0], sb[0]); atomicMax(&box4[1], sb[1]);
        atomicMin(&box4[2], sb[2]); atomicMax(&box4[3], sb[3]);
    }
}


// the same kernel writing the new density ALSO into a 3-D surface (the renderer's GL_R32F texture mapped
// through CUDA-GL interop, boundingBox.cpp:364-385) instead of going device -> host -> glTexSubImage3D (cu:814).  The
// surface must equal the "past" buffer everywhere, including the cells advection never writes (boundary shell, solids:
// they keep the buffer's stale value), so this variant visits every cell of the planes [za, zb).
__global__ void __launch_bounds__(256) k_advect_smoke_surf(GridP g, const float* __restrict__ s0, float* __restrict__ s1,
                                                           const float* __restrict__ u, const float* __restrict__ v,
                                                           const float* __restrict__ w, const unsigned char* __restrict__ code,
                                                           float dt, int za, int zb, int2 zv, int* __restrict__ flag,
                                                           cudaSurfaceObject_t surf)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int z = za + blockIdx.z * blockDim.z + threadIdx.z;
    if (x >= g.W || y >= g.H || z >= zb) return;
    const long long c = cell_index(g, x, y, z);
    float val;
    const bool interior = x >= 1 && y >= 1 && z >= 1 && x < g.W - 1 && y < g.H - 1 && z < g.D - 1;
    if (interior && (code[code_index(g, x, y, z)] & CODE_SELF)) {
        const long long n = node_index(g, x, y, z);
        const float mu = __fmul_rn(__fmul_rn(__fadd_rn(u[n], u[n + 1]), -0.5f), dt);
        const float mv = __fmul_rn(__fmul_rn(__fadd_rn(v[n], v[n + g.P]), -0.5f), dt);
        const float mw = __fmul_rn(__fmul_rn(__fadd_rn(w[n], w[n + g.nplane]), -0.5f), dt);
        const float px = __double2float_rn(__dadd_rn(__dadd_rn((double)x, 0.5), (double)mu));
        const float py = __double2float_rn(__dadd_rn(__dadd_rn((double)y, 0.5), (double)mv));
        const float pz = __double2float_rn(__dadd_rn(__dadd_rn((double)z, 0.5), (double)mw));
        const float bx = (float)(unsigned)(g.W - 1), by = (float)(unsigned)(g.H - 1), bz = (float)(unsigned)(g.D - 1);
        val = sample_global(s0, g.W, g.cplane, g.zlo, px, py, pz, .5f, .5f, .5f, bx, by, bz, zv, flag);
        s1[c] = val;
    } else {
        val = s1[c];
    }
    surf3Dwrite(val, surf, x * (int)sizeof(float), y, z);
}

// the solid mask as ONE BIT per cell at the boundary (bit i of byte k = cell 8k + i in the
// reference's cell order; 1 = fluid) -- an eighth of the bytes for voxelised obstacles (N3) uploaded or read back.
__global__ void __launch_bounds__(256) k_mask_unpack(const unsigned char* __restrict__ bits, unsigned char* __restrict__ mask, size_t first_cell, size_t ncells)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ncells) return;
    const size_t c = first_cell + i;
    mask[i] = (bits[c >> 3] >> (c & 7)) & 1u;
}
__global__ void __launch_bounds__(256) k_mask_pack(const unsigned char* __restrict__ mask, unsigned char* __restrict__ bits, size_t first_cell, size_t ncells)
{
    // one thread per output byte of the packed range (first_cell is a multiple of 8: the host checks)
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k * 8 >= ncells) return;
    unsigned b = 0;
    for (int j = 0; j < 8 && k * 8 + j < ncells; j++) b |= (mask[k * 8 + j] != 0 ? 1u : 0u) << j;
    bits[(first_cell >> 3) + k] = (unsigned char)b;
}

// ---------------------------------------------------------------------------------------------------
// max |div| over interior fluid cells (residual the reference never computes; formula cu:379-381).
// Warp-shuffle reduction, one atomicMax per block on the float's bit pattern (values are >= 0).
__global__ void __launch_bounds__(256) k_max_divergence(GridP g, const float* __restrict__ u, const float* __restrict__ v,
                                                        const float* __restrict__ w, const unsigned char* __restrict__ code,
                                                        unsigned* __restrict__ out, int za)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.f;
    if (i < g.W * g.H) {
        const int y = i / g.W, x = i - y * g.W;
        const int z = za + blockIdx.y;
        if (x >= 1 && y >= 1 && z >= 1 && x < g.W - 1 && y < g.H - 1 && z < g.D - 1 &&
            (code[code_index(g, x, y, z)] & CODE_SELF)) {
            const long long n = node_index(g, x, y, z);
            float div = __fadd_rn(-u[n], u[n + 1]);
            div = __fadd_rn(div, -v[n]);
            div = __fadd_rn(div, v[n + g.P]);
            div = __fadd_rn(div, -w[n]);
            div = __fadd_rn(div, w[n + g.nplane]);
            m = fabsf(div);
        }
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
        for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
    }
}

// ---------------------------------------------------------------------------------------------------
// Cross-GPU synchronisation for the peer-memory halo path: monotonically increasing epoch counters that live in
// each slab's arena and are written by its neighbours over NVLink.
//   k_epoch_signal: after everything enqueued before it on the stream (kernel boundary + system fence), publish
//                   `epoch` into the neighbour's counter  => "my planes are final for this epoch and I no longer read
//                   yours from the previous one"
//   k_epoch_wait:   spin until the neighbour published `epoch`; a clock-based timeout raises flag[1] instead of
//                   hanging the GPU if a peer died.
__global__ void k_epoch_signal(unsigned* peer_counter_a, unsigned* peer_counter_b, unsigned epoch)
{
    __threadfence_system();
    if (peer_counter_a) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_counter_a), "r"(epoch) : "memory");
    if (peer_counter_b) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_counter_b), "r"(epoch) : "memory");
}

__global__ void k_epoch_wait(const unsigned* my_counter_a, const unsigned* my_counter_b, unsigned epoch, int* flags,
                             long long timeout_cycles)
{
    const long long t0 = clock64();
    for (int i = 0; i < 2; i++) {
        const unsigned* c = i == 0 ? my_counter_a : my_counter_b;
        if (!c) continue;
        unsigned v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
            if ((int)(v - epoch) >= 0) break;
            if (clock64() - t0 > timeout_cycles) { flags[1] = 1; return; }
            __nanosleep(200);
        } while (true);
    }
}

// contiguous plane ranges pulled from a neighbour's memory (or any device memory): 16-byte grid-stride copy
__global__ void __launch_bounds__(256) k_copy16(float4* __restrict__ dst, const float4* __restrict__ src, size_t n16)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// 64-bit content hash of a plane range of one field, independent of how the domain is cut into slabs: every element
// contributes mix(bit pattern, GLOBAL reference-layout index) and the contributions are summed modulo 2^64 (commutative,
// so the order of the atomics does not matter).  Used by bench.py / tools to prove that N slabs hold exactly the bits of
// the single-GPU run at sizes where copying whole fields to the host would be the bottleneck.  -0.0f is hashed as +0.0f
// (the parity tests compare with ==, which does not tell them apart either).
__device__ __forceinline__ unsigned long long hash_mix(unsigned bits, unsigned long long idx)
{
    if (bits == 0x80000000u) bits = 0u;
    unsigned long long h = (idx + 0x9E3779B97F4A7C15ull) * 0xBF58476D1CE4E5B9ull;
    h ^= (unsigned long long)bits * 0x94D049BB133111EBull;
    h ^= h >> 31; h *= 0xD6E8FEB86659FD93ull; h ^= h >> 29;
    return h;
}
// f: internal layout with row pitch `pitch` and plane stride `plane`, first stored plane zlo; reference-layout row
// length nx, rows per plane ny; planes [za, zb) are hashed.  One thread per element of a row segment; grid.y = rows, grid.z = planes.
__global__ void __launch_bounds__(256) k_hash_field(const float* __restrict__ f, int pitch, long long plane, int zlo, int nx, int ny,
                                                    int za, unsigned long long* __restrict__ out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = za + blockIdx.z;
    unsigned long long h = 0;
    if (x < nx) h = hash_mix(__float_as_uint(f[(long long)(z - zlo) * plane + (long long)y * pitch + x]),
                             ((unsigned long long)z * ny + y) * nx + x);
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    __shared__ unsigned long long sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = h;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += sm[i];
        atomicAdd(out, t);
    }
}

// SURVEY 8(f) N4 (opt-in, off for every parity path): the density as IEEE binary16, round-to-nearest-even, for
// consumers that want half the device->host bytes.  Two cells per thread.
__global__ void __launch_bounds__(256) k_density_half(const float* __restrict__ in, __half* __restrict__ out, size_t n)
{
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i + 1 < n) {
        const float2 v = *reinterpret_cast<const float2*>(in + i);
        *reinterpret_cast<__half2*>(out + i) = __floats2half2_rn(v.x, v.y);
    } else if (i < n) {
        out[i] = __float2half_rn(in[i]);
    }
}

} // namespace smk

// kernels_advect_tma.cuh -- u,v,w self-advection with TMA-staged tiles (sm_100a).
//
// Same arithmetic as k_advect_velocity (kernels_basic.cuh; reference velocityAdvectionU/V/W cu:527-615) -- it calls
// the same rounding-exact helpers -- but the ~75 gathers of a node (three 8-point face sums + three trilinear
// samples) are served from shared memory: a CTA owns a TX x TY tile of nodes and marches along z; every z-step the
// TMA engine (cp.async.bulk.tensor.3d, one elected thread, completion on an mbarrier) drops the next (TX+2h) x (TY+2h)
// plane of u, v and w -- tile plus backtrace halo (2 nodes; 4 in x for 16-byte alignment) -- into a ring of 2h+2 planes.  Out-of-range box parts (domain
// edge, negative coordinates) are zero-filled by the TMA unit.  A backtrace that leaves the staged box (|vel|*dt >= 2
// cells; the clamp bounds it near 3*sqrt(dt)) falls back to the global-memory sampler, so results do not
// depend on the halo assumption.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include "grid.h"
#include "kernels_basic.cuh"

namespace smk {

struct AdvTma {
    static constexpr int TX = 32, TY = 8, HALO = 2; // halo of 2 nodes in y and z
    static constexpr int HX = 4;                    // x halo: the box must START on a 16-byte boundary (measured: a start
                                                    // coordinate that is not a multiple of 4 floats raises "illegal instruction")
    static constexpr int BX = TX + 2 * HX, BY = TY + 2 * HALO;     // 40 x 12 box (inner extent 160 B: multiple of 16)
    static constexpr int NSLOT = 2 * HALO + 2;                     // planes z-2..z+2 in use + one in flight
    static constexpr int PLANE = BX * BY;                          // floats per staged plane and field
    static constexpr int SLOT_BYTES = (PLANE * 4 + 127) / 128 * 128; // TMA destinations are 128-byte aligned
    static constexpr int SLOT_FLOATS = SLOT_BYTES / 4;
    static constexpr int THREADS = TX * TY;
    static constexpr size_t SMEM = (size_t)3 * NSLOT * SLOT_BYTES + NSLOT * 8 + 128;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: returns false after ~2^26 polls instead of hanging the GPU on a lost transaction
__device__ __forceinline__ bool mbar_wait(unsigned long long* bar, unsigned parity)
{
#pragma unroll 1
    for (int it = 0; it < (1 << 26); it++) {
        unsigned ok;
        // (suspend-time hint: a thread that finds the plane still in flight sleeps until the barrier flips instead of
        //  polling -- the profile showed 14 probes per thread and z-step, 6 % of the kernel's issue slots)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

// One 32-bit shared-memory load at lane address + immediate.  The kernel addresses its staging rings with explicit
// shared-window addresses: through generic pointers the compiler rebuilt the window base (S2R / LEA) at every use.
template <int OFF>
__device__ __forceinline__ float lds32(unsigned a)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF) : "memory");
    return v;
}

// The staged window of one z-step, shared by the three fields (they are staged with the same box): planes
// z-HALO .. z+HALO in ring slots sb, sb+1, ... (mod NSLOT), box origin (x0, y0).
struct Window {
    int x0, y0;        // global coordinates of box element (0,0)
    int zlo, zhi;      // planes a sample may read: the staged ones that also hold valid data (reach guard of slab runs)
    int zbase;         // plane in slot sb (= z - HALO)
    int sb;            // ring slot of that plane
};

// clamped trilinear sample (sampleSmoke cu:451-484): same index / weight / summation code as sample_global, the
// eight corners come from the staged planes when they are all resident.  fs = shared-window address of the field's
// staging ring.  Addressing: i1 - i0 is 0 or 1 on every axis (tri_axis), so one unsigned compare per axis covers both
// corners, the second row / plane / column is the first plus a selected stride, and nothing is multiplied twice.
__device__ __forceinline__ float sample_staged(const Window& w, unsigned fs, const float* __restrict__ f,
                                               long long sy, long long sz, int zlo, float px, float py, float pz,
                                               float dx, float dy, float dz, float bx, float by, float bz, int2 zv,
                                               int* __restrict__ flag)
{
    using A = AdvTma;
    Tri t;
    tri_axis(px, dx, bx, t.x0, t.x1, t.xw0, t.xw1);
    tri_axis(py, dy, by, t.y0, t.y1, t.yw0, t.yw1);
    tri_axis(pz, dz, bz, t.z0, t.z1, t.zw0, t.zw1);
    const unsigned rx = (unsigned)(t.x0 - w.x0), ry = (unsigned)(t.y0 - w.y0);
    // (a corner pair that is clamped onto the last staged column / row / plane takes the global path: domain edges only)
    if (rx < (unsigned)(A::BX - 1) && ry < (unsigned)(A::BY - 1) && t.z0 >= w.zlo && t.z0 < w.zhi) {
        int s0 = w.sb + (t.z0 - w.zbase);
        s0 -= s0 >= A::NSLOT ? A::NSLOT : 0;
        const unsigned r00 = fs + (unsigned)(s0 * A::SLOT_FLOATS + (int)ry * A::BX + (int)rx) * 4u;
        const unsigned oy = t.y1 != t.y0 ? A::BX * 4 : 0;
        const unsigned oz = t.z1 != t.z0 ? (s0 == A::NSLOT - 1 ? (unsigned)(-(A::NSLOT - 1) * A::SLOT_BYTES) : (unsigned)A::SLOT_BYTES) : 0u;
        const unsigned ox = (unsigned)(t.x1 - t.x0) * 4u;
        const unsigned r10 = r00 + oy, r01 = r00 + oz, r11 = r01 + oy;
        return tri_combine(t, lds32<0>(r00), lds32<0>(r00 + ox), lds32<0>(r10), lds32<0>(r10 + ox),
                           lds32<0>(r01), lds32<0>(r01 + ox), lds32<0>(r11), lds32<0>(r11 + ox));
    }
    return sample_global(f, sy, sz, zlo, px, py, pz, dx, dy, dz, bx, by, bz, zv, flag); // long backtrace: global path
}

__global__ void __launch_bounds__(AdvTma::THREADS)
k_advect_velocity_tma(GridP g, const __grid_constant__ CUtensorMap mu, const __grid_constant__ CUtensorMap mv,
                      const __grid_constant__ CUtensorMap mw, const float* __restrict__ u0, const float* __restrict__ v0,
                      const float* __restrict__ w0, float* __restrict__ u1, float* __restrict__ v1, float* __restrict__ w1,
                      const unsigned char* __restrict__ code, float dt, int za, int zb, int zchunk, int2 zv,
                      int* __restrict__ flag, DynRange dr)
{
    using A = AdvTma;
    dyn_range(dr, za, zb);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* su = reinterpret_cast<float*>(smem_raw);
    float* sv = su + A::NSLOT * A::SLOT_FLOATS;
    float* sw = sv + A::NSLOT * A::SLOT_FLOATS;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(sw + A::NSLOT * A::SLOT_FLOATS);
    const unsigned su_a = smem_u32(su), sv_a = smem_u32(sv), sw_a = smem_u32(sw);

    const int tid = threadIdx.x;
    const int lx = tid % A::TX, ly = tid / A::TX;
    const int tx0 = blockIdx.x * A::TX, ty0 = blockIdx.y * A::TY;      // first node of the tile
    const int z_first = za + blockIdx.z * zchunk, z_last = min(z_first + zchunk, zb); // node planes [z_first, z_last)
    if (z_first >= z_last) return;
    const int x = tx0 + lx, y = ty0 + ly;
    const int bx0 = tx0 - A::HX, by0 = ty0 - A::HALO;

    if (tid == 0) {
        for (int i = 0; i < A::NSLOT; i++) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // Planes are numbered from the first one this piece stages, i = p - (z_first - HALO), and live in slot i mod NSLOT:
    // slot and barrier parity advance by counting, without a division per step.
    const int p0 = z_first - A::HALO;
    auto issue = [&](int p, int slot) {
        mbar_expect_tx(&bars[slot], 3u * A::PLANE * 4u);
        tma_load_3d(su + slot * A::SLOT_FLOATS, &mu, bx0, by0, p - g.zlo, &bars[slot]);
        tma_load_3d(sv + slot * A::SLOT_FLOATS, &mv, bx0, by0, p - g.zlo, &bars[slot]);
        tma_load_3d(sw + slot * A::SLOT_FLOATS, &mw, bx0, by0, p - g.zlo, &bars[slot]);
    };
    if (tid == 0)
        for (int i = 0; i <= 2 * A::HALO; i++) issue(p0 + i, i);

    const float bx = (float)(unsigned)(g.W - 1), by = (float)(unsigned)(g.H - 1), bz = (float)(unsigned)(g.D - 1);
    const long long P = g.P, S = g.nplane;
    const bool xy_ok = x >= 1 && y >= 1 && x < g.W && y < g.H;
    // per-thread constants of the z march: which components this column can ever write, the node's corner (x-1, y-1)
    // inside a staged plane (every neighbour of the face sums is that address + a non-negative immediate), running
    // offsets of the node and of its stencil code
    const unsigned needU = (xy_ok && y < g.H - 1) ? (unsigned)(CODE_SELF | CODE_SX0) : 0x100u; // 0x100: never matches a byte
    const unsigned needV = (xy_ok && x < g.W - 1) ? (unsigned)(CODE_SELF | CODE_SY0) : 0x100u;
    const unsigned needW = (xy_ok && x < g.W - 1 && y < g.H - 1) ? (unsigned)(CODE_SELF | CODE_SZ0) : 0x100u;
    const unsigned corner = (unsigned)((y - by0 - 1) * A::BX + (x - bx0 - 1)) * 4u;
    const float xf = (float)x, yf = (float)y, xh = half_up(x), yh = half_up(y);
    long long n = xy_ok ? node_index(g, x, y, z_first) : 0;
    const unsigned char* pc = code + (xy_ok ? code_index(g, x, y, z_first) : 0);
    constexpr int BX4 = A::BX * 4;
    constexpr int C00 = BX4 + 4;                  // the node itself, relative to its corner
    auto load_code = [&](int z) -> unsigned { return (xy_ok && z >= 1 && z < g.D) ? (unsigned)*pc : 0u; };
    // the stencil code of the node is the one global load in front of everything else of a z-step: fetched a step ahead
    unsigned cd_next = load_code(z_first);
    int sb = 0;                                   // slot of plane z - HALO
    int ws = 2 * A::HALO;                         // slot of plane z + HALO, the one a step waits for ...
    unsigned wpar = 0;                            // ... and the parity of its barrier
    for (int z = z_first; z < z_last; z++) {
        const unsigned cd = cd_next;
        pc += g.kplane;
        if (z + 1 < z_last) cd_next = load_code(z + 1);
        // every plane is waited for exactly once, when it enters the window as z+HALO (the first step waits for all five)
        if (z == z_first) {
            for (int i = 0; i < 2 * A::HALO; i++)
                if (!mbar_wait(&bars[i], 0u)) { flag[2] = 1; return; }
        }
        if (!mbar_wait(&bars[ws], wpar)) { flag[2] = 1; return; }
        const bool zin = z < g.D - 1;
        const bool doU = (cd & needU) == needU && zin;
        const bool doV = (cd & needV) == needV && zin;
        const bool doW = (cd & needW) == needW;
        if (doU || doV || doW) {
            // planes z-1 and z of the three fields at this node's corner
            int s1 = sb + A::HALO - 1, s2 = sb + A::HALO;
            s1 -= s1 >= A::NSLOT ? A::NSLOT : 0;
            s2 -= s2 >= A::NSLOT ? A::NSLOT : 0;
            const unsigned om = (unsigned)(s1 * A::SLOT_BYTES) + corner, o0 = (unsigned)(s2 * A::SLOT_BYTES) + corner;
            const unsigned Um = su_a + om, U0 = su_a + o0, Vm = sv_a + om, V0 = sv_a + o0, Wm = sw_a + om, W0 = sw_a + o0;
            const Window win{bx0, by0, max(z - A::HALO, zv.x), min(z + A::HALO, zv.y), z - A::HALO, sb};
            // 8-point face sums in the reference's order (avgU/avgV/avgW cu:409-447), then *0.125
            float au = 0.f, av = 0.f, aw = 0.f;
            if (doV || doW) { // u at (x, y), (x+1, y), (x, y-1), (x+1, y-1) of planes z-1, z
                float a = lds32<C00>(Um);
                a = __fadd_rn(a, lds32<C00 + 4>(Um)); a = __fadd_rn(a, lds32<4>(Um)); a = __fadd_rn(a, lds32<8>(Um));
                a = __fadd_rn(a, lds32<C00>(U0)); a = __fadd_rn(a, lds32<C00 + 4>(U0)); a = __fadd_rn(a, lds32<4>(U0));
                a = __fadd_rn(a, lds32<8>(U0));
                au = __fmul_rn(a, 0.125f);
            }
            if (doU || doW) { // v at (x, y), (x-1, y), (x, y+1), (x-1, y+1) of planes z-1, z
                float a = lds32<C00>(Vm);
                a = __fadd_rn(a, lds32<C00 - 4>(Vm)); a = __fadd_rn(a, lds32<C00 + BX4>(Vm)); a = __fadd_rn(a, lds32<C00 + BX4 - 4>(Vm));
                a = __fadd_rn(a, lds32<C00>(V0)); a = __fadd_rn(a, lds32<C00 - 4>(V0)); a = __fadd_rn(a, lds32<C00 + BX4>(V0));
                a = __fadd_rn(a, lds32<C00 + BX4 - 4>(V0));
                av = __fmul_rn(a, 0.125f);
            }
            if (doU || doV) { // w at (x, y), (x-1, y), (x, y-1), (x-1, y-1) of planes z, z-1
                float a = lds32<C00>(W0);
                a = __fadd_rn(a, lds32<C00 - 4>(W0)); a = __fadd_rn(a, lds32<4>(W0)); a = __fadd_rn(a, lds32<0>(W0));
                a = __fadd_rn(a, lds32<C00>(Wm)); a = __fadd_rn(a, lds32<C00 - 4>(Wm)); a = __fadd_rn(a, lds32<4>(Wm));
                a = __fadd_rn(a, lds32<0>(Wm));
                aw = __fmul_rn(a, 0.125f);
            }
            const float zh = __fadd_rn((float)z, 0.5f); // = half_up(z): exact below 2^23
            if (doU) {
                const float px = __fmaf_rn(-lds32<C00>(U0), dt, xf);
                const float py = __fmaf_rn(-av, dt, yh);
                const float pz = __fmaf_rn(-aw, dt, zh);
                u1[n] = sample_staged(win, su_a, u0, P, S, g.zlo, px, py, pz, 0.f, .5f, .5f, bx, by, bz, zv, flag);
            }
            if (doV) {
                const float px = __fmaf_rn(-au, dt, xh);
                const float py = __fmaf_rn(-lds32<C00>(V0), dt, yf);
                const float pz = __fmaf_rn(-aw, dt, zh);
                v1[n] = sample_staged(win, sv_a, v0, P, S, g.zlo, px, py, pz, .5f, 0.f, .5f, bx, by, bz, zv, flag);
            }
            if (doW) {
                const float px = __fmaf_rn(-au, dt, xh);
                const float py = __fmaf_rn(-av, dt, yh);
                const float pz = __fmaf_rn(-lds32<C00>(W0), dt, (float)z);
                w1[n] = sample_staged(win, sw_a, w0, P, S, g.zlo, px, py, pz, .5f, .5f, 0.f, bx, by, bz, zv, flag);
            }
        }
        n += g.nplane;
        __syncthreads(); // every thread is done with step z: the free slot (the one in front of sb) takes plane z+HALO+1
        if (tid == 0 && z + 1 < z_last) issue(z + A::HALO + 1, sb == 0 ? A::NSLOT - 1 : sb - 1);
        sb = sb + 1 == A::NSLOT ? 0 : sb + 1;
        if (++ws == A::NSLOT) { ws = 0; wpar ^= 1u; }
    }
}

} // namespace smk

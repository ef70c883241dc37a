// ref_cpu_driver.cpp -- runs the REFERENCE's own kernel bodies as host code.  TEST INFRASTRUCTURE.
//
// oracle/Makefile extracts the kernel section of /root/reference/project/smokeSimulation.cu
// (from `__global__ void fillSmoke` up to, not including, `int indexNow = 1;`, i.e. cu:251-705)
// at build time and pipes it into this translation unit through REF_KERNELS_INC; nothing from the
// reference is stored in this repository.  The bodies are compiled unchanged: `__global__` /
// `__device__` become empty, threadIdx/blockIdx/blockDim become thread-local variables that the
// launcher below sets for every simulated thread, and OpenMP spreads the blocks over the host cores.
//
// The launch schedule is the one in simulate() / drawObjects() (cu:714-819): same grids, same
// 8x8x8 blocks, same order.  Semantics for buffers the reference leaves uninitialised: zero (H2).
//
// Outputs: oracle/_ref/libref_cpu.so (git-ignored).  Used (a) to pin oracle/smoke_oracle.c bit-exactly,
// (b) to generate tests/golden/, (c) as the reported CPU baseline (cpu_baseline.kind = "reference").
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <vector_types.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#undef __global__
#undef __device__
#undef __host__
#define __global__
#define __device__
#define __host__

static thread_local uint3 threadIdx, blockIdx;
static thread_local dim3 blockDim;

#define OVER_RELAXATION 1.9            /* cu:38 */
#define MAX_VELOCITY_PER_STEP 3.0f     /* cu:39 */

namespace refk {
#include REF_KERNELS_INC
}

namespace {

struct Obj { int type; float x, y, z, vx, vy, vz, r; };

struct State {
    uint3 dim{}, sdim{};
    float* smoke[2]{}; float* u[2]{}; float* v[2]{}; float* w[2]{};
    bool* s = nullptr;
    float* d_obst = nullptr; float* d_src = nullptr;
    int indexNow = 1, tempIndexPast = 0;
    float gravity = -9.82f, alpha = 2.0f;
    int iterations = 30;
    std::vector<Obj> objects;
};

template <class F>
void launch(dim3 grid, dim3 block, F&& body)
{
    const long nblocks = (long)grid.x * grid.y * grid.z;
#pragma omp parallel for schedule(static)
    for (long b = 0; b < nblocks; b++) {
        blockDim = block;
        blockIdx.x = (unsigned)(b % grid.x);
        blockIdx.y = (unsigned)((b / grid.x) % grid.y);
        blockIdx.z = (unsigned)(b / ((long)grid.x * grid.y));
        for (unsigned tz = 0; tz < block.z; tz++)
            for (unsigned ty = 0; ty < block.y; ty++)
                for (unsigned tx = 0; tx < block.x; tx++) {
                    threadIdx.x = tx; threadIdx.y = ty; threadIdx.z = tz;
                    body();
                }
    }
}

dim3 full_grid(const State& st)
{
    return dim3((unsigned)std::ceil(st.dim.x / 8.0), (unsigned)std::ceil(st.dim.y / 8.0),
                (unsigned)std::ceil(st.dim.z / 8.0));
}

void draw_objects(State& st) // cu:714-771
{
    std::vector<float> src, obs;
    for (auto& o : st.objects) {
        if (o.type == 1) { src.insert(src.end(), {o.x, o.y, o.z, o.r}); }
        else { obs.insert(obs.end(), {o.x, o.y, o.z, o.vx, o.vy, o.vz, o.r}); }
    }
    const int nsrc = (int)src.size() / 4, nobs = (int)obs.size() / 7;
    float* psrc = src.data(); float* pobs = obs.data();
    dim3 blk(8, 8, 8), grd = full_grid(st);
    launch(grd, blk, [&] { refk::fillSmoke(st.smoke[0], st.smoke[1], st.dim, psrc, nsrc); });
    launch(grd, blk, [&] { refk::fillObstacle(st.s, st.dim, pobs, nobs); });
}

} // namespace

extern "C" {

void* refcpu_create(unsigned W, unsigned H, unsigned D, const float* smoke0)
{
    State* st = new State;
    st->dim = {W, H, D};
    st->sdim = {W + 1, H + 1, D + 1};
    size_t nc = (size_t)W * H * D, ns = (size_t)(W + 1) * (H + 1) * (D + 1);
    for (int i = 0; i < 2; i++) {
        st->smoke[i] = (float*)calloc(nc, sizeof(float));
        st->u[i] = (float*)calloc(ns, sizeof(float));
        st->v[i] = (float*)calloc(ns, sizeof(float));
        st->w[i] = (float*)calloc(ns, sizeof(float));
    }
    if (smoke0) memcpy(st->smoke[0], smoke0, nc * sizeof(float));
    st->s = (bool*)malloc(nc);
    memset(st->s, 1, nc);
    for (unsigned z = 0; z < D; z++)
        for (unsigned x = 0; x < W; x++) st->s[x + (size_t)z * W * H] = 0; // cu:200-207
    return st;
}

void refcpu_destroy(void* h)
{
    State* st = (State*)h;
    for (int i = 0; i < 2; i++) { free(st->smoke[i]); free(st->u[i]); free(st->v[i]); free(st->w[i]); }
    free(st->s);
    delete st;
}

void refcpu_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
}
int refcpu_get_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int refcpu_add_obstacle(void* h, float x, float y, float z, float vx, float vy, float vz, float r)
{
    State* st = (State*)h;
    st->objects.push_back({0, x, y, z, vx, vy, vz, r});
    return (int)st->objects.size() - 1;
}
int refcpu_add_source(void* h, float x, float y, float z, float r)
{
    State* st = (State*)h;
    st->objects.push_back({1, x, y, z, 0, 0, 0, r});
    return (int)st->objects.size() - 1;
}
void refcpu_update_object_pos(void* h, int id, float x, float y, float z)
{
    State* st = (State*)h;
    st->objects[id].x = x; st->objects[id].y = y; st->objects[id].z = z;
}
void refcpu_set_params(void* h, float gravity, float alpha)
{
    State* st = (State*)h;
    st->gravity = gravity; st->alpha = alpha;
}
void refcpu_set_iterations(void* h, int it) { ((State*)h)->iterations = it; }

void refcpu_flip(void* h) // cu:777-779
{
    State* st = (State*)h;
    st->tempIndexPast = st->indexNow;
    st->indexNow = st->indexNow == 0 ? 1 : 0;
}
void refcpu_fill(void* h) { draw_objects(*(State*)h); }
void refcpu_integrate(void* h, float dt) // cu:789
{
    State& st = *(State*)h; int n = st.indexNow;
    launch(full_grid(st), dim3(8, 8, 8), [&] {
        refk::integrate(st.v[n], st.smoke[n], st.s, st.dim, st.sdim, dt, st.gravity, st.alpha); });
}
void refcpu_clamp(void* h, float dt) // cu:791
{
    State& st = *(State*)h; int n = st.indexNow;
    launch(full_grid(st), dim3(8, 8, 8), [&] {
        refk::velocityConfinement(st.u[n], st.v[n], st.w[n], st.dim, st.sdim, dt); });
}
void refcpu_pressure_halfsweep(void* h, int offset) // cu:795-800
{
    State& st = *(State*)h; int n = st.indexNow;
    dim3 grd((unsigned)std::ceil(std::ceil(st.dim.x / 2.0) / 8.0), (unsigned)std::ceil(st.dim.y / 8.0),
             (unsigned)std::ceil(st.dim.z / 8.0));
    launch(grd, dim3(8, 8, 8), [&] {
        refk::divergence(st.u[n], st.v[n], st.w[n], st.s, st.dim, st.sdim, (char)offset); });
}
void refcpu_advect_velocity(void* h, float dt) // cu:805-807
{
    State& st = *(State*)h; int n = st.indexNow, p = st.tempIndexPast;
    dim3 grd = full_grid(st), blk(8, 8, 8);
    launch(grd, blk, [&] { refk::velocityAdvectionU(st.u[n], st.u[p], st.v[n], st.w[n], st.s, st.dim, st.sdim, dt); });
    launch(grd, blk, [&] { refk::velocityAdvectionV(st.v[n], st.v[p], st.u[n], st.w[n], st.s, st.dim, st.sdim, dt); });
    launch(grd, blk, [&] { refk::velocityAdvectionW(st.w[n], st.w[p], st.u[n], st.v[n], st.s, st.dim, st.sdim, dt); });
}
void refcpu_advect_smoke(void* h, float dt) // cu:810
{
    State& st = *(State*)h; int n = st.indexNow, p = st.tempIndexPast;
    launch(full_grid(st), dim3(8, 8, 8), [&] {
        refk::advectSmoke(st.smoke[n], st.smoke[p], st.u[p], st.v[p], st.w[p], st.s, st.dim, st.sdim, dt); });
}

void refcpu_step(void* h, float dt) // cu:774-819 minus the D2H copy
{
    State& st = *(State*)h;
    refcpu_flip(h);
    refcpu_fill(h);
    refcpu_integrate(h, dt);
    refcpu_clamp(h, dt);
    for (int i = 0; i < st.iterations; i++) {
        refcpu_pressure_halfsweep(h, 0);
        refcpu_pressure_halfsweep(h, 1);
    }
    refcpu_advect_velocity(h, dt);
    refcpu_advect_smoke(h, dt);
}

static void* fptr(State& st, int field, int which, size_t* bytes)
{
    int b = which == 0 ? st.indexNow : which == 1 ? st.tempIndexPast : which - 2;
    size_t nc = (size_t)st.dim.x * st.dim.y * st.dim.z, ns = (size_t)st.sdim.x * st.sdim.y * st.sdim.z;
    switch (field) {
    case 0: *bytes = nc * 4; return st.smoke[b];
    case 1: *bytes = ns * 4; return st.u[b];
    case 2: *bytes = ns * 4; return st.v[b];
    case 3: *bytes = ns * 4; return st.w[b];
    case 4: *bytes = nc; return st.s;
    }
    *bytes = 0; return nullptr;
}
void refcpu_get_field(void* h, int field, int which, void* dst)
{
    size_t n; void* p = fptr(*(State*)h, field, which, &n);
    if (p) memcpy(dst, p, n);
}
void refcpu_set_field(void* h, int field, int which, const void* src)
{
    size_t n; void* p = fptr(*(State*)h, field, which, &n);
    if (p) memcpy(p, src, n);
}
int refcpu_index_now(void* h) { return ((State*)h)->indexNow; }

} // extern "C"

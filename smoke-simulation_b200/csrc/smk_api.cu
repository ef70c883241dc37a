// smk_api.cu -- host side of libsmoke_b200.so: state, per-step kernel schedule, C ABI (include/smoke_b200.h).
//
// Replaces the host half of the reference's project/smokeSimulation.cu (globals cu:16-58, alloc/free
// cu:124-248, drawObjects cu:714-771, simulate cu:774-819) with an explicit handle, one CUDA stream, kernel
// parameters instead of per-step H2D copies of the object arrays, and events for per-stage timing.
// There is no CPU fallback anywhere in this file: if CUDA is unavailable every call fails with SMK_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/smoke_b200.h"
#include "grid.h"
#include "kernels_basic.cuh"
#include "kernels_pressure_fused.cuh"
#include "kernels_pressure_reg.cuh"

namespace {

struct Sphere { // cu:45-52
    int type;   // 0 obstacle, 1 source
    float x, y, z, vx, vy, vz, r;
};

struct TimedSpan {
    int stage;
    cudaEvent_t e0, e1;
};

} // namespace

struct smk_sim {
    GridP g{};
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;

    float* smoke[2]{};
    float* u[2]{};
    float* v[2]{};
    float* w[2]{};
    float* scratch[3]{}; // second u,v,w set for the out-of-place fused pressure passes (lazy)
    unsigned char* mask = nullptr;
    unsigned char* code = nullptr;
    int num_sms = 148;
    unsigned* d_scalar = nullptr; // device scratch for reductions

    int now = 1, past = 0; // indexNow / tempIndexPast, cu:707-708
    float gravity = -9.82f; // cu:28
    float alpha = 2.0f;     // cu:29
    std::vector<Sphere> objects;

    int solver = SMK_SOLVER_RBGS;
    int iterations = 30; // cu:797
    int fuse = 0;

    // slab decomposition (single GPU: owns everything, no ghosts)
    int c0 = 0, c1 = 0, ghost = 0;
    smk_exchange_fn exchange = nullptr;
    void* exchange_ctx = nullptr;

    // host buffers registered for fast density readback
    std::vector<void*> registered;

    // timing
    std::vector<TimedSpan> spans;
    std::vector<cudaEvent_t> free_events;
    double stage_ms[SMK_STAGE_COUNT]{};
    long stage_launches[SMK_STAGE_COUNT]{};
    long launches = 0;

    std::string err;
};

namespace {

thread_local std::string g_last_global_error;

int fail(smk_sim* s, int code, const std::string& msg)
{
    if (s) s->err = msg;
    g_last_global_error = msg;
    return code;
}

#define CK(s, call)                                                                                     \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail((s), SMK_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
    } while (0)

size_t node_count(const GridP& g) { return (size_t)g.nplane * g.nzn; }
size_t cell_count(const GridP& g) { return (size_t)g.cplane * g.nzc; }
size_t code_count(const GridP& g) { return (size_t)g.kplane * g.nzc; }

cudaEvent_t get_event(smk_sim* s)
{
    if (!s->free_events.empty()) {
        cudaEvent_t e = s->free_events.back();
        s->free_events.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

int fold_timers(smk_sim* s)
{
    if (s->spans.empty()) return SMK_OK;
    CK(s, cudaStreamSynchronize(s->stream));
    for (auto& sp : s->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.e0, sp.e1) == cudaSuccess) s->stage_ms[sp.stage] += ms;
        s->free_events.push_back(sp.e0);
        s->free_events.push_back(sp.e1);
    }
    s->spans.clear();
    return SMK_OK;
}

// RAII span: records an event pair around a stage on the step's stream
struct Span {
    smk_sim* s;
    TimedSpan sp;
    Span(smk_sim* s_, int stage) : s(s_)
    {
        sp.stage = stage;
        sp.e0 = get_event(s);
        sp.e1 = get_event(s);
        cudaEventRecord(sp.e0, s->stream);
    }
    ~Span()
    {
        cudaEventRecord(sp.e1, s->stream);
        s->spans.push_back(sp);
    }
};

void count_launch(smk_sim* s, int stage, int n = 1)
{
    s->launches += n;
    s->stage_launches[stage] += n;
}

dim3 row_grid(long long per_plane, int planes) { return dim3((unsigned)((per_plane + 255) / 256), (unsigned)planes); }

ObjP pack_objects(const smk_sim* s)
{
    ObjP o{};
    for (const Sphere& sp : s->objects) {
        if (sp.type == 1 && o.nsrc < SMK_MAX_OBJ) {
            float* d = o.src[o.nsrc++];
            d[0] = sp.x; d[1] = sp.y; d[2] = sp.z; d[3] = sp.r;
        } else if (sp.type == 0 && o.nobs < SMK_MAX_OBJ) {
            float* d = o.obs[o.nobs++];
            d[0] = sp.x; d[1] = sp.y; d[2] = sp.z; d[3] = sp.r;
        }
    }
    return o;
}

// ---- stages (global plane ranges; a single GPU passes the whole stored range) -------------------------------

int stage_fill(smk_sim* s)
{
    const GridP& g = s->g;
    Span sp(s, SMK_STAGE_FILL);
    const ObjP o = pack_objects(s);
    if (g.nzc > 0) {
        if (o.nsrc > 0 || o.nobs > 0) {
            smk::k_fill<<<row_grid(g.cplane, g.nzc), 256, 0, s->stream>>>(g, s->smoke[0], s->smoke[1], s->mask, o, g.zlo);
            count_launch(s, SMK_STAGE_FILL);
        }
        smk::k_codes<<<row_grid(g.cplane, g.nzc), 256, 0, s->stream>>>(g, s->mask, s->code, g.zlo);
        count_launch(s, SMK_STAGE_FILL);
    }
    CK(s, cudaGetLastError());
    return SMK_OK;
}

int stage_force_clamp(smk_sim* s, float dt)
{
    const GridP& g = s->g;
    Span sp(s, SMK_STAGE_FORCE);
    const int n = s->now;
    smk::k_force_clamp<<<row_grid(g.nplane, g.nzc), 256, 0, s->stream>>>(g, s->u[n], s->v[n], s->w[n], s->smoke[n], s->code,
                                                                        dt, s->gravity, s->alpha, g.zlo);
    count_launch(s, SMK_STAGE_FORCE);
    CK(s, cudaGetLastError());
    return SMK_OK;
}

// cell planes a pressure sweep may update: interior planes of the stored range
void pressure_planes(const GridP& g, int& za, int& zb)
{
    za = std::max(1, g.zlo);
    zb = std::min(g.D - 1, g.zlo + g.nzc);
}

int launch_halfsweep(smk_sim* s, int offset, int zlo_req = INT32_MIN, int zhi_req = INT32_MAX)
{
    const GridP& g = s->g;
    int za, zb;
    pressure_planes(g, za, zb);
    za = std::max(za, zlo_req);
    zb = std::min(zb, zhi_req);
    if (zb <= za) return SMK_OK;
    const int n = s->now;
    const int halfW = (g.W + 1) >> 1;
    const dim3 grid((unsigned)((halfW + 63) / 64), (unsigned)((g.H + 3) / 4), (unsigned)(zb - za));
    smk::k_pressure_half<<<grid, dim3(64, 4, 1), 0, s->stream>>>(g, s->u[n], s->v[n], s->w[n], s->code, offset, za);
    count_launch(s, SMK_STAGE_PRESSURE);
    return SMK_OK;
}

// ---- fused pressure passes (kernels_pressure_fused.cuh) ---------------------------------------------------
// z-chunking shared by the fused kernels: enough CTAs to fill the SMs, as few lead-in/lead-out planes
// (2K per chunk) as possible
int pick_zchunk(const smk_sim* s, int tiles_xy, int K)
{
    const GridP& g = s->g;
    int best_n = 1;
    double best = -1.0;
    for (int n = 1; n <= std::max(1, g.nzn / 4); n++) {
        const int zc = (g.nzn + n - 1) / n;
        const long ctas = (long)tiles_xy * ((g.nzn + zc - 1) / zc);
        const long waves = (ctas + s->num_sms - 1) / s->num_sms;
        const double eff = (double)ctas / (double)(waves * s->num_sms) * (double)zc / (double)(zc + 2 * K);
        if (eff > best + 1e-9) { best = eff; best_n = n; }
    }
    return (g.nzn + best_n - 1) / best_n;
}

int ensure_scratch(smk_sim* s)
{
    if (s->scratch[0]) return SMK_OK;
    const size_t nb = node_count(s->g) * sizeof(float);
    for (int i = 0; i < 3; i++) {
        CK(s, cudaMalloc(&s->scratch[i], nb));
        CK(s, cudaMemsetAsync(s->scratch[i], 0, nb, s->stream));
    }
    return SMK_OK;
}

void swap_in_scratch(smk_sim* s)
{
    const int n = s->now;
    std::swap(s->u[n], s->scratch[0]);
    std::swap(s->v[n], s->scratch[1]);
    std::swap(s->w[n], s->scratch[2]);
}

template <int K, int FUSED_NW, int FUSED_RPW>
int launch_fused_pass_cfg(smk_sim* s, int sweep0)
{
    using C = smk::FusedCfg<K, FUSED_NW, FUSED_RPW>;
    const GridP& g = s->g;
    static bool configured = false;
    auto kern = smk::k_pressure_fused<K, FUSED_NW, FUSED_RPW>;
    if (!configured) {
        CK(s, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        configured = true;
    }
    int rc = ensure_scratch(s);
    if (rc) return rc;
    const int tx = (g.W + 1 + C::OX - 1) / C::OX, ty = (g.SY + C::OY - 1) / C::OY;
    const int zchunk = pick_zchunk(s, tx * ty, K);
    const dim3 grid((unsigned)tx, (unsigned)ty, (unsigned)((g.nzn + zchunk - 1) / zchunk));
    const int n = s->now;
    kern<<<grid, C::THREADS, C::SMEM, s->stream>>>(g, s->u[n], s->v[n], s->w[n], s->scratch[0], s->scratch[1], s->scratch[2],
                                                    s->code, sweep0, zchunk);
    swap_in_scratch(s);
    count_launch(s, SMK_STAGE_PRESSURE);
    return SMK_OK;
}

// register-resident fused pass (kernels_pressure_reg.cuh): u, w in registers, v in shared memory
template <int K, int NW>
int launch_reg_pass(smk_sim* s, int sweep0)
{
    using C = smk::RegCfg<K, NW>;
    const GridP& g = s->g;
    static bool configured = false;
    auto kern = smk::k_pressure_reg<K, NW>;
    if (!configured) {
        CK(s, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        configured = true;
    }
    int rc = ensure_scratch(s);
    if (rc) return rc;
    const int tx = (g.W + 1 + C::OX - 1) / C::OX, ty = (g.SY + C::OY - 1) / C::OY;
    const int zchunk = pick_zchunk(s, tx * ty, K);
    const dim3 grid((unsigned)tx, (unsigned)ty, (unsigned)((g.nzn + zchunk - 1) / zchunk));
    const int n = s->now;
    kern<<<grid, C::THREADS, C::SMEM, s->stream>>>(g, s->u[n], s->v[n], s->w[n], s->scratch[0], s->scratch[1], s->scratch[2],
                                                    s->code, sweep0, zchunk);
    swap_in_scratch(s);
    count_launch(s, SMK_STAGE_PRESSURE);
    return SMK_OK;
}

template <int K>
int launch_fused_pass(smk_sim* s, int sweep0)
{
    // tile shape (warps x rows per warp); tuning knob SMK_FUSED_CFG for experiments, default 24x2
    static const int cfg = getenv("SMK_FUSED_CFG") ? atoi(getenv("SMK_FUSED_CFG")) : 0;
    if (K == 4 && cfg == 0) return launch_reg_pass<4, 16>(s, sweep0);
    switch (cfg) {
    case 163: return launch_fused_pass_cfg<K, 16, 3>(s, sweep0);
    case 124: return launch_fused_pass_cfg<K, 12, 4>(s, sweep0);
    case 321: return launch_fused_pass_cfg<K, 32, 1>(s, sweep0);
    case 162: return launch_fused_pass_cfg<K, 16, 2>(s, sweep0);
    default: return launch_fused_pass_cfg<K, 24, 2>(s, sweep0);
    }
}

int stage_pressure(smk_sim* s)
{
    Span sp(s, SMK_STAGE_PRESSURE);
    const int total = 2 * s->iterations; // half-sweeps, alternating offset 0,1 (cu:797-801)
    int fuse = s->fuse;
    if (fuse == 0) fuse = (s->g.W + 1 < 32 || s->g.nzn < 8) ? 1 : 4; // tiny grids: tiles would be mostly halo
    int done = 0, rc = SMK_OK;
    // EXPERIMENT (env SMK_L2_ZC / SMK_L2_G): L2-resident wavefront of the unfused kernel.  Groups of G
    // half-sweeps are applied chunk by chunk (Zc planes), sweep j shifted down by j planes, so the planes
    // a group touches stay in the 126 MB L2 between launches.  Same per-cell order of updates => same bits.
    static const int l2_zc = getenv("SMK_L2_ZC") ? atoi(getenv("SMK_L2_ZC")) : 0;
    static const int l2_g = getenv("SMK_L2_G") ? atoi(getenv("SMK_L2_G")) : 20;
    if (fuse == 1 && l2_zc > 0) {
        const GridP& g = s->g;
        while (done < total) {
            const int G = std::min(l2_g, total - done);
            for (int c0 = 0; c0 - G < g.D; c0 += l2_zc)
                for (int j = 0; j < G; j++) launch_halfsweep(s, (done + j) & 1, c0 - j, c0 + l2_zc - j);
            done += G;
        }
    }
    while (done < total && rc == SMK_OK) {
        const int left = total - done;
        if (fuse >= 4 && left >= 4 && (done & 1) == 0) { rc = launch_fused_pass<4>(s, done); done += 4; }
        else if (fuse >= 2 && left >= 2 && (done & 1) == 0) { rc = launch_fused_pass<2>(s, done); done += 2; }
        else { rc = launch_halfsweep(s, done & 1); done += 1; }
    }
    if (rc) return rc;
    CK(s, cudaGetLastError());
    return SMK_OK;
}

int stage_advect_velocity(smk_sim* s, float dt)
{
    const GridP& g = s->g;
    Span sp(s, SMK_STAGE_ADVECT_VEL);
    const int n = s->now, p = s->past;
    const int za = std::max(1, g.zlo), zb = std::min(g.D, g.zlo + g.nzc);
    if (zb > za) {
        smk::k_advect_velocity<<<row_grid(g.nplane, zb - za), 256, 0, s->stream>>>(g, s->u[n], s->v[n], s->w[n], s->u[p],
                                                                               s->v[p], s->w[p], s->code, dt, za);
        count_launch(s, SMK_STAGE_ADVECT_VEL);
    }
    CK(s, cudaGetLastError());
    return SMK_OK;
}

int stage_advect_smoke(smk_sim* s, float dt)
{
    const GridP& g = s->g;
    Span sp(s, SMK_STAGE_ADVECT_SMOKE);
    const int n = s->now, p = s->past;
    int za, zb;
    pressure_planes(g, za, zb);
    if (zb > za) {
        smk::k_advect_smoke<<<row_grid(g.cplane, zb - za), 256, 0, s->stream>>>(g, s->smoke[n], s->smoke[p], s->u[p], s->v[p],
                                                                            s->w[p], s->code, dt, za);
        count_launch(s, SMK_STAGE_ADVECT_SMOKE);
    }
    CK(s, cudaGetLastError());
    return SMK_OK;
}

void flip(smk_sim* s) // cu:777-779
{
    s->past = s->now;
    s->now = s->now == 0 ? 1 : 0;
}

bool try_register(smk_sim* s, void* p, size_t bytes)
{
    if (std::find(s->registered.begin(), s->registered.end(), p) != s->registered.end()) return true;
    static const bool off = getenv("SMK_NO_HOST_REGISTER") != nullptr;
    if (off) return false;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type != cudaMemoryTypeUnregistered) return true; // already pinned
    cudaGetLastError();
    if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    s->registered.push_back(p);
    return true;
}

int enqueue_step(smk_sim* s, float dt, float* density_host)
{
    int rc;
    flip(s);
    if ((rc = stage_fill(s))) return rc;
    if ((rc = stage_force_clamp(s, dt))) return rc;
    if ((rc = stage_pressure(s))) return rc;
    if ((rc = stage_advect_velocity(s, dt))) return rc;
    if ((rc = stage_advect_smoke(s, dt))) return rc;
    if (density_host) {
        Span sp(s, SMK_STAGE_READBACK);
        const size_t bytes = cell_count(s->g) * sizeof(float);
        float* dst = density_host + (size_t)s->g.zlo * s->g.cplane;
        try_register(s, dst, bytes);
        CK(s, cudaMemcpyAsync(dst, s->smoke[s->past], bytes, cudaMemcpyDeviceToHost, s->stream));
    }
    return SMK_OK;
}

struct FieldRef {
    void* dev;
    size_t elem;
    bool staggered;
};

int field_ref(smk_sim* s, int field, int which, FieldRef* out)
{
    if (which < 0 || which > 3) return SMK_ERR_ARG;
    const int b = which == SMK_BUF_NOW ? s->now : which == SMK_BUF_PAST ? s->past : which - 2;
    switch (field) {
    case SMK_FIELD_SMOKE: *out = {s->smoke[b], 4, false}; return SMK_OK;
    case SMK_FIELD_U: *out = {s->u[b], 4, true}; return SMK_OK;
    case SMK_FIELD_V: *out = {s->v[b], 4, true}; return SMK_OK;
    case SMK_FIELD_W: *out = {s->w[b], 4, true}; return SMK_OK;
    case SMK_FIELD_MASK: *out = {s->mask, 1, false}; return SMK_OK;
    }
    return SMK_ERR_ARG;
}

// copy between the reference layout on the host and the internal layout on the device
int copy_field(smk_sim* s, const FieldRef& f, void* host, bool to_host)
{
    const GridP& g = s->g;
    cudaMemcpy3DParms p{};
    if (f.staggered) {
        const size_t hx = (size_t)(g.W + 1), hy = (size_t)(g.H + 1);
        char* hbase = (char*)host + (size_t)g.zlo * hx * hy * f.elem;
        cudaPitchedPtr hp = make_cudaPitchedPtr(hbase, hx * f.elem, hx, hy);
        cudaPitchedPtr dp = make_cudaPitchedPtr(f.dev, (size_t)g.P * f.elem, (size_t)g.P, (size_t)g.SY);
        p.srcPtr = to_host ? dp : hp;
        p.dstPtr = to_host ? hp : dp;
        p.extent = make_cudaExtent(hx * f.elem, hy, (size_t)g.nzn);
    } else {
        const size_t hx = (size_t)g.W, hy = (size_t)g.H;
        char* hbase = (char*)host + (size_t)g.zlo * hx * hy * f.elem;
        cudaPitchedPtr hp = make_cudaPitchedPtr(hbase, hx * f.elem, hx, hy);
        cudaPitchedPtr dp = make_cudaPitchedPtr(f.dev, hx * f.elem, hx, hy);
        p.srcPtr = to_host ? dp : hp;
        p.dstPtr = to_host ? hp : dp;
        p.extent = make_cudaExtent(hx * f.elem, hy, (size_t)g.nzc);
    }
    p.kind = to_host ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice;
    CK(s, cudaMemcpy3DAsync(&p, s->stream));
    CK(s, cudaStreamSynchronize(s->stream));
    return SMK_OK;
}

int create_common(smk_sim** out, unsigned W, unsigned H, unsigned D, int c0, int c1, int ghost, const float* smoke0_full)
{
    if (!out) return SMK_ERR_ARG;
    *out = nullptr;
    if (W < 3 || H < 3 || D < 3 || W > 4096 || H > 4096 || D > 65535) return fail(nullptr, SMK_ERR_ARG, "grid dimensions out of range");
    if (c0 < 0 || c1 > (int)D || c0 >= c1) return fail(nullptr, SMK_ERR_ARG, "bad slab range");
    smk_sim* s = new smk_sim;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        std::string m = std::string("cudaGetDevice: ") + cudaGetErrorString(e) + " (libsmoke_b200 has no CPU fallback)";
        delete s;
        return fail(nullptr, SMK_ERR_CUDA, m);
    }
    s->device = dev;
    GridP& g = s->g;
    g.W = (int)W; g.H = (int)H; g.D = (int)D;
    g.P = (int)((W + 1 + 7) / 8 * 8);
    g.SY = (int)H + 1;
    g.nplane = (long long)g.P * g.SY;
    g.cplane = (long long)W * H;
    s->c0 = c0; s->c1 = c1; s->ghost = ghost;
    g.zlo = std::max(0, c0 - ghost);
    const int zhc = std::min((int)D, c1 + ghost);
    g.nzc = zhc - g.zlo;
    g.nzn = g.nzc + 1;
    g.PC = (int)((W + 15) / 16 * 16);
    g.kplane = (long long)g.PC * g.H;
    cudaDeviceGetAttribute(&s->num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (s->num_sms <= 0) s->num_sms = 148;

#define CKN(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            std::string m_ = std::string(#call) + ": " + cudaGetErrorString(e_);                        \
            smk_destroy(s);                                                                             \
            return fail(nullptr, SMK_ERR_CUDA, m_);                                                     \
        }                                                                                               \
    } while (0)

    CKN(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    const size_t nb = node_count(g) * sizeof(float), cb = cell_count(g) * sizeof(float);
    for (int i = 0; i < 2; i++) {
        CKN(cudaMalloc(&s->smoke[i], cb));
        CKN(cudaMalloc(&s->u[i], nb));
        CKN(cudaMalloc(&s->v[i], nb));
        CKN(cudaMalloc(&s->w[i], nb));
        CKN(cudaMemsetAsync(s->smoke[i], 0, cb, s->stream));
        CKN(cudaMemsetAsync(s->u[i], 0, nb, s->stream));
        CKN(cudaMemsetAsync(s->v[i], 0, nb, s->stream));
        CKN(cudaMemsetAsync(s->w[i], 0, nb, s->stream));
    }
    CKN(cudaMalloc(&s->mask, cell_count(g)));
    CKN(cudaMalloc(&s->code, code_count(g)));
    CKN(cudaMalloc(&s->d_scalar, 64));
    CKN(cudaMemsetAsync(s->code, 0, code_count(g), s->stream));
    // mask: fluid everywhere, solid on the plane y == 0 (cu:200-207)
    CKN(cudaMemsetAsync(s->mask, 1, cell_count(g), s->stream));
    CKN(cudaMemset2DAsync(s->mask, (size_t)g.cplane, 0, (size_t)g.W, (size_t)g.nzc, s->stream));
    if (smoke0_full)
        CKN(cudaMemcpyAsync(s->smoke[0], smoke0_full + (size_t)g.zlo * g.cplane, cb, cudaMemcpyHostToDevice, s->stream));
    CKN(cudaStreamSynchronize(s->stream));
#undef CKN
    *out = s;
    return SMK_OK;
}

} // namespace

// ======================================================================================================
extern "C" {

int smk_abi_version(void) { return SMK_ABI_VERSION; }

const char* smk_last_error(smk_sim* s) { return s ? s->err.c_str() : g_last_global_error.c_str(); }

int smk_create(smk_sim** out, unsigned W, unsigned H, unsigned D, const float* smoke0_host)
{
    return create_common(out, W, H, D, 0, (int)D, 0, smoke0_host);
}

int smk_create_slab(smk_sim** out, unsigned W, unsigned H, unsigned D, unsigned z_begin, unsigned z_end, unsigned ghost,
                    const float* smoke0_host_full)
{
    return create_common(out, W, H, D, (int)z_begin, (int)z_end, (int)ghost, smoke0_host_full);
}

int smk_destroy(smk_sim* s)
{
    if (!s) return SMK_ERR_ARG;
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (void* p : s->registered) cudaHostUnregister(p);
    for (auto& sp : s->spans) { cudaEventDestroy(sp.e0); cudaEventDestroy(sp.e1); }
    for (auto e : s->free_events) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) {
        cudaFree(s->smoke[i]); cudaFree(s->u[i]); cudaFree(s->v[i]); cudaFree(s->w[i]);
    }
    for (int i = 0; i < 3; i++) cudaFree(s->scratch[i]);
    cudaFree(s->mask); cudaFree(s->code); cudaFree(s->d_scalar);
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    cudaGetLastError();
    delete s;
    return SMK_OK;
}

int smk_print_gpu_properties(void) // cu:62-85
{
    int devices = 0;
    cudaError_t err = cudaGetDeviceCount(&devices);
    if (err != cudaSuccess) return fail(nullptr, SMK_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(err));
    for (int i = 0; i < devices; i++) {
        cudaDeviceProp prop;
        printf("CUDA Device - ID %d\n", i);
        if (cudaGetDeviceProperties(&prop, i) == cudaSuccess) {
            printf("Name: \t\t\t\t%s (sm_%d%d, %d SMs)\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
            printf("Max threads per block: \t\t%d\n", prop.maxThreadsPerBlock);
            printf("Max block dimensions: \t\t(%d, %d, %d)\n", prop.maxThreadsDim[0], prop.maxThreadsDim[1], prop.maxThreadsDim[2]);
            printf("Max grid dimensions: \t\t(%d, %d, %d)\n", prop.maxGridSize[0], prop.maxGridSize[1], prop.maxGridSize[2]);
            printf("Shared memory per block: \t%.2lfKB\n", prop.sharedMemPerBlock / 1024.);
        }
        printf("\n");
    }
    return SMK_OK;
}

int smk_add_obstacle(smk_sim* s, float x, float y, float z, float vx, float vy, float vz, float r)
{
    if (!s) return -SMK_ERR_ARG;
    int n = 0;
    for (auto& o : s->objects) n += o.type == 0;
    if (n >= SMK_MAX_OBJECTS) return -SMK_ERR_LIMIT;
    s->objects.push_back({0, x, y, z, vx, vy, vz, r});
    return (int)s->objects.size() - 1;
}

int smk_add_source(smk_sim* s, float x, float y, float z, float r)
{
    if (!s) return -SMK_ERR_ARG;
    int n = 0;
    for (auto& o : s->objects) n += o.type == 1;
    if (n >= SMK_MAX_OBJECTS) return -SMK_ERR_LIMIT;
    s->objects.push_back({1, x, y, z, 0.f, 0.f, 0.f, r});
    return (int)s->objects.size() - 1;
}

int smk_update_object_pos(smk_sim* s, int id, float x, float y, float z)
{
    if (!s || id < 0 || id >= (int)s->objects.size()) return SMK_ERR_ARG;
    s->objects[id].x = x; s->objects[id].y = y; s->objects[id].z = z;
    return SMK_OK;
}

float* smk_gravity_ptr(smk_sim* s) { return s ? &s->gravity : nullptr; }
float* smk_buoyancy_ptr(smk_sim* s) { return s ? &s->alpha : nullptr; }

int smk_set_solver(smk_sim* s, int variant, int iterations, int fuse)
{
    if (!s || iterations < 0 || fuse < 0) return SMK_ERR_ARG;
    if (variant != SMK_SOLVER_RBGS) return fail(s, SMK_ERR_ARG, "solver variant not available");
    if (fuse != 0 && fuse != 1 && fuse != 2 && fuse != 4) return fail(s, SMK_ERR_ARG, "fuse must be 0, 1, 2 or 4");
    s->solver = variant; s->iterations = iterations; s->fuse = fuse;
    return SMK_OK;
}

int smk_set_stream(smk_sim* s, void* cuda_stream)
{
    if (!s) return SMK_ERR_ARG;
    CK(s, cudaStreamSynchronize(s->stream));
    int rc = fold_timers(s);
    if (rc) return rc;
    if (cuda_stream) {
        if (s->own_stream) cudaStreamDestroy(s->stream);
        s->stream = (cudaStream_t)cuda_stream;
        s->own_stream = false;
    } else if (!s->own_stream) {
        CK(s, cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        s->own_stream = true;
    }
    return SMK_OK;
}

int smk_step_async(smk_sim* s, float dt, float* density_host)
{
    if (!s) return SMK_ERR_ARG;
    return enqueue_step(s, dt, density_host);
}

int smk_sync(smk_sim* s)
{
    if (!s) return SMK_ERR_ARG;
    CK(s, cudaStreamSynchronize(s->stream));
    if (s->spans.size() > 4096) return fold_timers(s);
    return SMK_OK;
}

int smk_step(smk_sim* s, float dt, float* density_host)
{
    if (!s) return SMK_ERR_ARG;
    int rc = enqueue_step(s, dt, density_host);
    if (rc) return rc;
    return smk_sync(s);
}

const float* smk_density_device(smk_sim* s) { return s ? s->smoke[s->past] : nullptr; }

int smk_stage_flip(smk_sim* s) { if (!s) return SMK_ERR_ARG; flip(s); return SMK_OK; }
int smk_stage_fill(smk_sim* s) { return s ? stage_fill(s) : SMK_ERR_ARG; }
int smk_stage_force_clamp(smk_sim* s, float dt) { return s ? stage_force_clamp(s, dt) : SMK_ERR_ARG; }
int smk_stage_pressure_halfsweep(smk_sim* s, int offset)
{
    if (!s) return SMK_ERR_ARG;
    Span sp(s, SMK_STAGE_PRESSURE);
    launch_halfsweep(s, offset & 1);
    CK(s, cudaGetLastError());
    return SMK_OK;
}
int smk_stage_pressure(smk_sim* s) { return s ? stage_pressure(s) : SMK_ERR_ARG; }
int smk_stage_advect_velocity(smk_sim* s, float dt) { return s ? stage_advect_velocity(s, dt) : SMK_ERR_ARG; }
int smk_stage_advect_smoke(smk_sim* s, float dt) { return s ? stage_advect_smoke(s, dt) : SMK_ERR_ARG; }

int smk_get_field(smk_sim* s, int field, int which, void* host_dst)
{
    if (!s || !host_dst) return SMK_ERR_ARG;
    FieldRef f;
    if (field_ref(s, field, which, &f)) return fail(s, SMK_ERR_ARG, "bad field / buffer selector");
    return copy_field(s, f, host_dst, true);
}

int smk_set_field(smk_sim* s, int field, int which, const void* host_src)
{
    if (!s || !host_src) return SMK_ERR_ARG;
    FieldRef f;
    if (field_ref(s, field, which, &f)) return fail(s, SMK_ERR_ARG, "bad field / buffer selector");
    int rc = copy_field(s, f, const_cast<void*>(host_src), false);
    if (rc) return rc;
    if (field == SMK_FIELD_MASK) { // keep the stencil codes consistent with an injected mask
        const GridP& g = s->g;
        smk::k_codes<<<row_grid(g.cplane, g.nzc), 256, 0, s->stream>>>(g, s->mask, s->code, g.zlo);
        count_launch(s, SMK_STAGE_FILL);
        CK(s, cudaGetLastError());
        CK(s, cudaStreamSynchronize(s->stream));
    }
    return SMK_OK;
}

int smk_index_now(smk_sim* s) { return s ? s->now : -1; }

int smk_max_divergence(smk_sim* s, float* out)
{
    if (!s || !out) return SMK_ERR_ARG;
    const GridP& g = s->g;
    int za, zb;
    pressure_planes(g, za, zb);
    CK(s, cudaMemsetAsync(s->d_scalar, 0, 4, s->stream));
    if (zb > za) {
        const int n = s->now;
        smk::k_max_divergence<<<row_grid(g.cplane, zb - za), 256, 0, s->stream>>>(g, s->u[n], s->v[n], s->w[n], s->code,
                                                                              s->d_scalar, za);
        s->launches++;
    }
    unsigned bits = 0;
    CK(s, cudaMemcpyAsync(&bits, s->d_scalar, 4, cudaMemcpyDeviceToHost, s->stream));
    CK(s, cudaStreamSynchronize(s->stream));
    memcpy(out, &bits, 4);
    return SMK_OK;
}

int smk_stage_time(smk_sim* s, int stage, double* ms_total, long* launches)
{
    if (!s || stage < 0 || stage >= SMK_STAGE_COUNT) return SMK_ERR_ARG;
    int rc = fold_timers(s);
    if (rc) return rc;
    if (ms_total) *ms_total = s->stage_ms[stage];
    if (launches) *launches = s->stage_launches[stage];
    return SMK_OK;
}

int smk_reset_timers(smk_sim* s)
{
    if (!s) return SMK_ERR_ARG;
    int rc = fold_timers(s);
    if (rc) return rc;
    for (int i = 0; i < SMK_STAGE_COUNT; i++) { s->stage_ms[i] = 0; s->stage_launches[i] = 0; }
    return SMK_OK;
}

long smk_launch_count(smk_sim* s) { return s ? s->launches : -1; }

int smk_set_exchange(smk_sim* s, smk_exchange_fn fn, void* ctx)
{
    if (!s) return SMK_ERR_ARG;
    s->exchange = fn; s->exchange_ctx = ctx;
    return SMK_OK;
}

} // extern "C"

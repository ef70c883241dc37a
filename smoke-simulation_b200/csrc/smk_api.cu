// smk_api.cu -- host side of libsmoke_b200.so: state, per-step kernel schedule, C ABI (include/smoke_b200.h).
//
// Replaces the host half of the reference's project/smokeSimulation.cu (globals cu:16-58, alloc/free
// cu:124-248, drawObjects cu:714-771, simulate cu:774-819) with an explicit handle, one CUDA stream, kernel
// parameters instead of per-step H2D copies of the object arrays, and events for per-stage timing.
// There is no CPU fallback anywhere in this file: if CUDA is unavailable every call fails with SMK_ERR_CUDA.
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h> // header-only (NVTX v3): ranges per stage for Nsight Systems timelines

#include <algorithm>
#include <array>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <queue>
#include <string>
#include <vector>

#include "../../include/smoke_b200.h"
#include "grid.h"
#include "kernels_basic.cuh"
#include "kernels_pressure_fused.cuh"
#include "kernels_pressure_reg.cuh"
#include "kernels_pressure_tma.cuh"
#include "kernels_advect_tma.cuh"
#include "kernels_jacobi.cuh"
#include "slab_plan.h"
#include "pass_schedule.h"

namespace {

struct Sphere { // cu:45-52
    int type;   // 0 obstacle, 1 source
    float x, y, z, vx, vy, vz, r;
};

struct TimedSpan {
    int stage;
    cudaEvent_t e0, e1;
};

// byte offsets inside a slab's arena: [epoch counters][density x2][u x3][v x3][w x3] (x3 = ping, pong, scratch)
struct ArenaLayout {
    size_t counters = 0, smoke[2] = {0, 0}, u[3] = {0, 0, 0}, v[3] = {0, 0, 0}, w[3] = {0, 0, 0}, total = 0;
};

ArenaLayout make_layout(const slab::Geom& geom)
{
    const size_t P = (size_t)((geom.W + 1 + 7) / 8 * 8), SY = (size_t)geom.H + 1;
    const size_t nzc = (size_t)(geom.zhc - geom.zlo), nzn = nzc + 1;
    const size_t nb = (P * SY * nzn * sizeof(float) + 255) / 256 * 256;
    const size_t cb = ((size_t)geom.W * geom.H * nzc * sizeof(float) + 255) / 256 * 256;
    ArenaLayout l;
    size_t o = 4096;
    for (int i = 0; i < 2; i++) { l.smoke[i] = o; o += cb; }
    for (int i = 0; i < 3; i++) { l.u[i] = o; o += nb; }
    for (int i = 0; i < 3; i++) { l.v[i] = o; o += nb; }
    for (int i = 0; i < 3; i++) { l.w[i] = o; o += nb; }
    l.total = o;
    return l;
}

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
    unsigned char* pcode = nullptr; // the same stencil information in the encoding of the fused pressure passes (grid.h)
    unsigned char* cflag = nullptr; // [stored cell plane][tile y][tile x]: the plane holds a COMPLEX cell inside that K = 4 tile
    int cflag_tx = 0, cflag_ty = 0;
    int num_sms = 148;
    unsigned* d_scalar = nullptr; // device scratch for reductions

    int now = 1, past = 0; // indexNow / tempIndexPast, cu:707-708
    float gravity = -9.82f; // cu:28
    float alpha = 2.0f;     // cu:29
    std::vector<Sphere> objects;

    int solver = SMK_SOLVER_RBGS;
    int pass_epoch_next = -1;   // half-sweep index whose handshake epoch the previous pass kernel publishes itself
    long long* d_passdbg = nullptr; // SMK_PASS_DEBUG=1: {start clock, cycles, SM id, variant} of every CTA of the last fused pass
    int passdbg_ctas = 0;
    unsigned* d_dyn = nullptr;  // device words: [0] max |w| bits (atomicMax), [1] advection margin in planes
    bool pending_force = false; // forcing + clamp of this step are applied by the first pressure pass (fused)
    float pending_dt = 0.f;
    int pending_a = 0, pending_b = 0;
    int obstacle_union = 0;     // SURVEY N3 extension: 0 = reference semantics (last obstacle decides)
    bool mask_dirty = true; // the mask / stencil codes on the device do not reflect the current obstacle list yet
    int iterations = 30; // cu:797
    int fuse = 0;
    int pass_ctas = getenv("SMK_PASS_CTAS") ? atoi(getenv("SMK_PASS_CTAS")) : 0; // smk_set_pass_ctas (0: chunk grid)
    int last_pass_ctas = 0; // CTAs of the last balanced pass launch (0: it was a (tile, z-chunk) grid)
    int last_pass_kernel = 0; // SMK_PASS_REG / SMK_PASS_TMA: what the last fused pass ran on

    // slab decomposition (single GPU: owns everything, no ghosts); schedule and validity tracking in slab_plan.h
    slab::Geom geom{};
    slab::Carry carry{};
    smk_exchange_fn exchange = nullptr;
    void* exchange_ctx = nullptr;
    int* d_flags = nullptr; // [0]: a backtrace left the valid planes of a slab (SMK_ERR_REACH); [1]: peer wait timed out
    long exchanges = 0;
    unsigned long long readback_bytes = 0; // bytes the steps copied device -> host so far (smk_readback_bytes)

    // one allocation for every exchanged field so that a neighbour process can map it with ONE CUDA IPC handle
    char* arena = nullptr;
    ArenaLayout lay{};
    int vel_id[2] = {0, 1}; // physical buffer (0..2) behind u/v/w[0], u/v/w[1] ...
    int scratch_id = 2;     // ... and behind the scratch set; identical on every rank (same swap history)
    struct Peer {
        char* arena = nullptr;
        bool ipc = false;
        slab::Geom geom{};
        ArenaLayout lay{};
    } peer[2];              // [0] lower-z neighbour, [1] upper-z neighbour
    bool p2p = false;       // all existing neighbours attached: native peer-memory halo path
    unsigned epoch = 0;
    cudaStream_t aux_stream = nullptr; // boundary z-chunks of a peer-reading pass run here, behind the epoch wait,
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr; // while the interior chunks already run on the main stream

    // TMA descriptors of the three physical buffers of u, v, w (box = advection tile + halo), [field][physical id]
    CUtensorMap tmap[3][3];
    bool tma_ok = false;
    // ... of the fused pressure pass (box = one 64 x 32 tile plane; kernels_pressure_tma.cuh): [source][field][physical id]
    // with source 0 = this slab, 1 / 2 = the lower / upper neighbour's peer-mapped arena; the stencil codes; the density
    // [source][buffer]
    CUtensorMap pmap[3][3];   // [source][physical id]: u, v, w as one 4-D tensor (x, y, z, field)
    CUtensorMap pmap_code;
    CUtensorMap pmap_smoke[3][2];
    bool pass_tma_ok = false, pass_tma_smoke_ok = false;
    int pass_kernel = SMK_PASS_AUTO; // smk_set_pass_kernel
    bool pass_tma_peer_ok[2] = {false, false};

    // host ranges this handle page-locked for the density readback (smk_register_host, or implicitly by smk_step)
    struct HostReg { char* p; size_t bytes; bool implicit; };
    std::vector<HostReg> registered;
    // pipelined readback (smk_step_async with a host buffer): the new density is snapshotted device-to-device and
    // copied to the host on a second stream while the next step computes
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_snap = nullptr, ev_copied = nullptr;
    cudaEvent_t ev_chunk[8] = {}; // chunked blocking readback (enqueue_step)
    cudaSurfaceObject_t density_surf = 0; // smk_bind_density_array: the density advection also writes this 3-D surface (N1)
    cudaArray_t density_array = nullptr;
    float* snapshot = nullptr;
    void* half_stage = nullptr; // binary16 staging buffer of smk_read_density_half
    bool copy_pending = false;
    // sparse blocking readback (smk_set_readback_box / SMK_READBACK_BOX=1, single GPU): rows that can hold non-zero density only
    bool box_mode = getenv("SMK_READBACK_BOX") && atoi(getenv("SMK_READBACK_BOX")) != 0;
    int* d_box = nullptr;          // device {y0, y1, z0, z1} of the new density's non-zero rows
    int* h_box = nullptr;          // pinned mirror, valid after the step's stream has been synchronised
    int host_box[4] = {0, 0, 0, 0}; // non-zero rows of what the caller's buffer holds (after the last readback)
    const float* host_box_buf = nullptr; // the buffer host_box describes (nullptr: unknown -> full copy)
    float* box_pending_host = nullptr;   // readback whose box check is due at the next smk_sync
    int box_copied[4] = {0, 0, 0, 0};    // rows copied by that readback (y0 >= y1: nothing, z0 < 0: everything)

    // timing
    std::vector<TimedSpan> spans;
    std::vector<cudaEvent_t> free_events;
    double stage_ms[SMK_STAGE_COUNT]{};
    long stage_launches[SMK_STAGE_COUNT]{};
    long launches = 0;

    // balanced piece lists of the fused pressure passes (pass_schedule.h), one per (range, K) seen so far
    struct DevSchedule {
        int tx, ty, lo, hi, lead, nctas, zl, zh; // zl / zh: boundary-piece limits (INT_MIN / INT_MAX: none)
        int4* pieces;
        int* first;
        int launch_ctas, cost;
        int nboundary[2];
    };
    std::vector<DevSchedule> schedules;
    std::vector<std::array<int, 5>> zchunk_cache; // pick_zchunk_tma: {tx, ty, K, planes, chunk}
    std::vector<std::pair<std::array<int, 4>, std::vector<int>>> chunks_cache; // pick_chunks_tma: {tx, ty, K, planes} -> lengths
    std::vector<const void*> configured; // kernels whose dynamic shared-memory limit has been raised on this handle's device
    std::string err;
};

namespace {

// Every entry point that touches the device runs on the device the handle was created on and restores the caller's
// current device afterwards (a second handle on another GPU of the same process, or a caller that
// changed the current device between calls, used to launch into the wrong context).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(const smk_sim* s)
    {
        if (!s) return;
        if (cudaGetDevice(&prev) == cudaSuccess && prev != s->device) switched = cudaSetDevice(s->device) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if (switched) cudaSetDevice(prev);
    }
};

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
        static const char* const names[SMK_STAGE_COUNT] = {"smk:fill", "smk:force+clamp", "smk:pressure", "smk:advect u,v,w", "smk:advect density", "smk:readback"};
        nvtxRangePushA(names[stage]);
        sp.stage = stage;
        sp.e0 = get_event(s);
        sp.e1 = get_event(s);
        cudaEventRecord(sp.e0, s->stream);
    }
    ~Span()
    {
        cudaEventRecord(sp.e1, s->stream);
        s->spans.push_back(sp);
        nvtxRangePop();
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
    o.obstacle_union = s->obstacle_union;
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

void launch_codes(smk_sim* s)
{
    const GridP& g = s->g;
    const smk::TileFlags tf{s->cflag, s->cflag_tx, s->cflag_ty};
    if (s->cflag) cudaMemsetAsync(s->cflag, 0, (size_t)g.nzc * s->cflag_tx * s->cflag_ty, s->stream);
    if ((g.W & 3) == 0)
        smk::k_codes4<<<row_grid(g.cplane / 4, g.nzc), 256, 0, s->stream>>>(g, s->mask, s->code, s->pcode, g.zlo, tf);
    else
        smk::k_codes<<<row_grid(g.cplane, g.nzc), 256, 0, s->stream>>>(g, s->mask, s->code, s->pcode, g.zlo, tf);
    count_launch(s, SMK_STAGE_FILL);
}

// Bounding boxes of the sources (|d| < |r| + 1 per axis, clipped to the interior); false if a launch over them is
// not possible (non-finite parameters, grid.z limit) -- the caller then scans the whole grid.
bool source_boxes(const smk_sim* s, const ObjP& o, smk::SrcBoxes& b)
{
    const GridP& g = s->g;
    const int dim[3] = {g.W, g.H, g.D};
    b.n[0] = b.n[1] = b.n[2] = 0;
    for (int k = 0; k < o.nsrc; k++) {
        for (int a = 0; a < 3; a++) {
            const double c = o.src[k][a], r1 = std::fabs((double)o.src[k][3]) + 1.0;
            if (!std::isfinite(c) || !std::isfinite(r1)) return false;
            const double lo = std::max(1.0, std::floor(c - r1)), hi = std::min((double)dim[a] - 2.0, std::ceil(c + r1));
            b.lo[k][a] = (int)std::min(lo, (double)dim[a]);
            b.n[a] = std::max(b.n[a], hi >= lo ? (int)(hi - lo) + 1 : 0);
        }
    }
    return (long long)b.n[2] * o.nsrc <= 65535;
}

// Source + mask fill.  The reference rewrites the mask from the obstacle list on every step (cu:289-313, 714-771); the
// result is a pure function of that list, so it -- and the stencil codes derived from it -- is recomputed only when
// the list or the mask changed (s->mask_dirty).  The sources are stamped every step, over their bounding boxes.
int stage_fill(smk_sim* s)
{
    const GridP& g = s->g;
    Span sp(s, SMK_STAGE_FILL);
    const ObjP o = pack_objects(s);
    static const bool always_full = getenv("SMK_FILL_FULL") != nullptr;
    if (g.nzc > 0) {
        smk::SrcBoxes b;
        const bool dirty = s->mask_dirty || always_full;
        if ((dirty && o.nobs > 0) || (o.nsrc > 0 && !source_boxes(s, o, b))) {
            smk::k_fill<<<row_grid(g.cplane, g.nzm), 256, 0, s->stream>>>(g, s->smoke[0], s->smoke[1], s->mask, o, g.mzlo);
            count_launch(s, SMK_STAGE_FILL);
        } else if (o.nsrc > 0 && b.n[0] > 0 && b.n[1] > 0 && b.n[2] > 0) {
            const dim3 grid((unsigned)((b.n[0] + 31) / 32), (unsigned)((b.n[1] + 7) / 8), (unsigned)(b.n[2] * o.nsrc));
            smk::k_fill_sources<<<grid, dim3(32, 8), 0, s->stream>>>(g, s->smoke[0], s->smoke[1], o, b);
            count_launch(s, SMK_STAGE_FILL);
        }
        if (dirty) launch_codes(s);
        s->mask_dirty = false;
    }
    CK(s, cudaGetLastError());
    return SMK_OK;
}

// node planes [za, zb)
int stage_force_clamp(smk_sim* s, float dt, int za, int zb)
{
    const GridP& g = s->g;
    Span sp(s, SMK_STAGE_FORCE);
    const int n = s->now;
    za = std::max(za, g.zlo);
    zb = std::min(zb, std::min(g.D, g.zlo + g.nzc));
    if (zb > za) {
        smk::k_force_clamp<<<row_grid(g.nplane, zb - za), 256, 0, s->stream>>>(g, s->u[n], s->v[n], s->w[n], s->smoke[n],
                                                                            s->code, dt, s->gravity, s->alpha, za);
        count_launch(s, SMK_STAGE_FORCE);
    }
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
int pick_zchunk(const smk_sim* s, int tiles_xy, int K, int nzn)
{
    int best_n = 1;
    double best = -1.0;
    for (int n = 1; n <= std::max(1, nzn / 4); n++) {
        const int zc = (nzn + n - 1) / n;
        const long ctas = (long)tiles_xy * ((nzn + zc - 1) / zc);
        const long waves = (ctas + s->num_sms - 1) / s->num_sms;
        const double eff = (double)ctas / (double)(waves * s->num_sms) * (double)zc / (double)(zc + 2 * K);
        if (eff > best + 1e-9) { best = eff; best_n = n; }
    }
    return (nzn + best_n - 1) / best_n;
}

// ... for the TMA-staged pass (one CTA per SM; pieces of the bottom tile row -- the floor, COMPLEX cells -- run the
// compact general variant, ~1.45x per z-step, and are handed out first): list-schedule every candidate chunk count on
// the SMs the way the hardware dispatches the grid (next CTA to the first SM that frees up) and keep the shortest span.
int pick_zchunk_tma(smk_sim* s, int tx, int ty, int K, int nzn)
{
    for (const auto& e : s->zchunk_cache) // (the search costs milliseconds of host time: once per shape, not per launch)
        if (e[0] == tx && e[1] == ty && e[2] == K && e[3] == nzn) return e[4];
    const int sms = std::max(1, s->num_sms);
    int best_n = 1;
    double best = 1e300;
    // chunk counts worth looking at: at least 2K planes per chunk, and no more pieces than ~64 per SM (beyond that the
    // lead-in planes dominate anyway) -- keeps the search at a few hundred thousand heap operations even at 1024^3
    const int n_max = std::max(1, std::min(nzn / (2 * K), (64 * sms) / std::max(1, tx * ty) + 1));
    for (int n = 1; n <= n_max; n++) {
        const int zc = (nzn + n - 1) / n;
        const int nch = (nzn + zc - 1) / zc;
        if (nch != n) continue;
        std::priority_queue<double, std::vector<double>, std::greater<double>> busy; // time at which each SM frees up
        for (int i = 0; i < sms; i++) busy.push(0.0);
        double span = 0.0;
        auto put = [&](double cost) { // the next CTA of the grid goes to the first SM that frees up
            const double t = busy.top() + cost;
            busy.pop();
            busy.push(t);
            span = std::max(span, t);
        };
        const double setup = 1.5; // prologue + pipeline fill of a piece, in z-steps
        for (int pass = 0; pass < 2; pass++)
            for (int c = 0; c < nch; c++) {
                const int planes = std::min(zc, nzn - c * zc);
                const double steps = planes + 2 * K + setup;
                if (pass == 0) for (int i = 0; i < tx; i++) put(1.45 * steps);
                else for (int i = 0; i < tx * (ty - 1); i++) put(steps);
            }
        if (span < best - 1e-9) { best = span; best_n = n; }
    }
    s->zchunk_cache.push_back({tx, ty, K, nzn, (nzn + best_n - 1) / best_n});
    return s->zchunk_cache.back()[4];
}

// The same simulation over chunk lists of UNEQUAL length (one GPU): nb long chunks of equal size first, one short chunk
// last -- longest-processing-time-first at the granularity the L2 argument allows (all tiles of a chunk still march
// together).  In the model this takes 9 % (256^3) / 5 % (512^3) off the span of the best equal split.  Returns the chunk
// lengths in dispatch order (empty: keep the equal split); cached per shape like pick_zchunk_tma.
std::vector<int> pick_chunks_tma(smk_sim* s, int tx, int ty, int K, int nzn)
{
    for (const auto& e : s->chunks_cache)
        if (e.first[0] == tx && e.first[1] == ty && e.first[2] == K && e.first[3] == nzn) return e.second;
    const int sms = std::max(1, s->num_sms);
    auto span_of = [&](const std::vector<int>& chunks) {
        std::priority_queue<double, std::vector<double>, std::greater<double>> busy;
        for (int i = 0; i < sms; i++) busy.push(0.0);
        double span = 0.0;
        auto put = [&](double cost) {
            const double t = busy.top() + cost;
            busy.pop();
            busy.push(t);
            span = std::max(span, t);
        };
        for (int pass = 0; pass < 2; pass++)
            for (int planes : chunks) {
                const double steps = planes + 2 * K + 1.5;
                if (pass == 0) for (int i = 0; i < tx; i++) put(1.45 * steps);
                else for (int i = 0; i < tx * (ty - 1); i++) put(steps);
            }
        return span;
    };
    std::vector<int> best_chunks;
    double best = 1e300;
    const int nb_max = std::max(1, std::min(std::min(nzn / (2 * K), 16), (64 * sms) / std::max(1, tx * ty) + 1));
    for (int nb = 1; nb <= nb_max; nb++) {
        const int a_max = (nzn + nb - 1) / nb;
        const int step = std::max(2, a_max / 32);
        for (int small = 0; small <= a_max; small = small == 0 ? 2 * K : small + step) {
            const int rest = nzn - small;
            if (rest < nb * K) break;
            std::vector<int> ch;
            const int A = (rest + nb - 1) / nb;
            for (int left = rest; left > 0; left -= A) ch.push_back(std::min(A, left));
            if ((int)ch.size() != nb || ch.back() < K) continue;
            if (small) ch.push_back(small);
            if ((int)ch.size() > 17) continue;
            const double sp = span_of(ch);
            if (sp < best - 1e-9) { best = sp; best_chunks = ch; }
        }
    }
    s->chunks_cache.push_back({{tx, ty, K, nzn}, best_chunks});
    return best_chunks;
}

int ensure_scratch(smk_sim*) { return SMK_OK; } // the scratch set is part of the arena

// cudaFuncSetAttribute applies to the CURRENT device only: remembered per handle, not per process
template <typename F>
int ensure_smem(smk_sim* s, F kern, size_t bytes)
{
    const void* key = reinterpret_cast<const void*>(kern);
    if (std::find(s->configured.begin(), s->configured.end(), key) != s->configured.end()) return SMK_OK;
    CK(s, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    s->configured.push_back(key);
    return SMK_OK;
}

void swap_in_scratch(smk_sim* s)
{
    const int n = s->now;
    std::swap(s->u[n], s->scratch[0]);
    std::swap(s->v[n], s->scratch[1]);
    std::swap(s->w[n], s->scratch[2]);
    std::swap(s->vel_id[n], s->scratch_id);
}

template <int K, int FUSED_NW, int FUSED_RPW>
int launch_fused_pass_cfg(smk_sim* s, int sweep0)
{
    using C = smk::FusedCfg<K, FUSED_NW, FUSED_RPW>;
    const GridP& g = s->g;
    auto kern = smk::k_pressure_fused<K, FUSED_NW, FUSED_RPW>;
    int rc = ensure_smem(s, kern, C::SMEM);
    if (rc) return rc;
    const int tx = (g.W + 1 + C::OX - 1) / C::OX, ty = (g.SY + C::OY - 1) / C::OY;
    const int zchunk = pick_zchunk(s, tx * ty, K, g.nzn);
    const dim3 grid((unsigned)tx, (unsigned)ty, (unsigned)((g.nzn + zchunk - 1) / zchunk));
    const int n = s->now;
    kern<<<grid, C::THREADS, C::SMEM, s->stream>>>(g, s->u[n], s->v[n], s->w[n], s->scratch[0], s->scratch[1], s->scratch[2],
                                                    s->code, sweep0, zchunk);
    swap_in_scratch(s);
    count_launch(s, SMK_STAGE_PRESSURE);
    return SMK_OK;
}

// register-resident fused pass (kernels_pressure_reg.cuh): u, w in registers, v in shared memory
// epoch handshake with the neighbours (peer-memory halo path): publish my epoch, wait for theirs
void peer_counters(smk_sim* s, unsigned* theirs[2], const unsigned* mine[2])
{
    theirs[0] = theirs[1] = nullptr;
    mine[0] = mine[1] = nullptr;
    for (int side = 0; side < 2; side++) {
        if (!s->peer[side].arena) continue;
        // counter [k] of an arena is written by the neighbour on side k: I am my lower neighbour's upper neighbour
        theirs[side] = reinterpret_cast<unsigned*>(s->peer[side].arena + s->peer[side].lay.counters) + (1 - side) * 32;
        mine[side] = reinterpret_cast<const unsigned*>(s->arena + s->lay.counters) + side * 32;
    }
}

int peer_signal(smk_sim* s, cudaStream_t st) // publish the next epoch: everything enqueued before is final
{
    unsigned* theirs[2]; const unsigned* mine[2];
    peer_counters(s, theirs, mine);
    s->epoch++;
    smk::k_epoch_signal<<<1, 1, 0, st>>>(theirs[0], theirs[1], s->epoch);
    s->launches++;
    CK(s, cudaGetLastError());
    return SMK_OK;
}

int peer_wait(smk_sim* s, cudaStream_t st, unsigned target) // wait until both neighbours published `target` (or later)
{
    unsigned* theirs[2]; const unsigned* mine[2];
    peer_counters(s, theirs, mine);
    smk::k_epoch_wait<<<1, 1, 0, st>>>(mine[0], mine[1], target, s->d_flags, (long long)2e10);
    s->launches++;
    CK(s, cudaGetLastError());
    return SMK_OK;
}

// epoch handshake with the neighbours (peer-memory halo path): publish my epoch, wait for theirs
int peer_sync(smk_sim* s)
{
    static const bool nosync = getenv("SMK_DBG_NOSYNC") != nullptr; // timing experiments only (races!)
    if (nosync) return SMK_OK;
    int rc = peer_signal(s, s->stream);
    if (rc) return rc;
    return peer_wait(s, s->stream, s->epoch);
}

// Epoch accounting of a pressure pass that reads neighbour planes.  EVERY branch of launch_reg_pass consumes exactly two
// epoch values per pass -- E_pre = "everything I enqueued before this pass is final" and E_post = "my boundary planes of
// this pass are written and I no longer read yours" -- whatever schedule the rank picked for its own plane count
// (in-kernel handshake, boundary chunks on a second stream, or a plain stream-level handshake), so neighbours whose
// slabs differ by a plane and therefore chunk differently still agree on every value (the branches used to
// advance the epoch by 2, 1 and 1).  Waits are ">= target", and a later value implies every earlier one.
// The stream-level branches publish E_pre and leave E_post to be implied by the next value they publish; they wait for
// E_pre in front of the first pass of a step (the neighbour's advection and fill must be covered) and for the previous
// pass's E_post otherwise (an in-kernel neighbour publishes nothing else between two passes).
unsigned stream_pass_epochs(smk_sim* s, int sweep0, cudaStream_t wait_stream, int* rc)
{
    *rc = peer_signal(s, s->stream);                 // E_pre
    const unsigned pre = s->epoch;
    s->epoch++;                                      // E_post: implied by whatever this rank publishes next
    s->pass_epoch_next = -1;
    if (*rc == SMK_OK) *rc = peer_wait(s, wait_stream, sweep0 == 0 ? pre : pre - 1);
    return pre;
}

// pointer to the (virtual) plane 0 of a field whose first stored plane is zlo; never dereferenced outside the stored planes
const float* plane0(const float* first_stored, long long plane, int zlo)
{
    return reinterpret_cast<const float*>(reinterpret_cast<uintptr_t>(first_stored) - (uintptr_t)((long long)zlo * plane * 4));
}

smk::PeerPlanes peer_planes(const smk_sim* s, int side)
{
    smk::PeerPlanes p{nullptr, nullptr, nullptr, 0, nullptr};
    const auto& pe = s->peer[side];
    if (!pe.arena) return p;
    const int id = s->vel_id[s->now]; // same physical buffer on every rank (identical swap history)
    p.zlo = pe.geom.zlo;
    // virtual plane-0 bases (kernels_pressure_reg.cuh): plane strides are the same on every slab (same W, H)
    p.u = plane0(reinterpret_cast<const float*>(pe.arena + pe.lay.u[id]), s->g.nplane, p.zlo);
    p.v = plane0(reinterpret_cast<const float*>(pe.arena + pe.lay.v[id]), s->g.nplane, p.zlo);
    p.w = plane0(reinterpret_cast<const float*>(pe.arena + pe.lay.w[id]), s->g.nplane, p.zlo);
    p.smoke = plane0(reinterpret_cast<const float*>(pe.arena + pe.lay.smoke[s->now]), s->g.cplane, p.zlo);
    return p;
}

// piece lists of a balanced pass (pass_schedule.h): computed on the host once per (tiles, plane range, K), kept on the
// device for the life of the simulation (a step alternates between at most a handful of ranges)
int get_schedule(smk_sim* s, int tx, int ty, int lo, int hi, int lead, int nctas, const smk_sim::DevSchedule** out,
                 int zl = INT_MIN, int zh = INT_MAX)
{
    for (const auto& d : s->schedules)
        if (d.tx == tx && d.ty == ty && d.lo == lo && d.hi == hi && d.lead == lead && d.nctas == nctas && d.zl == zl && d.zh == zh) {
            *out = &d;
            return SMK_OK;
        }
    sched::PassSchedule ps = sched::balance(tx, ty, lo, hi, lead, nctas);
    int nb[2] = {0, 0};
    if (zl != INT_MIN || zh != INT_MAX) sched::boundary_first(ps, zl, zh, nb);
    static_assert(sizeof(sched::Piece) == sizeof(int4), "pieces are uploaded as int4");
    smk_sim::DevSchedule d{tx, ty, lo, hi, lead, nctas, zl, zh, nullptr, nullptr, ps.nctas(), ps.cost, {nb[0], nb[1]}};
    CK(s, cudaMalloc(&d.pieces, std::max<size_t>(1, ps.pieces.size()) * sizeof(int4)));
    CK(s, cudaMalloc(&d.first, ps.first.size() * sizeof(int)));
    // Uploaded IN STREAM ORDER (first use only): a plain cudaMemcpy from pageable memory may return before its DMA has
    // landed, and the step's stream is not ordered behind the default stream.  cudaMemcpyAsync stages pageable
    // sources before it returns, so the host vectors may go out of scope.
    if (!ps.pieces.empty())
        CK(s, cudaMemcpyAsync(d.pieces, ps.pieces.data(), ps.pieces.size() * sizeof(int4), cudaMemcpyHostToDevice, s->stream));
    CK(s, cudaMemcpyAsync(d.first, ps.first.data(), ps.first.size() * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    s->schedules.push_back(d);
    *out = &s->schedules.back();
    return SMK_OK;
}

// register-resident fused pass (kernels_pressure_reg.cuh): u, w in registers, v in shared memory.
// [out_lo, out_hi) = node planes to write; with peers attached the planes outside the owned range are read from the
// neighbours' memory inside the kernel (after an epoch handshake), otherwise from the local ghost planes.
template <int K, int NW>
int launch_reg_pass(smk_sim* s, int sweep0, int out_lo, int out_hi, bool from_peers)
{
    using C = smk::RegCfg<K, NW>;
    const GridP& g = s->g;
    // SMK_PASS_KERNEL=reg | tma: process-wide choice of the pass kernel for same-box A/B runs -- the round-1 kernel
    // (kernels_pressure_reg.cuh) or the TMA-staged kernel (kernels_pressure_tma.cuh)
    static const char* kenv = getenv("SMK_PASS_KERNEL");
    // per handle (smk_set_pass_kernel), else the process default from the environment, else automatic
    const int kind = s->pass_kernel != SMK_PASS_AUTO ? s->pass_kernel
                     : (kenv && strcmp(kenv, "reg") == 0) ? SMK_PASS_REG : (kenv && strcmp(kenv, "tma") == 0) ? SMK_PASS_TMA : SMK_PASS_AUTO;
    bool use_tma = NW == 16 && s->pass_tma_ok && kind != SMK_PASS_REG; // the default where its tensor maps exist
    // small grids are launch- and pipeline-fill-bound, and there the round-1 kernel's shorter prologue wins (measured
    // crossover between 160^3 and 192^3: profiles/r2_tma_pass_final.txt); smk_set_pass_kernel / SMK_PASS_KERNEL=tma force it anyway
    if (use_tma && kind != SMK_PASS_TMA && (long long)g.P * g.SY * (out_hi - out_lo) < 5000000ll) use_tma = false;
    if (use_tma && s->pending_force && !s->pass_tma_smoke_ok) use_tma = false;
    if (use_tma && from_peers && ((s->peer[0].arena && !s->pass_tma_peer_ok[0]) || (s->peer[1].arena && !s->pass_tma_peer_ok[1]))) use_tma = false;
    {
        int rc0;
        if ((rc0 = ensure_smem(s, smk::k_pressure_reg<K, NW, false>, C::SMEM)) || (rc0 = ensure_smem(s, smk::k_pressure_reg<K, NW, true>, C::SMEM)) ||
            (rc0 = ensure_smem(s, smk::k_pressure_reg_bal<K, NW, false>, C::SMEM)) || (rc0 = ensure_smem(s, smk::k_pressure_reg_bal<K, NW, true>, C::SMEM)))
            return rc0;
    }
    // forcing + clamp deferred to this pass (exec_op / stage_pressure): the first pass of the step applies them on load
    const bool force = s->pending_force;
    s->pending_force = false;
    const smk::ForceArgs fa{s->smoke[s->now], s->pending_dt, s->gravity, s->alpha};
    const int n = s->now;
    auto launch_k = [&](dim3 grid, cudaStream_t st, int zchunk_, const smk::PassRange& r) {
        if constexpr (NW == 16) {
            if (use_tma) {
                smk::PassMaps m;
                const int id = s->vel_id[n];
                for (int src = 0; src < 3; src++) m.uvw[src] = s->pmap[src][id];
                m.pcode = s->pmap_code;
                for (int src = 0; src < 3; src++) m.smoke[src] = s->pmap_smoke[src][n];
                static const bool nocf = getenv("SMK_PASS_NO_CFLAG") != nullptr; // ablation: always the general variant
                const unsigned char* cf = (K == 4 && !nocf && (int)grid.x == s->cflag_tx && (int)grid.y == s->cflag_ty) ? s->cflag : nullptr;
                using T0 = smk::TmaCfg<K, 16, false>;
                using T1 = smk::TmaCfg<K, 16, true>;
                if (force) {
                    auto k = smk::k_pressure_tma<K, 16, true>;
                    if (ensure_smem(s, k, T1::SMEM_END)) return;
                    k<<<grid, T1::THREADS, T1::SMEM_END, st>>>(g, m, s->scratch[0], s->scratch[1], s->scratch[2], sweep0, zchunk_, r, fa, s->d_flags, cf, s->d_passdbg);
                } else {
                    auto k = smk::k_pressure_tma<K, 16, false>;
                    if (ensure_smem(s, k, T0::SMEM_END)) return;
                    k<<<grid, T0::THREADS, T0::SMEM_END, st>>>(g, m, s->scratch[0], s->scratch[1], s->scratch[2], sweep0, zchunk_, r, fa, s->d_flags, cf, s->d_passdbg);
                }
                s->passdbg_ctas = (int)(grid.x * grid.y * grid.z);
                s->last_pass_kernel = SMK_PASS_TMA;
                return;
            }
        }
        s->last_pass_kernel = SMK_PASS_REG;
        auto k = force ? smk::k_pressure_reg<K, NW, true> : smk::k_pressure_reg<K, NW, false>;
        k<<<grid, C::THREADS, C::SMEM, st>>>(g, s->u[n], s->v[n], s->w[n], s->scratch[0], s->scratch[1], s->scratch[2], s->code, sweep0, zchunk_, r, fa);
    };
    smk::PassRange pr{};
    pr.out_lo = out_lo; pr.out_hi = out_hi;
    pr.own_lo = g.zlo; pr.own_hi = g.zlo + g.nzn - 1; // default: everything stored counts as "own" (local source)
    pr.chunk_first = 0; pr.chunk_step = 1;
    pr.local = smk::PeerPlanes{plane0(s->u[s->now], g.nplane, g.zlo), plane0(s->v[s->now], g.nplane, g.zlo),
                               plane0(s->w[s->now], g.nplane, g.zlo), g.zlo, plane0(s->smoke[s->now], g.cplane, g.zlo)};
    const int nz = out_hi - out_lo;
    const int tx = (g.W + 1 + C::OX - 1) / C::OX, ty = (g.SY + C::OY - 1) / C::OY;
    int zchunk = use_tma ? pick_zchunk_tma(s, tx, ty, K, nz) : pick_zchunk(s, tx * ty, K, nz);
    static const int force_chunks = getenv("SMK_PASS_NCHUNKS") ? atoi(getenv("SMK_PASS_NCHUNKS")) : 0; // experiments
    if (force_chunks > 0) zchunk = std::max(K, (nz + force_chunks - 1) / force_chunks);
    // (kept from a precursor kernel that addressed a chunk with 32-bit byte offsets; harmless: it only splits huge chunks)
    while ((long long)(zchunk + 2 * K + 2) * g.nplane * 4 >= (1ll << 32) && zchunk > 2 * K) zchunk = (zchunk + 1) / 2;
    // Only the first and the last z-chunk may touch planes within K of the slab ends (the neighbours read those / they
    // read the neighbours'): the LAST chunk must therefore hold at least K planes too (the others hold zchunk >= K)
    while (from_peers && nz > zchunk && zchunk >= K && nz - (nz - 1) / zchunk * zchunk < K) zchunk++;
    const int nchunks = (nz + zchunk - 1) / zchunk;
    static const bool overlap = getenv("SMK_P2P_NO_OVERLAP") == nullptr;
    if (from_peers) {
        pr.own_lo = s->geom.own_node_lo(); pr.own_hi = s->geom.own_node_hi();
        pr.lower = peer_planes(s, 0); pr.upper = peer_planes(s, 1);
    }
    static const bool inkernel = getenv("SMK_P2P_STREAM_SYNC") == nullptr;
    if (from_peers && overlap && inkernel && nchunks >= 2 && zchunk >= K) {
        // One launch per pass, handshake inside the kernel (PassSync): the boundary chunks wait for the neighbour's
        // epoch themselves; the last boundary CTA per side publishes mine.  Only the first pass of a
        // step needs a signal from the stream: it must cover the advection and the fill in front of it.
        unsigned* theirs[2]; const unsigned* mine[2];
        peer_counters(s, theirs, mine);
        // Every pass owns two epoch values: one for a stream-level signal in front of it (published by the first pass
        // of a step; for the later passes the neighbour's previous pass kernel has already said more) and one that the
        // kernel publishes when its boundary chunks are done.
        if (sweep0 == 0 || s->pass_epoch_next != sweep0) {
            int rc = peer_signal(s, s->stream);
            if (rc) return rc;
            pr.sync.wait_epoch = s->epoch;
        } else {
            pr.sync.wait_epoch = s->epoch; // the previous pass kernel's own epoch
            s->epoch++;                    // this pass's stream-level slot stays unused
        }
        unsigned* local = reinterpret_cast<unsigned*>(s->arena + s->lay.counters);
        pr.sync.wait_ctr[0] = mine[0]; pr.sync.wait_ctr[1] = mine[1];
        pr.sync.sig_ctr[0] = theirs[0]; pr.sync.sig_ctr[1] = theirs[1];
        pr.sync.done_ctr[0] = local + 64; pr.sync.done_ctr[1] = local + 96;
        pr.sync.sig_epoch = ++s->epoch;
        pr.sync.flags = s->d_flags;
        pr.sync.nchunks = nchunks;
        // boundary chunks FIRST: their neighbour reads pay an NVLink round trip per z-step, which then hides behind the
        // other CTAs instead of forming the tail of the pass, and their epoch is published early in the pass, so the
        // neighbour's next pass finds it waiting (measured at N=2: 3.07 ms of passes per step, against 3.18 with the
        // boundary chunks last and 2.92 on a single GPU)
        static const bool blast = getenv("SMK_P2P_BOUNDARY_LAST") != nullptr, nowait = getenv("SMK_DBG_NOWAIT") != nullptr;
        pr.sync.first = blast ? 0 : 1;
        if (nowait) pr.sync.wait_epoch = 0; // timing experiments only (races!)
        s->pass_epoch_next = sweep0 + K;
        // opt-in (smk_set_pass_ctas): balanced piece lists, boundary pieces first in every CTA
        const smk_sim::DevSchedule* ds = nullptr;
        if (s->pass_ctas > 0) {
            const int zl = pr.lower.u ? pr.own_lo + K : INT_MIN, zh = pr.upper.u ? pr.own_hi - K + 1 : INT_MAX;
            int rc = get_schedule(s, tx, ty, out_lo, out_hi, sched::pass_lead(K), s->pass_ctas, &ds, zl, zh);
            if (rc) return rc;
        }
        s->last_pass_ctas = ds ? ds->launch_ctas : 0;
        if (ds) {
            pr.sync.npieces[0] = ds->nboundary[0]; pr.sync.npieces[1] = ds->nboundary[1];
            if (!pr.lower.u) pr.sync.wait_ctr[0] = nullptr; // (no neighbour on that side: nothing to wait for or to publish)
            if (!pr.upper.u) pr.sync.wait_ctr[1] = nullptr;
            auto kb = force ? smk::k_pressure_reg_bal<K, NW, true> : smk::k_pressure_reg_bal<K, NW, false>;
            kb<<<dim3((unsigned)ds->launch_ctas), C::THREADS, C::SMEM, s->stream>>>(
                g, s->u[n], s->v[n], s->w[n], s->scratch[0], s->scratch[1], s->scratch[2], s->code, sweep0, pr, fa, ds->pieces,
                ds->first);
        } else {
            launch_k(dim3((unsigned)tx, (unsigned)ty, (unsigned)nchunks), s->stream, zchunk, pr);
        }
    } else if (from_peers && overlap && nchunks >= 3 && zchunk >= K) {
        // The first and last z-chunk read neighbour planes; the interior chunks do not and never write planes a
        // neighbour may still be reading.  So: publish my epoch, start the interior chunks at once on the main stream, and
        // run the two boundary chunks on a second stream behind the epoch wait -- the handshake latency and the skew
        // between the GPUs hide behind the interior work.
        if (!s->aux_stream) {
            CK(s, cudaStreamCreateWithFlags(&s->aux_stream, cudaStreamNonBlocking));
            CK(s, cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
            CK(s, cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
        }
        // (the signal is enqueued on the main stream in front of the fork; the wait goes to the second stream)
        int rc = peer_signal(s, s->stream);
        if (rc) return rc;
        const unsigned pre = s->epoch;
        s->epoch++; // E_post (see stream_pass_epochs)
        s->pass_epoch_next = -1;
        CK(s, cudaEventRecord(s->ev_fork, s->stream));
        CK(s, cudaStreamWaitEvent(s->aux_stream, s->ev_fork, 0));
        if ((rc = peer_wait(s, s->aux_stream, sweep0 == 0 ? pre : pre - 1))) return rc;
        smk::PassRange pb = pr;
        pb.chunk_first = 0; pb.chunk_step = nchunks - 1;
        launch_k(dim3((unsigned)tx, (unsigned)ty, 2u), s->aux_stream, zchunk, pb);
        CK(s, cudaEventRecord(s->ev_join, s->aux_stream));
        smk::PassRange pi = pr;
        pi.chunk_first = 1; pi.chunk_step = 1;
        launch_k(dim3((unsigned)tx, (unsigned)ty, (unsigned)(nchunks - 2)), s->stream, zchunk, pi);
        CK(s, cudaStreamWaitEvent(s->stream, s->ev_join, 0));
        s->launches++;
    } else {
        if (from_peers) {
            static const bool nosync = getenv("SMK_DBG_NOSYNC") != nullptr; // timing experiments only (races!)
            int rc = SMK_OK;
            if (!nosync) stream_pass_epochs(s, sweep0, s->stream, &rc);
            if (rc) return rc;
        }
        // Default: the (tile, z-chunk) grid.  Opt-in (smk_set_pass_ctas / SMK_PASS_CTAS=n): n CTAs working through piece
        // lists of equal cost (pass_schedule.h).  Measured at 256^3 / 512^3 with one CTA per SM: 104 instead of 116 z-steps
        // on the busiest SM, but neighbouring tiles are no longer at the same z at the same time, their halo re-reads
        // miss L2 (hit rate 4 % instead of 31 %, DRAM reads 506 MB instead of 250 MB per pass) and the pass is no
        // faster (256^3) or 15 % slower (512^3): profiles/r1_balanced_schedule.txt.
        const smk_sim::DevSchedule* ds = nullptr;
        if (s->pass_ctas > 0 && nz > 0) {
            int rc = get_schedule(s, tx, ty, out_lo, out_hi, sched::pass_lead(K), s->pass_ctas, &ds);
            if (rc) return rc;
        }
        s->last_pass_ctas = ds ? ds->launch_ctas : 0;
        if (ds) {
            auto kb = force ? smk::k_pressure_reg_bal<K, NW, true> : smk::k_pressure_reg_bal<K, NW, false>;
            kb<<<dim3((unsigned)ds->launch_ctas), C::THREADS, C::SMEM, s->stream>>>(
                g, s->u[n], s->v[n], s->w[n], s->scratch[0], s->scratch[1], s->scratch[2], s->code, sweep0, pr, fa, ds->pieces,
                ds->first);
        } else {
            static const bool equal_chunks = getenv("SMK_PASS_EQUAL_CHUNKS") != nullptr; // A/B switch
            unsigned gz = (unsigned)nchunks;
            if (use_tma && !from_peers && !equal_chunks && force_chunks <= 0) {
                const std::vector<int> ch = pick_chunks_tma(s, tx, ty, K, nz);
                if (ch.size() >= 2) {
                    pr.nzcut = (int)ch.size();
                    pr.zcut[0] = out_lo;
                    for (size_t i = 0; i < ch.size(); i++) pr.zcut[i + 1] = pr.zcut[i] + ch[i];
                    gz = (unsigned)ch.size();
                }
            }
            launch_k(dim3((unsigned)tx, (unsigned)ty, gz), s->stream, zchunk, pr);
        }
    }
    swap_in_scratch(s);
    count_launch(s, SMK_STAGE_PRESSURE);
    return SMK_OK;
}

bool peer_passes(const smk_sim* s) // pressure passes may read the neighbours directly (K = 4 register kernel only)
{
    static const int cfg = getenv("SMK_FUSED_CFG") ? atoi(getenv("SMK_FUSED_CFG")) : 0;
    return s->p2p && s->geom.world > 1 && cfg == 0;
}

int effective_fuse(const smk_sim* s);

// Forcing + clamp can ride on the first pressure pass when that pass is the K = 4 register kernel and every plane it
// reads is pre-forcing data: a single GPU, or slabs whose passes read the neighbours directly.  (With the callback
// transport the ghost planes are exchanged AFTER forcing, so forcing stays a stage of its own there.)
bool can_fuse_force(const smk_sim* s)
{
    static const bool off = getenv("SMK_NO_FUSED_FORCE") != nullptr;
    static const int cfg = getenv("SMK_FUSED_CFG") ? atoi(getenv("SMK_FUSED_CFG")) : 0;
    return !off && cfg == 0 && s->solver == SMK_SOLVER_RBGS && (effective_fuse(s) == 4 || (effective_fuse(s) == 2 && s->geom.world == 1)) && s->iterations >= 2 &&
           (s->g.W & 3) == 0 && (s->geom.world == 1 || peer_passes(s));
}

template <int K>
int launch_fused_pass(smk_sim* s, int sweep0, int out_lo, int out_hi, bool from_peers)
{
    // SMK_FUSED_CFG selects the older shared-memory kernel (warps x rows per warp) for experiments
    static const int cfg = getenv("SMK_FUSED_CFG") ? atoi(getenv("SMK_FUSED_CFG")) : 0;
    if (K == 4 && cfg == 0) return launch_reg_pass<4, 16>(s, sweep0, out_lo, out_hi, from_peers);
    if (K == 2 && cfg == 0) return launch_reg_pass<2, 16>(s, sweep0, out_lo, out_hi, false);
    if (K == 2 && cfg == 20) return launch_reg_pass<2, 20>(s, sweep0, out_lo, out_hi, false);
    if (K == 4 && cfg == 24) return launch_reg_pass<4, 24>(s, sweep0, out_lo, out_hi, from_peers);
    if (K == 4 && cfg == 20) return launch_reg_pass<4, 20>(s, sweep0, out_lo, out_hi, from_peers);
    if (K == 4 && cfg == 12) return launch_reg_pass<4, 12>(s, sweep0, out_lo, out_hi, from_peers);
    switch (cfg) {
    case 163: return launch_fused_pass_cfg<K, 16, 3>(s, sweep0);
    case 124: return launch_fused_pass_cfg<K, 12, 4>(s, sweep0);
    case 321: return launch_fused_pass_cfg<K, 32, 1>(s, sweep0);
    case 162: return launch_fused_pass_cfg<K, 16, 2>(s, sweep0);
    default: return launch_fused_pass_cfg<K, 24, 2>(s, sweep0);
    }
}

int effective_fuse(const smk_sim* s)
{
    if (s->fuse != 0) return s->fuse;
    return (s->g.W + 1 < 32 || s->g.nzn < 8) ? 1 : 4; // tiny grids: tiles would be mostly halo
}

// half-sweeps [sweep0, sweep0 + K) on every stored plane (offsets alternate 0,1: cu:797-801)
int run_pressure_pass(smk_sim* s, int sweep0, int K, int out_lo = INT32_MIN, int out_hi = INT32_MAX, bool from_peers = false)
{
    out_lo = std::max(out_lo, s->g.zlo);
    out_hi = std::min(out_hi, s->g.zlo + s->g.nzn);
    if (K == 4) return launch_fused_pass<4>(s, sweep0, out_lo, out_hi, from_peers);
    if (K == 2) return launch_fused_pass<2>(s, sweep0, out_lo, out_hi, false);
    return launch_halfsweep(s, sweep0 & 1);
}

// extension: `iterations` damped-Jacobi iterations (kernels_jacobi.cuh), out of place through the scratch set
int stage_pressure_jacobi(smk_sim* s)
{
    const GridP& g = s->g;
    const int n = s->now;
    const int xtiles = (g.W + smk::JTX - 1) / smk::JTX, ytiles = (g.SY + smk::JTY - 1) / smk::JTY;
    // z chunks: the prologue of a chunk (p of the plane below, first loads: about three plane-times per chunk)
    // is redundant work; pick the split with the best (wave quantisation x useful planes) product at 2 CTAs per SM
    int nchunk = 1;
    {
        const long slots = 2L * s->num_sms, tiles = (long)xtiles * ytiles;
        double best = -1.0;
        for (int nc = 1; nc <= std::max(1, g.nzn / 8); nc++) {
            const int zc = (g.nzn + nc - 1) / nc;
            const long ctas = tiles * ((g.nzn + zc - 1) / zc), waves = (ctas + slots - 1) / slots;
            const double eff = (double)ctas / (double)(waves * slots) * (double)zc / (double)(zc + 3);
            if (eff > best + 1e-9) { best = eff; nchunk = nc; }
        }
    }
    const int zchunk = (g.nzn + nchunk - 1) / nchunk;
    nchunk = (g.nzn + zchunk - 1) / zchunk;
    dim3 block(smk::JTHREADS), grid(xtiles, ytiles, nchunk);
    // opt-in (smk_set_pass_ctas): balanced piece lists (pass_schedule.h); measured slower than the grid for the same
    // reason as the RBGS passes (neighbouring tiles out of step -> their shared rows miss L2)
    const smk_sim::DevSchedule* ds = nullptr;
    if (s->pass_ctas > 0) {
        int rc = get_schedule(s, xtiles, ytiles, 0, g.nzn, sched::JACOBI_LEAD, s->pass_ctas, &ds);
        if (rc) return rc;
    }
    s->last_pass_ctas = ds ? ds->launch_ctas : 0;
    for (int it = 0; it < s->iterations; it++) {
        if (ds)
            smk::k_jacobi_bal<<<dim3((unsigned)ds->launch_ctas), block, 0, s->stream>>>(
                g, s->u[n], s->v[n], s->w[n], s->scratch[0], s->scratch[1], s->scratch[2], s->code, xtiles, ds->pieces, ds->first);
        else
            smk::k_jacobi<<<grid, block, 0, s->stream>>>(g, s->u[n], s->v[n], s->w[n], s->scratch[0], s->scratch[1],
                                                        s->scratch[2], s->code, zchunk, xtiles);
        count_launch(s, SMK_STAGE_PRESSURE);
        swap_in_scratch(s);
    }
    CK(s, cudaGetLastError());
    return SMK_OK;
}

int stage_pressure(smk_sim* s)
{
    Span sp(s, SMK_STAGE_PRESSURE);
    if (s->solver == SMK_SOLVER_JACOBI) return stage_pressure_jacobi(s);
    const int total = 2 * s->iterations;
    const int fuse = effective_fuse(s);
    int done = 0, rc = SMK_OK;
    while (done < total && rc == SMK_OK) {
        const int left = total - done;
        const int K = (fuse >= 4 && left >= 4 && (done & 1) == 0) ? 4 : (fuse >= 2 && left >= 2 && (done & 1) == 0) ? 2 : 1;
        rc = run_pressure_pass(s, done, K);
        done += K;
    }
    if (rc) return rc;
    CK(s, cudaGetLastError());
    return SMK_OK;
}

// node planes [za, zb); [vlo, vhi] = planes of the "now" velocities that hold valid data (reach guard)
int stage_advect_velocity(smk_sim* s, float dt, int za, int zb, int vlo, int vhi, int dyn_mode = 0)
{
    const GridP& g = s->g;
    Span sp(s, SMK_STAGE_ADVECT_VEL);
    const int n = s->now, p = s->past;
    // dynamic plane range (adaptive margin, kernels_basic.cuh): sides without a neighbour never shrink
    const smk::DynRange dr{s->d_dyn, dyn_mode, s->geom.has_lower() ? s->geom.own_node_lo() : -(1 << 28),
                           s->geom.has_upper() ? s->geom.own_node_hi() : (1 << 28)};
    za = std::max(za, std::max(1, g.zlo));
    zb = std::min(zb, std::min(g.D, g.zlo + g.nzc));
    if (zb > za) {
        static const int use_tma = getenv("SMK_ADVECT_TMA") ? atoi(getenv("SMK_ADVECT_TMA")) : 1;
        if (use_tma && s->tma_ok) {
            // TMA-staged tiles (kernels_advect_tma.cuh): tile + halo planes land in shared memory, gathers are LDS
            using A = smk::AdvTma;
            if (int rc0 = ensure_smem(s, smk::k_advect_velocity_tma, A::SMEM)) return rc0;
            const int id = s->vel_id[n];
            const int tiles = ((g.W + A::TX - 1) / A::TX) * ((g.H + A::TY - 1) / A::TY);
            int zchunk = zb - za; // enough CTAs for ~8 per SM, chunks of at least 16 planes (5 lead-in planes each)
            while (zchunk > 16 && (long)tiles * ((zb - za + zchunk - 1) / zchunk) < 8L * s->num_sms) zchunk = (zchunk + 1) / 2;
            const dim3 grd((g.W + A::TX - 1) / A::TX, (g.H + A::TY - 1) / A::TY, (zb - za + zchunk - 1) / zchunk);
            smk::k_advect_velocity_tma<<<grd, A::THREADS, A::SMEM, s->stream>>>(
                g, s->tmap[0][id], s->tmap[1][id], s->tmap[2][id], s->u[n], s->v[n], s->w[n], s->u[p], s->v[p], s->w[p], s->code, dt,
                za, zb, zchunk, make_int2(vlo, vhi), s->d_flags, dr);
        } else {
            const int by = 4, bz = 2; // block shape is immaterial (issue-bound; measured 32x4x2 .. 32x8x1 within 1 %)
            const dim3 blk(32, by, bz), grd((g.W + 31) / 32, (g.H + by - 1) / by, (zb - za + bz - 1) / bz);
            smk::k_advect_velocity<<<grd, blk, 0, s->stream>>>(
                g, s->u[n], s->v[n], s->w[n], s->u[p], s->v[p], s->w[p], s->code, dt, za, zb, make_int2(vlo, vhi), s->d_flags, dr);
        }
        count_launch(s, SMK_STAGE_ADVECT_VEL);
    }
    CK(s, cudaGetLastError());
    return SMK_OK;
}

// cell planes [za, zb); [vlo, vhi] = cell planes of the "now" density that hold valid data
int stage_advect_smoke(smk_sim* s, float dt, int za, int zb, int vlo, int vhi)
{
    const GridP& g = s->g;
    Span sp(s, SMK_STAGE_ADVECT_SMOKE);
    const int n = s->now, p = s->past;
    // (surface variant: the requested planes widened to the boundary planes 0 and D-1 when the request touches plane 1 / D-2)
    const int za0 = za <= 1 ? 0 : za, zb0 = zb >= g.D - 1 ? g.D : zb;
    za = std::max(za, std::max(1, g.zlo));
    zb = std::min(zb, std::min(g.D - 1, g.zlo + g.nzc));
    if (zb > za) {
        const int by = 4, bz = 2;
        const dim3 blk(32, by, bz), grd((g.W + 31) / 32, (g.H + by - 1) / by, (zb - za + bz - 1) / bz);
        // 32-bit element indices whenever every stored field has fewer than 2^31 elements (SMK_ADVECT_IDX64=1: never)
        static const bool idx64 = getenv("SMK_ADVECT_IDX64") != nullptr;
        const bool small = !idx64 && (size_t)g.nplane * g.nzn < ((size_t)1 << 31) && (size_t)g.kplane * g.nzc < ((size_t)1 << 31);
        auto kern = small ? smk::k_advect_smoke32 : smk::k_advect_smoke;
        if (!s->density_surf)
            kern<<<grd, blk, 0, s->stream>>>(
                g, s->smoke[n], s->smoke[p], s->u[p], s->v[p], s->w[p], s->code, dt, za, zb, make_int2(vlo, vhi), s->d_flags);
        count_launch(s, SMK_STAGE_ADVECT_SMOKE);
    }
    if (s->density_surf) { // N1: every owned cell plane goes to the bound surface too (incl. the planes advection never writes)
        const int sa = std::max(za0, s->geom.c0), sb = std::min(zb0, s->geom.c1);
        if (sb > sa) {
            const int by = 4, bz = 2;
            const dim3 blk(32, by, bz), grd((g.W + 31) / 32, (g.H + by - 1) / by, (sb - sa + bz - 1) / bz);
            smk::k_advect_smoke_surf<<<grd, blk, 0, s->stream>>>(g, s->smoke[n], s->smoke[p], s->u[p], s->v[p], s->w[p], s->code, dt, sa, sb,
                                                               make_int2(vlo, vhi), s->d_flags, s->density_surf);
            if (zb <= za) count_launch(s, SMK_STAGE_ADVECT_SMOKE);
        }
    }
    CK(s, cudaGetLastError());
    return SMK_OK;
}

void flip(smk_sim* s) // cu:777-779
{
    s->past = s->now;
    s->now = s->now == 0 ? 1 : 0;
}

// Page-lock [p, p + bytes) so that the device->host copy of the density runs at PCIe speed and asynchronously.
// The library cannot know the lifetime of a caller's buffer: an entry is trusted only if it covers the
// requested range AND the driver still reports the range as registered host memory; a stale entry (the caller freed
// the buffer, or reuses the address with another size) is dropped and the range registered afresh.  Implicit
// registrations are released by smk_destroy or by smk_unregister_host; the contract is in smoke_b200.h.
bool range_is_pinned(const void* p)
{
    cudaPointerAttributes at{};
    const bool ok = cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    return ok;
}

void drop_registration(smk_sim* s, size_t i)
{
    if (range_is_pinned(s->registered[i].p)) cudaHostUnregister(s->registered[i].p);
    cudaGetLastError();
    s->registered.erase(s->registered.begin() + (long)i);
}

bool try_register(smk_sim* s, void* vp, size_t bytes, bool implicit = true)
{
    char* p = static_cast<char*>(vp);
    if (!p || bytes == 0) return false;
    for (size_t i = 0; i < s->registered.size();) {
        auto& r = s->registered[i];
        const bool overlaps = p < r.p + r.bytes && r.p < p + bytes;
        if (!overlaps) { i++; continue; }
        const bool covers = r.p <= p && p + bytes <= r.p + r.bytes;
        if (covers && range_is_pinned(p) && range_is_pinned(p + bytes - 1)) return true;
        drop_registration(s, i); // stale or partial: register the requested range afresh
    }
    static const bool off = getenv("SMK_NO_HOST_REGISTER") != nullptr;
    if (off && implicit) return false;
    if (range_is_pinned(p) && range_is_pinned(p + bytes - 1)) return true; // pinned by the caller (cudaHostAlloc / torch pin_memory)
    if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    s->registered.push_back({p, bytes, implicit});
    return true;
}

// halo exchange of one field set through the caller's transport (smk_set_exchange)
// native halo exchange over peer-mapped memory: handshake, then PULL the neighbours' owned boundary planes into my
// ghost planes (the neighbour's send region == my receive region, same global planes)
int p2p_pull(smk_sim* s, int set, cudaStream_t st) // the copies only; the caller has done the handshake
{
    const GridP& g = s->g;
    for (const slab::Region& r : slab::regions(s->geom, set)) {
        const auto& pe = s->peer[r.side];
        if (!pe.arena || r.recv_n <= 0) continue;
        const size_t plane = set == slab::SET_VEL_NOW ? (size_t)g.nplane : (size_t)g.cplane;
        const size_t n16 = (size_t)r.recv_n * plane * sizeof(float) / 16;
        const int nf = set == slab::SET_VEL_NOW ? 3 : 1;
        for (int i = 0; i < nf; i++) {
            float* dst; const float* src;
            if (set == slab::SET_VEL_NOW) {
                float* mine[3] = {s->u[s->now], s->v[s->now], s->w[s->now]};
                const int id = s->vel_id[s->now];
                const size_t off[3] = {pe.lay.u[id], pe.lay.v[id], pe.lay.w[id]};
                dst = mine[i] + (size_t)(r.recv_lo - g.zlo) * plane;
                src = reinterpret_cast<const float*>(pe.arena + off[i]) + (size_t)(r.recv_lo - pe.geom.zlo) * plane;
            } else {
                dst = s->smoke[s->now] + (size_t)(r.recv_lo - g.zlo) * plane;
                src = reinterpret_cast<const float*>(pe.arena + pe.lay.smoke[s->now]) + (size_t)(r.recv_lo - pe.geom.zlo) * plane;
            }
            const unsigned blocks = (unsigned)std::min<size_t>((n16 + 255) / 256, 4 * (size_t)s->num_sms);
            smk::k_copy16<<<blocks, 256, 0, st>>>(reinterpret_cast<float4*>(dst), reinterpret_cast<const float4*>(src), n16);
            s->launches++;
        }
    }
    s->exchanges++;
    CK(s, cudaGetLastError());
    return SMK_OK;
}

int run_exchange_p2p(smk_sim* s, int set)
{
    int rc = peer_sync(s);
    if (rc) return rc;
    return p2p_pull(s, set, s->stream);
}

int run_exchange(smk_sim* s, int set)
{
    if (s->p2p) return run_exchange_p2p(s, set);
    if (!s->exchange) return fail(s, SMK_ERR_TRANSPORT, "slab step needs a halo transport: call smk_set_exchange() or attach the peers");
    const GridP& g = s->g;
    std::vector<smk_halo_region> out;
    for (const slab::Region& r : slab::regions(s->geom, set)) {
        if (set == slab::SET_VEL_NOW) {
            float* f[3] = {s->u[s->now], s->v[s->now], s->w[s->now]};
            for (int i = 0; i < 3; i++)
                out.push_back({r.side, f[i] + (size_t)(r.send_lo - g.zlo) * g.nplane, f[i] + (size_t)(r.recv_lo - g.zlo) * g.nplane,
                               (size_t)r.send_n * g.nplane * sizeof(float), (size_t)r.recv_n * g.nplane * sizeof(float)});
        } else {
            float* f = s->smoke[s->now];
            out.push_back({r.side, f + (size_t)(r.send_lo - g.zlo) * g.cplane, f + (size_t)(r.recv_lo - g.zlo) * g.cplane,
                           (size_t)r.send_n * g.cplane * sizeof(float), (size_t)r.recv_n * g.cplane * sizeof(float)});
        }
    }
    s->exchanges++;
    if (s->exchange(s->exchange_ctx, set, out.data(), (int)out.size(), (void*)s->stream) != 0)
        return fail(s, SMK_ERR_TRANSPORT, "halo transport callback failed");
    return SMK_OK;
}

int exec_op(smk_sim* s, const slab::Op& op, float dt)
{
    switch (op.kind) {
    case slab::OP_FLIP: flip(s); return SMK_OK;
    case slab::OP_FILL: return stage_fill(s);
    case slab::OP_FORCE:
        if (can_fuse_force(s)) { // applied by the first pressure pass on load (kernels_pressure_reg.cuh, ForceArgs)
            s->pending_force = true; s->pending_dt = dt; s->pending_a = op.a; s->pending_b = op.b;
            return SMK_OK;
        }
        return stage_force_clamp(s, dt, op.a, op.b);
    case slab::OP_PRESSURE: {
        Span sp(s, SMK_STAGE_PRESSURE);
        // the Jacobi extension (single GPU, so the plan holds no exchange) runs whole at the plan's first pressure op
        if (s->pending_force && !((op.p1 == 4 || op.p1 == 2) && s->solver == SMK_SOLVER_RBGS)) { // not the pass that can apply it
            s->pending_force = false;
            int rc = stage_force_clamp(s, s->pending_dt, s->pending_a, s->pending_b);
            if (rc) return rc;
        }
        if (s->solver == SMK_SOLVER_JACOBI) return op.p0 == 0 ? stage_pressure_jacobi(s) : SMK_OK;
        const bool from_peers = peer_passes(s) && op.p1 == 4;
        int rc = from_peers ? run_pressure_pass(s, op.p0, op.p1, op.a, op.b, true) : run_pressure_pass(s, op.p0, op.p1);
        if (rc == SMK_OK) CK(s, cudaGetLastError());
        return rc;
    }
    case slab::OP_ADVECT_VEL: return stage_advect_velocity(s, dt, op.a, op.b, op.p0, op.p1);
    case slab::OP_ADVECT_SMOKE: return stage_advect_smoke(s, dt, op.a, op.b, op.p0, op.p1);
    case slab::OP_EXCHANGE: return run_exchange(s, op.a);
    }
    return fail(s, SMK_ERR_ARG, "unknown op kind");
}

// one step = the plan of slab_plan.h executed with CUDA kernels (a single GPU is the 1-slab case: no exchanges)
// Peer-memory transport: the halo pull in front of the velocity advection hides behind the advection of the planes
// that need no ghost data.  Publish my epoch; a second stream waits for the neighbours' epoch and pulls their boundary
// planes (u, v, w and -- if the plan asks for it later in the step -- the density, which has been final since the
// fill) into my ghost planes; the main stream advects the interior planes meanwhile, then joins and advects the
// MARGIN planes at each slab end.
bool overlap_exchange_ok(const smk_sim* s, const slab::Op& adv)
{
    static const bool off = getenv("SMK_P2P_NO_OVERLAP") != nullptr;
    // (rank-invariant: every rank must take the same decision, the handshakes of the two variants differ)
    return s->p2p && !off && s->geom.world > 1 && s->geom.D / s->geom.world - 1 >= 8 + 2 * slab::MARGIN && adv.b - adv.a >= 8 + 2 * slab::MARGIN;
}

// max |w| over the owned planes of the "now" velocities -> advection margin in d_dyn[1] (kernels_basic.cuh)
int compute_margin(smk_sim* s, float dt)
{
    const GridP& g = s->g;
    const int lo = s->geom.own_node_lo(), hi = s->geom.own_node_hi();
    const size_t n4 = (size_t)(hi - lo + 1) * g.nplane / 4;
    const float4* w4 = reinterpret_cast<const float4*>(s->w[s->now] + (size_t)(lo - g.zlo) * g.nplane);
    smk::k_absmax_w<<<(unsigned)std::min<size_t>((n4 + 255) / 256, (size_t)8 * s->num_sms), 256, 0, s->stream>>>(w4, n4, s->d_dyn);
    smk::k_margin<<<1, 1, 0, s->stream>>>(s->d_dyn, dt, s->geom.ghost);
    s->launches += 2;
    CK(s, cudaGetLastError());
    return SMK_OK;
}

int exchange_overlapped_with_advect(smk_sim* s, const slab::Op& adv, float dt, bool with_smoke)
{
    int rc;
    if (!s->aux_stream) {
        CK(s, cudaStreamCreateWithFlags(&s->aux_stream, cudaStreamNonBlocking));
        CK(s, cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
        CK(s, cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
    }
    // One epoch value per exchange op of the plan, as in the non-overlapped execution (a neighbour whose slab is too
    // thin to overlap runs the two exchanges separately and must find the same values): pulling the density together
    // with u, v, w consumes the density exchange's value too.  It is published at once -- the density "now" has been
    // final since the fill -- and a later value implies the earlier one for a neighbour that waits for it.
    const unsigned first = s->epoch + 1;
    if (with_smoke) s->epoch++;
    if ((rc = peer_signal(s, s->stream))) return rc;
    CK(s, cudaEventRecord(s->ev_fork, s->stream));
    CK(s, cudaStreamWaitEvent(s->aux_stream, s->ev_fork, 0));
    if ((rc = peer_wait(s, s->aux_stream, first))) return rc;
    if ((rc = p2p_pull(s, slab::SET_VEL_NOW, s->aux_stream))) return rc;
    if (with_smoke && (rc = p2p_pull(s, slab::SET_SMOKE_NOW, s->aux_stream))) return rc;
    CK(s, cudaEventRecord(s->ev_join, s->aux_stream));
    // While the pull is in flight: the planes at least M inside the slab, M = the margin the projected velocities ask for
    // (compute_margin: on the device, read by the launches below -- a frame hitch with a large dt widens the strips
    // instead of failing).  Every plane such a node reads is an owned plane, never written by the pull.
    if ((rc = compute_margin(s, dt))) return rc;
    const slab::Geom& ge = s->geom;
    const int G = ge.ghost;
    if ((rc = stage_advect_velocity(s, dt, adv.a, adv.b, ge.own_node_lo(), ge.own_node_hi(), 1))) return rc;
    CK(s, cudaStreamWaitEvent(s->stream, s->ev_join, 0));
    // ... then the strips at the slab ends (at most `ghost` planes each; the launches size themselves from M), with every
    // stored plane valid
    if (ge.has_lower() && (rc = stage_advect_velocity(s, dt, adv.a, std::min(adv.b, ge.own_node_lo() + G), adv.p0, adv.p1, 2))) return rc;
    if (ge.has_upper() && (rc = stage_advect_velocity(s, dt, std::max(adv.a, ge.own_node_hi() - G + 1), adv.b, adv.p0, adv.p1, 3))) return rc;
    return SMK_OK;
}

// Sparse blocking readback (opt-in, single GPU).  The density is exactly zero outside the plume, and non-zero density
// moves at most (backtrace reach + 1) cells per step.  The caller's buffer is known to hold non-zero values only in the
// rows `host_box` (what the previous readback found); so copying the rows grow(host_box, 2) leaves the buffer
// bit-identical to a full copy PROVIDED the new density's non-zero rows lie inside them -- which a small reduction
// kernel reports and smk_sync checks after the step; if not (a source moved, a long backtrace, first call, another
// buffer), the full copy follows.  The reference always copies everything (cu:814).
int readback_box(smk_sim* s, float* density_host)
{
    const GridP& g = s->g;
    Span sp(s, SMK_STAGE_READBACK);
    if (!s->d_box) {
        CK(s, cudaMalloc(&s->d_box, 4 * sizeof(int)));
        CK(s, cudaMallocHost(&s->h_box, 4 * sizeof(int)));
    }
    const float* src = s->smoke[s->past];
    smk::k_box_init<<<1, 1, 0, s->stream>>>(s->d_box);
    smk::k_density_rows_box<<<4 * s->num_sms, 256, 0, s->stream>>>(src, g.W, g.H, g.nzc, g.zlo, s->d_box);
    s->launches += 2;
    CK(s, cudaMemcpyAsync(s->h_box, s->d_box, 4 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    const size_t bytes = (size_t)g.nzc * g.cplane * sizeof(float);
    try_register(s, density_host, bytes);
    int c[4] = {0, 0, -1, -1}; // default: everything
    if (s->host_box_buf == density_host) {
        const int G = 2;
        if (s->host_box[0] >= s->host_box[1]) { c[0] = c[1] = c[2] = c[3] = 0; } // the buffer is all zero: copy nothing
        else {
            c[0] = std::max(0, s->host_box[0] - G); c[1] = std::min(g.H, s->host_box[1] + G);
            c[2] = std::max(0, s->host_box[2] - G); c[3] = std::min(g.D, s->host_box[3] + G);
            if ((double)(c[1] - c[0]) * (c[3] - c[2]) > 0.6 * (double)g.H * g.D) { c[0] = c[1] = 0; c[2] = c[3] = -1; } // not worth it
        }
    }
    s->readback_bytes += c[2] < 0 ? bytes : (size_t)std::max(0, c[1] - c[0]) * g.W * 4 * (size_t)std::max(0, c[3] - c[2]);
    if (c[2] < 0) {
        CK(s, cudaMemcpyAsync(density_host, src, bytes, cudaMemcpyDeviceToHost, s->stream));
    } else if (c[0] < c[1]) {
        const size_t off = (size_t)c[2] * g.cplane + (size_t)c[0] * g.W;
        CK(s, cudaMemcpy2DAsync(density_host + off, (size_t)g.cplane * 4, src + off, (size_t)g.cplane * 4,
                                (size_t)(c[1] - c[0]) * g.W * 4, (size_t)(c[3] - c[2]), cudaMemcpyDeviceToHost, s->stream));
    }
    for (int i = 0; i < 4; i++) s->box_copied[i] = c[i];
    s->box_pending_host = density_host;
    return SMK_OK;
}

// after the step's stream has been synchronised: the new density's non-zero rows are known; complete the readback
int finish_readback_box(smk_sim* s)
{
    float* host = s->box_pending_host;
    s->box_pending_host = nullptr;
    int e[4] = {s->h_box[0], s->h_box[1], s->h_box[2], s->h_box[3]};
    if (e[1] == INT_MIN) { e[0] = e[1] = e[2] = e[3] = 0; } // all zero
    const int* c = s->box_copied;
    const bool covered = c[2] < 0 || e[0] >= e[1] || (c[0] < c[1] && e[0] >= c[0] && e[1] <= c[1] && e[2] >= c[2] && e[3] <= c[3]);
    if (!covered) {
        const GridP& g = s->g;
        CK(s, cudaMemcpy(host, s->smoke[s->past], (size_t)g.nzc * g.cplane * sizeof(float), cudaMemcpyDeviceToHost));
        s->readback_bytes += (size_t)g.nzc * g.cplane * sizeof(float);
    }
    for (int i = 0; i < 4; i++) s->host_box[i] = e[i];
    s->host_box_buf = host;
    return SMK_OK;
}

int enqueue_step(smk_sim* s, float dt, float* density_host, bool pipelined = false)
{
    int rc = SMK_OK;
    const bool pp = peer_passes(s) && effective_fuse(s) == 4;
    const std::vector<slab::Op> ops = slab::plan_step(s->geom, s->iterations, effective_fuse(s), s->carry, pp);
    // Blocking readback: the density advection is the last stage and nothing after it reads its output, so it is run
    // in z-chunks and every finished chunk starts its way to the host on the copy stream while the next one is
    // computed; the stream of the step then waits for the last copy.
    static const int split_n = getenv("SMK_READBACK_CHUNKS") ? atoi(getenv("SMK_READBACK_CHUNKS")) : 4;
    const bool boxed = s->box_mode && density_host && !pipelined && s->geom.world == 1 && (s->g.W & 3) == 0;
    const bool split = !boxed && density_host && !pipelined && split_n > 1 && !ops.empty() && ops.back().kind == slab::OP_ADVECT_SMOKE &&
                       ops.back().b - ops.back().a >= 8 * split_n;
    const size_t nops = split ? ops.size() - 1 : ops.size();
    bool smoke_pulled = false;
    for (size_t i = 0; i < nops && rc == SMK_OK; i++) {
        // Peer-memory path: the fill stamps the sources into BOTH density buffers, one of which the neighbours pulled
        // ghost planes from at the end of their previous step.  One handshake in front of the fill -- "my pulls of your
        // previous-step planes are done" -- keeps a neighbour that runs ahead from stamping a moved source under a pull
        // that is still in flight.  (Every rank does it, so the epoch values stay aligned.)
        if (ops[i].kind == slab::OP_FILL && s->p2p && s->geom.world > 1 && (rc = peer_sync(s))) break;
        if (ops[i].kind == slab::OP_EXCHANGE && ops[i].a == slab::SET_SMOKE_NOW && smoke_pulled) continue;
        if (ops[i].kind == slab::OP_EXCHANGE && ops[i].a == slab::SET_VEL_NOW && i + 1 < ops.size() &&
            ops[i + 1].kind == slab::OP_ADVECT_VEL && overlap_exchange_ok(s, ops[i + 1])) {
            const bool with_smoke = i + 2 < ops.size() && ops[i + 2].kind == slab::OP_EXCHANGE && ops[i + 2].a == slab::SET_SMOKE_NOW;
            rc = exchange_overlapped_with_advect(s, ops[i + 1], dt, with_smoke);
            smoke_pulled = with_smoke;
            i++; // the advection is done
            continue;
        }
        rc = exec_op(s, ops[i], dt);
    }
    if (rc) return rc;
    if (boxed) return readback_box(s, density_host);
    if (split) {
        const GridP& g = s->g;
        const slab::Op& op = ops.back();
        const int c0 = s->geom.c0, c1 = s->geom.c1;
        try_register(s, density_host + (size_t)c0 * g.cplane, (size_t)(c1 - c0) * g.cplane * sizeof(float));
        if (!s->copy_stream) {
            CK(s, cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
            CK(s, cudaEventCreateWithFlags(&s->ev_snap, cudaEventDisableTiming));
            CK(s, cudaEventCreateWithFlags(&s->ev_copied, cudaEventDisableTiming));
        }
        if (!s->ev_chunk[0])
            for (auto& e : s->ev_chunk) CK(s, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        const int nch = std::min(split_n, (int)(sizeof(s->ev_chunk) / sizeof(s->ev_chunk[0])));
        for (int i = 0; i < nch; i++) {
            const int a = op.a + (int)((long long)(op.b - op.a) * i / nch), b = op.a + (int)((long long)(op.b - op.a) * (i + 1) / nch);
            if ((rc = stage_advect_smoke(s, dt, a, b, op.p0, op.p1))) return rc;
            CK(s, cudaEventRecord(s->ev_chunk[i], s->stream));
            CK(s, cudaStreamWaitEvent(s->copy_stream, s->ev_chunk[i], 0));
            // the owned planes outside [op.a, op.b) are boundary planes the advection never writes: they ride along
            const int pa = i == 0 ? c0 : a, pb = i == nch - 1 ? c1 : b;
            CK(s, cudaMemcpyAsync(density_host + (size_t)pa * g.cplane, s->smoke[s->past] + (size_t)(pa - g.zlo) * g.cplane,
                                  (size_t)(pb - pa) * g.cplane * sizeof(float), cudaMemcpyDeviceToHost, s->copy_stream));
            s->readback_bytes += (size_t)(pb - pa) * g.cplane * sizeof(float);
        }
        Span sp(s, SMK_STAGE_READBACK);
        CK(s, cudaEventRecord(s->ev_copied, s->copy_stream));
        CK(s, cudaStreamWaitEvent(s->stream, s->ev_copied, 0));
        return SMK_OK;
    }
    if (density_host) { // this slab's OWNED planes of the new density (a single GPU owns everything)
        Span sp(s, SMK_STAGE_READBACK);
        const GridP& g = s->g;
        const size_t bytes = (size_t)(s->geom.c1 - s->geom.c0) * g.cplane * sizeof(float);
        float* dst = density_host + (size_t)s->geom.c0 * g.cplane;
        try_register(s, dst, bytes);
        const float* src = s->smoke[s->past] + (size_t)(s->geom.c0 - g.zlo) * g.cplane;
        s->readback_bytes += bytes;
        if (!pipelined) {
            CK(s, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s->stream));
        } else {
            // the next step's source fill writes into this buffer: snapshot it (device to device, ~45 us at 256^3), then
            // let the copy engine move the snapshot to the host while the next step runs
            if (!s->copy_stream) {
                CK(s, cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
                CK(s, cudaEventCreateWithFlags(&s->ev_snap, cudaEventDisableTiming));
                CK(s, cudaEventCreateWithFlags(&s->ev_copied, cudaEventDisableTiming));
            }
            if (!s->snapshot) CK(s, cudaMalloc(&s->snapshot, bytes));
            if (s->copy_pending) CK(s, cudaStreamWaitEvent(s->stream, s->ev_copied, 0)); // previous snapshot fully on the host
            CK(s, cudaMemcpyAsync(s->snapshot, src, bytes, cudaMemcpyDeviceToDevice, s->stream));
            CK(s, cudaEventRecord(s->ev_snap, s->stream));
            CK(s, cudaStreamWaitEvent(s->copy_stream, s->ev_snap, 0));
            CK(s, cudaMemcpyAsync(dst, s->snapshot, bytes, cudaMemcpyDeviceToHost, s->copy_stream));
            CK(s, cudaEventRecord(s->ev_copied, s->copy_stream));
            s->copy_pending = true;
        }
    }
    return SMK_OK;
}

struct FieldRef {
    void* dev;
    size_t elem;
    bool staggered;
    bool is_mask;
};

int field_ref(smk_sim* s, int field, int which, FieldRef* out)
{
    if (which < 0 || which > 3) return SMK_ERR_ARG;
    const int b = which == SMK_BUF_NOW ? s->now : which == SMK_BUF_PAST ? s->past : which - 2;
    switch (field) {
    case SMK_FIELD_SMOKE: *out = {s->smoke[b], 4, false, false}; return SMK_OK;
    case SMK_FIELD_U: *out = {s->u[b], 4, true, false}; return SMK_OK;
    case SMK_FIELD_V: *out = {s->v[b], 4, true, false}; return SMK_OK;
    case SMK_FIELD_W: *out = {s->w[b], 4, true, false}; return SMK_OK;
    case SMK_FIELD_MASK: *out = {s->mask, 1, false, true}; return SMK_OK;
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
        const int z0 = f.is_mask ? g.mzlo : g.zlo, nz = f.is_mask ? g.nzm : g.nzc;
        char* hbase = (char*)host + (size_t)z0 * hx * hy * f.elem;
        cudaPitchedPtr hp = make_cudaPitchedPtr(hbase, hx * f.elem, hx, hy);
        cudaPitchedPtr dp = make_cudaPitchedPtr(f.dev, hx * f.elem, hx, hy);
        p.srcPtr = to_host ? dp : hp;
        p.dstPtr = to_host ? hp : dp;
        p.extent = make_cudaExtent(hx * f.elem, hy, (size_t)nz);
    }
    p.kind = to_host ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice;
    CK(s, cudaMemcpy3DAsync(&p, s->stream));
    CK(s, cudaStreamSynchronize(s->stream));
    return SMK_OK;
}

// cuTensorMapEncodeTiled is a driver-API entry point: fetched through the runtime so that the library does not link libcuda
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn()
{
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return (EncodeTiledFn)fn;
}

// tensor maps of the fused pressure pass over ONE arena (this slab's or a neighbour's): per physical buffer id one 4-D
// tensor (x, y, z, field) over u, v, w -- the three arrays of a buffer set lie lay.v[id] - lay.u[id] bytes apart -- with
// box 64 x 32 x 1 x 3, and the two density buffers (box 64 x 32 x 1)
bool build_pass_maps(const GridP& g, char* arena, const ArenaLayout& lay, int nzn, int nzc, CUtensorMap* node, CUtensorMap* smoke, bool* smoke_ok)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    for (int id = 0; id < 3; id++) {
        const size_t fs = lay.v[id] - lay.u[id];
        if (lay.w[id] - lay.v[id] != fs || (fs & 15)) return false;
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        const cuuint32_t box[4] = {64, 32, 1, 3};
        const cuuint64_t dims[4] = {(cuuint64_t)g.P, (cuuint64_t)g.SY, (cuuint64_t)nzn, 3};
        const cuuint64_t str[3] = {(cuuint64_t)g.P * 4, (cuuint64_t)g.nplane * 4, (cuuint64_t)fs};
        if (enc(&node[id], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, arena + lay.u[id], dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    const cuuint32_t estr[3] = {1, 1, 1};
    const cuuint32_t box[3] = {64, 32, 1};
    *smoke_ok = (g.W & 3) == 0 && nzc > 0;
    const cuuint64_t cdims[3] = {(cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)std::max(nzc, 1)};
    const cuuint64_t cstr[2] = {(cuuint64_t)g.W * 4, (cuuint64_t)g.cplane * 4};
    for (int b = 0; b < 2; b++) {
        if (*smoke_ok && enc(&smoke[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, arena + lay.smoke[b], cdims, cstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            *smoke_ok = false;
        if (!*smoke_ok) smoke[b] = node[0]; // never dereferenced (forcing then runs as its own kernel)
    }
    return true;
}

bool build_pass_code_map(smk_sim* s)
{
    EncodeTiledFn enc = encode_tiled_fn();
    const GridP& g = s->g;
    if (!enc || g.nzc <= 0) return false;
    const cuuint32_t estr[3] = {1, 1, 1};
    const cuuint32_t box[3] = {80, 32, 1};
    const cuuint64_t dims[3] = {(cuuint64_t)g.PC, (cuuint64_t)g.H, (cuuint64_t)g.nzc};
    const cuuint64_t str[2] = {(cuuint64_t)g.PC, (cuuint64_t)g.kplane};
    return enc(&s->pmap_code, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, s->pcode, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool build_tensor_maps(smk_sim* s)
{
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return false;
    }
    const GridP& g = s->g;
    const cuuint64_t dims[3] = {(cuuint64_t)g.P, (cuuint64_t)g.SY, (cuuint64_t)g.nzn};
    const cuuint64_t strides[2] = {(cuuint64_t)g.P * 4, (cuuint64_t)g.nplane * 4}; // bytes, dims 1 and 2
    const cuuint32_t box[3] = {smk::AdvTma::BX, smk::AdvTma::BY, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    for (int f = 0; f < 3; f++)
        for (int id = 0; id < 3; id++) {
            const size_t off = f == 0 ? s->lay.u[id] : f == 1 ? s->lay.v[id] : s->lay.w[id];
            CUresult r = ((EncodeTiledFn)fn)(&s->tmap[f][id], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, s->arena + off, dims, strides, box, estr,
                                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return false;
        }
    return true;
}

int create_common(smk_sim** out, unsigned W, unsigned H, unsigned D, int rank, int world, int ghost, const float* smoke0_full)
{
    if (!out) return SMK_ERR_ARG;
    *out = nullptr;
    if (W < 3 || H < 3 || D < 3 || W > 4096 || H > 4096 || D > 65535) return fail(nullptr, SMK_ERR_ARG, "grid dimensions out of range");
    if (world < 1 || rank < 0 || rank >= world) return fail(nullptr, SMK_ERR_ARG, "bad rank / world size");
    const slab::Geom geom = slab::make_geom((int)W, (int)H, (int)D, world, rank, ghost);
    if (!slab::geom_ok(geom)) return fail(nullptr, SMK_ERR_ARG, "slab too thin for its ghost depth (need D/world >= ghost+1, ghost >= 4)");
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
    s->geom = geom;
    s->carry = slab::initial_carry(geom);
    g.zlo = geom.zlo;
    g.nzc = geom.zhc - geom.zlo;
    g.nzn = g.nzc + 1;
    g.mzlo = std::max(0, geom.zlo - (geom.has_lower() ? 1 : 0));
    g.nzm = std::min((int)D, geom.zhc + (geom.has_upper() ? 1 : 0)) - g.mzlo;
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
    const size_t cb = cell_count(g) * sizeof(float);
    s->lay = make_layout(geom);
    CKN(cudaMalloc(&s->arena, s->lay.total));
    CKN(cudaMemsetAsync(s->arena, 0, s->lay.total, s->stream)); // every buffer zero-initialised, counters 0
    for (int i = 0; i < 2; i++) {
        s->smoke[i] = reinterpret_cast<float*>(s->arena + s->lay.smoke[i]);
        s->u[i] = reinterpret_cast<float*>(s->arena + s->lay.u[i]);
        s->v[i] = reinterpret_cast<float*>(s->arena + s->lay.v[i]);
        s->w[i] = reinterpret_cast<float*>(s->arena + s->lay.w[i]);
    }
    s->scratch[0] = reinterpret_cast<float*>(s->arena + s->lay.u[2]);
    s->scratch[1] = reinterpret_cast<float*>(s->arena + s->lay.v[2]);
    s->scratch[2] = reinterpret_cast<float*>(s->arena + s->lay.w[2]);
    const size_t mask_bytes = (size_t)g.cplane * g.nzm;
    CKN(cudaMalloc(&s->mask, mask_bytes));
    CKN(cudaMalloc(&s->d_flags, 64));
    CKN(cudaMemsetAsync(s->d_flags, 0, 64, s->stream));
    CKN(cudaMalloc(&s->code, code_count(g)));
    CKN(cudaMalloc(&s->pcode, code_count(g)));
    CKN(cudaMemsetAsync(s->pcode, 0, code_count(g), s->stream));
    {
        using T = smk::TmaCfg<4, 16, false>;
        s->cflag_tx = ((int)W + 1 + T::OX - 1) / T::OX;
        s->cflag_ty = ((int)H + 1 + T::OY - 1) / T::OY;
        const size_t nb = std::max<size_t>(1, (size_t)g.nzc * s->cflag_tx * s->cflag_ty);
        CKN(cudaMalloc(&s->cflag, nb));
        CKN(cudaMemsetAsync(s->cflag, 1, nb, s->stream)); // until the first k_codes run: assume COMPLEX cells everywhere
    }
    CKN(cudaMalloc(&s->d_scalar, 64));
    if (getenv("SMK_PASS_DEBUG")) { CKN(cudaMalloc(&s->d_passdbg, (size_t)1 << 22)); CKN(cudaMemset(s->d_passdbg, 0, (size_t)1 << 22));
        long long tile[3] = {1, 2, 1};
        if (const char* e = getenv("SMK_TRACE_TILE")) sscanf(e, "%lld,%lld,%lld", &tile[0], &tile[1], &tile[2]);
        CKN(cudaMemcpy(s->d_passdbg + (1 << 17) - 3, tile, sizeof(tile), cudaMemcpyHostToDevice)); }
    CKN(cudaMalloc(&s->d_dyn, 64));
    CKN(cudaMemsetAsync(s->d_dyn, 0, 64, s->stream));
    CKN(cudaMemsetAsync(s->code, 0, code_count(g), s->stream));
    // mask: fluid everywhere, solid on the plane y == 0 (cu:200-207)
    CKN(cudaMemsetAsync(s->mask, 1, mask_bytes, s->stream));
    CKN(cudaMemset2DAsync(s->mask, (size_t)g.cplane, 0, (size_t)g.W, (size_t)g.nzm, s->stream));
    if (smoke0_full)
        CKN(cudaMemcpyAsync(s->smoke[0], smoke0_full + (size_t)g.zlo * g.cplane, cb, cudaMemcpyHostToDevice, s->stream));
    CKN(cudaStreamSynchronize(s->stream));
    s->tma_ok = build_tensor_maps(s);
    s->pass_tma_ok = build_pass_maps(g, s->arena, s->lay, g.nzn, g.nzc, s->pmap[0], s->pmap_smoke[0], &s->pass_tma_smoke_ok) && build_pass_code_map(s);
    for (int side = 1; side <= 2; side++) { // placeholders until a neighbour is attached (never dereferenced)
        memcpy(s->pmap[side], s->pmap[0], sizeof(s->pmap[0]));
        memcpy(s->pmap_smoke[side], s->pmap_smoke[0], sizeof(s->pmap_smoke[0]));
    }
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
    return create_common(out, W, H, D, 0, 1, 0, smoke0_host);
}

int smk_create_slab(smk_sim** out, unsigned W, unsigned H, unsigned D, unsigned rank, unsigned world, unsigned ghost,
                    const float* smoke0_host_full)
{
    return create_common(out, W, H, D, (int)rank, (int)world, (int)ghost, smoke0_host_full);
}

// host-only helpers (no CUDA): the slab geometry and the schedule of one step, as executed by smk_step
int smk_slab_geometry(unsigned W, unsigned H, unsigned D, unsigned world, unsigned rank, unsigned ghost, int* out8)
{
    if (!out8 || world < 1 || rank >= world) return SMK_ERR_ARG;
    const slab::Geom g = slab::make_geom((int)W, (int)H, (int)D, (int)world, (int)rank, (int)ghost);
    out8[0] = g.c0; out8[1] = g.c1; out8[2] = g.zlo; out8[3] = g.zhc; out8[4] = g.ghost; out8[5] = slab::geom_ok(g) ? 1 : 0;
    out8[6] = g.own_node_lo(); out8[7] = g.own_node_hi();
    return SMK_OK;
}

int smk_slab_plan(unsigned W, unsigned H, unsigned D, unsigned world, unsigned rank, unsigned ghost, int iterations, int fuse,
                  int steps, int* ops5, int max_ops)
{
    if (world < 1 || rank >= world || steps < 1) return -SMK_ERR_ARG;
    const slab::Geom g = slab::make_geom((int)W, (int)H, (int)D, (int)world, (int)rank, (int)ghost);
    slab::Carry carry = slab::initial_carry(g);
    int n = 0;
    for (int st = 0; st < steps; st++)
        for (const slab::Op& op : slab::plan_step(g, iterations, fuse, carry)) {
            if (ops5 && n < max_ops) { int* o = ops5 + 5 * n; o[0] = op.kind; o[1] = op.a; o[2] = op.b; o[3] = op.p0; o[4] = op.p1; }
            n++;
        }
    return n;
}

int smk_slab_plan_p2p(unsigned W, unsigned H, unsigned D, unsigned world, unsigned rank, unsigned ghost, int iterations, int fuse,
                      int steps, int* ops5, int max_ops)
{
    if (world < 1 || rank >= world || steps < 1) return -SMK_ERR_ARG;
    const slab::Geom g = slab::make_geom((int)W, (int)H, (int)D, (int)world, (int)rank, (int)ghost);
    slab::Carry carry = slab::initial_carry(g);
    int n = 0;
    for (int st = 0; st < steps; st++)
        for (const slab::Op& op : slab::plan_step(g, iterations, fuse, carry, fuse == 4)) {
            if (ops5 && n < max_ops) { int* o = ops5 + 5 * n; o[0] = op.kind; o[1] = op.a; o[2] = op.b; o[3] = op.p0; o[4] = op.p1; }
            n++;
        }
    return n;
}

int smk_slab_regions(unsigned W, unsigned H, unsigned D, unsigned world, unsigned rank, unsigned ghost, int set, int* out5, int max_regions)
{
    if (world < 1 || rank >= world) return -SMK_ERR_ARG;
    const slab::Geom g = slab::make_geom((int)W, (int)H, (int)D, (int)world, (int)rank, (int)ghost);
    int n = 0;
    for (const slab::Region& r : slab::regions(g, set)) {
        if (out5 && n < max_regions) { int* o = out5 + 5 * n; o[0] = r.side; o[1] = r.send_lo; o[2] = r.send_n; o[3] = r.recv_lo; o[4] = r.recv_n; }
        n++;
    }
    return n;
}

int smk_pass_schedule(unsigned W, unsigned H, int out_lo, int out_hi, int K, int nctas, int* pieces4, int max_pieces, int* first,
                      int max_first, int* info4)
{
    if (K != 2 && K != 4) return -SMK_ERR_ARG;
    using C = smk::RegCfg<4, 16>; // OX does not depend on K (whole quads); OY does
    const int OY = C::LY - 2 * K;
    const int tx = ((int)W + 1 + C::OX - 1) / C::OX, ty = ((int)H + 1 + OY - 1) / OY;
    const sched::PassSchedule ps = sched::balance(tx, ty, out_lo, out_hi, sched::pass_lead(K), nctas);
    for (size_t i = 0; i < ps.pieces.size() && pieces4 && (int)i < max_pieces; i++) {
        pieces4[4 * i] = ps.pieces[i].bx; pieces4[4 * i + 1] = ps.pieces[i].by;
        pieces4[4 * i + 2] = ps.pieces[i].zo0; pieces4[4 * i + 3] = ps.pieces[i].zo1;
    }
    for (size_t i = 0; i < ps.first.size() && first && (int)i < max_first; i++) first[i] = ps.first[i];
    if (info4) { info4[0] = tx; info4[1] = ty; info4[2] = ps.cost; info4[3] = ps.nctas(); }
    return (int)ps.pieces.size();
}

int smk_destroy(smk_sim* s)
{ DeviceGuard dg(s);
    if (!s) return SMK_ERR_ARG;
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->copy_stream) { cudaStreamSynchronize(s->copy_stream); cudaStreamDestroy(s->copy_stream); }
    if (s->aux_stream) { cudaStreamSynchronize(s->aux_stream); cudaStreamDestroy(s->aux_stream); }
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    if (s->ev_snap) cudaEventDestroy(s->ev_snap);
    for (auto& e : s->ev_chunk) if (e) cudaEventDestroy(e);
    if (s->ev_copied) cudaEventDestroy(s->ev_copied);
    if (s->density_surf) cudaDestroySurfaceObject(s->density_surf);
    cudaFree(s->snapshot);
    cudaFree(s->half_stage);
    cudaFree(s->d_box);
    if (s->h_box) cudaFreeHost(s->h_box);
    while (!s->registered.empty()) drop_registration(s, s->registered.size() - 1);
    for (auto& sp : s->spans) { cudaEventDestroy(sp.e0); cudaEventDestroy(sp.e1); }
    for (auto e : s->free_events) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++)
        if (s->peer[i].arena && s->peer[i].ipc) cudaIpcCloseMemHandle(s->peer[i].arena);
    for (auto& d : s->schedules) { cudaFree(d.pieces); cudaFree(d.first); }
    cudaFree(s->arena);
    cudaFree(s->mask); cudaFree(s->code); cudaFree(s->pcode); cudaFree(s->cflag); cudaFree(s->d_scalar); cudaFree(s->d_flags); cudaFree(s->d_dyn); cudaFree(s->d_passdbg);
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
    s->mask_dirty = true;
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
    if (s->objects[id].type == 0) s->mask_dirty = true;
    return SMK_OK;
}

int smk_set_obstacle_mode(smk_sim* s, int mode)
{
    if (!s || (mode != SMK_OBSTACLES_LAST_WINS && mode != SMK_OBSTACLES_UNION)) return SMK_ERR_ARG;
    if (mode != s->obstacle_union) s->mask_dirty = true;
    s->obstacle_union = mode;
    return SMK_OK;
}

float* smk_gravity_ptr(smk_sim* s) { return s ? &s->gravity : nullptr; }
float* smk_buoyancy_ptr(smk_sim* s) { return s ? &s->alpha : nullptr; }

int smk_set_pass_kernel(smk_sim* s, int kind)
{
    if (!s || kind < SMK_PASS_AUTO || kind > SMK_PASS_TMA) return SMK_ERR_ARG;
    s->pass_kernel = kind;
    return SMK_OK;
}

int smk_set_solver(smk_sim* s, int variant, int iterations, int fuse)
{
    if (!s || iterations < 0 || fuse < 0) return SMK_ERR_ARG;
    if (variant != SMK_SOLVER_RBGS && variant != SMK_SOLVER_JACOBI) return fail(s, SMK_ERR_ARG, "solver variant not available");
    if (variant == SMK_SOLVER_JACOBI && s->geom.world > 1) return fail(s, SMK_ERR_ARG, "the Jacobi extension is single-GPU only");
    if (fuse != 0 && fuse != 1 && fuse != 2 && fuse != 4) return fail(s, SMK_ERR_ARG, "fuse must be 0, 1, 2 or 4");
    s->solver = variant; s->iterations = iterations; s->fuse = fuse;
    return SMK_OK;
}

int smk_set_pass_ctas(smk_sim* s, int nctas)
{
    if (!s || nctas < 0) return SMK_ERR_ARG;
    s->pass_ctas = nctas;
    return SMK_OK;
}

int smk_last_pass_ctas(smk_sim* s) { return s ? s->last_pass_ctas : -SMK_ERR_ARG; }
int smk_last_pass_kernel(smk_sim* s) { return s ? s->last_pass_kernel : -SMK_ERR_ARG; }

int smk_set_readback_box(smk_sim* s, int on)
{
    if (!s) return SMK_ERR_ARG;
    s->box_mode = on != 0;
    s->host_box_buf = nullptr; // whatever the caller's buffer holds is unknown again
    return SMK_OK;
}

int smk_set_stream(smk_sim* s, void* cuda_stream)
{ DeviceGuard dg(s);
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
{ DeviceGuard dg(s);
    if (!s) return SMK_ERR_ARG;
    return enqueue_step(s, dt, density_host, true);
}

int smk_sync(smk_sim* s)
{ DeviceGuard dg(s);
    if (!s) return SMK_ERR_ARG;
    CK(s, cudaStreamSynchronize(s->stream));
    if (s->copy_pending) { CK(s, cudaStreamSynchronize(s->copy_stream)); s->copy_pending = false; }
    if (s->box_pending_host) {
        int rc = finish_readback_box(s);
        if (rc) return rc;
    }
    { // device-side error flags: [0] backtrace left the valid planes (slab runs), [1] peer wait timed out, [2] TMA timeout (any run)
        int flag[3] = {0, 0, 0};
        CK(s, cudaMemcpy(flag, s->d_flags, sizeof(flag), cudaMemcpyDeviceToHost));
        if (flag[0] || flag[1] || flag[2]) cudaMemset(s->d_flags, 0, sizeof(flag));
        if (flag[2]) return fail(s, SMK_ERR_CUDA, "a TMA transaction of the advection kernel did not complete");
        if (flag[1]) return fail(s, SMK_ERR_TRANSPORT, "timed out waiting for a neighbour GPU (peer-memory halo path)");
        if (flag[0])
            return fail(s, SMK_ERR_REACH, "a backtrace reached beyond the slab's valid ghost planes (|w|*dt >= 1 cell): increase ghost");
    }
    if (s->spans.size() > 4096) return fold_timers(s);
    return SMK_OK;
}

int smk_step(smk_sim* s, float dt, float* density_host)
{ DeviceGuard dg(s);
    if (!s) return SMK_ERR_ARG;
    int rc = enqueue_step(s, dt, density_host);
    if (rc) return rc;
    return smk_sync(s);
}

const float* smk_density_device(smk_sim* s) { return s ? s->smoke[s->past] : nullptr; }

int smk_register_host(smk_sim* s, void* host, size_t bytes)
{
    if (!s || !host || bytes == 0) return SMK_ERR_ARG;
    DeviceGuard dg(s);
    if (!try_register(s, host, bytes, false)) return fail(s, SMK_ERR_CUDA, "cudaHostRegister failed for the caller's buffer");
    return SMK_OK;
}

int smk_unregister_host(smk_sim* s, void* host)
{
    if (!s || !host) return SMK_ERR_ARG;
    DeviceGuard dg(s);
    CK(s, cudaStreamSynchronize(s->stream));
    if (s->copy_stream) CK(s, cudaStreamSynchronize(s->copy_stream));
    char* p = static_cast<char*>(host);
    for (size_t i = 0; i < s->registered.size(); i++)
        if (s->registered[i].p <= p && p < s->registered[i].p + s->registered[i].bytes) {
            drop_registration(s, i);
            return SMK_OK;
        }
    return SMK_OK; // not registered by this handle: nothing to do
}

// this slab's owned planes of the last step's density as binary16 -- converted on the device
// (round to nearest even), half the bytes over PCIe.  Blocking.  Never used by the drop-in entry points.
int smk_read_density_half(smk_sim* s, void* host_half)
{ DeviceGuard dg(s);
    if (!s || !host_half) return SMK_ERR_ARG;
    const GridP& g = s->g;
    const size_t n = (size_t)(s->geom.c1 - s->geom.c0) * g.cplane;
    if (!s->half_stage) CK(s, cudaMalloc(&s->half_stage, n * 2));
    const float* src = s->smoke[s->past] + (size_t)(s->geom.c0 - g.zlo) * g.cplane;
    smk::k_density_half<<<(unsigned)((n / 2 + 256) / 256), 256, 0, s->stream>>>(src, static_cast<__half*>(s->half_stage), n);
    s->launches++;
    CK(s, cudaGetLastError());
    char* dst = static_cast<char*>(host_half) + (size_t)s->geom.c0 * g.cplane * 2;
    try_register(s, dst, n * 2);
    CK(s, cudaMemcpyAsync(dst, s->half_stage, n * 2, cudaMemcpyDeviceToHost, s->stream));
    return smk_sync(s);
}

// the renderer samples the density as an R32F 3-D texture (boundingBox.cpp:364-385).  With CUDA-GL interop
// (cudaGraphicsGLRegisterImage on m_gridTex -> cudaGraphicsSubResourceGetMappedArray) the new density goes straight
// into that array: one device-to-device 3-D copy on the step's stream, no host round trip.
int smk_copy_density_to_array(smk_sim* s, void* cuda_array)
{ DeviceGuard dg(s);
    if (!s || !cuda_array) return SMK_ERR_ARG;
    const GridP& g = s->g;
    cudaMemcpy3DParms p{};
    p.srcPtr = make_cudaPitchedPtr(s->smoke[s->past] + (size_t)(s->geom.c0 - g.zlo) * g.cplane, (size_t)g.W * 4, (size_t)g.W, (size_t)g.H);
    p.dstArray = static_cast<cudaArray_t>(cuda_array);
    p.dstPos = make_cudaPos(0, 0, (size_t)s->geom.c0);
    p.extent = make_cudaExtent((size_t)g.W, (size_t)g.H, (size_t)(s->geom.c1 - s->geom.c0));
    p.kind = cudaMemcpyDeviceToDevice;
    CK(s, cudaMemcpy3DAsync(&p, s->stream));
    return SMK_OK;
}

// SURVEY N1 as specified: bind the array ONCE; from then on the density advection itself writes every new density value
// into it with surf3Dwrite (k_advect_smoke_surf) -- no device-to-device pass, no host round trip.  NULL unbinds.
int smk_bind_density_array(smk_sim* s, void* cuda_array)
{
    if (!s) return SMK_ERR_ARG;
    DeviceGuard dg(s);
    CK(s, cudaStreamSynchronize(s->stream));
    if (s->density_surf) { cudaDestroySurfaceObject(s->density_surf); s->density_surf = 0; s->density_array = nullptr; }
    if (!cuda_array) return SMK_OK;
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = static_cast<cudaArray_t>(cuda_array);
    CK(s, cudaCreateSurfaceObject(&s->density_surf, &rd));
    s->density_array = static_cast<cudaArray_t>(cuda_array);
    return SMK_OK;
}

// the mask as one bit per cell (bit i of byte k = cell 8k + i, 1 = fluid), W*H*D/8 bytes (rounded up).
// smk_set_mask_bits == smk_set_field(SMK_FIELD_MASK) of the unpacked bytes (the stencil codes are rebuilt, the obstacle
// list is re-applied by the next fill like the reference would, cu:289-313); a slab touches the planes it stores.
int smk_set_mask_bits(smk_sim* s, const unsigned char* bits)
{
    if (!s || !bits) return SMK_ERR_ARG;
    DeviceGuard dg(s);
    const GridP& g = s->g;
    const size_t total = (size_t)g.cplane * g.D, first = (size_t)g.mzlo * g.cplane, n = (size_t)g.nzm * g.cplane;
    unsigned char* d = nullptr;
    CK(s, cudaMalloc(&d, (total + 7) / 8));
    cudaError_t e = cudaMemcpyAsync(d, bits, (total + 7) / 8, cudaMemcpyHostToDevice, s->stream);
    if (e == cudaSuccess) {
        smk::k_mask_unpack<<<(unsigned)((n + 255) / 256), 256, 0, s->stream>>>(d, s->mask, first, n);
        s->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(s, SMK_ERR_CUDA, std::string("smk_set_mask_bits: ") + cudaGetErrorString(e));
    s->carry = slab::initial_carry(s->geom);
    s->mask_dirty = true;
    launch_codes(s);
    CK(s, cudaGetLastError());
    CK(s, cudaStreamSynchronize(s->stream));
    return SMK_OK;
}

int smk_get_mask_bits(smk_sim* s, unsigned char* bits)
{
    if (!s || !bits) return SMK_ERR_ARG;
    DeviceGuard dg(s);
    const GridP& g = s->g;
    const size_t first = (size_t)g.mzlo * g.cplane, n = (size_t)g.nzm * g.cplane;
    if (first % 8) return fail(s, SMK_ERR_ARG, "smk_get_mask_bits: the slab's first stored cell is not on a byte boundary of the packed mask");
    unsigned char* d = nullptr;
    const size_t nb = (n + 7) / 8;
    CK(s, cudaMalloc(&d, nb));
    smk::k_mask_pack<<<(unsigned)((nb + 255) / 256), 256, 0, s->stream>>>(s->mask, d - (first >> 3), first, n);
    s->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(bits + (first >> 3), d, nb, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(s, SMK_ERR_CUDA, std::string("smk_get_mask_bits: ") + cudaGetErrorString(e));
    return SMK_OK;
}

// test helpers for the above (a plain 3-D float array stands in for the mapped GL texture)
void* smk_test_array_create(unsigned W, unsigned H, unsigned D)
{
    cudaArray_t a = nullptr;
    cudaChannelFormatDesc d = cudaCreateChannelDesc<float>();
    if (cudaMalloc3DArray(&a, &d, make_cudaExtent(W, H, D), cudaArraySurfaceLoadStore) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return a;
}
int smk_test_array_read(void* cuda_array, float* host, unsigned W, unsigned H, unsigned D)
{
    cudaMemcpy3DParms p{};
    p.srcArray = static_cast<cudaArray_t>(cuda_array);
    p.dstPtr = make_cudaPitchedPtr(host, (size_t)W * 4, W, H);
    p.extent = make_cudaExtent(W, H, D);
    p.kind = cudaMemcpyDeviceToHost;
    return cudaMemcpy3D(&p) == cudaSuccess ? SMK_OK : SMK_ERR_CUDA;
}
void smk_test_array_destroy(void* cuda_array) { cudaFreeArray(static_cast<cudaArray_t>(cuda_array)); }

int smk_stage_flip(smk_sim* s) { DeviceGuard dg(s); if (!s) return SMK_ERR_ARG; flip(s); return SMK_OK; }
int smk_stage_fill(smk_sim* s) { DeviceGuard dg(s); return s ? stage_fill(s) : SMK_ERR_ARG; }
int smk_stage_force_clamp(smk_sim* s, float dt) { DeviceGuard dg(s); return s ? stage_force_clamp(s, dt, s->g.zlo, s->g.zlo + s->g.nzn) : SMK_ERR_ARG; }
int smk_stage_pressure_halfsweep(smk_sim* s, int offset)
{ DeviceGuard dg(s);
    if (!s) return SMK_ERR_ARG;
    Span sp(s, SMK_STAGE_PRESSURE);
    launch_halfsweep(s, offset & 1);
    CK(s, cudaGetLastError());
    return SMK_OK;
}
int smk_stage_pressure(smk_sim* s) { DeviceGuard dg(s); return s ? stage_pressure(s) : SMK_ERR_ARG; }
int smk_stage_advect_velocity(smk_sim* s, float dt) { DeviceGuard dg(s); return s ? stage_advect_velocity(s, dt, s->g.zlo, s->g.zlo + s->g.nzn, s->g.zlo, s->g.zlo + s->g.nzn - 1) : SMK_ERR_ARG; }
int smk_stage_advect_smoke(smk_sim* s, float dt) { DeviceGuard dg(s); return s ? stage_advect_smoke(s, dt, s->g.zlo, s->g.zlo + s->g.nzc, s->g.zlo, s->g.zlo + s->g.nzc - 1) : SMK_ERR_ARG; }

int smk_get_field(smk_sim* s, int field, int which, void* host_dst)
{ DeviceGuard dg(s);
    if (!s || !host_dst) return SMK_ERR_ARG;
    FieldRef f;
    if (field_ref(s, field, which, &f)) return fail(s, SMK_ERR_ARG, "bad field / buffer selector");
    return copy_field(s, f, host_dst, true);
}

int smk_set_field(smk_sim* s, int field, int which, const void* host_src)
{ DeviceGuard dg(s);
    if (!s || !host_src) return SMK_ERR_ARG;
    FieldRef f;
    if (field_ref(s, field, which, &f)) return fail(s, SMK_ERR_ARG, "bad field / buffer selector");
    int rc = copy_field(s, f, const_cast<void*>(host_src), false);
    if (rc) return rc;
    s->carry = slab::initial_carry(s->geom); // injected fields are full-size on every rank: all stored planes valid
    if (field == SMK_FIELD_MASK) { // keep the stencil codes consistent with an injected mask
        s->mask_dirty = true;      // ... and let the next fill apply the obstacle list to it, as the reference would
        launch_codes(s);
        CK(s, cudaGetLastError());
        CK(s, cudaStreamSynchronize(s->stream));
    }
    return SMK_OK;
}

int smk_index_now(smk_sim* s) { return s ? s->now : -1; }

// ---- peer-memory halo path (CUDA IPC between the per-GPU processes, or plain pointers inside one process) ----
int smk_p2p_export(smk_sim* s, unsigned char* handle64)
{ DeviceGuard dg(s);
    if (!s || !handle64) return SMK_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaIpcMemHandle_t h;
    CK(s, cudaIpcGetMemHandle(&h, s->arena));
    memcpy(handle64, &h, 64);
    return SMK_OK;
}

void* smk_p2p_arena(smk_sim* s) { return s ? s->arena : nullptr; }

static int attach_common(smk_sim* s, int side, char* arena, bool ipc)
{
    auto& pe = s->peer[side];
    pe.arena = arena; pe.ipc = ipc;
    pe.geom = slab::make_geom(s->geom.W, s->geom.H, s->geom.D, s->geom.world, s->geom.rank + (side == 0 ? -1 : 1), s->geom.ghost);
    pe.lay = make_layout(pe.geom);
    {   // tensor maps over the neighbour's arena: the fused passes read its boundary planes by TMA over NVLink
        GridP pg = s->g;
        const int pnzc = pe.geom.zhc - pe.geom.zlo;
        bool smoke_ok = false;
        s->pass_tma_peer_ok[side] = build_pass_maps(pg, arena, pe.lay, pnzc + 1, pnzc, s->pmap[side + 1], s->pmap_smoke[side + 1], &smoke_ok) &&
                                    (smoke_ok || !s->pass_tma_smoke_ok);
    }
    s->p2p = (!s->geom.has_lower() || s->peer[0].arena) && (!s->geom.has_upper() || s->peer[1].arena);
    return SMK_OK;
}

int smk_p2p_attach_ipc(smk_sim* s, int side, const unsigned char* handle64)
{ DeviceGuard dg(s);
    if (!s || !handle64 || side < 0 || side > 1) return SMK_ERR_ARG;
    if ((side == 0 && !s->geom.has_lower()) || (side == 1 && !s->geom.has_upper())) return fail(s, SMK_ERR_ARG, "no neighbour on that side");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    CK(s, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    return attach_common(s, side, static_cast<char*>(p), true);
}

int smk_p2p_attach_ptr(smk_sim* s, int side, void* peer_arena)
{
    if (!s || !peer_arena || side < 0 || side > 1) return SMK_ERR_ARG;
    if ((side == 0 && !s->geom.has_lower()) || (side == 1 && !s->geom.has_upper())) return fail(s, SMK_ERR_ARG, "no neighbour on that side");
    return attach_common(s, side, static_cast<char*>(peer_arena), false);
}

long smk_exchange_count(smk_sim* s) { return s ? s->exchanges : -1; }

// Publish the NEXT epoch to the neighbours without waiting.  Only needed when several slabs are driven from one host
// thread on one GPU (tests): there all signals of a synchronisation point must be enqueued before any wait, because
// streams may share a hardware queue and a spinning wait would block a signal queued behind it.
int smk_p2p_presignal(smk_sim* s)
{ DeviceGuard dg(s);
    if (!s || !s->p2p) return SMK_ERR_ARG;
    unsigned* theirs[2]; const unsigned* mine[2];
    peer_counters(s, theirs, mine);
    smk::k_epoch_signal<<<1, 1, 0, s->stream>>>(theirs[0], theirs[1], s->epoch + 1);
    s->launches++;
    CK(s, cudaGetLastError());
    return SMK_OK;
}

int smk_exec_op(smk_sim* s, const int* op5, float dt)
{ DeviceGuard dg(s);
    if (!s || !op5) return SMK_ERR_ARG;
    return exec_op(s, slab::Op{op5[0], op5[1], op5[2], op5[3], op5[4]}, dt);
}

int smk_max_divergence(smk_sim* s, float* out)
{ DeviceGuard dg(s);
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

// 64-bit content hashes of this slab's OWNED planes (k_hash_field): out7 = {u, v, w "now", u, v, w "past", density "past"}.
// Summing the values of all slabs (mod 2^64) gives the hash of the whole domain, whatever the decomposition.
int smk_hash_range(smk_sim* s, int nlo, int nhi, int clo, int chi, unsigned long long* out7)
{
    if (!s || !out7) return SMK_ERR_ARG;
    DeviceGuard dg(s);
    const GridP& g = s->g;
    if (nlo < g.zlo || nhi > g.zlo + g.nzn || clo < g.zlo || chi > g.zlo + g.nzc) return fail(s, SMK_ERR_ARG, "smk_hash_range: planes not stored by this handle");
    unsigned long long* d = nullptr;
    CK(s, cudaMalloc(&d, 7 * sizeof(unsigned long long)));
    CK(s, cudaMemsetAsync(d, 0, 7 * sizeof(unsigned long long), s->stream));
    const float* nodes[6] = {s->u[s->now], s->v[s->now], s->w[s->now], s->u[s->past], s->v[s->past], s->w[s->past]};
    for (int i = 0; i < 6 && nhi > nlo; i++) {
        const dim3 grid((unsigned)((g.W + 1 + 255) / 256), (unsigned)(g.H + 1), (unsigned)(nhi - nlo));
        smk::k_hash_field<<<grid, 256, 0, s->stream>>>(nodes[i], g.P, g.nplane, g.zlo, g.W + 1, g.H + 1, nlo, d + i);
    }
    if (chi > clo) {
        const dim3 grid((unsigned)((g.W + 255) / 256), (unsigned)g.H, (unsigned)(chi - clo));
        smk::k_hash_field<<<grid, 256, 0, s->stream>>>(s->smoke[s->past], g.W, g.cplane, g.zlo, g.W, g.H, clo, d + 6);
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out7, d, 7 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(s, SMK_ERR_CUDA, std::string("smk_hash_range: ") + cudaGetErrorString(e));
    return SMK_OK;
}

int smk_hash_owned(smk_sim* s, unsigned long long* out7)
{
    if (!s) return SMK_ERR_ARG;
    return smk_hash_range(s, s->geom.own_node_lo(), s->geom.own_node_hi() + 1, s->geom.c0, s->geom.c1, out7);
}

int smk_stage_time(smk_sim* s, int stage, double* ms_total, long* launches)
{ DeviceGuard dg(s);
    if (!s || stage < 0 || stage >= SMK_STAGE_COUNT) return SMK_ERR_ARG;
    int rc = fold_timers(s);
    if (rc) return rc;
    if (ms_total) *ms_total = s->stage_ms[stage];
    if (launches) *launches = s->stage_launches[stage];
    return SMK_OK;
}

int smk_reset_timers(smk_sim* s)
{ DeviceGuard dg(s);
    if (!s) return SMK_ERR_ARG;
    int rc = fold_timers(s);
    if (rc) return rc;
    for (int i = 0; i < SMK_STAGE_COUNT; i++) { s->stage_ms[i] = 0; s->stage_launches[i] = 0; }
    return SMK_OK;
}

long smk_launch_count(smk_sim* s) { return s ? s->launches : -1; }

// development aid: per-CTA {start clock, cycles, SM id, variant * 1000 + planes} of the last fused pass (SMK_PASS_DEBUG=1)
int smk_debug_pass_ctas(smk_sim* s, long long* out4, int max_ctas)
{
    if (!s || !out4) return -SMK_ERR_ARG;
    DeviceGuard dg(s);
    if (!s->d_passdbg) return 0;
    if (max_ctas < 0) { // trace region: 16 warps x 80 steps x 4 timestamps of one CTA
        if (cudaStreamSynchronize(s->stream) != cudaSuccess || cudaMemcpy(out4, s->d_passdbg + (1 << 17), (size_t)34 * 80 * 8 * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -SMK_ERR_CUDA;
        return 34 * 80;
    }
    const int n = std::min(max_ctas, std::min(s->passdbg_ctas, 4096));
    if (cudaStreamSynchronize(s->stream) != cudaSuccess || cudaMemcpy(out4, s->d_passdbg, (size_t)n * 32, cudaMemcpyDeviceToHost) != cudaSuccess) return -SMK_ERR_CUDA;
    return n;
}
unsigned long long smk_readback_bytes(smk_sim* s) { return s ? s->readback_bytes : 0; }

int smk_selfcheck_omega(int device, unsigned first, unsigned long long count, unsigned long long* mismatches, unsigned long long* ties)
{
    if (!mismatches || count > (1ull << 32)) return SMK_ERR_ARG;
    int prev = 0;
    if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) return SMK_ERR_CUDA;
    unsigned long long* d = nullptr;
    unsigned long long h[2] = {0, 0};
    int rc = SMK_OK;
    if (cudaMalloc(&d, 16) != cudaSuccess || cudaMemset(d, 0, 16) != cudaSuccess) rc = SMK_ERR_CUDA;
    if (rc == SMK_OK) {
        smk::k_omega_check<<<148 * 8, 256>>>(first, count, d);
        if (cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost) != cudaSuccess) rc = SMK_ERR_CUDA;
    }
    cudaFree(d);
    cudaSetDevice(prev);
    *mismatches = h[0];
    if (ties) *ties = h[1];
    return rc;
}

int smk_set_exchange(smk_sim* s, smk_exchange_fn fn, void* ctx)
{
    if (!s) return SMK_ERR_ARG;
    s->exchange = fn; s->exchange_ctx = ctx;
    return SMK_OK;
}

} // extern "C"

// slab_plan.h -- z-slab decomposition and per-step schedule of the smoke step (pure host C++, no CUDA).
//
// The reference is single-GPU; large grids are partitioned here as z-slabs, one per GPU / process
// (SURVEY 8(e)).  z is the slowest index of every field, so a slab and each of its halos is one contiguous range.
// A slab owns cell planes [c0, c1) and stores `ghost` more planes on each interior side.  This header decides
//   * the geometry of a slab (which planes it owns / stores),
//   * the list of operations of one step with their plane ranges, and WHERE halo exchanges are needed, by tracking
//     the interval of planes on which each field set is still correct ("valid"):
//        pointwise stages keep it, a pressure pass of K half-sweeps shrinks it by K planes at every interior end,
//        advection needs a margin of MARGIN planes around its output,
//        an exchange restores it to the whole stored range,
//   * the halo regions of an exchange (offsets in planes; the executor turns them into pointers).
// The same plan is executed by the CUDA library (smk_api.cu) and, in the CPU tests, by the oracle under a gloo process
// group (tests/test_slab_cpu.py), so the multi-GPU schedule is verified bit-exactly without GPUs.
#pragma once
#include <algorithm>
#include <vector>

namespace slab {

constexpr int MARGIN = 2; // the LEAST planes of "now" data needed beyond an advected plane: 1 (8-point averages), and
                          // floor/ceil of a backtrace shorter than one cell + the trilinear corner.
                          // dt is wall-clock time in the reference (main.cpp:891-895): longer backtraces are served by
                          // refreshing the WHOLE ghost depth in front of advection (below) and by the adaptive margin of
                          // the overlapped peer-memory path (smk_api.cu compute_margin); SMK_ERR_REACH is left for reaches
                          // beyond the ghost allocation itself (device-side guard of the samplers)

enum OpKind { OP_FLIP = 0, OP_FILL = 1, OP_FORCE = 2, OP_PRESSURE = 3, OP_ADVECT_VEL = 4, OP_ADVECT_SMOKE = 5, OP_EXCHANGE = 6 };
enum SetId { SET_VEL_NOW = 0, SET_SMOKE_NOW = 1 };

struct Op {
    int kind;
    int a, b;   // plane range [a, b) the op writes (cells for FILL / ADVECT_SMOKE, nodes otherwise); EXCHANGE: a = set id
    int p0, p1; // PRESSURE: first half-sweep index, number of half-sweeps; ADVECT_*: valid input range [p0, p1] (guard)
};

struct Geom {
    int W, H, D, rank, world, ghost;
    int c0, c1;   // owned cell planes [c0, c1)
    int zlo, zhc; // stored cell planes [zlo, zhc); stored node planes [zlo, zhc]
    bool has_lower() const { return rank > 0; }
    bool has_upper() const { return rank < world - 1; }
    int own_node_lo() const { return c0; }
    int own_node_hi() const { return has_upper() ? c1 - 1 : D; } // inclusive; the top rank also owns node plane D
};

inline Geom make_geom(int W, int H, int D, int world, int rank, int ghost)
{
    Geom g{W, H, D, rank, world, world > 1 ? ghost : 0, 0, 0, 0, 0};
    g.c0 = (int)((long long)D * rank / world);
    g.c1 = (int)((long long)D * (rank + 1) / world);
    g.zlo = std::max(0, g.c0 - g.ghost);
    g.zhc = std::min(D, g.c1 + g.ghost);
    return g;
}

// a slab must be thick enough to serve its neighbours' ghosts from owned planes, and deep enough for one pass
inline bool geom_ok(const Geom& g) { return g.world == 1 || (g.c1 - g.c0 >= g.ghost + 1 && g.ghost >= 4); }

struct Interval { int lo, hi; }; // inclusive plane range

// Validity is tracked as a DEPTH: how many planes beyond the owned ones (at an interior slab end) still hold correct
// data.  The depth is the same number on every rank, so every rank takes the same exchange decisions (the exchange
// is collective) even though the first and last slab have a domain boundary -- where nothing ever becomes invalid --
// on one side.  Node fields have one more stored plane above the slab than below; the depth is the smaller count.
struct Carry {
    int vel_d;   // u,v,w "now"
    int smoke_d; // density "now"
};

inline Carry initial_carry(const Geom& g) { return Carry{g.ghost, g.ghost}; } // all fields replicated at start

inline Interval node_interval(const Geom& g, int d) { return {g.has_lower() ? g.c0 - d : 0, g.has_upper() ? g.c1 - 1 + d : g.D}; }
inline Interval cell_interval(const Geom& g, int d) { return {g.has_lower() ? g.c0 - d : 0, g.has_upper() ? g.c1 - 1 + d : g.D - 1}; }

// One halo region of an exchange, in planes relative to nothing (global plane indices).
struct Region {
    int side;           // 0 = lower-z neighbour, 1 = upper-z neighbour
    int send_lo, send_n; // my owned planes that go out
    int recv_lo, recv_n; // my ghost planes that come in
};

// u,v,w live on node planes (one more than cells): the upper ghost has ghost+1 planes.
inline std::vector<Region> regions(const Geom& g, int set)
{
    std::vector<Region> r;
    const int extra = set == SET_VEL_NOW ? 1 : 0;
    if (g.has_lower()) {
        const int nb_hi = std::min(g.D - 1 + extra, g.c0 + g.ghost - 1 + extra); // last plane of the lower neighbour's upper ghost
        r.push_back({0, g.c0, nb_hi - g.c0 + 1, g.zlo, g.c0 - g.zlo});
    }
    if (g.has_upper()) {
        const int nb_lo = std::max(0, g.c1 - g.ghost); // first plane of the upper neighbour's lower ghost
        r.push_back({1, nb_lo, g.c1 - nb_lo, g.c1, g.zhc - 1 + extra - g.c1 + 1});
    }
    return r;
}

// The schedule of ONE step (cu:774-819) on one slab.  `fuse` = half-sweeps per pressure pass (1, 2 or 4).
// `peer_passes`: the pressure passes read the planes outside the owned range straight from the neighbours' memory
// (kernels_pressure_reg.cuh, PassRange): they consume no ghost depth and need no exchange, but leave the local ghost
// planes of u,v,w stale.
inline std::vector<Op> plan_step(const Geom& g, int iterations, int fuse, Carry& carry, bool peer_passes = false)
{
    std::vector<Op> ops;
    const int D = g.D, G = g.ghost;
    int vel_d = carry.vel_d, smoke_d = carry.smoke_d;
    auto exchange = [&](int set) {
        if (g.world > 1) ops.push_back({OP_EXCHANGE, set, 0, 0, 0});
        if (set == SET_VEL_NOW) vel_d = G; else smoke_d = G;
    };

    ops.push_back({OP_FLIP, 0, 0, 0, 0});
    ops.push_back({OP_FILL, g.zlo, g.zhc, 0, 0}); // analytic: every stored cell plane, never exchanged

    // forcing + clamp are pointwise: run where both inputs are valid; the result is valid there
    {
        vel_d = std::min(vel_d, smoke_d);
        const Interval r = node_interval(g, vel_d);
        ops.push_back({OP_FORCE, r.lo, std::min(r.hi, D - 1) + 1, 0, 0}); // node plane D has no cell: never forced
    }

    // pressure passes: K half-sweeps consume K planes of valid ghost depth at every interior end
    const int total = 2 * iterations;
    for (int done = 0; done < total;) {
        int K = 1;
        if (fuse >= 4 && total - done >= 4 && (done & 1) == 0) K = 4;
        else if (fuse >= 2 && total - done >= 2 && (done & 1) == 0) K = 2;
        if (peer_passes && g.world > 1 && K == 4) { // only the K = 4 register kernel reads peers
            ops.push_back({OP_PRESSURE, g.own_node_lo(), g.own_node_hi() + 1, done, K});
            vel_d = 0;
        } else {
            if (vel_d < K) exchange(SET_VEL_NOW);
            ops.push_back({OP_PRESSURE, g.zlo, g.zhc + 1, done, K});
            vel_d -= K;
        }
        done += K;
    }

    // u,v,w advection: node planes [1, D); the plane above the slab is computed redundantly because the density
    // advection of the top owned cell plane reads the NEW w on its upper face (node fields store one more plane above)
    const int va = std::max(1, g.c0), vb = std::min(D - 1, g.has_upper() ? g.c1 : D - 1); // inclusive
    {
        if (vel_d < G) exchange(SET_VEL_NOW); // the full ghost depth: a backtrace may reach ghost - 1 planes
        const Interval v = node_interval(g, vel_d);
        ops.push_back({OP_ADVECT_VEL, va, vb + 1, v.lo, g.has_upper() ? v.hi + 1 : v.hi});
    }
    // density advection: interior cell planes of the slab
    const int sa = std::max(1, g.c0), sb = std::min(D - 2, g.c1 - 1); // inclusive
    {
        if (smoke_d < G) exchange(SET_SMOKE_NOW);
        const Interval v = cell_interval(g, smoke_d);
        ops.push_back({OP_ADVECT_SMOKE, sa, sb + 1, v.lo, v.hi});
    }
    // the next step starts from what advection wrote: the owned planes only
    carry.vel_d = 0;
    carry.smoke_d = 0;
    return ops;
}

} // namespace slab

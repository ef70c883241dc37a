// dropin.cu -- the reference's host entry points (project/smokeSimulation.cuh:4-17) on top of the C ABI.
//
// Behaviour kept from the reference: process-global singleton state; initializeVolume() may run during
// static initialisation, before main() (main.cpp:87-88 -> boundingBox.cpp:32), so nothing here depends on
// other translation units' static objects; failures print to stderr and exit(-1) (cu:131-135, 815-818);
// simulate() returns with the caller's buffer filled; objects may be added before or after the volume
// exists; gravity / buoyancy are plain floats the GUI mutates through the returned pointers.
#include <cstdio>
#include <cstdlib>

#include "../../include/smoke_b200.h"
#include "../host/smokeSimulation.cuh"

namespace {

struct PendingObject {
    int type;
    float x, y, z, vx, vy, vz, r;
};

// function-local statics: safe to use from other TUs' static constructors
struct Global {
    smk_sim* sim = nullptr;
    float gravity = -9.82f;     // cu:28
    float buoyancy = 2.0f;      // cu:29
    PendingObject pending[2 * SMK_MAX_OBJECTS];
    int npending = 0;           // objects known so far (ids = index), replayed into every new volume
};
Global& G()
{
    static Global g;
    return g;
}

[[noreturn]] void die(const char* what, smk_sim* s)
{
    fprintf(stderr, "%s: %s\n", what, smk_last_error(s));
    exit(-1);
}

void push_object(const PendingObject& o)
{
    Global& g = G();
    if (g.npending >= 2 * SMK_MAX_OBJECTS) {
        fprintf(stderr, "too many scene objects (max %d)\n", 2 * SMK_MAX_OBJECTS);
        exit(-1);
    }
    g.pending[g.npending++] = o;
    if (g.sim) {
        int id = o.type == 0 ? smk_add_obstacle(g.sim, o.x, o.y, o.z, o.vx, o.vy, o.vz, o.r)
                             : smk_add_source(g.sim, o.x, o.y, o.z, o.r);
        if (id < 0) die("Error adding scene object", g.sim);
    }
}

} // namespace

void getGPUProperties(void)
{
    if (smk_print_gpu_properties() != SMK_OK) {
        fprintf(stderr, "Error finding available GPUs, now exiting\n");
        exit(-1);
    }
}

void initializeVolume(float* smoke_grid, unsigned int width, unsigned int heigth, unsigned int depth)
{
    Global& g = G();
    if (g.sim) { smk_destroy(g.sim); g.sim = nullptr; }
    if (smk_create(&g.sim, width, heigth, depth, smoke_grid) != SMK_OK) die("Error allocating smoke grid on GPU", nullptr);
    for (int i = 0; i < g.npending; i++) {
        const PendingObject& o = g.pending[i];
        int id = o.type == 0 ? smk_add_obstacle(g.sim, o.x, o.y, o.z, o.vx, o.vy, o.vz, o.r)
                             : smk_add_source(g.sim, o.x, o.y, o.z, o.r);
        if (id < 0) die("Error adding scene object", g.sim);
    }
}

void deleteVolume()
{
    Global& g = G();
    if (g.sim) { smk_destroy(g.sim); g.sim = nullptr; }
}

int addObstacle(float x, float y, float z, float vx, float vy, float vz, float r)
{
    push_object({0, x, y, z, vx, vy, vz, r});
    return G().npending - 1;
}

int addSmokeSource(float x, float y, float z, float r)
{
    push_object({1, x, y, z, 0.f, 0.f, 0.f, r});
    return G().npending - 1;
}

void updateObjectPos(int id, float x, float y, float z)
{
    Global& g = G();
    if (id < 0 || id >= g.npending) return; // the reference indexes out of bounds here (UB, cu:106-109)
    g.pending[id].x = x; g.pending[id].y = y; g.pending[id].z = z;
    if (g.sim) smk_update_object_pos(g.sim, id, x, y, z);
}

float* getBuoyancy() { return &G().buoyancy; }
float* getGravity() { return &G().gravity; }

void simulate(float* smoke_grid, float dt)
{
    Global& g = G();
    if (!g.sim) {
        fprintf(stderr, "simulate() called before initializeVolume()\n");
        exit(-1);
    }
    // the GUI writes through getGravity()/getBuoyancy(); the reference reads the globals at launch (cu:789)
    *smk_gravity_ptr(g.sim) = g.gravity;
    *smk_buoyancy_ptr(g.sim) = g.buoyancy;
    if (smk_step(g.sim, dt, smoke_grid) != SMK_OK) die("Error in simulation step", g.sim);
}

// test hook (not part of the reference interface): the handle behind the wrappers
extern "C" smk_sim* smk_dropin_handle(void) { return G().sim; }

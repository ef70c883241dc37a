// headless_main.cpp -- replays the reference application's call sequence against the drop-in entry points
// without a window: static-initialisation-time initializeVolume (main.cpp:87-88 -> boundingBox.cpp:32),
// getGPUProperties / addSmokeSource / addObstacle / first step with dt = 0.01 (main.cpp:257, 288-293), then
// fixed 1/20 s ticks (main.cpp:93, 891-895) with the GUI-style parameter writes through getGravity()/getBuoyancy()
// (main.cpp:832-833) and a moving source (main.cpp:609-618).  Compiled by plain g++ against
// host/smokeSimulation.cuh only -- no CUDA header -- exactly like the reference's main.cpp / boundingBox.cpp.
//
//   g++ -O2 headless_main.cpp -I. -L.. -lsmoke_b200 -Wl,-rpath,'$ORIGIN/..' -o headless
//   ./headless [N=80] [ticks=20]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "smokeSimulation.cuh"

namespace {
struct Volume { // stands in for BoundingBox (boundingBox.h:41, boundingBox.cpp:17-67, 387-403)
    std::vector<float> grid;
    unsigned n;
    explicit Volume(unsigned n_) : grid((size_t)n_ * n_ * n_, 0.f), n(n_) { initializeVolume(grid.data(), n, n, n); }
    void update(float dt) { simulate(grid.data(), dt); }
    ~Volume() { deleteVolume(); }
};
unsigned env_n() { const char* e = getenv("SMOKE_N"); return e ? (unsigned)atoi(e) : 80u; }
Volume volume(env_n()); // constructed before main(), like the reference's global BoundingBox
} // namespace

int main(int argc, char** argv)
{
    const int ticks = argc > 1 ? atoi(argv[1]) : 20;
    const float n = (float)volume.n, s = n / 80.f;
    getGPUProperties();
    const int src = addSmokeSource(40 * s, 40 * s, 40 * s, 5 * s);
    addObstacle(60 * s, 10 * s, 60 * s, 0, 0, 0, 13 * s);
    volume.update(0.01f);
    for (int t = 1; t < ticks; t++) {
        if (t == ticks / 2) { *getBuoyancy() = 4.0f; *getGravity() = -9.0f; } // slider moves
        if (t > ticks / 2) updateObjectPos(src, 40 * s + 0.25f * (t - ticks / 2), 40 * s, 40 * s);
        volume.update(0.05f);
    }
    double sum = 0, mx = 0;
    for (float v : volume.grid) { sum += v; mx = std::fmax(mx, v); }
    printf("headless: %ux%ux%u, %d ticks, sum(density) = %.4f, max = %.4f\n", volume.n, volume.n, volume.n, ticks, sum, mx);
    return sum > 0 ? 0 : 1;
}

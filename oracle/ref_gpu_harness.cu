// ref_gpu_harness.cu -- the REFERENCE step itself, built headless for sm_100a.  TEST INFRASTRUCTURE.
//
// This translation unit #includes /root/reference/project/smokeSimulation.cu verbatim (path given
// by oracle/Makefile through REF_CU; nothing from the reference is stored in this repository) and
// adds a C ABI around its host entry points and its non-static globals (dev_u/v/w, dev_smoke,
// dev_s, indexNow, tempIndexPast; cu:16-24, 707-708) so tests can
//   * zero the [1] buffers the reference leaves uninitialised,
//   * inject identical initial fields, read every field back after each simulate(),
//   * launch individual reference kernels (per-kernel parity),
//   * time the reference's own kernels on the same B200 ("beat THAT kernel on the same box").
// Built with nvcc's default floating-point flags (fmad on, IEEE div), as the reference's CMake does.
// Output: oracle/_ref/libref_gpu.so (git-ignored; travels to the GPU box with the snapshot).
#include REF_CU

#include <cstring>

namespace {
size_t ref_ncell() { return (size_t)smokeDim.x * smokeDim.y * smokeDim.z; }
size_t ref_nstag() { return (size_t)smokeStaggeredDim.x * smokeStaggeredDim.y * smokeStaggeredDim.z; }

void* ref_field(int field, int which, size_t* bytes)
{
    int b = which == 0 ? indexNow : which == 1 ? tempIndexPast : which - 2;
    switch (field) {
    case 0: *bytes = ref_ncell() * 4; return dev_smoke[b];
    case 1: *bytes = ref_nstag() * 4; return dev_u[b];
    case 2: *bytes = ref_nstag() * 4; return dev_v[b];
    case 3: *bytes = ref_nstag() * 4; return dev_w[b];
    case 4: *bytes = ref_ncell(); return dev_s;
    }
    *bytes = 0;
    return nullptr;
}
dim3 ref_grid() { return dim3(ceil(smokeDim.x / 8.0), ceil(smokeDim.y / 8.0), ceil(smokeDim.z / 8.0)); }
} // namespace

extern "C" {

// initializeVolume (cu:124-238) + zero-fill of the [1] buffers + reset of the process-global
// scene/index state so that several scenes can run in one process.
int refgpu_init(const float* smoke0, unsigned W, unsigned H, unsigned D)
{
    objects.clear();
    currentId = 0;
    indexNow = 1;
    tempIndexPast = 0;
    gravity = -9.82;
    buoyancy_alpha = 2.0f;
    initializeVolume(const_cast<float*>(smoke0), W, H, D);
    cudaMemset(dev_smoke[1], 0, ref_ncell() * 4);
    cudaMemset(dev_u[1], 0, ref_nstag() * 4);
    cudaMemset(dev_v[1], 0, ref_nstag() * 4);
    cudaMemset(dev_w[1], 0, ref_nstag() * 4);
    return (int)cudaDeviceSynchronize();
}
void refgpu_free(void)
{
    deleteVolume();
    cudaFree(dev_obstacles);    // leaked by the reference (cu:240-248)
    cudaFree(dev_smokeSources);
}
int refgpu_add_obstacle(float x, float y, float z, float vx, float vy, float vz, float r) { return addObstacle(x, y, z, vx, vy, vz, r); }
int refgpu_add_source(float x, float y, float z, float r) { return addSmokeSource(x, y, z, r); }
void refgpu_update_object_pos(int id, float x, float y, float z) { updateObjectPos(id, x, y, z); }
void refgpu_set_params(float g, float a) { *getGravity() = g; *getBuoyancy() = a; }

// the reference's public step, host buffer and blocking D2H included (cu:774-819)
void refgpu_simulate(float* smoke_grid, float dt) { simulate(smoke_grid, dt); }

// individual stages on the current indices, for per-kernel parity and per-kernel timing
void refgpu_flip(void)
{
    tempIndexPast = indexNow;
    indexNow = indexNow == 0 ? 1 : 0;
}
void refgpu_fill(void) { drawObjects(); }
void refgpu_integrate(float dt)
{
    integrate<<<ref_grid(), dim3(8, 8, 8)>>>(dev_v[indexNow], dev_smoke[indexNow], dev_s, smokeDim, smokeStaggeredDim, dt, gravity, buoyancy_alpha);
}
void refgpu_clamp(float dt)
{
    velocityConfinement<<<ref_grid(), dim3(8, 8, 8)>>>(dev_u[indexNow], dev_v[indexNow], dev_w[indexNow], smokeDim, smokeStaggeredDim, dt);
}
void refgpu_pressure_halfsweep(int offset)
{
    dim3 g(ceil(ceil(smokeDim.x / 2.0) / 8.0), ceil(smokeDim.y / 8.0), ceil(smokeDim.z / 8.0));
    divergence<<<g, dim3(8, 8, 8)>>>(dev_u[indexNow], dev_v[indexNow], dev_w[indexNow], dev_s, smokeDim, smokeStaggeredDim, (char)offset);
}
void refgpu_advect_velocity(float dt)
{
    dim3 g = ref_grid(), b(8, 8, 8);
    int n = indexNow, p = tempIndexPast;
    velocityAdvectionU<<<g, b>>>(dev_u[n], dev_u[p], dev_v[n], dev_w[n], dev_s, smokeDim, smokeStaggeredDim, dt);
    velocityAdvectionV<<<g, b>>>(dev_v[n], dev_v[p], dev_u[n], dev_w[n], dev_s, smokeDim, smokeStaggeredDim, dt);
    velocityAdvectionW<<<g, b>>>(dev_w[n], dev_w[p], dev_u[n], dev_v[n], dev_s, smokeDim, smokeStaggeredDim, dt);
}
void refgpu_advect_smoke(float dt)
{
    int n = indexNow, p = tempIndexPast;
    advectSmoke<<<ref_grid(), dim3(8, 8, 8)>>>(dev_smoke[n], dev_smoke[p], dev_u[p], dev_v[p], dev_w[p], dev_s, smokeDim, smokeStaggeredDim, dt);
}
int refgpu_sync(void) { return (int)cudaDeviceSynchronize(); }

int refgpu_get_field(int field, int which, void* dst)
{
    size_t n; void* p = ref_field(field, which, &n);
    if (!p) return -1;
    return (int)cudaMemcpy(dst, p, n, cudaMemcpyDeviceToHost);
}
int refgpu_set_field(int field, int which, const void* src)
{
    size_t n; void* p = ref_field(field, which, &n);
    if (!p) return -1;
    return (int)cudaMemcpy(p, src, n, cudaMemcpyHostToDevice);
}
int refgpu_index_now(void) { return indexNow; }

// device-time of `ticks` reference steps without the host round-trip (kernels only), in ms
float refgpu_time_kernels(float dt, int ticks)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int t = 0; t < ticks; t++) {
        refgpu_flip();
        drawObjects();
        refgpu_integrate(dt);
        refgpu_clamp(dt);
        for (int i = 0; i < 30; i++) { refgpu_pressure_halfsweep(0); refgpu_pressure_halfsweep(1); }
        refgpu_advect_velocity(dt);
        refgpu_advect_smoke(dt);
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ms;
}

} // extern "C"

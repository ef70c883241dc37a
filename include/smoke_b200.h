/*
 * smoke_b200.h -- C ABI of the B200-native smoke step (libsmoke_b200.so).
 *
 * This is the drop-in boundary underneath the reference's host entry points
 * (reference: project/smokeSimulation.cuh:4-17).  Plain pointers and sizes only; no CUDA, torch
 * or C++ types.  Every entry point names the reference interface it replaces (file:line under
 * /root/reference, "cu" = project/smokeSimulation.cu).  The C++ wrappers with the reference's
 * exact signatures live in smoke-simulation_b200/host/smokeSimulation.cuh + csrc/dropin.cu and are
 * ~10-line forwards to these functions on one process-global handle.
 *
 * There is no CPU fallback: every compute entry point launches sm_100a kernels and returns
 * SMK_ERR_CUDA if no usable device is present.
 *
 * Array layout at this boundary is the reference's: cell fields x + y*W + z*W*H,
 * staggered fields x + y*(W+1) + z*(W+1)*(H+1), x fastest (cu:146-147).  Internally the
 * staggered fields are stored with a padded row pitch; smk_get_field / smk_set_field convert.
 */
#ifndef SMOKE_B200_H
#define SMOKE_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMK_ABI_VERSION 1

typedef struct smk_sim smk_sim; /* opaque */

/* status codes (the reference has none: it prints and exit(-1)s, cu:131-135; the .cuh wrappers keep
 * that behaviour on top of these) */
enum {
    SMK_OK = 0,
    SMK_ERR_CUDA = 1,      /* a CUDA runtime call failed; see smk_last_error()      */
    SMK_ERR_ARG = 2,       /* bad argument (null handle, id out of range, ...)      */
    SMK_ERR_LIMIT = 3,     /* more objects than SMK_MAX_OBJECTS                     */
    SMK_ERR_REACH = 4,     /* slab mode: backtrace reach exceeded the ghost depth   */
    SMK_ERR_TRANSPORT = 5  /* slab mode: halo transport missing or failed           */
};

/* field ids for smk_get_field / smk_set_field (same numbering as oracle/smoke_oracle.h) */
enum { SMK_FIELD_SMOKE = 0, SMK_FIELD_U = 1, SMK_FIELD_V = 2, SMK_FIELD_W = 3, SMK_FIELD_MASK = 4 };
/* buffer selectors: the reference ping-pongs two buffers per field (cu:16-24, 707-708) */
enum { SMK_BUF_NOW = 0, SMK_BUF_PAST = 1, SMK_BUF_0 = 2, SMK_BUF_1 = 3 };

/* pressure-solver variants.  Only RBGS is the reference's (cu:356-394, 797-801). */
enum {
    SMK_SOLVER_RBGS = 0,       /* red-black SOR on the face velocities, omega 1.9 -- reference parity */
    SMK_SOLVER_JACOBI = 1      /* damped Jacobi on the same velocity form -- extension, no reference  */
};

/* stage ids for smk_stage_time() */
enum {
    SMK_STAGE_FILL = 0, SMK_STAGE_FORCE = 1, SMK_STAGE_PRESSURE = 2, SMK_STAGE_ADVECT_VEL = 3,
    SMK_STAGE_ADVECT_SMOKE = 4, SMK_STAGE_READBACK = 5, SMK_STAGE_COUNT = 6
};

#define SMK_MAX_OBJECTS 16 /* per type; the reference's device arrays hold 3 (cu:56, 221-233) */

/* Threading and device contract (the reference is one thread, one GPU, process-global state, cu:16-58): a handle
 * belongs to the CUDA device that was current in smk_create(); every entry point switches to that device and restores
 * the caller's current device before returning, so several handles on several GPUs may live in one process.  A handle
 * is not thread-safe: serialise calls on the same handle.  Different handles may be used from different threads. */

/* ---- lifetime -------------------------------------------------------------------------------------- */

/* replaces initializeVolume(float* smoke_grid, w, h, d)  (smokeSimulation.cuh:8, cu:124-238).
 * smoke0_host: W*H*D floats or NULL (= zeros).  All second buffers are zero-initialised (the
 * reference leaves them uninitialised).  Mask: fluid everywhere, solid on y == 0. */
int smk_create(smk_sim** out, unsigned W, unsigned H, unsigned D, const float* smoke0_host);

/* one z-slab of a W x H x D grid for multi-GPU runs (no reference counterpart): slab `rank` of
 * `world` owns cell planes [D*rank/world, D*(rank+1)/world) and keeps `ghost` (>= 4) extra planes on each interior
 * side.  smoke0_host_full is the FULL W*H*D initial density (or NULL).  The step then needs a halo transport
 * (smk_set_exchange); results are bit-identical to the single-GPU run. */
int smk_create_slab(smk_sim** out, unsigned W, unsigned H, unsigned D, unsigned rank, unsigned world,
                    unsigned ghost, const float* smoke0_host_full);

/* host-only (no CUDA call): slab geometry, the operation list smk_step executes for `steps` consecutive steps
 * starting from replicated initial fields, and the halo regions of an exchange -- the same code smk_step uses.
 *   out8 = {c0, c1, zlo, zhc, ghost, ok, own_node_lo, own_node_hi}
 *   ops5: 5 ints per op {kind, a, b, p0, p1}; kind 0 flip, 1 fill, 2 force+clamp [a,b), 3 pressure pass (first
 *         half-sweep p0, count p1), 4 advect u,v,w nodes [a,b) (valid input planes [p0,p1]), 5 advect density cells
 *         [a,b), 6 exchange (a = set: 0 u,v,w "now", 1 density "now").  Returns the number of ops.
 *   out5: 5 ints per region {side, send_lo, send_n, recv_lo, recv_n} in global plane indices. */
int smk_slab_geometry(unsigned W, unsigned H, unsigned D, unsigned world, unsigned rank, unsigned ghost, int* out8);
int smk_slab_plan(unsigned W, unsigned H, unsigned D, unsigned world, unsigned rank, unsigned ghost, int iterations,
                  int fuse, int steps, int* ops5, int max_ops);
int smk_slab_regions(unsigned W, unsigned H, unsigned D, unsigned world, unsigned rank, unsigned ghost, int set,
                     int* out5, int max_regions);

/* host-only (no CUDA call): how a fused pressure pass of K (2 or 4) half-sweeps over the node planes [out_lo, out_hi)
 * of a W x H grid is cut into pieces for `nctas` CTAs of equal cost (csrc/pass_schedule.h; no reference counterpart --
 * the reference launches one thread per cell, cu:795-801).  pieces4: {tile x, tile y, zo0, zo1} per piece; first:
 * CTA b works through pieces [first[b], first[b+1]); info4 = {tiles in x, tiles in y, z-steps of the busiest CTA, CTAs used}.
 * Returns the number of pieces. */
int smk_pass_schedule(unsigned W, unsigned H, int out_lo, int out_hi, int K, int nctas, int* pieces4, int max_pieces,
                      int* first, int max_first, int* info4);

/* replaces deleteVolume()  (smokeSimulation.cuh:9, cu:240-248) */
int smk_destroy(smk_sim* s);

/* replaces getGPUProperties()  (smokeSimulation.cuh:3, cu:62-85): prints the same property lines */
int smk_print_gpu_properties(void);

/* ---- scene and parameters -------------------------------------------------------------------------- */

/* replace addObstacle / addSmokeSource / updateObjectPos  (smokeSimulation.cuh:12-14, cu:88-109).
 * Return the new object's id (dense, creation order, shared between both types) or -SMK_ERR_*. */
int smk_add_obstacle(smk_sim* s, float x, float y, float z, float vx, float vy, float vz, float r);
int smk_add_source(smk_sim* s, float x, float y, float z, float r);
int smk_update_object_pos(smk_sim* s, int id, float x, float y, float z);

/* Extension.  The reference rewrites the mask once per obstacle, so where spheres overlap or follow one
 * another the LAST obstacle decides (cu:304-310: a cell inside obstacle 0 but outside obstacle 1 ends up fluid).
 * SMK_OBSTACLES_UNION makes a cell solid when it lies inside ANY obstacle.  Arbitrary voxelised solids: upload a mask
 * with smk_set_field(SMK_FIELD_MASK, ...) and add no sphere obstacle -- with an empty obstacle list the reference (and
 * this library) never touches the mask (cu:289-313), so it persists.  Up to SMK_MAX_OBJECTS spheres per type. */
enum { SMK_OBSTACLES_LAST_WINS = 0, SMK_OBSTACLES_UNION = 1 };
int smk_set_obstacle_mode(smk_sim* s, int mode);

/* replace getGravity() / getBuoyancy()  (smokeSimulation.cuh:16-17, cu:31-36): stable pointers to
 * library-owned floats that the caller may write between steps; re-read at every step. */
float* smk_gravity_ptr(smk_sim* s);
float* smk_buoyancy_ptr(smk_sim* s);

/* solver selection (extension; defaults = the reference: RBGS, 30 iterations, cu:797).
 * fuse = number of half-sweeps fused per kernel launch (temporal blocking); 0 = library default,
 * 1 = one launch per half-sweep.  Results are identical for every fuse value. */
int smk_set_solver(smk_sim* s, int variant, int iterations, int fuse);
/* Which kernel runs the fused red-black passes (same bits either way; tests run both):
 *   SMK_PASS_AUTO  TMA-staged kernel (kernels_pressure_tma.cuh) on grids of >= 5 M nodes per slab, the round-1 kernel
 *                  (kernels_pressure_reg.cuh) below -- small grids are launch-bound and its prologue is shorter;
 *   SMK_PASS_REG / SMK_PASS_TMA  force one.  Process-wide default: SMK_PASS_KERNEL=reg|tma in the environment. */
enum { SMK_PASS_AUTO = 0, SMK_PASS_REG = 1, SMK_PASS_TMA = 2 };
int smk_set_pass_kernel(smk_sim* s, int kind);
/* SMK_PASS_REG / SMK_PASS_TMA: the kernel the most recent fused pass ran on (0: none yet) -- bench.py names it */
int smk_last_pass_kernel(smk_sim* s);
/* scheduling of the fused pressure passes (no effect on results): 0 = default, a (tile, z-chunk) grid of CTAs;
 * nctas > 0 = that many CTAs working through balanced piece lists (csrc/pass_schedule.h; an experiment kept for
 * ablation: measured no faster). */
int smk_set_pass_ctas(smk_sim* s, int nctas);
/* CTAs of the most recent pressure pass if it ran on balanced piece lists, 0 if it ran as a (tile, z-chunk) grid */
int smk_last_pass_ctas(smk_sim* s);

/* sparse blocking readback (extension, single GPU, off by default; SMK_READBACK_BOX=1 turns it on for every simulation):
 * smk_step(sim, dt, host) copies only the rows that can hold non-zero density -- the rows the previous readback into
 * the SAME buffer found non-zero, grown by two cells -- and falls back to the full copy when the new density is not
 * covered (checked on the device every step).  The buffer ends up identical to the reference's full copy (cu:814) as long
 * as the caller does not write to it between steps.  smk_readback_bytes() counts what was actually moved. */
int smk_set_readback_box(smk_sim* s, int on);

/* ---- the step ---------------------------------------------------------------------------------------- */

/* replaces simulate(float* smoke_grid, float dt)  (smokeSimulation.cuh:5, cu:774-819).
 * Runs flip, source/mask fill, forcing, clamp, the pressure sweeps, velocity and density advection.
 * density_host: W*H*D floats that receive the new density (blocking, like cu:814), or NULL to skip
 * the device->host round trip (the result stays on the device; smk_density_device()). */
int smk_step(smk_sim* s, float dt, float* density_host);

/* Host-buffer contract of smk_step / smk_step_async / smk_read_density_half.  The reference copies into a pageable
 * std::vector with a blocking cudaMemcpy (cu:814; boundingBox.h:41).  To run that copy at PCIe speed and overlapped with
 * the last kernels, the library PAGE-LOCKS the caller's range (cudaHostRegister) the first time it sees it and keeps it
 * locked until smk_unregister_host() or smk_destroy().  A caller that frees or reallocates the buffer while the
 * handle lives must call smk_unregister_host() first (the drop-in wrappers do: deleteVolume destroys the handle).
 * A cached registration is re-validated against the driver and against the requested size on every use; buffers the
 * caller pinned itself (cudaHostAlloc, torch pin_memory) are used as they are.  SMK_NO_HOST_REGISTER=1 in the environment
 * disables implicit registration (the copy then goes through the driver's pageable path).
 * smk_register_host() makes the registration explicit (e.g. once, right after allocating the buffer). */
int smk_register_host(smk_sim* s, void* host, size_t bytes);
int smk_unregister_host(smk_sim* s, void* host);

/* same, without waiting: returns after enqueueing; smk_sync() waits.  density_host may be NULL.  With a host buffer
 * the readback is PIPELINED: the new density is snapshotted on the device and copied to the host on a second
 * stream while the next step computes (the buffer holds step n's density once step n+1 has been enqueued and
 * smk_sync() returned, or after the next smk_step_async has waited for it). */
int smk_step_async(smk_sim* s, float dt, float* density_host);
int smk_sync(smk_sim* s);

/* device pointer to the density produced by the last step, reference layout (W*H*D floats).
 * Replaces the D2H + glTexSubImage3D round trip of cu:814 / boundingBox.cpp:380-385. */
const float* smk_density_device(smk_sim* s);

/* Extension (opt-in, never used on a parity path): the density of the last step as IEEE binary16
 * (converted on the device, round to nearest even) into host_half (W*H*D 16-bit values; a slab writes its owned
 * planes at their global offset).  Half the device->host bytes of smk_step's float readback.  Blocking. */
int smk_read_density_half(smk_sim* s, void* host_half);

/* Copy the density of the last step into a 3-D CUDA array (cudaArray_t passed as void*; W x H x D, one float
 * channel) on the step's stream -- e.g. the array of the renderer's GL_R32F texture mapped through CUDA-GL interop.
 * Together with simulate(nullptr, dt) this removes the D2H + glTexSubImage3D round trip. */
int smk_copy_density_to_array(smk_sim* s, void* cuda_array);
/* SURVEY N1 as specified: bind the array once (NULL unbinds); from then on the density advection kernel writes every new
 * density value into it with surf3Dwrite -- the whole W x H x D volume equals the "past" density buffer after each step,
 * with no device-to-device pass and no host round trip (replaces cu:814 + boundingBox.cpp:380-385).  The array must allow
 * surface load/store (cudaArraySurfaceLoadStore / a GL texture registered with cudaGraphicsRegisterFlagsSurfaceLoadStore). */
int smk_bind_density_array(smk_sim* s, void* cuda_array);
/* the solid mask as ONE BIT per cell (bit i of byte k = cell 8k + i in cell order, 1 = fluid;
 * (W*H*D + 7) / 8 bytes).  smk_set_mask_bits has the semantics of smk_set_field(SMK_FIELD_MASK, ...) of the unpacked mask. */
int smk_set_mask_bits(smk_sim* s, const unsigned char* bits);
int smk_get_mask_bits(smk_sim* s, unsigned char* bits);
/* test helpers: a plain 3-D float array standing in for the mapped texture */
void* smk_test_array_create(unsigned W, unsigned H, unsigned D);
int smk_test_array_read(void* cuda_array, float* host, unsigned W, unsigned H, unsigned D);
void smk_test_array_destroy(void* cuda_array);

/* run the whole step on a caller-owned CUDA stream (cudaStream_t passed as void*); NULL = own stream */
int smk_set_stream(smk_sim* s, void* cuda_stream);

/* ---- stage-level entry points (per-kernel parity tests; same order as cu:777-810) ------------------- */
int smk_stage_flip(smk_sim* s);                         /* cu:777-779 */
int smk_stage_fill(smk_sim* s);                         /* cu:714-771 */
int smk_stage_force_clamp(smk_sim* s, float dt);        /* cu:789 + cu:791 */
int smk_stage_pressure_halfsweep(smk_sim* s, int offset); /* cu:799-800, one launch */
int smk_stage_pressure(smk_sim* s);                     /* cu:797-801, all iterations, fused as configured */
int smk_stage_advect_velocity(smk_sim* s, float dt);    /* cu:805-807 */
int smk_stage_advect_smoke(smk_sim* s, float dt);       /* cu:810 */

/* ---- field access (tests, checkpoints) --------------------------------------------------------------- */
/* copy a whole field in the reference layout.  In slab mode only the planes this slab stores are
 * touched (host arrays are still full-size). */
int smk_get_field(smk_sim* s, int field, int which, void* host_dst);
int smk_set_field(smk_sim* s, int field, int which, const void* host_src);
int smk_index_now(smk_sim* s);

/* max |div| over interior fluid cells of the "now" velocities (formula cu:379-381), by a device
 * reduction (warp shuffles + one atomic per block).  The reference computes no residual. */
int smk_max_divergence(smk_sim* s, float* out);

/* 64-bit content hashes of the planes this slab OWNS, computed on the device: out7 = {u, v, w "now", u, v, w "past",
 * density "past"}.  Every element contributes a mix of its bit pattern and its GLOBAL reference-layout index, summed
 * modulo 2^64, so the sum over all slabs equals the hash of the single-GPU run iff the fields are bit-identical
 * (-0 counted as +0).  No reference counterpart (the reference is single-GPU and has no checks). */
int smk_hash_owned(smk_sim* s, unsigned long long* out7);
/* same over the node planes [node_lo, node_hi) and cell planes [cell_lo, cell_hi) of a handle that stores them (to
 * hash the slab-shaped parts of a single-GPU run) */
int smk_hash_range(smk_sim* s, int node_lo, int node_hi, int cell_lo, int cell_hi, unsigned long long* out7);

/* ---- measurement -------------------------------------------------------------------------------------- */
/* accumulated device time (CUDA events on the step's stream) and launch count per stage since the
 * last smk_reset_timers().  Events are always recorded; reading them synchronises. */
int smk_stage_time(smk_sim* s, int stage, double* ms_total, long* launches);
int smk_reset_timers(smk_sim* s);
/* kernels launched by this handle since creation */
long smk_launch_count(smk_sim* s);
/* development aid (SMK_PASS_DEBUG=1 in the environment at smk_create): {start clock, cycles, SM id, variant * 1000 + planes}
 * of every CTA of the most recent fused pressure pass; returns the number of CTAs written (0: not enabled) */
int smk_debug_pass_ctas(smk_sim* s, long long* out4, int max_ctas);
/* bytes the steps of this handle have copied device -> host so far (the density readback of cu:814; counted where the
 * copies are enqueued, so bench.py reports measured, not assumed, bytes per step) */
unsigned long long smk_readback_bytes(smk_sim* s);
/* Self-check of the pressure pass's binary32 evaluation of the over-relaxation product `(float)((double)q * -1.9)`
 * (cu:384) against the double-precision sequence, on the device, for the `count` binary32 bit patterns from `first`
 * (count = 2^32: all of them, ~1 s).  *mismatches must come back 0; *ties (may be NULL) = inputs resolved by the
 * tie-to-even rule (kernels_pressure_tma.cuh, p_from_q; CPU twin: tools/experiments/omega_fp32_exhaustive.c). */
int smk_selfcheck_omega(int device, unsigned first, unsigned long long count, unsigned long long* mismatches, unsigned long long* ties);

/* ---- multi-GPU halo transport (slab mode) --------------------------------------------------------------- */
/* Caller-provided exchange: called from smk_step when ghost planes of a field set must be refreshed.
 * set: 0 = u,v,w "now" (three regions per neighbour, in the order u, v, w), 1 = density "now".  For every listed
 * region the callee must send `send_ptr` (device, contiguous, `send_bytes`) to the neighbour on `side` (0 = lower z,
 * 1 = upper z) and receive the neighbour's matching planes (`recv_bytes`) into `recv_ptr`.  All work must be ordered
 * on `cuda_stream`. */
typedef struct {
    int side;          /* 0 = lower-z neighbour, 1 = upper-z neighbour */
    void* send_ptr;    /* my owned boundary planes */
    void* recv_ptr;    /* my ghost planes          */
    size_t send_bytes; /* (u,v,w have one more upper-ghost plane than they send up: sizes differ per direction) */
    size_t recv_bytes;
} smk_halo_region;
typedef int (*smk_exchange_fn)(void* ctx, int set, const smk_halo_region* regions, int nregions, void* cuda_stream);
int smk_set_exchange(smk_sim* s, smk_exchange_fn fn, void* ctx);

/* Peer-memory halo path (the B200-native transport): every slab's exchanged fields live in one allocation ("arena").
 * A neighbour process maps it through a 64-byte CUDA IPC handle (smk_p2p_export -> smk_p2p_attach_ipc); slabs that
 * live in one process pass the arena pointer (smk_p2p_arena -> smk_p2p_attach_ptr).  Once every existing neighbour is
 * attached, smk_step needs no transport callback: the fused pressure passes read the neighbours' boundary planes
 * directly over NVLink inside the kernel (one epoch handshake per pass, no ghost copies), and the remaining halo
 * refreshes (before advection) are pulls over the mapped memory.  side: 0 = lower-z neighbour, 1 = upper-z. */
int smk_p2p_export(smk_sim* s, unsigned char* handle64);
int smk_p2p_attach_ipc(smk_sim* s, int side, const unsigned char* handle64);
void* smk_p2p_arena(smk_sim* s);
int smk_p2p_attach_ptr(smk_sim* s, int side, void* peer_arena);
long smk_exchange_count(smk_sim* s);
/* test helper for several slabs driven by ONE host thread on ONE GPU: publish the next epoch without waiting */
int smk_p2p_presignal(smk_sim* s);
/* like smk_slab_plan, for the peer-memory path (pressure passes consume no ghost depth) */
int smk_slab_plan_p2p(unsigned W, unsigned H, unsigned D, unsigned world, unsigned rank, unsigned ghost, int iterations,
                      int fuse, int steps, int* ops5, int max_ops);

/* execute ONE operation of a plan returned by smk_slab_plan (op5 = {kind, a, b, p0, p1}); an exchange op calls the
 * transport.  smk_step == every op of plan_step in order.  Used by tests to drive several slabs in lock-step. */
int smk_exec_op(smk_sim* s, const int* op5, float dt);

const char* smk_last_error(smk_sim* s);
int smk_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif

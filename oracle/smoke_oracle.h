/*
 * smoke_oracle.h -- CPU restatement of the reference smoke step.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the parity oracle for the hot path of RasmusAlmryd/smoke-simulation
 * (project/smokeSimulation.cu, "cu" below).  It is NOT part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product library (libsmoke_b200.so) never links, loads or calls anything in oracle/.
 *
 * Pinning: the reference ships no tests or golden vectors, so this
 * restatement is pinned against the reference's OWN kernel bodies compiled as host code
 * (oracle/_ref/libref_cpu.so, built by oracle/Makefile from /root/reference in place) --
 * bit-exact in contract=0 mode -- and against fixtures generated from that build
 * (tests/golden/, generator: oracle/make_golden.py).  On the GPU box it is additionally compared
 * with the reference step itself built headless for sm_100a (oracle/_ref/libref_gpu.so).
 *
 * Arrays use the reference layout: cell index x + y*W + z*W*H, staggered index
 * x + y*(W+1) + z*(W+1)*(H+1), x fastest (cu:146-147).
 */
#ifndef SMOKE_ORACLE_H
#define SMOKE_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct smk_oracle smk_oracle;

/* field ids shared with include/smoke_b200.h */
enum {
    ORC_FIELD_SMOKE = 0, /* W*H*D floats            */
    ORC_FIELD_U = 1,     /* (W+1)(H+1)(D+1) floats  */
    ORC_FIELD_V = 2,
    ORC_FIELD_W = 3,
    ORC_FIELD_MASK = 4   /* W*H*D bytes, 1 = fluid  */
};

/* contract = 0: every a*b+c is two roundings (what g++ -O2 makes of the reference bodies).
 * contract = 1: mirror the FMA contraction nvcc applies to the reference kernels for sm_100a
 *               (read off the SASS; the exact pattern is documented at each function). */
smk_oracle* orc_create(unsigned W, unsigned H, unsigned D, const float* smoke0, int contract);
void orc_destroy(smk_oracle* o);

void orc_set_threads(int n);          /* OpenMP threads used by the loops (0 = default)      */
int orc_get_threads(void);

int orc_add_obstacle(smk_oracle* o, float x, float y, float z, float vx, float vy, float vz, float r);
int orc_add_source(smk_oracle* o, float x, float y, float z, float r);
void orc_update_object_pos(smk_oracle* o, int id, float x, float y, float z);
void orc_set_params(smk_oracle* o, float gravity, float buoyancy_alpha);
void orc_set_iterations(smk_oracle* o, int iterations); /* reference: 30 (cu:797)            */
void orc_set_solver(smk_oracle* o, int solver);         /* 0 = RBGS (reference), 1 = damped Jacobi (extension) */
void orc_jacobi_iteration(smk_oracle* o, float* p_scratch); /* extension, see smoke_oracle.c */
void orc_set_obstacle_mode(smk_oracle* o, int union_mode); /* 0 = last obstacle decides (reference), 1 = union (extension) */

/* one full step, cu:774-819 */
void orc_step(smk_oracle* o, float dt);

/* individual stages on the CURRENT "now" buffers (for per-kernel parity tests).
 * orc_flip() performs the index swap simulate() starts with (cu:777-779). */
void orc_flip(smk_oracle* o);
void orc_fill(smk_oracle* o);                  /* cu:714-771 */
void orc_integrate(smk_oracle* o, float dt);   /* cu:315-329 */
void orc_clamp(smk_oracle* o, float dt);       /* cu:331-352 */
void orc_pressure_halfsweep(smk_oracle* o, int offset); /* cu:356-394 */
void orc_advect_velocity(smk_oracle* o, float dt);      /* cu:527-615 */
void orc_advect_smoke(smk_oracle* o, float dt);         /* cu:617-638 */

/* the same stages restricted to a plane range [za, zb) (nodes for integrate / clamp / advect_velocity, cells for
 * the others): used by tests/test_slab_cpu.py to execute the multi-GPU z-slab schedule on the CPU */
void orc_integrate_r(smk_oracle* o, float dt, int za, int zb);
void orc_clamp_r(smk_oracle* o, float dt, int za, int zb);
void orc_pressure_halfsweep_r(smk_oracle* o, int offset, int za, int zb);
void orc_advect_velocity_r(smk_oracle* o, float dt, int za, int zb);
void orc_advect_smoke_r(smk_oracle* o, float dt, int za, int zb);

/* which: 0 = buffer indexNow, 1 = buffer tempIndexPast, 2/3 = physical buffer 0/1 */
void orc_get_field(const smk_oracle* o, int field, int which, void* dst);
void orc_set_field(smk_oracle* o, int field, int which, const void* src);
int orc_index_now(const smk_oracle* o);

/* max |div| over interior fluid cells of the "now" velocities, formula cu:379-381 */
float orc_max_divergence(const smk_oracle* o);

#ifdef __cplusplus
}
#endif
#endif

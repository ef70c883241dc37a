/*
 * smoke_oracle.c -- CPU restatement of the reference smoke step.  TEST INFRASTRUCTURE ONLY
 * (see smoke_oracle.h for who may use it and how it is pinned).
 *
 * Every function cites the lines of /root/reference/project/smokeSimulation.cu ("cu") whose
 * arithmetic it restates.  Written as plain loops over cells / faces; no code is shared with
 * the reference.  Build: gcc -O2 -fopenmp -mfma -ffp-contract=off (the compiler must not
 * contract on its own -- contraction is explicit through the fm() helper below).
 */
#include "smoke_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_OBJECTS 64

typedef struct {
    int type; /* 0 obstacle, 1 source (cu:45-52) */
    float x, y, z, vx, vy, vz, r;
} orc_sphere;

struct smk_oracle {
    int W, H, D;    /* cells */
    int SX, SY, SZ; /* staggered dims = cells + 1 (cu:147) */
    size_t ncell, nstag;
    float* smoke[2];
    float* u[2];
    float* v[2];
    float* w[2];
    unsigned char* s; /* 1 = fluid, 0 = solid (cu:24, 193-207) */
    int now, past;    /* indexNow = 1, tempIndexPast = 0 at start (cu:707-708) */
    float gravity, alpha;
    int iterations;
    int solver; /* 0 = red-black SOR (reference), 1 = damped Jacobi (extension) */
    int obstacle_union; /* 0 = the last obstacle decides (reference, cu:304-310), 1 = solid inside ANY obstacle (extension) */
    int contract;
    int nobj;
    orc_sphere obj[ORC_MAX_OBJECTS];
};

static int g_threads = 0;

void orc_set_threads(int n)
{
    g_threads = n;
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
}

int orc_get_threads(void)
{
#ifdef _OPENMP
    return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
    return 1;
#endif
}

/* a*b + c with either one rounding (contract) or two */
static inline float fm(int contract, float a, float b, float c)
{
    if (contract) return fmaf(a, b, c);
    float p = a * b;
    return p + c;
}

static inline size_t cidx(const smk_oracle* o, int x, int y, int z)
{
    return (size_t)x + (size_t)y * o->W + (size_t)z * o->W * o->H;
}
static inline size_t sidx(const smk_oracle* o, int x, int y, int z)
{
    return (size_t)x + (size_t)y * o->SX + (size_t)z * o->SX * o->SY;
}

/* cu:124-238: buffers, mask = fluid everywhere except the plane y == 0 (cu:200-207).
 * Semantics for the buffers the reference leaves uninitialised (index 1): zero. */
smk_oracle* orc_create(unsigned W, unsigned H, unsigned D, const float* smoke0, int contract)
{
    smk_oracle* o = (smk_oracle*)calloc(1, sizeof(*o));
    if (!o) return NULL;
    o->W = (int)W; o->H = (int)H; o->D = (int)D;
    o->SX = o->W + 1; o->SY = o->H + 1; o->SZ = o->D + 1;
    o->ncell = (size_t)W * H * D;
    o->nstag = (size_t)o->SX * o->SY * o->SZ;
    for (int i = 0; i < 2; i++) {
        o->smoke[i] = (float*)calloc(o->ncell, sizeof(float));
        o->u[i] = (float*)calloc(o->nstag, sizeof(float));
        o->v[i] = (float*)calloc(o->nstag, sizeof(float));
        o->w[i] = (float*)calloc(o->nstag, sizeof(float));
    }
    if (smoke0) memcpy(o->smoke[0], smoke0, o->ncell * sizeof(float));
    o->s = (unsigned char*)malloc(o->ncell);
    memset(o->s, 1, o->ncell);
    for (int z = 0; z < o->D; z++)
        for (int x = 0; x < o->W; x++) o->s[cidx(o, x, 0, z)] = 0;
    o->now = 1; o->past = 0;
    o->gravity = -9.82f; /* cu:28 */
    o->alpha = 2.0f;     /* cu:29 */
    o->iterations = 30;  /* cu:797 */
    o->contract = contract;
    return o;
}

void orc_destroy(smk_oracle* o)
{
    if (!o) return;
    for (int i = 0; i < 2; i++) { free(o->smoke[i]); free(o->u[i]); free(o->v[i]); free(o->w[i]); }
    free(o->s);
    free(o);
}

/* cu:88-109: ids are dense, in creation order, shared between both object types */
int orc_add_obstacle(smk_oracle* o, float x, float y, float z, float vx, float vy, float vz, float r)
{
    if (o->nobj >= ORC_MAX_OBJECTS) return -1;
    orc_sphere sp = {0, x, y, z, vx, vy, vz, r};
    o->obj[o->nobj] = sp;
    return o->nobj++;
}
int orc_add_source(smk_oracle* o, float x, float y, float z, float r)
{
    if (o->nobj >= ORC_MAX_OBJECTS) return -1;
    orc_sphere sp = {1, x, y, z, 0.f, 0.f, 0.f, r};
    o->obj[o->nobj] = sp;
    return o->nobj++;
}
void orc_update_object_pos(smk_oracle* o, int id, float x, float y, float z)
{
    o->obj[id].x = x; o->obj[id].y = y; o->obj[id].z = z;
}
void orc_set_params(smk_oracle* o, float gravity, float buoyancy_alpha)
{
    o->gravity = gravity; o->alpha = buoyancy_alpha;
}
void orc_set_iterations(smk_oracle* o, int iterations) { o->iterations = iterations; }
void orc_set_solver(smk_oracle* o, int solver) { o->solver = solver; }
void orc_set_obstacle_mode(smk_oracle* o, int union_mode) { o->obstacle_union = union_mode; }
int orc_index_now(const smk_oracle* o) { return o->now; }

void orc_flip(smk_oracle* o) /* cu:777-779 */
{
    o->past = o->now;
    o->now = o->now ? 0 : 1;
}

/* squared distance exactly as cu:265 / cu:303: three powf(d, 2) summed left to right, the
 * cell coordinate converted int -> float before the subtraction */
static inline float sphere_dist(int x, int y, int z, const orc_sphere* sp)
{
    float dx = powf((float)x - sp->x, 2.f);
    float dy = powf((float)y - sp->y, 2.f);
    float dz = powf((float)z - sp->z, 2.f);
    return dx + dy + dz;
}

/* cu:714-771 (drawObjects) = fillSmoke cu:251-273 then fillObstacle cu:289-313.
 * Sources: interior cells inside any source sphere get density 1 in BOTH buffers.
 * Obstacles: every interior cell is rewritten per obstacle, so the LAST obstacle decides
 * (cu:304-310); with no obstacle the mask is left untouched. */
void orc_fill(smk_oracle* o)
{
    const int W = o->W, H = o->H, D = o->D;
#pragma omp parallel for schedule(static)
    for (int z = 1; z < D - 1; z++)
        for (int y = 1; y < H - 1; y++)
            for (int x = 1; x < W - 1; x++) {
                size_t c = cidx(o, x, y, z);
                int have_obstacle = 0;
                unsigned char sval = 1;
                for (int i = 0; i < o->nobj; i++) {
                    const orc_sphere* sp = &o->obj[i];
                    float dist = sphere_dist(x, y, z, sp);
                    int inside = dist < sp->r * sp->r;
                    if (sp->type == 1) {
                        if (inside) { o->smoke[0][c] = 1.0f; o->smoke[1][c] = 1.0f; }
                    } else {
                        have_obstacle = 1;
                        sval = inside ? 0 : (o->obstacle_union ? sval : 1); /* reference: the last obstacle decides */
                    }
                }
                if (have_obstacle) o->s[c] = sval;
            }
}

/* cu:315-329.  v-face (x,y,z), x in [0,W), y in [1,H), z in [0,D), both cells fluid:
 *   v += smoke*gravity*dt + (alpha*smoke)*dt        (smoke of the upper cell)
 * nvcc (SASS): t = fma(smoke*gravity, dt, (smoke*alpha)*dt); v = t + v. */
void orc_integrate(smk_oracle* o, float dt) { orc_integrate_r(o, dt, 0, o->D); }

/* the same on node planes [za, zb) only (z-slab emulation in tests/test_slab_cpu.py) */
void orc_integrate_r(smk_oracle* o, float dt, int za, int zb)
{
    const int W = o->W, H = o->H, D = o->D, ct = o->contract;
    if (za < 0) za = 0;
    if (zb > D) zb = D;
    float* v = o->v[o->now];
    const float* smoke = o->smoke[o->now];
    const float g = o->gravity, alpha = o->alpha;
#pragma omp parallel for schedule(static)
    for (int z = za; z < zb; z++)
        for (int y = 1; y < H; y++)
            for (int x = 0; x < W; x++) {
                if (!o->s[cidx(o, x, y, z)] || !o->s[cidx(o, x, y - 1, z)]) continue;
                float d = smoke[cidx(o, x, y, z)];
                float buoy = (alpha * d) * dt;
                float t = fm(ct, d * g, dt, buoy);
                size_t f = sidx(o, x, y, z);
                v[f] = v[f] + t;
            }
}

/* cu:331-352 (the max-velocity clamp).  Same staggered index for the three components,
 * x,y,z in [1,W) x [1,H) x [1,D):  L = u^2+v^2+w^2;  if L*dt > 9: scale by 9/(L*dt).
 * nvcc (SASS): L = fma(w,w, fma(u,u, v*v)); one IEEE divide shared by the three products. */
void orc_clamp(smk_oracle* o, float dt) { orc_clamp_r(o, dt, 0, o->D); }

void orc_clamp_r(smk_oracle* o, float dt, int za, int zb)
{
    const int W = o->W, H = o->H, D = o->D, ct = o->contract;
    float *u = o->u[o->now], *v = o->v[o->now], *w = o->w[o->now];
    if (za < 1) za = 1;
    if (zb > D) zb = D;
#pragma omp parallel for schedule(static)
    for (int z = za; z < zb; z++)
        for (int y = 1; y < H; y++)
            for (int x = 1; x < W; x++) {
                size_t f = sidx(o, x, y, z);
                float a = u[f], b = v[f], c = w[f];
                float L;
                if (ct) L = fmaf(c, c, fmaf(a, a, b * b));
                else { float aa = a * a, bb = b * b, cc = c * c; L = (aa + bb) + cc; }
                float t = L * dt;
                if (t > 9.0f) {
                    float k = 9.0f / t;
                    u[f] = a * k; v[f] = b * k; w[f] = c * k;
                }
            }
}

/* cu:356-394.  One colour of SOR Gauss-Seidel directly on the face velocities.
 * offset 0 <-> (x+y+z) even, offset 1 <-> odd (x = 2x' - (y+z+offset)%2, cu:362).
 * Interior fluid cells with at least one fluid neighbour:
 *   div = ((((-u0 + u1) - v0) + v1) - w0) + w1          (cu:379-381, left to right)
 *   p   = (float)((double)(-div / (float)acc) * 1.9)     (1.9 is a double literal, cu:38/384)
 *   u0 -= p*sx0; u1 += p*sx1; v0 -= p*sy0; v1 += p*sy1; w0 -= p*sz0; w1 += p*sz1
 * p*s is exact (s in {0,1}), so contraction does not change the result here.
 * Same-colour cells share no face, so the loop is race-free in any order. */
void orc_pressure_halfsweep(smk_oracle* o, int offset) { orc_pressure_halfsweep_r(o, offset, 0, o->D); }

/* the same on cell planes [za, zb) only */
void orc_pressure_halfsweep_r(smk_oracle* o, int offset, int za, int zb)
{
    const int W = o->W, H = o->H, D = o->D;
    float *u = o->u[o->now], *v = o->v[o->now], *w = o->w[o->now];
    const unsigned char* s = o->s;
    if (za < 1) za = 1;
    if (zb > D - 1) zb = D - 1;
#pragma omp parallel for schedule(static)
    for (int z = za; z < zb; z++)
        for (int y = 1; y < H - 1; y++) {
            int x = 1 + ((1 + y + z + offset) & 1); /* first interior x with (x+y+z+offset) even */
            for (; x < W - 1; x += 2) {
                if (!s[cidx(o, x, y, z)]) continue;
                int sx0 = s[cidx(o, x - 1, y, z)], sx1 = s[cidx(o, x + 1, y, z)];
                int sy0 = s[cidx(o, x, y - 1, z)], sy1 = s[cidx(o, x, y + 1, z)];
                int sz0 = s[cidx(o, x, y, z - 1)], sz1 = s[cidx(o, x, y, z + 1)];
                int acc = sx0 + sx1 + sy0 + sy1 + sz0 + sz1;
                if (acc == 0) continue;
                size_t iu0 = sidx(o, x, y, z), iu1 = sidx(o, x + 1, y, z);
                size_t iv1 = sidx(o, x, y + 1, z), iw1 = sidx(o, x, y, z + 1);
                float div = -u[iu0] + u[iu1];
                div = div + -v[iu0];
                div = div + v[iv1];
                div = div + -w[iu0];
                div = div + w[iw1];
                float q = -div / (float)acc;
                float p = (float)((double)q * 1.9);
                u[iu0] = u[iu0] - p * (float)sx0;
                u[iu1] = u[iu1] + p * (float)sx1;
                v[iu0] = v[iu0] - p * (float)sy0;
                v[iv1] = v[iv1] + p * (float)sy1;
                w[iu0] = w[iu0] - p * (float)sz0;
                w[iw1] = w[iw1] + p * (float)sz1;
            }
        }
}

/* EXTENSION without a reference counterpart (SURVEY point 1 / H9; BASELINE configs[1] "Jacobi"): one damped-Jacobi
 * iteration on the same velocity form.  Every interior fluid cell with a fluid neighbour computes, from the OLD fields,
 *   p_c = (float)((double)(-div_c / (float)acc_c) * (2.0/3.0))          (div and acc exactly as in the half-sweep)
 * and every face then receives the correction of BOTH adjacent cells, low-side cell first:
 *   u[x] = (u[x] - p_c * s(x-1)) + p_(x-1) * s_c      (c = cell x; likewise v, w)
 * p*s is exact, so the two roundings are the two additions.  Weight 2/3 damps the checkerboard mode (plain
 * simultaneous application leaves it undamped).  `p` is a caller-provided scratch of W*H*D floats. */
void orc_jacobi_iteration(smk_oracle* o, float* p)
{
    const int W = o->W, H = o->H, D = o->D;
    float *u = o->u[o->now], *v = o->v[o->now], *w = o->w[o->now];
    const unsigned char* s = o->s;
#pragma omp parallel for schedule(static)
    for (int z = 0; z < D; z++)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                size_t c = cidx(o, x, y, z);
                p[c] = 0.f;
                if (x < 1 || y < 1 || z < 1 || x >= W - 1 || y >= H - 1 || z >= D - 1 || !s[c]) continue;
                int acc = s[cidx(o, x - 1, y, z)] + s[cidx(o, x + 1, y, z)] + s[cidx(o, x, y - 1, z)] + s[cidx(o, x, y + 1, z)] +
                          s[cidx(o, x, y, z - 1)] + s[cidx(o, x, y, z + 1)];
                if (acc == 0) continue;
                size_t f = sidx(o, x, y, z);
                float div = -u[f] + u[sidx(o, x + 1, y, z)];
                div = div + -v[f];
                div = div + v[sidx(o, x, y + 1, z)];
                div = div + -w[f];
                div = div + w[sidx(o, x, y, z + 1)];
                float q = -div / (float)acc;
                p[c] = (float)((double)q * (2.0 / 3.0));
            }
#pragma omp parallel for schedule(static)
    for (int z = 0; z < D; z++)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                size_t c = cidx(o, x, y, z), f = sidx(o, x, y, z);
                float sc = (float)s[c], pc = p[c];
                if (x > 0) { float t = u[f] - pc * (float)s[cidx(o, x - 1, y, z)]; u[f] = t + p[cidx(o, x - 1, y, z)] * sc; }
                if (y > 0) { float t = v[f] - pc * (float)s[cidx(o, x, y - 1, z)]; v[f] = t + p[cidx(o, x, y - 1, z)] * sc; }
                if (z > 0) { float t = w[f] - pc * (float)s[cidx(o, x, y, z - 1)]; w[f] = t + p[cidx(o, x, y, z - 1)] * sc; }
            }
}

/* cu:409-447.  The 8-point face sums in source order, then /8.  They reach one plane
 * down in z (and y-1 / y+1, x-1 / x+1 as listed) -- not the textbook 4-point averages. */
static inline float avg_u(const smk_oracle* o, const float* f, int x, int y, int z)
{
    float a = f[sidx(o, x, y, z - 1)];
    a = a + f[sidx(o, x + 1, y, z - 1)];
    a = a + f[sidx(o, x, y - 1, z - 1)];
    a = a + f[sidx(o, x + 1, y - 1, z - 1)];
    a = a + f[sidx(o, x, y, z)];
    a = a + f[sidx(o, x + 1, y, z)];
    a = a + f[sidx(o, x, y - 1, z)];
    a = a + f[sidx(o, x + 1, y - 1, z)];
    return a / 8;
}
static inline float avg_v(const smk_oracle* o, const float* f, int x, int y, int z)
{
    float a = f[sidx(o, x, y, z - 1)];
    a = a + f[sidx(o, x - 1, y, z - 1)];
    a = a + f[sidx(o, x, y + 1, z - 1)];
    a = a + f[sidx(o, x - 1, y + 1, z - 1)];
    a = a + f[sidx(o, x, y, z)];
    a = a + f[sidx(o, x - 1, y, z)];
    a = a + f[sidx(o, x, y + 1, z)];
    a = a + f[sidx(o, x - 1, y + 1, z)];
    return a / 8;
}
static inline float avg_w(const smk_oracle* o, const float* f, int x, int y, int z)
{
    float a = f[sidx(o, x, y, z)];
    a = a + f[sidx(o, x - 1, y, z)];
    a = a + f[sidx(o, x, y - 1, z)];
    a = a + f[sidx(o, x - 1, y - 1, z)];
    a = a + f[sidx(o, x, y, z - 1)];
    a = a + f[sidx(o, x - 1, y, z - 1)];
    a = a + f[sidx(o, x, y - 1, z - 1)];
    a = a + f[sidx(o, x - 1, y - 1, z - 1)];
    return a / 8;
}

/* cu:451-484.  Clamped trilinear sample.  Clamp bounds always come from the CELL dims
 * (even for staggered fields); strides (px,py) are those of the sampled array.
 *   p  = max(min(pos, dim-1), 1);  q = p - delta;  i0 = (int)min(floor(q), dim-1)
 *   w1 = q - i0; w0 = 1 - w1;      i1 = (int)min(i0+1, dim-1)
 * Sum order 000,100,010,110,001,101,011,111 with each weight ((xw*yw)*zw).
 * nvcc (SASS): acc = RN(w100*f100); then acc = fma(w, f, acc) for 000,010,110,001,101,011,111. */
static inline float sample(const smk_oracle* o, const float* f, int px, int py,
                           float posx, float posy, float posz, float dx, float dy, float dz)
{
    const int ct = o->contract;
    float bx = (float)(unsigned)(o->W - 1), by = (float)(unsigned)(o->H - 1), bz = (float)(unsigned)(o->D - 1);
    float x = fmaxf(fminf(posx, bx), 1.f);
    float y = fmaxf(fminf(posy, by), 1.f);
    float z = fmaxf(fminf(posz, bz), 1.f);
    float qx = x - dx, qy = y - dy, qz = z - dz;
    int x0 = (int)fminf(floorf(qx), bx);
    int y0 = (int)fminf(floorf(qy), by);
    int z0 = (int)fminf(floorf(qz), bz);
    float xw1 = qx - (float)x0, yw1 = qy - (float)y0, zw1 = qz - (float)z0;
    float xw0 = 1.f - xw1, yw0 = 1.f - yw1, zw0 = 1.f - zw1;
    int x1 = (int)fminf((float)(x0 + 1), bx);
    int y1 = (int)fminf((float)(y0 + 1), by);
    int z1 = (int)fminf((float)(z0 + 1), bz);
    size_t sy = (size_t)px, sz = (size_t)px * py;
    float f000 = f[x0 + y0 * sy + z0 * sz], f100 = f[x1 + y0 * sy + z0 * sz];
    float f010 = f[x0 + y1 * sy + z0 * sz], f110 = f[x1 + y1 * sy + z0 * sz];
    float f001 = f[x0 + y0 * sy + z1 * sz], f101 = f[x1 + y0 * sy + z1 * sz];
    float f011 = f[x0 + y1 * sy + z1 * sz], f111 = f[x1 + y1 * sy + z1 * sz];
    float w000 = (xw0 * yw0) * zw0, w100 = (xw1 * yw0) * zw0;
    float w010 = (xw0 * yw1) * zw0, w110 = (xw1 * yw1) * zw0;
    float w001 = (xw0 * yw0) * zw1, w101 = (xw1 * yw0) * zw1;
    float w011 = (xw0 * yw1) * zw1, w111 = (xw1 * yw1) * zw1;
    float acc;
    if (ct) {
        acc = w100 * f100;
        acc = fmaf(w000, f000, acc);
    } else {
        float t0 = w000 * f000, t1 = w100 * f100;
        acc = t0 + t1;
    }
    acc = fm(ct, w010, f010, acc);
    acc = fm(ct, w110, f110, acc);
    acc = fm(ct, w001, f001, acc);
    acc = fm(ct, w101, f101, acc);
    acc = fm(ct, w011, f011, acc);
    acc = fm(ct, w111, f111, acc);
    return acc;
}

/* pos0 - dt*vel: nvcc contracts to fma(-vel, dt, pos0) (cu:543-545 and siblings) */
static inline float backtrace(int ct, float pos0, float vel, float dt)
{
    if (ct) return fmaf(-vel, dt, pos0);
    float p = dt * vel;
    return pos0 - p;
}

/* cu:527-615.  Semi-Lagrangian self-advection of the three MAC components,
 * now -> past.  Faces whose two adjacent cells are not both fluid, and the outer ranges,
 * are NOT written (they keep whatever the destination buffer held).
 *   U (cu:527-555): x in [1,W), y in [1,H-1), z in [1,D-1); s[x]&&s[x-1]; pos (x, y+.5, z+.5)
 *   V (cu:557-585): x in [1,W-1), y in [1,H), z in [1,D-1); s[y]&&s[y-1]; pos (x+.5, y, z+.5)
 *   W (cu:587-615): x in [1,W-1), y in [1,H-1), z in [1,D); s[z]&&s[z-1]; pos (x+.5, y+.5, z)
 * "i + 0.5" is evaluated in double and narrowed (cu:536-537 etc.). */
void orc_advect_velocity(smk_oracle* o, float dt) { orc_advect_velocity_r(o, dt, 0, o->D); }

/* the same on node planes [za, zb) only */
void orc_advect_velocity_r(smk_oracle* o, float dt, int za, int zb)
{
    const int W = o->W, H = o->H, D = o->D, SX = o->SX, SY = o->SY, ct = o->contract;
    if (za < 1) za = 1;
    if (zb > D) zb = D;
    const float *u0 = o->u[o->now], *v0 = o->v[o->now], *w0 = o->w[o->now];
    float *u1 = o->u[o->past], *v1 = o->v[o->past], *w1 = o->w[o->past];
    const unsigned char* s = o->s;
#pragma omp parallel for schedule(static)
    for (int z = za; z < zb; z++)
        for (int y = 1; y < H; y++)
            for (int x = 1; x < W; x++) {
                size_t f = sidx(o, x, y, z);
                int here = s[cidx(o, x, y, z)];
                if (y < H - 1 && z < D - 1 && here && s[cidx(o, x - 1, y, z)]) {
                    float px = backtrace(ct, (float)x, u0[f], dt);
                    float py = backtrace(ct, (float)((double)y + 0.5), avg_v(o, v0, x, y, z), dt);
                    float pz = backtrace(ct, (float)((double)z + 0.5), avg_w(o, w0, x, y, z), dt);
                    u1[f] = sample(o, u0, SX, SY, px, py, pz, 0.f, .5f, .5f);
                }
                if (x < W - 1 && z < D - 1 && here && s[cidx(o, x, y - 1, z)]) {
                    float px = backtrace(ct, (float)((double)x + 0.5), avg_u(o, u0, x, y, z), dt);
                    float py = backtrace(ct, (float)y, v0[f], dt);
                    float pz = backtrace(ct, (float)((double)z + 0.5), avg_w(o, w0, x, y, z), dt);
                    v1[f] = sample(o, v0, SX, SY, px, py, pz, .5f, 0.f, .5f);
                }
                if (x < W - 1 && y < H - 1 && here && s[cidx(o, x, y, z - 1)]) {
                    float px = backtrace(ct, (float)((double)x + 0.5), avg_u(o, u0, x, y, z), dt);
                    float py = backtrace(ct, (float)((double)y + 0.5), avg_v(o, v0, x, y, z), dt);
                    float pz = backtrace(ct, (float)z, w0[f], dt);
                    w1[f] = sample(o, w0, SX, SY, px, py, pz, .5f, .5f, 0.f);
                }
            }
}

/* cu:617-638.  Density advection with the NEW velocities (the "past" buffers written by
 * orc_advect_velocity; cu:810).  Interior fluid cells only.
 *   u_t = (u[x] + u[x+1]) / 2 ...;  pos = (float)((double)x + 0.5 - (double)(u_t*dt))
 * (u_t*dt is a float product, widened; the subtraction is in double: cu:628-630). */
void orc_advect_smoke(smk_oracle* o, float dt) { orc_advect_smoke_r(o, dt, 0, o->D); }

/* the same on cell planes [za, zb) only */
void orc_advect_smoke_r(smk_oracle* o, float dt, int za, int zb)
{
    const int W = o->W, H = o->H, D = o->D;
    const float *u = o->u[o->past], *v = o->v[o->past], *w = o->w[o->past];
    const float* s0 = o->smoke[o->now];
    float* s1 = o->smoke[o->past];
    if (za < 1) za = 1;
    if (zb > D - 1) zb = D - 1;
#pragma omp parallel for schedule(static)
    for (int z = za; z < zb; z++)
        for (int y = 1; y < H - 1; y++)
            for (int x = 1; x < W - 1; x++) {
                size_t c = cidx(o, x, y, z);
                if (!o->s[c]) continue;
                size_t f = sidx(o, x, y, z);
                float ut = (u[f] + u[sidx(o, x + 1, y, z)]) / 2;
                float vt = (v[f] + v[sidx(o, x, y + 1, z)]) / 2;
                float wt = (w[f] + w[sidx(o, x, y, z + 1)]) / 2;
                float ud = ut * dt, vd = vt * dt, wd = wt * dt;
                float px = (float)(((double)x + 0.5) - (double)ud);
                float py = (float)(((double)y + 0.5) - (double)vd);
                float pz = (float)(((double)z + 0.5) - (double)wd);
                s1[c] = sample(o, s0, W, H, px, py, pz, .5f, .5f, .5f);
            }
}

/* cu:774-819 without the device->host copy: flip, fill, force, clamp,
 * iterations x (even, odd) half-sweeps, velocity advection, density advection. */
void orc_step(smk_oracle* o, float dt)
{
    orc_flip(o);
    orc_fill(o);
    orc_integrate(o, dt);
    orc_clamp(o, dt);
    if (o->solver == 1) { /* extension: damped Jacobi */
        float* p = (float*)malloc(o->ncell * sizeof(float));
        for (int i = 0; i < o->iterations; i++) orc_jacobi_iteration(o, p);
        free(p);
    } else
    for (int i = 0; i < o->iterations; i++) {
        orc_pressure_halfsweep(o, 0);
        orc_pressure_halfsweep(o, 1);
    }
    orc_advect_velocity(o, dt);
    orc_advect_smoke(o, dt);
}

static void* field_ptr(const smk_oracle* o, int field, int which, size_t* bytes)
{
    int b = which == 0 ? o->now : which == 1 ? o->past : which - 2;
    switch (field) {
    case ORC_FIELD_SMOKE: *bytes = o->ncell * sizeof(float); return o->smoke[b];
    case ORC_FIELD_U: *bytes = o->nstag * sizeof(float); return o->u[b];
    case ORC_FIELD_V: *bytes = o->nstag * sizeof(float); return o->v[b];
    case ORC_FIELD_W: *bytes = o->nstag * sizeof(float); return o->w[b];
    case ORC_FIELD_MASK: *bytes = o->ncell; return o->s;
    }
    *bytes = 0;
    return NULL;
}

void orc_get_field(const smk_oracle* o, int field, int which, void* dst)
{
    size_t n; void* p = field_ptr(o, field, which, &n);
    if (p) memcpy(dst, p, n);
}
void orc_set_field(smk_oracle* o, int field, int which, const void* src)
{
    size_t n; void* p = field_ptr(o, field, which, &n);
    if (p) memcpy(p, src, n);
}

float orc_max_divergence(const smk_oracle* o)
{
    const int W = o->W, H = o->H, D = o->D;
    const float *u = o->u[o->now], *v = o->v[o->now], *w = o->w[o->now];
    float m = 0.f;
#pragma omp parallel for reduction(max : m) schedule(static)
    for (int z = 1; z < D - 1; z++)
        for (int y = 1; y < H - 1; y++)
            for (int x = 1; x < W - 1; x++) {
                if (!o->s[cidx(o, x, y, z)]) continue;
                size_t f = sidx(o, x, y, z);
                float div = -u[f] + u[sidx(o, x + 1, y, z)];
                div = div + -v[f];
                div = div + v[sidx(o, x, y + 1, z)];
                div = div + -w[f];
                div = div + w[sidx(o, x, y, z + 1)];
                float a = fabsf(div);
                if (a > m) m = a;
            }
    return m;
}

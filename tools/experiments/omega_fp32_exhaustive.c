// Exhaustive check (all 2^32 binary32 inputs) of the fp32-only evaluation of the reference's over-relaxation product
//     p = (float)((double)q * -1.9)                       (project/smokeSimulation.cu:384; a double multiply, rounded twice)
// used by smoke-simulation_b200/csrc/kernels_pressure_tma.cuh (p_from_q):
//     c = -1.9 (binary64) = ch + cl + cr,  ch = (float)c,  cl = (float)(c - ch)
//     P+ = fma(q, ch, q * (cl + 2^-39))        P- = fma(q, ch, q * (cl - 2^-39))       (cl +- 2^-39 are binary32 constants)
//     p  = P+ if P+ == P-, else the one of the two (adjacent) values whose mantissa is even.
// Why: q*ch + RN(q*cl) equals q*c up to ~2^-48 |p| (the rounding of the small product, the dropped cr) and RN53(q*c) up to 2^-53 |p|; the shift
// 2^-39 |q| = 2^-40 |p| dominates both, so RN53(q*c) lies between the arguments of the two fused roundings, and rounding to binary32
// is monotonic: P+ == P- is the answer.  P+ != P- means a rounding boundary of binary32 lies within 2^-40 |p| of q*c.
// 1.9 (binary64) is 19/10 - 8.9e-17, and 19 q / 10 lies on a lattice of tenths of an ulp: such a boundary is then an
// EXACT tie of 19 q / 10 (5.26 % = 1/19 of all inputs), q*c misses it by less than half an ulp of binary64, the double
// product rounds ONTO the midpoint and the conversion to binary32 breaks the tie to even.
// Needs the small products free of underflow: the kernel sends 0 < |d| < 2^-96 (d = n q, n <= 6) to the F2F / DMUL / F2F path, so
// here every |q| >= 2^-99 (and q = +-0, sign included) must match; smaller inputs are only counted.
// Build / run:  gcc -O2 -march=native -fopenmp -ffp-contract=off -o /tmp/omega_check tools/experiments/omega_fp32_exhaustive.c -lm && /tmp/omega_check [stride]
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

int main(int argc, char** argv)
{
    const long long stride = argc > 1 ? atoll(argv[1]) : 1; // 1 = every bit pattern (about a minute on 8 cores); tests sample
    const double c = -1.9;
    const float ch = (float)c, cl = (float)(c - (double)ch);
    const float eps = 0x1p-39f, clp = cl + eps, clm = cl - eps;
    if ((double)clp != (double)cl + (double)eps || (double)clm != (double)cl - (double)eps) { printf("cl +- eps not exact\n"); return 2; }
    printf("ch = %a (0x%08x)  cl = %a (0x%08x)  cl+eps = %a (0x%08x)  cl-eps = %a (0x%08x)\n", ch, f2u(ch), cl, f2u(cl), clp, f2u(clp), clm, f2u(clm));
    unsigned long long ties = 0, wrong = 0, wrong_small = 0, checked = 0, not_adjacent = 0;
#pragma omp parallel for reduction(+ : ties, wrong, wrong_small, checked, not_adjacent) schedule(static)
    for (long long i = 0; i < (1ll << 32); i += stride) {
        const uint32_t b = (uint32_t)i;
        if ((b & 0x7f800000u) == 0x7f800000u) continue; // inf / nan: never produced (clamped fields)
        const float q = u2f(b);
        const float ref = (float)((double)q * c);
        const uint32_t bp = f2u(fmaf(q, ch, q * clp)), bm = f2u(fmaf(q, ch, q * clm));
        const int32_t d = (int32_t)(bp - bm);
        const uint32_t hi = (int32_t)bp > (int32_t)bm ? bp : bm;
        const uint32_t res = hi & ~((uint32_t)d & 1u);
        checked++;
        const uint32_t a = b & 0x7fffffffu;
        const int small = a != 0 && a < 0x0E000000u; // 0 < |q| < 2^-99
        if (!small) { ties += d != 0; not_adjacent += (d > 1 || d < -1); }
        if (res != f2u(ref)) { if (small) wrong_small++; else wrong++; }
    }
    printf("checked %llu finite inputs\n", checked);
    printf("|q| >= 2^-99 or q == 0: different from the reference: %llu   <- must be 0\n", wrong);
    printf("|q| >= 2^-99: P+ != P- (ties broken to even): %llu = %.4f %%, of which not adjacent: %llu\n", ties, 100.0 * (double)ties / (double)checked, not_adjacent);
    printf("0 < |q| < 2^-99 (F2F/DMUL path in the kernel): different: %llu\n", wrong_small);
    return wrong != 0;
}

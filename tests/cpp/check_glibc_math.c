/* check_glibc_math.c -- pins minerva_b200/csrc/glibc_math.h (the exact-mode math of the CUDA kernels) against THIS machine's libm: expf, logf, tanhf and the reference's sigmoid
 * formula compared bit for bit on every one of the 2^32 float inputs (or every `stride`-th one).  TEST INFRASTRUCTURE.
 *   gcc -O2 -fopenmp -mfma -ffp-contract=off -I minerva_b200/csrc tests/cpp/check_glibc_math.c -lm && ./a.out [stride]
 * prints "<fn>: <mismatches> mismatches of <tested>" per function; exit status 0 iff every function matches everywhere.
 * NaN results compare equal to NaN (payload / sign of a NaN is not part of the contract). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#define MNV_INFF INFINITY
#define MNV_NANF NAN
#define MNV_MUL(a, b) ((a) * (b))
#define MNV_ADD(a, b) ((a) + (b))
#define MNV_SUB(a, b) ((a) - (b))
#define MNV_DIV(a, b) ((a) / (b))
#define MNV_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define MNV_FMULF(a, b) ((a) * (b))
#define MNV_FADDF(a, b) ((a) + (b))
#define MNV_FSUBF(a, b) ((a) - (b))
#define MNV_FDIVF(a, b) ((a) / (b))
#include "glibc_math.h"

static float ref_sigmoid(float x) { return 1.0 / (1.0 + expf(-x)); }   /* basic.cpp:416, verbatim formula */

int main(int argc, char** argv) {
  const unsigned long long stride = argc > 1 ? strtoull(argv[1], 0, 10) : 1;
  const char* names[4] = {"expf", "logf", "tanhf", "sigmoid"};
  unsigned long long bad[4] = {0, 0, 0, 0}, first[4] = {~0ull, ~0ull, ~0ull, ~0ull}, tested = 0;
#pragma omp parallel for reduction(+ : bad[:4], tested) schedule(static)
  for (long long ii = 0; ii < (long long)((0x100000000ull + stride - 1) / stride); ++ii) {
    const uint32_t u = (uint32_t)((unsigned long long)ii * stride);
    const float x = mnv_u2f(u);
    const float want[4] = {expf(x), logf(x), tanhf(x), ref_sigmoid(x)};
    const float got[4] = {mnv_glibc_expf(x), mnv_glibc_logf(x), mnv_glibc_tanhf(x), mnv_ref_sigmoidf(x)};
    ++tested;
    for (int f = 0; f < 4; ++f) {
      const int same = (want[f] != want[f] && got[f] != got[f]) || mnv_f2u(want[f]) == mnv_f2u(got[f]);
      if (!same) {
        ++bad[f];
#pragma omp critical
        if (u < first[f]) first[f] = u;
      }
    }
  }
  int rc = 0;
  for (int f = 0; f < 4; ++f) {
    printf("%s: %llu mismatches of %llu", names[f], bad[f], tested);
    if (bad[f]) { printf(" (first at 0x%08llx)", first[f]); rc = 1; }
    printf("\n");
  }
  return rc;
}

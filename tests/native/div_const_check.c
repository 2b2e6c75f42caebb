/* Exhaustive check of the constant-divisor division used inside the bit-exact zone
 * (vampire_b200/csrc/vb_common.cuh: sdiv_const).  For each divisor y given on the command line and EVERY fp32
 * mantissa of x (2^23 values) at several exponents and both signs, the sequence
 *     r  = RN(1/y)                       (host)
 *     q0 = RN(x * r)
 *     q1 = fma(fma(-q0, y, x), r, q0)
 *     q2 = fma(fma(-q1, y, x), r, q1)
 * must equal the IEEE-754 quotient RN(x / y) bit for bit.  Prints the number of mismatches per divisor.
 * Build: gcc -O2 -ffp-contract=off [-mfma] div_const_check.c -lm      (test infrastructure only) */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

int main(int argc, char** argv) {
  static const int exps[] = {-100, -20, -1, 0, 1, 9, 30, 100};
  long long total_bad = 0;
  for (int a = 1; a < argc; ++a) {
    const float y = strtof(argv[a], NULL);
    volatile float rv = 1.0f / y;
    const float r = rv;
    long long bad = 0;
    for (size_t ei = 0; ei < sizeof(exps) / sizeof(exps[0]); ++ei) {
      for (uint32_t m = 0; m < (1u << 23); ++m) {
        const float x0 = ldexpf(1.0f + (float)m * 0x1p-23f, exps[ei]);
        for (int sgn = 0; sgn < 2; ++sgn) {
          const float x = sgn ? -x0 : x0;
          volatile float q0v = x * r;
          const float q0 = q0v;
          const float q1 = fmaf(fmaf(-q0, y, x), r, q0);
          const float q2 = fmaf(fmaf(-q1, y, x), r, q1);
          volatile float refv = x / y;
          if (bits(q2) != bits(refv)) ++bad;
        }
      }
    }
    printf("%.9g %lld\n", (double)y, bad);
    total_bad += bad;
  }
  return total_bad ? 1 : 0;
}

/*
 * div_by_scalefactor_model.c -- TEST INFRASTRUCTURE: k_pack's quantiser divides every subband sample by a scalefactor
 * (encode_new.c:500-533: d = sample / scalefactor[idx]).  The kernel forms the IEEE quotient without the division
 * subroutine: y = RN(1 / b), q0 = RN(a y), r = a - q0 b (exact, one FMA), q = RN(q0 + r y) (Markstein's sequence).
 * Here that sequence is compared bit for bit with a / b for all 64 scalefactors over N random operands in the range
 * subband samples live in and N operands built to straddle the rounding boundaries of the quotient.
 * usage: div_by_scalefactor_model [CASES_PER_SCALEFACTOR]     prints "... bad N"
 */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
#include <stdint.h>
#define MP2_TABLE_QUAL static const
#include "mp2_tables.h"
static uint64_t rs = 88172645463325252ULL;
static uint64_t rnd(void){ rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return rs; }
int main(int argc, char **argv){
  long bad = 0, n = 0;
  const long per = argc > 1 ? atol(argv[1]) : 20000000;
  for (int i = 0; i < 64; i++) {
    const double b = MP2_SCALEFACTOR[i], y = 1.0 / b;
    for (long t = 0; t < per; t++) {
      uint64_t u = rnd();
      double a;
      if (t & 1) { // random mantissa, exponent in the range subband samples live in (|x| <= ~2, down to 1e-12), both signs
        uint64_t m = u & 0xFFFFFFFFFFFFFull; int e = 1023 - (int)((u >> 52) % 42); uint64_t bits = ((uint64_t)(u >> 63) << 63) | ((uint64_t)e << 52) | m; memcpy(&a, &bits, 8);
      } else { // adversarial: a = q*b for q with few mantissa bits +- tiny -> quotients near rounding boundaries
        double q = (double)((u >> 11) & 0xFFFFFF) + 0.5; q = ldexp(q, -((int)(u & 31))); a = q * b; uint64_t bits; memcpy(&bits, &a, 8); bits += (int)((u >> 40) % 5) - 2; memcpy(&a, &bits, 8);
      }
      const double want = a / b;
      const double q0 = a * y, r = fma(-q0, b, a), got = fma(r, y, q0);
      n++;
      if (memcmp(&want, &got, 8)) { if (bad < 10) printf("sf %d a %a: %a vs %a\n", i, a, want, got); bad++; }
    }
  }
  printf("%ld cases, bad %ld\n", n, bad); return 0; }

/*
 * sf_index_model.c -- TEST INFRASTRUCTURE: k_filterbank's direct scalefactor-index computation against the reference's
 * binary search (encode_new.c:207-219, restated as sf_index_of in oracle/mp2_oracle.c which this file includes).
 * The search finds the largest index whose table value is >= the block maximum; the table is 2^(1 - i/3) written as
 * rounded decimals, so the binade of the maximum pins the index to four neighbouring entries, whose comparisons are
 * independent of each other; maxima outside the table's range, or within rounding of a binade edge, take the search.
 * Checked on every table value and every power of two +- 3 ulp, 0, subnormals, values above 2, and random doubles.
 * usage: sf_index_model N_RANDOM      prints "bad N"
 */
#include "../oracle/mp2_oracle.c"
#include <stdio.h>

static unsigned direct(double mx)
{
    uint64_t bits;
    memcpy(&bits, &mx, 8);
    const int e = (int)((bits >> 52) & 0x7ff) - 1023;
    if (e < -19 || e > 0) return sf_index_of(mx);
    const int g = 3 * (1 - e); /* table[g] ~ 2^e <= mx < 2^(e+1) ~ table[g-3] */
    if (!(mx <= MP2_SCALEFACTOR[g - 3])) return sf_index_of(mx);
    return (unsigned)(g - 3 + (mx <= MP2_SCALEFACTOR[g - 2]) + (mx <= MP2_SCALEFACTOR[g - 1]) + (mx <= MP2_SCALEFACTOR[g]));
}

static uint64_t rs = 0x9E3779B97F4A7C15ull;
static uint64_t rnd(void) { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return rs; }

int main(int argc, char **argv)
{
    long n = argc > 1 ? atol(argv[1]) : 10000000, bad = 0, fallback = 0;
    #define CHECK(v) do { const double v_ = (v); if (direct(v_) != sf_index_of(v_)) { if (bad < 10) printf("%a: %u vs %u\n", v_, direct(v_), sf_index_of(v_)); bad++; } } while (0)
    for (int i = 0; i < 64; i++)
        for (int d = -3; d <= 3; d++) { double v = MP2_SCALEFACTOR[i]; uint64_t b; memcpy(&b, &v, 8); b += d; memcpy(&v, &b, 8); CHECK(v); }
    for (int e = -30; e <= 3; e++)
        for (int d = -3; d <= 3; d++) { double v = ldexp(1.0, e); uint64_t b; memcpy(&b, &v, 8); b += d; memcpy(&v, &b, 8); CHECK(v); }
    CHECK(0.0); CHECK(1e-320); CHECK(1e-25); CHECK(1e-20); CHECK(2.0); CHECK(2.5); CHECK(100.0);
    for (long t = 0; t < n; t++) {
        uint64_t u = rnd(), m = u & 0xFFFFFFFFFFFFFull;
        int e = 1023 + 2 - (int)((u >> 52) % 26);
        uint64_t b = ((uint64_t)e << 52) | m;
        double v; memcpy(&v, &b, 8);
        CHECK(v);
    }
    (void)fallback;
    printf("bad %ld\n", bad);
    return 0;
}

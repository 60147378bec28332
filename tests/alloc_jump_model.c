/*
 * alloc_jump_model.c -- TEST INFRASTRUCTURE: CPU model of k_alloc's jump start (odr_audioenc_b200/csrc/mp2_kernels.cu)
 * against the oracle's verbatim greedy loop (greedy_alloc in oracle/mp2_oracle.c = encode_new.c:1078-1187).
 *
 * The greedy loop always raises the entry with the smallest mask-to-noise ratio; an entry's ratio grows with every
 * step.  So the steps happen in the order of their keys (ratio before the step), and as long as every step is
 * affordable, the state after "all steps with key < lambda" does not depend on that order: entry e stands at the
 * first allocation b whose ratio snr[b] - smr_e is >= lambda (joint bands: the smaller of the two channels' ratios
 * drives the shared allocation).  The jump start bisects lambda for the highest level whose total cost still fits the
 * budget, sets that state directly and lets the exact loop finish from there.
 * Random SMR vectors (with many exact ties), scfsi codes, budgets and joint-stereo bounds.  Prints "bad N".
 * usage: alloc_jump_model N_TRIALS SEED
 */
#include "../oracle/mp2_oracle.c"
#include <stdio.h>

#define MIN_STEP_BITS 12 /* 12 granule-triplets x 1 bit: the cheapest step in the tables (9 -> 10 bits per triplet) */
static int jump_alloc(const mp2o_cfg *c, double smr[2][32], uint8_t scfsi[2][32], int jsbound, int adb, uint8_t bit_alloc[2][32],
                      int steps, long *rounds_left)
{
    int nch = c->nch, sblimit = c->sblimit, bbal = 0;
    for (int sb = 0; sb < jsbound; sb++) bbal += nch * MP2_ROW_NBAL[row_of(c, sb)];
    for (int sb = jsbound; sb < sblimit; sb++) bbal += MP2_ROW_NBAL[row_of(c, sb)];
    const int ad = adb - (bbal + 16 + 32);
    /* state at level lambda, and its cost */
    int best_b[2][32];
    memset(best_b, 0, sizeof best_b); /* level 'nothing granted yet' is where the plain loop starts */
    double lo = 1e300, hi = -1e300;
    for (int sb = 0; sb < sblimit; sb++)
        for (int ch = 0; ch < nch; ch++) {
            const double m0 = MP2_QC_SNR[0] - smr[ch][sb];
            if (m0 < lo) lo = m0;
            const int row = row_of(c, sb), top = (1 << MP2_ROW_NBAL[row]) - 1;
            const double m1 = MP2_QC_SNR[MP2_ROW_QC[row][top]] - smr[ch][sb];
            if (m1 > hi) hi = m1;
        }
    if (getenv("HI_SPAN") && hi > lo + atof(getenv("HI_SPAN"))) hi = lo + atof(getenv("HI_SPAN"));
    for (int it = 0; it < steps; it++) {
        const double lambda = it == 0 ? lo : 0.5 * (lo + hi); /* level lo: nothing granted, always affordable */
        int b_at[2][32], cost = 0;
        for (int sb = 0; sb < sblimit; sb++) {
            const int row = row_of(c, sb), top = (1 << MP2_ROW_NBAL[row]) - 1;
            const int joint = nch == 2 && sb >= jsbound;
            for (int ch = 0; ch < nch; ch++) {
                double s = smr[ch][sb];
                if (joint) { s = smr[0][sb] > smr[1][sb] ? smr[0][sb] : smr[1][sb]; if (ch == 1) { b_at[1][sb] = b_at[0][sb]; continue; } }
                int b = 0;
                while (b < top && MP2_QC_SNR[MP2_ROW_QC[row][b]] - s < lambda) b++;
                b_at[ch][sb] = b;
                if (b > 0) {
                    const int q = MP2_ROW_QC[row][b];
                    cost += 12 * MP2_QC_NCODE[q] * MP2_QC_BITS[q] + 2 + 6 * MP2_SCFSI_NSF[scfsi[ch][sb]];
                    if (joint) cost += 2 + 6 * MP2_SCFSI_NSF[scfsi[1][sb]];
                }
            }
        }
        if (cost <= ad) { memcpy(best_b, b_at, sizeof b_at); if (it) lo = lambda; }
        else hi = lambda;
    }
    /* continue with the verbatim loop from that state */
    double mnr[2][32];
    char used[2][32];
    int spent = 0;
    memset(bit_alloc, 0, 64);
    for (int sb = 0; sb < sblimit; sb++) {
        const int row = row_of(c, sb), top = (1 << MP2_ROW_NBAL[row]) - 1, joint = nch == 2 && sb >= jsbound;
        for (int ch = 0; ch < nch; ch++) {
            const int b = best_b[ch][sb];
            bit_alloc[ch][sb] = (uint8_t)b;
            mnr[ch][sb] = MP2_QC_SNR[MP2_ROW_QC[row][b]] - smr[ch][sb];
            used[ch][sb] = b == 0 ? 0 : (b >= top ? 2 : 1);
            if (b > 0 && !(joint && ch == 1)) {
                const int q = MP2_ROW_QC[row][b];
                spent += 12 * MP2_QC_NCODE[q] * MP2_QC_BITS[q] + 2 + 6 * MP2_SCFSI_NSF[scfsi[ch][sb]];
                if (joint) spent += 2 + 6 * MP2_SCFSI_NSF[scfsi[1][sb]];
            }
        }
    }
    for (;;) {
        int min_sb = -1, min_ch = -1;
        double small = 999999.0;
        for (int ch = 0; ch < nch; ch++)
            for (int sb = 0; sb < sblimit; sb++)
                if (used[ch][sb] != 2 && small > mnr[ch][sb]) { small = mnr[ch][sb]; min_sb = sb; min_ch = ch; }
        if (min_sb < 0) break;
        if (ad - spent < MIN_STEP_BITS) break; /* no step of any entry costs less: nothing can be granted any more */
        (*rounds_left)++;
        int row = row_of(c, min_sb), oth = 1 - min_ch;
        int qn = MP2_ROW_QC[row][bit_alloc[min_ch][min_sb] + 1];
        int cost = 12 * MP2_QC_NCODE[qn] * MP2_QC_BITS[qn];
        if (used[min_ch][min_sb]) {
            int q = MP2_ROW_QC[row][bit_alloc[min_ch][min_sb]];
            cost -= 12 * MP2_QC_NCODE[q] * MP2_QC_BITS[q];
        } else {
            cost += 2 + 6 * MP2_SCFSI_NSF[scfsi[min_ch][min_sb]];
            if (nch == 2 && min_sb >= jsbound) cost += 2 + 6 * MP2_SCFSI_NSF[scfsi[oth][min_sb]];
        }
        if (ad >= spent + cost) {
            int ba = ++bit_alloc[min_ch][min_sb];
            spent += cost;
            used[min_ch][min_sb] = 1;
            mnr[min_ch][min_sb] = MP2_QC_SNR[MP2_ROW_QC[row][ba]] - smr[min_ch][min_sb];
            if (ba >= (1 << MP2_ROW_NBAL[row]) - 1) used[min_ch][min_sb] = 2;
        } else used[min_ch][min_sb] = 2;
        if (min_sb >= jsbound && nch == 2) {
            int ba = bit_alloc[oth][min_sb] = bit_alloc[min_ch][min_sb];
            used[oth][min_sb] = used[min_ch][min_sb];
            mnr[oth][min_sb] = MP2_QC_SNR[MP2_ROW_QC[row][ba]] - smr[oth][min_sb];
        }
    }
    return ad - spent;
}

static unsigned long long rs;
static unsigned rnd(void) { rs = rs * 6364136223846793005ULL + 1442695040888963407ULL; return (unsigned)(rs >> 33); }

int main(int argc, char **argv)
{
    long n = argc > 1 ? atol(argv[1]) : 100000, bad = 0, rounds_ref = 0, rounds_jump = 0;
    rs = argc > 2 ? (unsigned long long)atoll(argv[2]) : 1;
    static const struct { long fs; char mode; int br; } cfgs[] = {{48000, 'j', 192}, {48000, 's', 192}, {48000, 'j', 128}, {24000, 'm', 64},
        {48000, 's', 96}, {48000, 'm', 96}, {48000, 'j', 256}, {48000, 's', 384}, {24000, 'j', 144}, {48000, 'j', 64}, {32000, 's', 192}, {32000, 'm', 48}, {24000, 'm', 8}};
    for (long t = 0; t < n; t++) {
        mp2o_cfg c;
        const int k = rnd() % (sizeof cfgs / sizeof cfgs[0]);
        mp2o_configure(&c, cfgs[k].fs, cfgs[k].mode, cfgs[k].br, 1, 0);
        double smr[2][32];
        uint8_t scfsi[2][32], a[2][32], b[2][32];
        const int style = rnd() % 4;
        for (int ch = 0; ch < 2; ch++)
            for (int sb = 0; sb < 32; sb++) {
                const double u = (double)(rnd() % 100000) / 100000.0;
                double v = -25.0 + 70.0 * u - 0.8 * sb;
                if (style == 1) v = (double)((int)v);             /* whole dB: many exact ties */
                if (style == 2) v = (double)((int)(v / 6.0)) * 6.0 + 0.5 * (rnd() % 2);
                if (style == 3 && ch == 1) v = smr[0][sb];        /* identical channels */
                smr[ch][sb] = v;
                scfsi[ch][sb] = (uint8_t)(rnd() % 4);
            }
        int jsbound = c.sblimit;
        if (c.mode == 1 && rnd() % 2) { static const int jb[4] = {4, 8, 12, 16}; jsbound = jb[rnd() % 4]; }
        int adb = 8 * c.lg_frame - (c.dab_ext * 8 + 16);
        if (rnd() % 4 == 0) adb -= 8 * (int)(rnd() % 40);          /* X-PAD takes bits away */
        long r1 = 0;
        /* reference round count: bits granted = rounds (plus failures); count via a copy of the loop is not needed: use jump with 1 step (lambda = lo: empty state) */
        const int left_ref = greedy_alloc(&c, smr, scfsi, jsbound, adb, a);
        long r0 = 0;
        uint8_t z[2][32];
        jump_alloc(&c, smr, scfsi, jsbound, adb, z, 1, &r0);
        const int left = jump_alloc(&c, smr, scfsi, jsbound, adb, b, argc > 3 ? atoi(argv[3]) : 5, &r1);
        rounds_ref += r0; rounds_jump += r1;
        if (left != left_ref || memcmp(a, b, 64) || memcmp(a, z, 64)) {
            if (bad < 5) printf("trial %ld cfg %d jsbound %d adb %d: left %d vs %d\n", t, k, jsbound, adb, left, left_ref);
            bad++;
        }
    }
    printf("rounds per frame: plain %.1f, after the jump start %.1f\n", (double)rounds_ref / n, (double)rounds_jump / n);
    printf("bad %ld\n", bad);
    return 0;
}

// mp2_psy2_init.h -- TEST INFRASTRUCTURE (part of the oracle): start-up tables of psychoacoustic model 2 and of
// model 0, evaluated with libm exactly as the reference does in psycho_2_init (ref: psycho_2.c:259-420) and
// psycho_0 (ref: psycho_0.c:36-47, ath.c:7-49).  The product does not compute these: it carries them frozen from the
// compiled reference (csrc/mp2_psy2_tables.h); tests/test_tables.py checks that both hold the same bits.
#pragma once
#include <math.h>
#include <string.h>

#define MP2_P2_CBANDS 64
#define MP2_P2_HBLK 513

typedef struct {
    int n_part;                           /* partitions in use = partition[512] + 1 */
    int partition[MP2_P2_HBLK];           /* FFT line -> partition */
    int numlines[MP2_P2_CBANDS];
    int first_line[MP2_P2_CBANDS + 1];    /* partition p covers lines first_line[p] .. first_line[p+1]-1 */
    double cbval[MP2_P2_CBANDS], tmn[MP2_P2_CBANDS], rnorm[MP2_P2_CBANDS], bmax_of[MP2_P2_CBANDS];
    double s[MP2_P2_CBANDS][MP2_P2_CBANDS]; /* s[j][k]: spreading of partition k's energy into partition j */
    int absthr_table;                     /* index into MP2_ABSTHR (psycho_2.c:292-309) */
} mp2_psy2_tables;

static inline int mp2_psy2_init(mp2_psy2_tables *T, double sfreq)
{
    /* ref: psycho_2.c:24-34 */
    static const double crit_band[27] = {0, 100, 200, 300, 400, 510, 630, 770, 920, 1080, 1270, 1480, 1720, 2000,
                                         2320, 2700, 3150, 3700, 4400, 5300, 6400, 7700, 9500, 12000, 15500, 25000, 30000};
    static const double bmax[27] = {20.0, 20.0, 20.0, 20.0, 20.0, 17.0, 15.0, 10.0, 7.0, 4.4, 4.5, 4.5, 4.5, 4.5,
                                    4.5, 4.5, 4.5, 4.5, 4.5, 4.5, 4.5, 4.5, 4.5, 4.5, 3.5, 3.5, 3.5};
    const double LN_TO_LOG10 = 0.2302585093; /* ref: common.h:31 */
    double fthr[MP2_P2_HBLK];
    int i, j;
    memset(T, 0, sizeof *T);
    i = (int)(sfreq + 0.5);
    switch (i) { /* ref: psycho_2.c:292-309 */
    case 32000: case 16000: T->absthr_table = 0; break;
    case 44100: case 22050: T->absthr_table = 1; break;
    case 48000: case 24000: T->absthr_table = 2; break;
    default: return -1;
    }
    {   /* ref: psycho_2.c:340-372: bark value of each line, partitions at least 0.33 bark wide */
        const double freq_mult = sfreq / 1024;
        double temp1, temp2, bval_lo;
        for (i = 0; i < MP2_P2_HBLK; i++) {
            temp1 = i * freq_mult;
            j = 1;
            while (temp1 > crit_band[j]) j++;
            fthr[i] = j - 1 + (temp1 - crit_band[j - 1]) / (crit_band[j] - crit_band[j - 1]);
        }
        T->partition[0] = 0;
        temp2 = 1;
        T->cbval[0] = fthr[0];
        bval_lo = fthr[0];
        for (i = 1; i < MP2_P2_HBLK; i++) {
            if ((fthr[i] - bval_lo) > 0.33) {
                T->partition[i] = T->partition[i - 1] + 1;
                T->cbval[T->partition[i - 1]] = T->cbval[T->partition[i - 1]] / temp2;
                T->cbval[T->partition[i]] = fthr[i];
                bval_lo = fthr[i];
                T->numlines[T->partition[i - 1]] = (int)temp2;
                temp2 = 1;
            } else {
                T->partition[i] = T->partition[i - 1];
                T->cbval[T->partition[i]] += fthr[i];
                temp2++;
            }
        }
        T->numlines[T->partition[i - 1]] = (int)temp2;
        T->cbval[T->partition[i - 1]] = T->cbval[T->partition[i - 1]] / temp2;
        T->n_part = T->partition[MP2_P2_HBLK - 1] + 1;
    }
    for (j = 0; j < MP2_P2_CBANDS; j++) /* ref: psycho_2.c:378-397; the reference fills s[i][j] with j outermost */
        for (i = 0; i < MP2_P2_CBANDS; i++) {
            double temp1 = (T->cbval[i] - T->cbval[j]) * 1.05, temp2, temp3;
            if (temp1 >= 0.5 && temp1 <= 2.5) {
                temp2 = temp1 - 0.5;
                temp2 = 8.0 * (temp2 * temp2 - 2.0 * temp2);
            } else temp2 = 0;
            temp1 += 0.474;
            temp3 = 15.811389 + 7.5 * temp1 - 17.5 * sqrt((double)(1.0 + temp1 * temp1));
            if (temp3 <= -100) T->s[i][j] = 0;
            else {
                temp3 = (temp2 + temp3) * LN_TO_LOG10;
                T->s[i][j] = exp(temp3);
            }
        }
    for (j = 0; j < MP2_P2_CBANDS; j++) { /* ref: psycho_2.c:400-408 */
        double temp1 = 15.5 + T->cbval[j];
        T->tmn[j] = (temp1 > 24.5) ? temp1 : 24.5;
        T->rnorm[j] = 0;
        for (i = 0; i < MP2_P2_CBANDS; i++) T->rnorm[j] += T->s[j][i];
        T->bmax_of[j] = bmax[(unsigned int)(T->cbval[j] + 0.5)]; /* ref: psycho_2.c:195-196 */
    }
    for (j = 0, i = 0; j <= T->n_part; j++) { /* partitions are runs of consecutive lines */
        T->first_line[j] = i;
        while (i < MP2_P2_HBLK && T->partition[i] == j) i++;
    }
    for (j = T->n_part + 1; j <= MP2_P2_CBANDS; j++) T->first_line[j] = MP2_P2_HBLK;
    return 0;
}

/* Psychoacoustic model 0 (ref: psycho_0.c:27-47, ath.c:7-49): lowest absolute threshold of hearing within each
 * subband, in dB, evaluated on the host with libm as the reference does on its first call. */
static inline double mp2_ath_db(double f)
{   /* ref: ath.c:7-49 with value = 0 */
    double ath;
    if (f < -.3) f = 3410;
    f /= 1000;
    f = (0.01 > f) ? 0.01 : f;
    f = (18.0 < f) ? 18.0 : f;
    ath = 3.640 * pow(f, -0.8) - 6.800 * exp(-0.6 * pow(f - 3.4, 2.0)) + 6.000 * exp(-0.15 * pow(f - 8.7, 2.0)) +
          (0.6 + 0.04 * 0.0) * 0.001 * pow(f, 4.0);
    return ath + 0;
}

static inline void mp2_psy0_init(double ath_min[32], double sfreq)
{
    const double freqperline = sfreq / 1024.0;
    int sb, i;
    for (sb = 0; sb < 32; sb++) ath_min[sb] = 1000;
    for (i = 0; i < 512; i++) {
        const double thisfreq = i * freqperline;
        const double ath_val = mp2_ath_db(thisfreq);
        if (ath_val < ath_min[i >> 4]) ath_min[i >> 4] = ath_val;
    }
}

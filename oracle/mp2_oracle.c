/*
 * mp2_oracle.c -- TEST INFRASTRUCTURE ONLY: the parity oracle (see mp2_oracle.h).
 *
 * Plain sequential C restatement of libtoolame-dab's Layer II DAB encode path,
 * stateless per frame.  Build with -ffp-contract=off (the reference is built
 * -std=c99, i.e. without FMA contraction: SURVEY.md section 1).
 * All "ref:" citations are relative to /root/reference/libtoolame-dab/.
 */
#include "mp2_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "mp2_tables.h"
#include "mp2_alloc_tables.h"
#include "mp2_psy2_tables.h"
#include "mp2_psy2_init.h"

#define DBMIN (-200.0)      /* ref: encoder.h:31 */
#define POWERNORM 90.3090   /* ref: encoder.h:34 */
#define T_TONE 20           /* ref: encoder.h:30 */
#define T_NOISE 10          /* ref: encoder.h:29 */
#define L_LAST (-1)         /* ref: encoder.h:32 */
#define L_STOP (-100)       /* ref: encoder.h:33 */

/* ------------------------------------------------------------------ configuration */

/* ref: toolame.c:120-262 (defaults + setters), common.c:76-144, encode_new.c:104-156, availbits.c:37-67 */
int mp2o_configure(mp2o_cfg *c, long fs_hz, char mode, int bitrate_kbps, int psy, int pad_len)
{
    memset(c, 0, sizeof *c);
    switch (fs_hz) { /* ref: common.c:118-144 */
    case 44100: c->version = 1; c->sfreq_idx = 0; break;
    case 48000: c->version = 1; c->sfreq_idx = 1; break;
    case 32000: c->version = 1; c->sfreq_idx = 2; break;
    case 22050: c->version = 0; c->sfreq_idx = 0; break;
    case 24000: c->version = 0; c->sfreq_idx = 1; break;
    case 16000: c->version = 0; c->sfreq_idx = 2; break;
    default: return -1;
    }
    c->fs_hz = (int)fs_hz;
    if (psy < 0 || psy > 3) return -2; /* ref: toolame.c:204 */
    c->psy = psy;
    switch (mode) { /* ref: toolame.c:174-200 */
    case 's': c->mode = 0; c->mode_ext = 0; break;
    case 'j': c->mode = 1; c->mode_ext = 2; break;
    case 'd': c->mode = 2; c->mode_ext = 0; break;
    case 'm': c->mode = 3; c->mode_ext = 0; break;
    default: return -3;
    }
    c->nch = (c->mode == 3) ? 1 : 2;
    if (bitrate_kbps == 0) bitrate_kbps = MP2_BITRATE[c->version][10]; /* ref: toolame.c:217-218 */
    c->bitrate_index = -1;
    for (int i = 0; i < 15; i++) /* ref: common.c:95-116 (first match, index 0 included) */
        if (MP2_BITRATE[c->version][i] == bitrate_kbps) { c->bitrate_index = i; break; }
    if (c->bitrate_index < 0) return -4;
    c->bitrate_kbps = bitrate_kbps;
    c->dab_ext = 4; /* ref: toolame.c:147,225-232 */
    if (c->version == 1 && bitrate_kbps / (c->mode == 3 ? 1 : 2) < 56) c->dab_ext = 2;
    if (pad_len < 0) return -5;
    c->pad_len = pad_len;
    {   /* ref: encode_new.c:104-124 == tables.c:18-39 */
        int br_per_ch = bitrate_kbps / c->nch;
        static const double s_freq[2][4] = {{22.05, 24, 16, 0}, {44.1, 48, 32, 0}};
        int sfrq = (int)s_freq[c->version][c->sfreq_idx];
        if (c->version == 1) {
            if ((sfrq == 48 && br_per_ch >= 56) || (br_per_ch >= 56 && br_per_ch <= 80)) c->tablenum = 0;
            else if (sfrq != 48 && br_per_ch >= 96) c->tablenum = 1;
            else if (sfrq != 32 && br_per_ch <= 48) c->tablenum = 2;
            else c->tablenum = 3;
        } else c->tablenum = 4;
        c->sblimit = MP2_TAB_SBLIMIT[c->tablenum];
        /* ref: availbits.c:42-46: whole slots of (1152/fs_kHz)*(kbps/8); the padding logic (frac != 0) only
           fires at 44.1/22.05 kHz, which the oracle does not model (DAB uses 48/24 kHz) */
        double average = (1152.0 / s_freq[c->version][c->sfreq_idx]) * ((double)bitrate_kbps / 8.0);
        c->lg_frame = (int)average;
        if (average - (double)c->lg_frame != 0) return -6;
    }
    c->jsbound = (c->mode == 1) ? MP2_JSBOUND[c->mode_ext] : c->sblimit; /* ref: common.c:87-91 */
    c->psy_freq = c->version == 1 ? c->sfreq_idx : c->sfreq_idx + 4;      /* ref: psycho_1.c:42-48 */
    return 0;
}

/* sample `idx` of channel ch, scaled like the reference ((double)s/SCALE); zero before the stream start */
static inline double pcm_at(const int16_t *pcm, int nch, int ch, long idx)
{
    return idx < 0 ? 0.0 : (double)pcm[idx * nch + ch] / 32768.0;
}

/* ------------------------------------------------------------------ polyphase analysis filterbank */

/* ref: subband.c:201-310 (WindowFilterSubband), linear-history form:
 * X[32a+r] = sample r' = 31-r of the block a blocks before the current one, i.e. X[j] = pcm[start+31-j]. */
static void fb_block(const int16_t *pcm, int nch, int ch, long start, double s[32])
{
    double X[512], y[64], yp[32];
    for (int j = 0; j < 512; j++) X[j] = pcm_at(pcm, nch, ch, start + 31 - j);
    for (int i = 0; i < 64; i++) { /* ref: subband.c:246-258,272-283: products added left to right */
        double t = X[i] * MP2_ENWINDOW[i];
        for (int j = 1; j < 8; j++) t += X[i + 64 * j] * MP2_ENWINDOW[i + 64 * j];
        y[i] = t;
    }
    yp[0] = y[16]; /* ref: subband.c:260,285-291 */
    for (int i = 1; i <= 16; i++) yp[i] = y[i + 16] + y[16 - i];
    for (int i = 17; i < 32; i++) yp[i] = y[i + 16] - y[80 - i];
    for (int i = 15; i >= 0; i--) { /* ref: subband.c:293-305: even and odd k accumulated separately from 0.0 */
        double s0 = 0.0, s1 = 0.0;
        for (int k = 0; k < 32; k += 2) {
            s0 += MP2_DCT[i][k] * yp[k];
            s1 += MP2_DCT[i][k + 1] * yp[k + 1];
        }
        s[i] = s0 + s1;
        s[31 - i] = s0 - s1;
    }
}

void mp2o_filterbank_frame(const int16_t *pcm, int nch, int ch, long frame, double sb[36][32])
{
    for (int b = 0; b < 36; b++) /* ref: toolame.c:308-312, block = gr*12+bl */
        fb_block(pcm, nch, ch, frame * 1152 + 32L * b, sb[b]);
}

/* ------------------------------------------------------------------ scalefactors */

/* ref: encode_new.c:179-230 for one (ch,gr,sb): max |x| of the 12 samples, 5-step binary search + fix-up */
static unsigned sf_index_of(double cur_max)
{
    unsigned sf = 32;
    for (unsigned l = 16; l; l >>= 1) {
        if (cur_max <= MP2_SCALEFACTOR[sf]) sf += l;
        else sf -= l;
    }
    if (cur_max > MP2_SCALEFACTOR[sf]) sf--;
    return sf;
}

static void scalefactors(const double sb[36][32], int sblimit, uint8_t sf[3][32])
{
    for (int gr = 0; gr < 3; gr++)
        for (int k = 0; k < 32; k++) {
            if (k >= sblimit) { sf[gr][k] = 0; continue; } /* never written by the reference, stays 0 */
            double mx = fabs(sb[gr * 12 + 11][k]);
            for (int j = 10; j >= 0; j--) {
                double t = fabs(sb[gr * 12 + j][k]);
                if (t > mx) mx = t;
            }
            sf[gr][k] = (uint8_t)sf_index_of(mx);
        }
}

/* ------------------------------------------------------------------ FHT-1024 (psy model FFT) */

/* ref: fft.c:78-1185.  The 496-entry swap table at fft.c:87-1078 is the 10-bit bit-reversal permutation
 * (every pair (k, rev(k)) with k < rev(k)); the twiddles per (stage,i) come from the Buneman recurrence
 * frozen in MP2_FHT_TW. */
void mp2o_fht1024(double *fz)
{
    for (int k = 0; k < 1024; k++) {
        int r = 0;
        for (int b = 0; b < 10; b++) r |= ((k >> b) & 1) << (9 - b);
        if (k < r) { double a = fz[k]; fz[k] = fz[r]; fz[r] = a; }
    }
    for (double *fi = fz; fi < fz + 1024; fi += 4) { /* ref: fft.c:1092-1101 */
        double f1 = fi[0] - fi[1], f0 = fi[0] + fi[1];
        double f3 = fi[2] - fi[3], f2 = fi[2] + fi[3];
        fi[2] = f0 - f2; fi[0] = f0 + f2;
        fi[3] = f1 - f3; fi[1] = f1 + f3;
    }
    const double SQRT2 = 1.4142135623730951454746218587388284504414; /* ref: fft.c:35 */
    int stage = 0;
    for (int k = 2; k <= 8; k += 2, stage++) { /* ref: fft.c:1103-1184 */
        int k1 = 1 << k, k2 = k1 << 1, k4 = k2 << 1, k3 = k2 + k1, kx = k1 >> 1;
        for (double *fi = fz, *gi = fz + kx; fi < fz + 1024; fi += k4, gi += k4) {
            double f1 = fi[0] - fi[k1], f0 = fi[0] + fi[k1];
            double f3 = fi[k2] - fi[k3], f2 = fi[k2] + fi[k3];
            fi[k2] = f0 - f2; fi[0] = f0 + f2; fi[k3] = f1 - f3; fi[k1] = f1 + f3;
            double g1 = gi[0] - gi[k1], g0 = gi[0] + gi[k1];
            double g3 = SQRT2 * gi[k3], g2 = SQRT2 * gi[k2];
            gi[k2] = g0 - g2; gi[0] = g0 + g2; gi[k3] = g1 - g3; gi[k1] = g1 + g3;
        }
        for (int i = 1; i < kx; i++) {
            const double *tw = MP2_FHT_TW[MP2_FHT_TW_OFFSET[stage] + i - 1];
            double c1 = tw[0], s1 = tw[1], c2 = tw[2], s2 = tw[3];
            for (double *fi = fz + i, *gi = fz + k1 - i; fi < fz + 1024; fi += k4, gi += k4) {
                double a, b, f0, f1, f2, f3, g0, g1, g2, g3;
                b = s2 * fi[k1] - c2 * gi[k1]; a = c2 * fi[k1] + s2 * gi[k1];
                f1 = fi[0] - a; f0 = fi[0] + a; g1 = gi[0] - b; g0 = gi[0] + b;
                b = s2 * fi[k3] - c2 * gi[k3]; a = c2 * fi[k3] + s2 * gi[k3];
                f3 = fi[k2] - a; f2 = fi[k2] + a; g3 = gi[k2] - b; g2 = gi[k2] + b;
                b = s1 * f2 - c1 * g3; a = c1 * f2 + s1 * g3;
                fi[k2] = f0 - a; fi[0] = f0 + a; gi[k3] = g1 - b; gi[k1] = g1 + b;
                b = c1 * g2 - s1 * f3; a = s1 * g2 + c1 * f3;
                gi[k2] = g0 - a; gi[0] = g0 + a; fi[k3] = f1 - b; fi[k1] = f1 + b;
            }
        }
    }
}

/* ------------------------------------------------------------------ psychoacoustic model 1 */

/* ref: psycho_1.c:180-205 */
static double add_db(double a, double b)
{
    double fdiff = 10.0 * (a - b);
    if (fdiff > 990.0) return a;
    if (fdiff < -990.0) return b;
    int idiff = (int)fdiff;
    if (idiff >= 0) return a + MP2_DBTABLE[idiff];
    return b + MP2_DBTABLE[-idiff];
}

typedef struct {
    double x[512];
    int type[512], next[512];
    int map[512];
} psy1_lines;

/* ref: psycho_1.c:160-168 -- later partitions overwrite the shared boundary line; bins above the last
 * line keep the allocator's zero (mem.c:21) */
static void psy1_make_map(int fq, int map[512])
{
    memset(map, 0, 512 * sizeof(int));
    for (int i = 1; i < MP2_SUB_SIZE[fq]; i++)
        for (int j = MP2_LTG_LINE[fq][i - 1]; j <= MP2_LTG_LINE[fq][i]; j++) map[j] = i;
}

/* ref: psycho_1.c:267-340 */
static void psy1_tonal(psy1_lines *p, int *tone)
{
    int last = L_LAST, first = L_LAST, run, last_but_one = L_LAST;
    *tone = L_LAST;
    for (int i = 2; i < 512 - 12; i++) {
        if (p->x[i] > p->x[i - 1] && p->x[i] >= p->x[i + 1]) {
            p->type[i] = T_TONE;
            p->next[i] = L_LAST;
            if (last != L_LAST) p->next[last] = i;
            else first = *tone = i;
            last = i;
        }
    }
    last = L_LAST;
    first = *tone;
    *tone = L_LAST;
    while (first != L_LAST && first != L_STOP) {
        if (first < 3 || first > 500) run = 0;
        else if (first < 63) run = 2;
        else if (first < 127) run = 3;
        else if (first < 255) run = 6;
        else run = 12;
        double max = p->x[first] - 7;
        for (int j = 2; j <= run; j++)
            if (max < p->x[first - j] || max < p->x[first + j]) { p->type[first] = 0; break; }
        if (p->type[first] == T_TONE) {
            int help = first;
            if (*tone == L_LAST) *tone = first;
            while (p->next[help] != L_LAST && (p->next[help] - first) <= run) help = p->next[help];
            help = p->next[help];
            p->next[first] = help;
            if ((first - last) <= run) {
                if (last_but_one != L_LAST) p->next[last_but_one] = first;
            }
            if (first > 1 && first < 500) {
                double tmp = add_db(p->x[first - 1], p->x[first + 1]);
                p->x[first] = add_db(p->x[first], tmp);
            }
            for (int j = 1; j <= run; j++) {
                p->x[first - j] = p->x[first + j] = DBMIN;
                p->next[first - j] = p->next[first + j] = L_STOP;
                p->type[first - j] = p->type[first + j] = 0;
            }
            last_but_one = last;
            last = first;
            first = p->next[first];
        } else {
            if (last != L_LAST) p->next[last] = p->next[first];
            int ll = first;
            first = p->next[first];
            p->next[ll] = L_STOP;
        }
    }
}

/* ref: psycho_1.c:350-400 */
static void psy1_noise(psy1_lines *p, int *noise, int fq, const double *energy)
{
    const int *cbound = MP2_CBOUND[fq];
    int last = L_LAST;
    for (int i = 0; i < MP2_CB_COUNT[fq] - 1; i++) {
        double weight = 0.0, sum = DBMIN;
        for (int j = cbound[i]; j < cbound[i + 1]; j++) {
            if (p->type[j] != T_TONE && p->x[j] != DBMIN) {
                sum = add_db(p->x[j], sum);
                weight += 1073741824 * energy[j] * (double)(j - cbound[i]) / (double)(cbound[i + 1] - cbound[i]);
                p->x[j] = DBMIN;
            }
        }
        int centre;
        if (sum <= DBMIN) centre = (cbound[i + 1] + cbound[i]) / 2;
        else {
            double index = weight * pow(10.0, -0.1 * sum);
            centre = cbound[i] + (int)(index * (double)(cbound[i + 1] - cbound[i]));
        }
        if (p->type[centre] == T_TONE) {
            if (p->type[centre + 1] == T_TONE) centre++;
            else centre--;
        }
        if (last == L_LAST) *noise = centre;
        else {
            p->next[centre] = L_LAST;
            p->next[last] = centre;
        }
        p->x[centre] = sum;
        p->type[centre] = T_NOISE;
        last = centre;
    }
}

/* ref: psycho_1.c:409-470 */
static void psy1_subsample(psy1_lines *p, int fq, int *tone, int *noise)
{
    const double *hear = MP2_LTG_HEAR[fq], *bark = MP2_LTG_BARK[fq];
    for (int pass = 0; pass < 2; pass++) {
        int *head = pass == 0 ? tone : noise;
        int i = *head, old = L_STOP;
        while (i != L_LAST && i != L_STOP) {
            if (p->x[i] < hear[p->map[i]]) {
                p->type[i] = 0;
                p->x[i] = DBMIN;
                if (old == L_STOP) *head = p->next[i];
                else p->next[old] = p->next[i];
            } else old = i;
            i = p->next[i];
        }
    }
    int i = *tone, old = L_STOP;
    while (i != L_LAST && i != L_STOP) {
        if (p->next[i] == L_LAST) break;
        /* NB: like the reference, next may be STOP here; it indexes x/map out of range there as well.
           In the reference power[-100] is undefined behaviour; we stop the walk instead. */
        if (p->next[i] == L_STOP) break;
        int nx = p->next[i];
        if (bark[p->map[nx]] - bark[p->map[i]] < 0.5) {
            if (p->x[nx] > p->x[i]) {
                if (old == L_STOP) *tone = nx;
                else p->next[old] = nx;
                p->type[i] = 0;
                p->x[i] = DBMIN;
                i = nx;
            } else {
                p->type[nx] = 0;
                p->x[nx] = DBMIN;
                p->next[i] = p->next[nx];
                old = i;
            }
        } else {
            old = i;
            i = nx;
        }
    }
}

/* one masker's contribution at threshold line k (ref: psycho_1.c:489-506 tonal, :508-525 noise) */
static inline double psy1_spread(double ltg_x, double dz, double barkm, double xm, int tonal)
{
    double tmps = tonal ? -1.525 - 0.275 * barkm - 4.5 + xm : -1.525 - 0.175 * barkm - 0.5 + xm;
    double vf;
    if (dz < -1) vf = 17 * (dz + 1) - (0.4 * xm + 6);
    else if (dz < 0) vf = (0.4 * xm + 6) * dz;
    else if (dz < 1) vf = (-17 * dz);
    else vf = -(dz - 1) * (17 - 0.15 * xm) - 17;
    return add_db(ltg_x, tmps + vf);
}

/* ref: psycho_1.c:480-532 */
static void psy1_threshold(const psy1_lines *p, int fq, int tone, int noise, int bit_rate, double ltg_x[134])
{
    const double *hear = MP2_LTG_HEAR[fq], *bark = MP2_LTG_BARK[fq];
    for (int k = 1; k < MP2_SUB_SIZE[fq]; k++) {
        double x = DBMIN;
        for (int t = tone; t != L_LAST && t != L_STOP; t = p->next[t]) {
            double dz = bark[k] - bark[p->map[t]];
            if (dz >= -3.0 && dz < 8.0) x = psy1_spread(x, dz, bark[p->map[t]], p->x[t], 1);
        }
        for (int t = noise; t != L_LAST && t != L_STOP; t = p->next[t]) {
            double dz = bark[k] - bark[p->map[t]];
            if (dz >= -3.0 && dz < 8.0) x = psy1_spread(x, dz, bark[p->map[t]], p->x[t], 0);
        }
        if (bit_rate < 96) x = add_db(hear[k], x);
        else x = add_db(hear[k] - 12.0, x);
        ltg_x[k] = x;
    }
}

/* ref: psycho_1.c:22-87 for one channel of one frame.  FFT window = samples [1152n-192, 1152n+832)
 * (ring buffer of 1408 with write offset 256 and read offset +1216: psycho_1.c:30,61-74). */
void mp2o_psy1_frame(const mp2o_cfg *c, const int16_t *pcm, int ch, long frame,
                     const uint8_t scalar_pre[3][32], double smr[32], double ltmin[32], double spike[32])
{
    static _Thread_local psy1_lines P; /* big working set: static, one per thread */
    double xr[1024], energy[513];
    int fq = c->psy_freq, tone, noise;
    psy1_make_map(fq, P.map);
    for (int i = 0; i < 1024; i++) /* ref: psycho_1.c:236-237 */
        xr[i] = pcm_at(pcm, c->nch, ch, frame * 1152 - 192 + i) * MP2_HANN[i];
    mp2o_fht1024(xr);
    energy[0] = xr[0] * xr[0]; /* ref: fft.c:1278-1296 */
    for (int i = 1; i < 512; i++) energy[i] = (xr[i] * xr[i] + xr[1024 - i] * xr[1024 - i]) / 2.0;
    energy[512] = xr[512] * xr[512];
    for (int i = 0; i < 512; i++) { /* ref: psycho_1.c:241-248 */
        P.x[i] = energy[i] < 1E-20 ? -200.0 + POWERNORM : 10 * log10(energy[i]) + POWERNORM;
        P.next[i] = L_STOP;
        P.type[i] = 0;
    }
    for (int sbn = 0; sbn < 32; sbn++) { /* ref: psycho_1.c:252-257 */
        double sum = 1E-20;
        for (int j = 0; j < 16; j++) sum += 1073741824 * energy[sbn * 16 + j];
        spike[sbn] = 10.0 * log10(sum);
    }
    psy1_tonal(&P, &tone);
    psy1_noise(&P, &noise, fq, energy);
    psy1_subsample(&P, fq, &tone, &noise);
    double ltg_x[134];
    psy1_threshold(&P, fq, tone, noise, c->bitrate_kbps / c->nch, ltg_x);
    {   /* ref: psycho_1.c:541-559 */
        int sub_size = MP2_SUB_SIZE[fq], j = 1;
        for (int i = 0; i < c->sblimit; i++) {
            if (j >= sub_size - 1) ltmin[i] = MP2_LTG_HEAR[fq][sub_size - 1];
            else {
                double mn = ltg_x[j];
                while (j < sub_size && (MP2_LTG_LINE[fq][j] >> 4) == i) {
                    if (mn > ltg_x[j]) mn = ltg_x[j];
                    j++;
                }
                ltmin[i] = mn;
            }
        }
    }
    for (int i = 0; i < c->sblimit; i++) { /* ref: psycho_1.c:568-581 with encode_new.c:260-277 folded in */
        unsigned lo = scalar_pre[0][i];
        if (scalar_pre[1][i] < lo) lo = scalar_pre[1][i];
        if (scalar_pre[2][i] < lo) lo = scalar_pre[2][i];
        double mx = MP2_SF_DB[lo];
        if (spike[i] > mx) mx = spike[i];
        smr[i] = mx - ltmin[i];
    }
    for (int i = c->sblimit; i < 32; i++) smr[i] = 0, ltmin[i] = 0;
}

/* ------------------------------------------------------------------ psychoacoustic model 2 */

typedef struct { double energy[513], phi[513]; } p2_spec;

/* ref: psycho_2.c:80-92 (window of block B = raw samples [576B-480, 576B+544), zeros before the stream start)
 * and fft.c:1230-1275 (psycho_2_fft, built without NEWATAN) */
static void psy2_spectrum(const int16_t *pcm, int nch, int ch, long block, p2_spec *S)
{
    const double PI = 3.14159265358979; /* ref: common.h:26 */
    double w[1024];
    for (int j = 0; j < 1024; j++) {
        long idx = 576 * block - 480 + j;
        w[j] = MP2_P2_WINDOW[j] * (idx < 0 ? 0.0 : (double)pcm[idx * nch + ch]);
    }
    mp2o_fht1024(w);
    S->energy[0] = w[0] * w[0];
    S->phi[0] = 0; /* never written by the reference: stays the allocator's zero */
    for (int i = 1, j = 1023; i < 512; i++, j--) {
        double a = w[i], b = w[j];
        S->energy[i] = (a * a + b * b) / 2.0;
        if (S->energy[i] < 0.0005) {
            S->energy[i] = 0.0005;
            S->phi[i] = 0;
        } else S->phi[i] = atan2(-(double)a, (double)b) + PI / 4;
    }
    S->energy[512] = w[512] * w[512];
    S->phi[512] = atan2(0.0, (double)w[512]);
}

/* ref: psycho_2.c:52-254 for one channel of one frame, stateless: the unpredictability measure of block B needs
 * r = sqrt(energy) and phi of blocks B-1 and B-2, which are zeros (not the FFT of silence) before the stream start
 * (psycho_2.c:322-326). */
void mp2o_psy2_frame(const mp2o_cfg *c, const int16_t *pcm, int ch, long frame, double smr[32])
{
    static _Thread_local mp2_psy2_tables T;
    static _Thread_local double T_for = 0;
    static _Thread_local p2_spec S[3];
    const double nmt = 5.5, LN_TO_LOG10 = 0.2302585093;
    double sfreq = (double)c->fs_hz; /* (FLOAT) s_freq[version][idx] * 1000 */
    if (T_for != sfreq) { mp2_psy2_init(&T, sfreq); T_for = sfreq; }
    const double *absthr = MP2_ABSTHR[T.absthr_table];
    double snrtmp[2][32];
    for (int i = 0; i < 2; i++) {
        long B = 2 * frame + i;
        double cm[513], grouped_e[64], grouped_c[64], ecb[64], cb[64], bc[64], nb[64], fthr[513];
        psy2_spectrum(pcm, c->nch, ch, B, &S[0]);
        for (int a = 1; a <= 2; a++)
            if (B - a >= 0) psy2_spectrum(pcm, c->nch, ch, B - a, &S[a]);
        for (int j = 0; j < 513; j++) {
            double r1 = B - 1 >= 0 ? sqrt(S[1].energy[j]) : 0, p1 = B - 1 >= 0 ? S[1].phi[j] : 0;
            double r2 = B - 2 >= 0 ? sqrt(S[2].energy[j]) : 0, p2 = B - 2 >= 0 ? S[2].phi[j] : 0;
            double r_prime = 2.0 * r1 - r2, phi_prime = 2.0 * p1 - p2;
            double rn = sqrt(S[0].energy[j]), phi = S[0].phi[j];
            double temp1 = rn * cos(phi) - r_prime * cos(phi_prime);
            double temp2 = rn * sin(phi) - r_prime * sin(phi_prime);
            double temp3 = rn + fabs(r_prime);
            cm[j] = temp3 != 0 ? sqrt(temp1 * temp1 + temp2 * temp2) / temp3 : 0;
        }
        const double *energy = S[0].energy;
        for (int j = 1; j < 64; j++) grouped_e[j] = grouped_c[j] = 0;
        grouped_e[0] = energy[0];
        grouped_c[0] = energy[0] * cm[0];
        for (int j = 1; j < 513; j++) {
            grouped_e[T.partition[j]] += energy[j];
            grouped_c[T.partition[j]] += energy[j] * cm[j];
        }
        for (int j = 0; j < 64; j++) { /* ref: psycho_2.c:163-177 */
            ecb[j] = 0;
            cb[j] = 0;
            for (int k = 0; k < 64; k++)
                if (T.s[j][k] != 0.0) {
                    ecb[j] += T.s[j][k] * grouped_e[k];
                    cb[j] += T.s[j][k] * grouped_c[k];
                }
            if (ecb[j] != 0) cb[j] = cb[j] / ecb[j];
            else cb[j] = 0;
        }
        for (int j = 0; j < 64; j++) { /* ref: psycho_2.c:183-198 */
            if (cb[j] < .05) cb[j] = 0.05;
            else if (cb[j] > .5) cb[j] = 0.5;
            double tb = -0.434294482 * log((double)cb[j]) - 0.301029996;
            bc[j] = T.tmn[j] * tb + nmt * (1.0 - tb);
            bc[j] = (bc[j] > T.bmax_of[j]) ? bc[j] : T.bmax_of[j];
            bc[j] = exp((double)-bc[j] * LN_TO_LOG10);
        }
        for (int j = 0; j < 64; j++) /* ref: psycho_2.c:205-209 */
            nb[j] = (T.rnorm[j] && T.numlines[j]) ? ecb[j] * bc[j] / (T.rnorm[j] * T.numlines[j]) : 0;
        for (int j = 0; j < 513; j++) {
            double t = nb[T.partition[j]];
            fthr[j] = (t > absthr[j]) ? t : absthr[j];
        }
        for (int j = 0; j < 193; j += 16) { /* ref: psycho_2.c:231-241 */
            double minthres = 60802371420160.0, sum_energy = 0.0;
            for (int k = 0; k < 17; k++) {
                if (minthres > fthr[j + k]) minthres = fthr[j + k];
                sum_energy += energy[j + k];
            }
            snrtmp[i][j / 16] = sum_energy / (minthres * 17.0);
            snrtmp[i][j / 16] = 4.342944819 * log((double)snrtmp[i][j / 16]);
        }
        for (int j = 208; j < 512; j += 16) { /* ref: psycho_2.c:242-251 */
            double minthres = 0.0, sum_energy = 0.0;
            for (int k = 0; k < 17; k++) {
                minthres += fthr[j + k];
                sum_energy += energy[j + k];
            }
            snrtmp[i][j / 16] = sum_energy / minthres;
            snrtmp[i][j / 16] = 4.342944819 * log((double)snrtmp[i][j / 16]);
        }
    }
    for (int i = 0; i < 32; i++) smr[i] = (snrtmp[0][i] > snrtmp[1][i]) ? snrtmp[0][i] : snrtmp[1][i];
}

/* ------------------------------------------------------------------ scalefactor select information */

/* ref: encode_new.c:288-354; rewrites sf[] in place */
static void scfsi_pattern(uint8_t sf[3][32], int sblimit, uint8_t scfsi[32])
{
    static const int pattern[5][5] = {{0x123, 0x122, 0x122, 0x133, 0x123},
                                      {0x113, 0x111, 0x111, 0x444, 0x113},
                                      {0x111, 0x111, 0x111, 0x333, 0x113},
                                      {0x222, 0x222, 0x222, 0x333, 0x123},
                                      {0x123, 0x122, 0x122, 0x133, 0x123}};
    for (int i = 0; i < 32; i++) scfsi[i] = 0;
    for (int i = 0; i < sblimit; i++) {
        int d[2] = {(int)sf[0][i] - (int)sf[1][i], (int)sf[1][i] - (int)sf[2][i]}, cls[2];
        for (int j = 0; j < 2; j++) cls[j] = d[j] <= -3 ? 0 : d[j] < 0 ? 1 : d[j] == 0 ? 2 : d[j] < 3 ? 3 : 4;
        switch (pattern[cls[0]][cls[1]]) {
        case 0x123: scfsi[i] = 0; break;
        case 0x122: scfsi[i] = 3; sf[2][i] = sf[1][i]; break;
        case 0x133: scfsi[i] = 3; sf[1][i] = sf[2][i]; break;
        case 0x113: scfsi[i] = 1; sf[1][i] = sf[0][i]; break;
        case 0x111: scfsi[i] = 2; sf[1][i] = sf[2][i] = sf[0][i]; break;
        case 0x222: scfsi[i] = 2; sf[0][i] = sf[2][i] = sf[1][i]; break;
        case 0x333: scfsi[i] = 2; sf[0][i] = sf[1][i] = sf[2][i]; break;
        case 0x444:
            scfsi[i] = 2;
            if (sf[0][i] > sf[2][i]) sf[0][i] = sf[2][i];
            sf[1][i] = sf[2][i] = sf[0][i];
        }
    }
}

/* ------------------------------------------------------------------ bit allocation */

static inline int row_of(const mp2o_cfg *c, int sb) { return MP2_TAB_ROW[c->tablenum][sb]; }

/* ref: encode_new.c:634-705 with min_mnr = 0 */
static int bits_for_nonoise(const mp2o_cfg *c, double smr[2][32], uint8_t scfsi[2][32], int jsbound)
{
    int nch = c->nch, sblimit = c->sblimit, bbal = 0;
    for (int sb = 0; sb < jsbound; sb++) bbal += nch * MP2_ROW_NBAL[row_of(c, sb)];
    for (int sb = jsbound; sb < sblimit; sb++) bbal += MP2_ROW_NBAL[row_of(c, sb)];
    int req = 32 + bbal + 16; /* banc + bbal + berr (error protection is always on: toolame.c:146) */
    for (int sb = 0; sb < sblimit; sb++)
        for (int ch = 0; ch < (sb < jsbound ? nch : 1); ch++) {
            int row = row_of(c, sb), maxAlloc = (1 << MP2_ROW_NBAL[row]) - 1, ba;
            for (ba = 0; ba < maxAlloc - 1; ba++)
                if (MP2_QC_SNR[MP2_ROW_QC[row][ba]] - smr[ch][sb] >= 0.0) break;
            if (nch == 2 && sb >= jsbound)
                for (; ba < maxAlloc - 1; ba++)
                    if (MP2_QC_SNR[MP2_ROW_QC[row][ba]] - smr[1 - ch][sb] >= 0.0) break;
            if (ba > 0) {
                int q = MP2_ROW_QC[row][ba];
                int smp = 12 * MP2_QC_NCODE[q] * MP2_QC_BITS[q];
                int sel = 2, sc = 6 * MP2_SCFSI_NSF[scfsi[ch][sb]];
                if (nch == 2 && sb >= jsbound) { sel += 2; sc += 6 * MP2_SCFSI_NSF[scfsi[1 - ch][sb]]; }
                req += smp + sel + sc;
            }
        }
    return req;
}

/* ref: encode_new.c:1078-1187 (greedy loop) with maxmnr_new :1061-1077 inlined; returns leftover bits */
static int greedy_alloc(const mp2o_cfg *c, double smr[2][32], uint8_t scfsi[2][32], int jsbound, int adb,
                        uint8_t bit_alloc[2][32])
{
    int nch = c->nch, sblimit = c->sblimit, bbal = 0;
    double mnr[2][32];
    char used[2][32];
    for (int sb = 0; sb < jsbound; sb++) bbal += nch * MP2_ROW_NBAL[row_of(c, sb)];
    for (int sb = jsbound; sb < sblimit; sb++) bbal += MP2_ROW_NBAL[row_of(c, sb)];
    int ad = adb - (bbal + 16 + 32);
    memset(bit_alloc, 0, 64);
    for (int sb = 0; sb < sblimit; sb++)
        for (int ch = 0; ch < nch; ch++) { mnr[ch][sb] = MP2_QC_SNR[0] - smr[ch][sb]; used[ch][sb] = 0; }
    int bspl = 0, bscf = 0, bsel = 0;
    for (;;) {
        int min_sb = -1, min_ch = -1;
        double small = 999999.0;
        for (int ch = 0; ch < nch; ch++)
            for (int sb = 0; sb < sblimit; sb++)
                if (used[ch][sb] != 2 && small > mnr[ch][sb]) { small = mnr[ch][sb]; min_sb = sb; min_ch = ch; }
        if (min_sb < 0) break;
        int row = row_of(c, min_sb), oth = 1 - min_ch;
        int qn = MP2_ROW_QC[row][bit_alloc[min_ch][min_sb] + 1];
        int increment = 12 * MP2_QC_NCODE[qn] * MP2_QC_BITS[qn], scale = 0, seli = 0;
        if (used[min_ch][min_sb]) {
            int q = MP2_ROW_QC[row][bit_alloc[min_ch][min_sb]];
            increment -= 12 * MP2_QC_NCODE[q] * MP2_QC_BITS[q];
        } else {
            seli = 2;
            scale = 6 * MP2_SCFSI_NSF[scfsi[min_ch][min_sb]];
            if (nch == 2 && min_sb >= jsbound) { seli += 2; scale += 6 * MP2_SCFSI_NSF[scfsi[oth][min_sb]]; }
        }
        if (ad >= bspl + bscf + bsel + seli + scale + increment) {
            int ba = ++bit_alloc[min_ch][min_sb];
            bspl += increment; bscf += scale; bsel += seli;
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
    return ad - (bspl + bscf + bsel);
}

/* ------------------------------------------------------------------ CRCs */

static void crc_update(unsigned data, unsigned length, unsigned *crc, unsigned top, unsigned poly, unsigned mask)
{   /* ref: crc.c:43-56 (16 bit, 0x8005) and :100-113 (8 bit, 0x1D) */
    unsigned masking = 1u << length;
    while ((masking >>= 1)) {
        unsigned carry = *crc & top;
        *crc <<= 1;
        if (!carry ^ !(data & masking)) *crc ^= poly;
    }
    *crc &= mask;
}

/* ref: crc.c:58-98, `packed` = group index 0..3 */
static unsigned scf_crc(const mp2o_cfg *c, uint8_t bit_alloc[2][32], uint8_t scfsi[2][32], uint8_t sf[2][3][32], int packed)
{
    static const int f[5] = {0, 4, 8, 16, 30};
    int first = f[packed], last = f[packed + 1];
    if (last > c->sblimit) last = c->sblimit;
    unsigned crc = 0;
    for (int i = first; i < last; i++)
        for (int k = 0; k < c->nch; k++)
            if (bit_alloc[k][i]) switch (scfsi[k][i]) {
                case 0: for (int j = 0; j < 3; j++) crc_update(sf[k][j][i] >> 3, 3, &crc, 0x80, 0x1D, 0xff); break;
                case 1: case 3:
                    crc_update(sf[k][0][i] >> 3, 3, &crc, 0x80, 0x1D, 0xff);
                    crc_update(sf[k][2][i] >> 3, 3, &crc, 0x80, 0x1D, 0xff);
                    break;
                case 2: crc_update(sf[k][0][i] >> 3, 3, &crc, 0x80, 0x1D, 0xff);
                }
    return crc;
}

/* ------------------------------------------------------------------ bit writer */

typedef struct { uint8_t *buf; long bitpos; } bitw;
static void putbits(bitw *w, unsigned val, int n) /* MSB first (ref: bitstream.c:127-150) */
{
    for (int i = n - 1; i >= 0; i--, w->bitpos++)
        if ((val >> i) & 1) w->buf[w->bitpos >> 3] |= (uint8_t)(0x80 >> (w->bitpos & 7));
}

/* ------------------------------------------------------------------ one frame, up to its own ScF-CRC */

typedef struct {
    uint8_t bytes[2048];
    uint8_t scfcrc_own[4];
} frame_out;

static void encode_one(const mp2o_cfg *c, const int16_t *pcm, long n, const uint8_t *xpad_rec, frame_out *fo, mp2o_tap *tap)
{
    static _Thread_local mp2o_tap T; /* big: static, one per thread */
    mp2o_tap *t = tap ? tap : &T;
    memset(t, 0, sizeof *t);
    int nch = c->nch, sblimit = c->sblimit;
    int xpad_len = (xpad_rec && c->pad_len) ? xpad_rec[c->pad_len] : 0;
    /* ref: toolame.c:292-302 */
    int adb = 8 * c->lg_frame - (c->dab_ext * 8 + (xpad_len ? xpad_len : 2) * 8);

    static _Thread_local double jsamp[36][32];
    for (int ch = 0; ch < nch; ch++) {
        mp2o_filterbank_frame(pcm, nch, ch, n, t->sb_sample[ch]);
        scalefactors(t->sb_sample[ch], sblimit, t->scalar_pre[ch]);
    }
    if (c->mode == 1) { /* ref: toolame.c:332-337, encode_new.c:237-246 */
        for (int b = 0; b < 36; b++)
            for (int sb = 0; sb < 32; sb++)
                jsamp[b][sb] = sb < sblimit ? .5 * (t->sb_sample[0][b][sb] + t->sb_sample[1][b][sb]) : 0.0;
        scalefactors(jsamp, sblimit, t->j_scale);
    }
    /* ref: toolame.c:361-452 (psy model switch); models 1 and 2 are restated */
    for (int ch = 0; ch < nch; ch++) {
        if (c->psy == 0) { /* ref: psycho_0.c:27-69: SMR from the smallest scalefactor index and the subband's lowest ATH */
            static _Thread_local double ath_min[32], ath_for = 0;
            if (ath_for != (double)c->fs_hz) { mp2_psy0_init(ath_min, (double)c->fs_hz); ath_for = (double)c->fs_hz; }
            for (int sb = 0; sb < 32; sb++) {
                int mn = t->scalar_pre[ch][0][sb];
                for (int gr = 1; gr < 3; gr++)
                    if (mn > t->scalar_pre[ch][gr][sb]) mn = t->scalar_pre[ch][gr][sb];
                t->smr[ch][sb] = 2.0 * (30.0 - mn) - ath_min[sb];
            }
        } else if (c->psy == 2) mp2o_psy2_frame(c, pcm, ch, n, t->smr[ch]);
        else mp2o_psy1_frame(c, pcm, ch, n, t->scalar_pre[ch], t->smr[ch], t->ltmin[ch], t->spike[ch]);
    }

    memcpy(t->scalar, t->scalar_pre, sizeof t->scalar);
    for (int ch = 0; ch < nch; ch++) scfsi_pattern(t->scalar[ch], sblimit, t->scfsi[ch]);

    /* ref: encode_new.c:803-819 */
    int mode = c->mode, mode_ext = c->mode_ext, jsbound = c->jsbound;
    if (c->mode == 1) {
        mode = 0; mode_ext = 0; jsbound = sblimit;
        if (bits_for_nonoise(c, t->smr, t->scfsi, jsbound) > adb) {
            mode = 1;
            mode_ext = 4;
            int rq;
            do {
                --mode_ext;
                jsbound = MP2_JSBOUND[mode_ext];
                rq = bits_for_nonoise(c, t->smr, t->scfsi, jsbound);
            } while (rq > adb && mode_ext > 0);
        }
    }
    t->mode = mode; t->mode_ext = mode_ext; t->jsbound = jsbound;
    int left = greedy_alloc(c, t->smr, t->scfsi, jsbound, adb, t->bit_alloc);
    t->adb_left = left;

    /* ref: crc.c:12-41 */
    unsigned crc = 0xffff;
    crc_update((unsigned)c->bitrate_index, 4, &crc, 0x8000, 0x8005, 0xffff);
    crc_update((unsigned)c->sfreq_idx, 2, &crc, 0x8000, 0x8005, 0xffff);
    crc_update(0, 1, &crc, 0x8000, 0x8005, 0xffff); /* padding */
    crc_update(0, 1, &crc, 0x8000, 0x8005, 0xffff); /* extension */
    crc_update((unsigned)mode, 2, &crc, 0x8000, 0x8005, 0xffff);
    crc_update((unsigned)mode_ext, 2, &crc, 0x8000, 0x8005, 0xffff);
    crc_update(0, 1, &crc, 0x8000, 0x8005, 0xffff); /* copyright */
    crc_update(0, 1, &crc, 0x8000, 0x8005, 0xffff); /* original */
    crc_update(0, 2, &crc, 0x8000, 0x8005, 0xffff); /* emphasis */
    for (int i = 0; i < sblimit; i++)
        for (int k = 0; k < (i < jsbound ? nch : 1); k++)
            crc_update(t->bit_alloc[k][i], (unsigned)MP2_ROW_NBAL[row_of(c, i)], &crc, 0x8000, 0x8005, 0xffff);
    for (int i = 0; i < sblimit; i++)
        for (int k = 0; k < nch; k++)
            if (t->bit_alloc[k][i]) crc_update(t->scfsi[k][i], 2, &crc, 0x8000, 0x8005, 0xffff);
    t->crc16 = crc;

    memset(fo->bytes, 0, sizeof fo->bytes);
    bitw w = {fo->bytes, 0};
    /* ref: encode_new.c:356-373 */
    putbits(&w, 0xfff, 12); putbits(&w, (unsigned)c->version, 1); putbits(&w, 4 - 2, 2); putbits(&w, 0, 1);
    putbits(&w, (unsigned)c->bitrate_index, 4); putbits(&w, (unsigned)c->sfreq_idx, 2);
    putbits(&w, 0, 1); putbits(&w, 0, 1);
    putbits(&w, (unsigned)mode, 2); putbits(&w, (unsigned)mode_ext, 2);
    putbits(&w, 0, 1); putbits(&w, 0, 1); putbits(&w, 0, 2);
    putbits(&w, crc, 16); /* ref: toolame.c:478-480 */
    for (int sb = 0; sb < sblimit; sb++) /* ref: encode_new.c:383-399 */
        for (int ch = 0; ch < (sb < jsbound ? nch : 1); ch++)
            putbits(&w, t->bit_alloc[ch][sb], MP2_ROW_NBAL[row_of(c, sb)]);
    for (int sb = 0; sb < sblimit; sb++) /* ref: encode_new.c:413-444 */
        for (int ch = 0; ch < nch; ch++)
            if (t->bit_alloc[ch][sb]) putbits(&w, t->scfsi[ch][sb], 2);
    for (int sb = 0; sb < sblimit; sb++)
        for (int ch = 0; ch < nch; ch++)
            if (t->bit_alloc[ch][sb]) switch (t->scfsi[ch][sb]) {
                case 0: for (int gr = 0; gr < 3; gr++) putbits(&w, t->scalar[ch][gr][sb], 6); break;
                case 1: case 3: putbits(&w, t->scalar[ch][0][sb], 6); putbits(&w, t->scalar[ch][2][sb], 6); break;
                case 2: putbits(&w, t->scalar[ch][0][sb], 6);
                }
    /* ref: encode_new.c:479-547 */
    for (int b = 0; b < 36; b++)
        for (int sb = 0; sb < sblimit; sb++)
            for (int ch = 0; ch < (sb < jsbound ? nch : 1); ch++)
                if (t->bit_alloc[ch][sb]) {
                    int gr = b / 12;
                    double d;
                    if (nch == 2 && sb >= jsbound) d = jsamp[b][sb] / MP2_SCALEFACTOR[t->j_scale[gr][sb]];
                    else d = t->sb_sample[ch][b][sb] / MP2_SCALEFACTOR[t->scalar[ch][gr][sb]];
                    int q = MP2_ROW_QC[row_of(c, sb)][t->bit_alloc[ch][sb]];
                    d = d * MP2_QC_A[q] + MP2_QC_B[q];
                    int sig = 1;
                    if (!(d >= 0)) { sig = 0; d += 1.0; }
                    uint32_t v = (uint32_t)(d * (double)MP2_QC_MSB[q]);
                    if (sig) v |= (uint32_t)MP2_QC_MSB[q];
                    t->q[ch][b][sb] = v;
                }
    /* ref: encode_new.c:560-598 */
    for (int gr = 0; gr < 3; gr++)
        for (int j = 0; j < 12; j += 3)
            for (int sb = 0; sb < sblimit; sb++)
                for (int ch = 0; ch < (sb < jsbound ? nch : 1); ch++)
                    if (t->bit_alloc[ch][sb]) {
                        int q = MP2_ROW_QC[row_of(c, sb)][t->bit_alloc[ch][sb]];
                        const uint32_t *s0 = &t->q[ch][gr * 12 + j][sb];
                        if (MP2_QC_NCODE[q] == 3) {
                            for (int x = 0; x < 3; x++) putbits(&w, s0[32 * x], MP2_QC_BITS[q]);
                        } else {
                            unsigned y = (unsigned)MP2_QC_STEPS[q];
                            putbits(&w, s0[0] + s0[32] * y + s0[64] * y * y, MP2_QC_BITS[q]);
                        }
                    }
    w.bitpos += left; /* ref: toolame.c:509-512, zero stuffing */
    if (xpad_len) /* ref: toolame.c:515-524 */
        for (int i = c->pad_len - xpad_len; i < c->pad_len - 2; i++) putbits(&w, xpad_rec[i], 8);
    for (int i = c->dab_ext - 1; i >= 0; i--) { /* ref: toolame.c:527-542 (own CRC; the shift is applied by the caller) */
        unsigned v = scf_crc(c, t->bit_alloc, t->scfsi, t->scalar, i);
        fo->scfcrc_own[i] = t->scfcrc_own[i] = (uint8_t)v;
        putbits(&w, v, 8);
    }
    if (xpad_len) { putbits(&w, xpad_rec[c->pad_len - 2], 8); putbits(&w, xpad_rec[c->pad_len - 1], 8); }
    else putbits(&w, 0, 16); /* ref: toolame.c:544-551 */
    if (w.bitpos != 8L * c->lg_frame) abort(); /* every frame is exactly lg_frame bytes */
}

int mp2o_encode(const mp2o_cfg *c, const int16_t *pcm, long n_frames_total, long f0, long f1,
                const uint8_t *xpad, uint8_t *out, mp2o_tap *taps)
{
    if (c->psy < 0 || c->psy > 2) return -1;
    static _Thread_local frame_out fo;
    for (long n = f0; n < f1 + 1 && n < n_frames_total; n++) {
        const uint8_t *rec = xpad ? xpad + (size_t)n * (c->pad_len + 1) : NULL;
        encode_one(c, pcm, n, rec, &fo, (taps && n < f1) ? &taps[n - f0] : NULL);
        if (n < f1) memcpy(out + (size_t)(n - f0) * c->lg_frame, fo.bytes, (size_t)c->lg_frame);
        if (n > f0) { /* ref: toolame.c:527-539: frame n overwrites the ScF-CRC field of frame n-1 */
            uint8_t *prev = out + (size_t)(n - 1 - f0) * c->lg_frame;
            for (int i = c->dab_ext - 1, k = 0; i >= 0; i--, k++) prev[c->lg_frame - 2 - c->dab_ext + k] = fo.scfcrc_own[i];
        }
    }
    return 0;
}

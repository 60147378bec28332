/*
 * ref_driver.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Drives the unmodified reference libtoolame-dab (compiled from
 * /root/reference by oracle/Makefile) with the call sequence odr-audioenc
 * uses for `--dab`:
 *   start-up order   src/odr-audioenc.cpp:687-721
 *                    (init -> samplerate -> psy -> mode -> bitrate -> pad)
 *   per-frame        src/odr-audioenc.cpp:1139-1158 (de-interleave into
 *                    short[2][1152], toolame_encode_frame)
 *   end              toolame_finish (src/odr-audioenc.cpp:1161)
 * and concatenates every returned chunk, so the output file is the stream
 * n_frames x lg_frame bytes long the reference produces for that PCM.
 *
 * usage: ref_driver FS MODE BITRATE PSY PADLEN IN.pcm OUT.mp2 [--xpad F] [--tap F] [--tapbig F] [--bench]
 *   IN.pcm   interleaved s16le, nch channels (nch = 1 for MODE m, else 2)
 *   --xpad   file of n_frames records of PADLEN+1 bytes; the last byte of a
 *            record is the used length handed to toolame_encode_frame as
 *            xpad_len (convention of src/odr-audioenc.cpp:823-852)
 *   --tap    per-frame ref_tap_small records (oracle/ref_tap.c)
 *   --tapbig per-frame sb_sample (double[2][3][12][32]) then subband (u32[2][3][12][32])
 *   --bench  print {"frames":N,"seconds":T} for the encode loop only (PCM preloaded)
 *
 * One process per stream: the reference keeps its state in statics and
 * cannot be re-initialised.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>
#include "toolame.h"

typedef struct {
    int32_t mode, mode_ext, jsbound, sblimit, nch, tablenum, bitrate_index, dab_extension;
    uint32_t scalar[2][3][32];
    uint32_t j_scale[3][32];
    uint32_t scfsi[2][32];
    uint32_t bit_alloc[2][32];
    double smr[2][32];
    double max_sc[2][32];
} ref_tap_small;
void ref_tap_read_small(ref_tap_small *t);
const double *ref_tap_sb_sample(void);
const unsigned int *ref_tap_subband(void);

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char **argv)
{
    if (argc < 8) {
        fprintf(stderr, "usage: %s FS MODE BITRATE PSY PADLEN IN.pcm OUT.mp2 [--xpad F] [--tap F] [--tapbig F] [--bench]\n", argv[0]);
        return 2;
    }
    long fs = atol(argv[1]);
    char mode = argv[2][0];
    int brate = atoi(argv[3]);
    int psy = atoi(argv[4]);
    int padlen = atoi(argv[5]);
    const char *in_path = argv[6], *out_path = argv[7];
    const char *xpad_path = NULL, *tap_path = NULL, *tapbig_path = NULL;
    int bench = 0, repeat = 1;
    for (int i = 8; i < argc; i++) {
        if (!strcmp(argv[i], "--xpad") && i + 1 < argc) xpad_path = argv[++i];
        else if (!strcmp(argv[i], "--tap") && i + 1 < argc) tap_path = argv[++i];
        else if (!strcmp(argv[i], "--tapbig") && i + 1 < argc) tapbig_path = argv[++i];
        else if (!strcmp(argv[i], "--repeat") && i + 1 < argc) repeat = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--bench")) bench = 1;
        else { fprintf(stderr, "bad arg %s\n", argv[i]); return 2; }
    }
    int nch = (mode == 'm') ? 1 : 2;

    FILE *fi = fopen(in_path, "rb");
    if (!fi) { perror(in_path); return 1; }
    fseek(fi, 0, SEEK_END);
    long nbytes = ftell(fi);
    fseek(fi, 0, SEEK_SET);
    long n_frames = nbytes / (2L * nch * 1152);
    int16_t *pcm = (int16_t *)malloc((size_t)n_frames * nch * 1152 * 2 + 16);
    if (fread(pcm, 2, (size_t)n_frames * nch * 1152, fi) != (size_t)n_frames * nch * 1152) { perror("read"); return 1; }
    fclose(fi);

    unsigned char *xpad = NULL;
    if (xpad_path && padlen > 0) {
        FILE *fx = fopen(xpad_path, "rb");
        if (!fx) { perror(xpad_path); return 1; }
        xpad = (unsigned char *)malloc((size_t)n_frames * (padlen + 1));
        if (fread(xpad, (size_t)padlen + 1, (size_t)n_frames, fx) != (size_t)n_frames) { perror("xpad read"); return 1; }
        fclose(fx);
    }

    if (toolame_init()) return 3;
    if (toolame_set_samplerate(fs)) return 3;
    if (toolame_set_psy_model(psy)) return 3;
    if (toolame_set_channel_mode(mode)) return 3;
    if (toolame_set_bitrate(brate)) return 3;
    if (toolame_set_pad(padlen)) return 3;

    FILE *fo = fopen(out_path, "wb");
    if (!fo) { perror(out_path); return 1; }
    FILE *ft = tap_path ? fopen(tap_path, "wb") : NULL;
    FILE *fb = tapbig_path ? fopen(tapbig_path, "wb") : NULL;

    static short buf[2][1152];
    static unsigned char out[8192];
    unsigned char zero_pad[8] = {0};
    long total = 0;
    double t0 = now_s();
    for (int rep = 0; rep < repeat; rep++)
    for (long f = 0; f < n_frames; f++) {
        const int16_t *p = pcm + (size_t)f * nch * 1152;
        if (nch == 1) {
            memcpy(buf[0], p, 1152 * 2);
        } else {
            for (int i = 0; i < 1152; i++) { buf[0][i] = p[2 * i]; buf[1][i] = p[2 * i + 1]; }
        }
        unsigned char *xp = zero_pad;
        size_t xlen = 0;
        if (xpad) { xp = xpad + (size_t)f * (padlen + 1); xlen = xp[padlen]; }
        int n = toolame_encode_frame(buf, xp, xlen, out, 4092);
        if (n > 0 && !bench) fwrite(out, 1, (size_t)n, fo);
        total += n;
        if (ft) { ref_tap_small t; ref_tap_read_small(&t); fwrite(&t, sizeof t, 1, ft); }
        if (fb) {
            fwrite(ref_tap_sb_sample(), sizeof(double), 2 * 3 * 12 * 32, fb);
            fwrite(ref_tap_subband(), sizeof(unsigned int), 2 * 3 * 12 * 32, fb);
        }
    }
    double t1 = now_s();
    int n = toolame_finish(out, 4092);
    if (n > 0 && !bench) fwrite(out, 1, (size_t)n, fo);
    total += n;
    fclose(fo);
    if (ft) fclose(ft);
    if (fb) fclose(fb);
    if (bench)
        printf("{\"frames\": %ld, \"seconds\": %.6f, \"bytes\": %ld}\n", n_frames * repeat, t1 - t0, total);
    return 0;
}

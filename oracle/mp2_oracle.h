/*
 * mp2_oracle.h -- TEST INFRASTRUCTURE ONLY (the parity oracle; never linked into the product).
 *
 * CPU restatement, in plain C, of the libtoolame-dab MPEG Layer II DAB encode
 * path (toolame_encode_frame, toolame.c:267-554, and everything it calls with
 * -DNEWENCODE), re-formulated STATELESS: every frame is computed from the PCM
 * of the stream alone (zero history before sample 0), so any frame range can
 * be produced independently -- the property the B200 batch path relies on.
 * Each function in mp2_oracle.c cites the reference lines it follows.
 *
 * Pinning: tests/test_oracle_vs_ref.py checks this oracle byte-for-byte and
 * tap-for-tap against the reference compiled unmodified from /root/reference
 * (oracle/_ref, see oracle/Makefile) and against the committed fixtures in
 * tests/golden/ generated from that build (the reference itself ships no
 * tests or golden vectors: SURVEY.md section 4).
 */
#ifndef MP2_ORACLE_H
#define MP2_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int fs_hz, version, sfreq_idx;   /* version 1 = MPEG-1, 0 = LSF (common.c:118-144) */
    int mode, mode_ext;              /* header.mode as configured (0 s,1 j,2 d,3 m) */
    int nch, bitrate_kbps, bitrate_index;
    int tablenum, sblimit, jsbound;  /* encode_new.c:104-156, common.c:76-93 */
    int dab_ext, lg_frame;           /* toolame.c:225-232, availbits.c:37-67 */
    int psy, pad_len;
    int psy_freq;                    /* psy-1 table index (psycho_1.c:42-48) */
} mp2o_cfg;

/* per-frame intermediate results, for stage-level parity */
typedef struct {
    double sb_sample[2][36][32];     /* [ch][gr*12+bl][sb] */
    uint8_t scalar_pre[2][3][32];    /* before sf_transmission_pattern */
    uint8_t scalar[2][3][32];        /* after */
    uint8_t j_scale[3][32];
    uint8_t scfsi[2][32];
    uint8_t bit_alloc[2][32];
    double smr[2][32];
    double ltmin[2][32], spike[2][32];
    uint32_t q[2][36][32];           /* quantised samples */
    int32_t mode, mode_ext, jsbound, adb_left;
    uint32_t crc16;
    uint8_t scfcrc_own[4];           /* CRC_calcDAB of THIS frame, index = group */
} mp2o_tap;

/* returns 0, or <0 for an illegal parameter (the reference exit()s on a bad bitrate: common.c:110-115) */
int mp2o_configure(mp2o_cfg *c, long fs_hz, char mode, int bitrate_kbps, int psy, int pad_len);

/*
 * Encode frames [f0,f1) of a stream whose PCM (interleaved s16, nch channels) starts at sample 0
 * and holds n_frames_total frames.  xpad: NULL or n_frames_total records of pad_len+1 bytes, last
 * byte = used length (src/odr-audioenc.cpp:823-852).  out: (f1-f0)*lg_frame bytes in FINAL stream
 * form, i.e. frame n carries the ScF-CRC of frame n+1 and the last frame of the stream its own
 * (toolame.c:527-542).  taps: NULL or f1-f0 records.
 */
int mp2o_encode(const mp2o_cfg *c, const int16_t *pcm, long n_frames_total, long f0, long f1,
                const uint8_t *xpad, uint8_t *out, mp2o_tap *taps);

/* stage entry points used by the stage-level tests */
void mp2o_psy2_frame(const mp2o_cfg *c, const int16_t *pcm, int ch, long frame, double smr[32]);
void mp2o_filterbank_frame(const int16_t *pcm, int nch, int ch, long frame, double sb[36][32]);
void mp2o_fht1024(double *x);
void mp2o_psy1_frame(const mp2o_cfg *c, const int16_t *pcm, int ch, long frame,
                     const uint8_t scalar_pre[3][32], double smr[32], double ltmin[32], double spike[32]);

#ifdef __cplusplus
}
#endif
#endif

/*
 * toolame_b200.h -- C ABI of the B200-native MPEG Layer II (MP2) DAB encode path.
 *
 * Two groups of entry points, all `extern "C"`, plain pointers and sizes:
 *
 *  (1) the libtoolame-dab API, unchanged (drop-in for the reference's
 *      libtoolame-dab/toolame.h:13-48, export list libtoolame-dab.sym:1-9): the same nine
 *      symbols with the same argument meaning, return convention and output chunking.  They are
 *      declared in include/toolame.h (same prototypes as the reference header) and implemented on
 *      top of group (2) with a one-frame batch per call.
 *
 *  (2) the batch API (new): encodes many frames of one stream per call, data-parallel over frames
 *      and channels on the GPU.  It replaces N successive calls of toolame_encode_frame
 *      (libtoolame-dab/toolame.c:267-554) and produces the same bytes, frame-aligned: frame n of
 *      the output carries the ScF-CRC of frame n+1 (toolame.c:527-542), the last frame of the
 *      stream its own.
 *
 * Every function returns 0 on success or a negative TLB_E_* code; nothing here calls exit()
 * (the reference does for an illegal bitrate: common.c:110-115).  There is no CPU fallback: without a
 * CUDA device tlb_batch_create fails with TLB_E_CUDA.
 */
#ifndef TOOLAME_B200_H
#define TOOLAME_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define TLB_API __attribute__((visibility("default")))
#else
#define TLB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define TLB_E_PARAM   (-1)  /* illegal sample rate / mode / bitrate / psy model / pad length */
#define TLB_E_CUDA    (-2)  /* CUDA runtime error (tlb_last_error() has the text) */
#define TLB_E_ARG     (-3)  /* NULL pointer, bad sizes, bad history */
#define TLB_E_UNSUPP  (-4)  /* legal for the reference, not (yet) built here (psy model 3; 44.1/22.05 kHz padding) */

/* Stream parameters: what the reference takes through toolame_set_samplerate / _set_channel_mode /
 * _set_bitrate / _set_psy_model / _set_pad (toolame.c:168-262). */
typedef struct {
    int32_t sample_rate;   /* Hz: 48000, 24000 (DAB); 32000 / 16000 also accepted */
    int32_t channel_mode;  /* 's', 'd', 'j' or 'm' */
    int32_t bitrate;       /* kbit/s, 0 = default of the reference (toolame.c:217-218) */
    int32_t psy_model;     /* 0, 1 (the odr-audioenc default) or 2 */
    int32_t pad_len;       /* X-PAD + F-PAD bytes reserved per record, 0 = none (toolame_set_pad) */
} tlb_config;

/* Derived per-stream constants (common.c:76-93, availbits.c:37-67, encode_new.c:104-156). */
typedef struct {
    int32_t nch, lg_frame, sblimit, tablenum, dab_ext, version, bitrate_index, sfreq_idx;
    int32_t samples_per_frame;  /* 1152 */
    int32_t halo_samples;       /* PCM history a frame needs before its first sample (480; 1632 with psy model 2) */
} tlb_info;

/* Validate a configuration and derive its constants on the host -- no CUDA call, usable without a GPU (what
 * toolame_set_bitrate needs: toolame.c:212-237 / BitrateIndex common.c:95-116).  info may be NULL. */
TLB_API int tlb_config_check(const tlb_config *cfg, tlb_info *info);

typedef struct tlb_batch tlb_batch;

/* Create an encoder for one stream configuration on CUDA device `device`.
 * max_chunk_frames = frames per kernel launch (0 = default: 75 776); device working memory is
 * about 40 kB per chunk frame (stereo, psy model 1) for each chunk slot in use.  Slots are allocated on first use
 * and sized for the path that uses them: two of max_chunk_frames on the device-resident path, three of
 * min(max_chunk_frames, 14 208) on the host-buffer path (never more than the call's n_frames). */
TLB_API int tlb_batch_create(tlb_batch **out, const tlb_config *cfg, int device, size_t max_chunk_frames);
TLB_API void tlb_batch_destroy(tlb_batch *b);
TLB_API int tlb_batch_info(const tlb_batch *b, tlb_info *info);
TLB_API const char *tlb_last_error(void);

/*
 * Encode n_frames frames from HOST memory (host<->device copies included, chunks double-buffered).
 *   pcm               interleaved s16 (WAV order), nch channels; pcm[0] = first sample of the first
 *                     frame to encode
 *   history_samples   samples per channel that are valid BEFORE pcm (pcm[-history_samples*nch ..]);
 *                     0 = stream start (the reference's zero history), otherwise >= halo_samples
 *                     (tlb_info: 480, or 1632 with psy model 2)
 *   has_next          1: pcm holds n_frames+1 frames and frame n_frames only lends its ScF-CRC to
 *                        the last emitted frame (a time chunk in the middle of a stream);
 *                     0: stream end, the last frame keeps its own ScF-CRC
 *   xpad              NULL, or n_frames+has_next records of pad_len+1 bytes in odr-audioenc's layout
 *                     (src/odr-audioenc.cpp:823-852): data right-aligned in the first pad_len bytes,
 *                     last byte = used length (0 or 2..pad_len; 1 is taken as 0 and larger values as
 *                     pad_len), i.e. what the caller passes to toolame_encode_frame as xpad_data / xpad_len
 *   out               n_frames * lg_frame bytes
 * pcm / out may be pageable or pinned (tlb_host_alloc); pinned memory makes the copies asynchronous.
 */
TLB_API int tlb_batch_encode(tlb_batch *b, const int16_t *pcm, size_t n_frames, size_t history_samples,
                     int has_next, const uint8_t *xpad, uint8_t *out);

/* As tlb_batch_encode, but returns once the work is queued; tlb_batch_sync waits.  With pageable host memory the
 * copies still block; pinned memory (tlb_host_alloc) makes the call fully asynchronous. */
TLB_API int tlb_batch_encode_async(tlb_batch *b, const int16_t *pcm, size_t n_frames, size_t history_samples,
                           int has_next, const uint8_t *xpad, uint8_t *out);

/* Many services (streams) at once: the multi-service form of the loop in src/odr-audioenc.cpp:819-1276 run once per
 * service.  Services of different configurations run side by side on the GPU.  An entry is a whole stream from its
 * start (history_samples = 0, has_next = 0) or a time piece of one (as tlb_batch_encode: pcm[0] = first sample of the
 * piece's first frame, history_samples valid samples before it, has_next = one more frame follows in pcm) -- what a
 * rank of a multi-GPU feeder is handed by a plan like odr_audioenc_b200/sharding.py's ensemble_shards.
 * xpad may be NULL per service; out receives n_frames * lg_frame bytes per service. */
typedef struct {
    tlb_config cfg;
    const int16_t *pcm;   /* interleaved s16, (n_frames + has_next) * 1152 * nch samples from pcm[0] on */
    size_t n_frames;
    const uint8_t *xpad;  /* NULL or n_frames + has_next records of pad_len + 1 bytes */
    uint8_t *out;
    size_t history_samples;
    int32_t has_next;
} tlb_service;
TLB_API int tlb_encode_services(const tlb_service *sv, size_t n, int device, size_t chunk_frames);

/* Same, with pcm / xpad / out already in DEVICE memory of the encoder's GPU; d_pcm must be readable
 * from d_pcm - history_samples*nch.  Asynchronous on the encoder's stream; tlb_batch_sync waits. */
TLB_API int tlb_batch_encode_device(tlb_batch *b, const int16_t *d_pcm, size_t n_frames, size_t history_samples,
                            int has_next, const uint8_t *d_xpad, uint8_t *d_out);
TLB_API int tlb_batch_sync(tlb_batch *b);

/* The step before the encoder in odr-audioenc's loop, on the device (src/odr-audioenc.cpp:1020-1055): gain
 * correction of interleaved s16 PCM in place (gain_db = 0 leaves it untouched) and the per-frame peak levels
 * d_peaks[frame][2] = (left, right), computed on (left, right) sample pairs also in mono, as the reference does.
 * Asynchronous on the encoder's stream (the one tlb_batch_encode_device uses), so it can be queued right before it. */
TLB_API int tlb_batch_gain_peak_device(tlb_batch *b, int16_t *d_pcm, size_t n_frames, double gain_db, int16_t *d_peaks);

/* The same for the host-buffer entry points: from now on tlb_batch_encode / _async apply gain_db on the device to
 * the staged PCM (the caller's buffer is not modified) and, if peaks is not NULL, store the per-frame peak levels
 * peaks[frame][2] there (frame counted from the first frame of each encode call).  gain_db = 0, peaks = NULL: off. */
TLB_API int tlb_batch_set_gain(tlb_batch *b, double gain_db, int16_t *peaks);

/* CUDA stream (cudaStream_t) the device-resident calls run on, for event timing by the caller. */
TLB_API void *tlb_batch_stream(tlb_batch *b);
/* Number of kernel launches issued by this encoder so far. */
TLB_API uint64_t tlb_batch_launches(const tlb_batch *b);

/* Per-kernel device time: with profiling enabled every chunk records CUDA events around its kernels on the
 * launching stream; tlb_batch_kernel_times syncs, returns the summed milliseconds and launch counts of the
 * tlb_kernel_count() kernels (names: tlb_kernel_name(k)) since the last call, and clears them.  ms and launches
 * point at tlb_kernel_count() elements each. */
TLB_API int tlb_batch_profile(tlb_batch *b, int enable);
TLB_API int tlb_batch_kernel_times(tlb_batch *b, double *ms, uint64_t *launches);
TLB_API int tlb_kernel_count(void);
TLB_API const char *tlb_kernel_name(int k);                          /* psy model 1 kernel set */
TLB_API const char *tlb_batch_kernel_name(const tlb_batch *b, int k); /* this encoder's kernel set ("" = unused slot) */
/* Measured FP64 rate of the device, TFLOP/s with mul+add = 2 flop: DFMA chains, and DMUL+DADD chains (the only
 * form this path may use: the reference is built without FMA contraction). */
TLB_API int tlb_fp64_peak(int device, double *dfma_tflops, double *dmul_dadd_tflops);
/* Device self-test of the spectrum kernel's log10 (CUDA's own algorithm without the exits for zero, negative,
 * subnormal, infinite and NaN arguments, which an energy >= 1e-20 never takes): compares bit patterns with CUDA's
 * log10 on n values (every binade from 2^-67 to 2^60, binade edges and the reduction boundary over-sampled).  Returns
 * the number of values that differ (0 = identical), or a negative TLB_E_* code; *first_bad = one differing value. */
TLB_API long long tlb_selftest_log10(int device, unsigned long long n, double *first_bad);

/* Pinned host memory for pcm / out buffers. */
TLB_API void *tlb_host_alloc(size_t bytes);
TLB_API void tlb_host_free(void *p);

/* Per-frame intermediate results of the most recent chunk (parity tests read these; layouts below).
 * Copies min(bytes, available) bytes and returns the number copied, or a negative error. */
enum {
    TLB_TAP_SB_SAMPLE = 0,  /* double [frames][nch][36][32]  subband samples (subband.c:201-310); subbands >= sblimit,
                             * which the reference computes but never reads, are not kept and read as 0 */
    TLB_TAP_SCALAR_PRE = 1, /* uint8  [frames][2][3][32]     scalefactor indices before the scfsi pattern */
    TLB_TAP_J_SCALE = 2,    /* uint8  [frames][3][32]        joint-stereo scalefactor indices */
    TLB_TAP_SMR = 3,        /* double [frames][2][32]        signal-to-mask ratios (psycho_1.c:568-581) */
    TLB_TAP_SIDE = 4        /* tlb_side [frames] */
};
typedef struct {
    uint8_t bit_alloc[2][32];
    uint8_t scfsi[2][32];
    uint8_t scalar[2][3][32];  /* after sf_transmission_pattern (encode_new.c:288-354) */
    uint8_t scfcrc_own[4];     /* CRC_calcDAB of this frame, index = subband group (crc.c:58-98) */
    uint8_t mode, mode_ext, jsbound, xpad_len;
    int32_t adb_left;          /* zero-stuffing bits (toolame.c:509-512) */
    uint32_t crc16;            /* crc.c:12-41 */
} tlb_side;
TLB_API long tlb_batch_tap(tlb_batch *b, int what, void *dst, size_t bytes);

/* State of the libtoolame-dab drop-in stream (include/toolame.h): 0 while it is healthy, otherwise the TLB_E_* code
 * that stopped it.  The reference's API has no error return on toolame_encode_frame (it returns bytes written), so a
 * batch that cannot be encoded -- after one retry on a fresh encoder -- stops the stream instead of leaving a gap
 * in it: every later call returns 0 bytes, toolame_finish hands out what was complete, toolame_init starts again. */
TLB_API int toolame_b200_status(void);

#ifdef __cplusplus
}
#endif
#endif

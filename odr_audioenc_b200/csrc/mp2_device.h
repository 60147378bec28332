// mp2_device.h -- host <-> kernel interface of the MP2 DAB encode path (internal; not the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/toolame_b200.h"

// Per-stream constants handed to every kernel by value.
struct Mp2Params {
    int nch, sblimit, tablenum;
    int sbw;                            // subband samples kept per block row in HBM: 8, 16 or 32 (>= sblimit)
    int mode, mode_ext, jsbound;        // as configured (toolame_set_channel_mode); JS frames re-decide per frame
    int version, bitrate_index, sfreq_idx;
    int dab_ext, lg_frame, pad_len;
    int psy_freq, sub_size, cb_count;   // psy-1 table selectors (psycho_1.c:42-56)
    int bitrate_per_ch;                 // kbit/s per channel (psycho_1_threshold's ATH offset switch)
    int psy;                            // psychoacoustic model: 0, 1 or 2
};

// Per-stream lookup tables of the psychoacoustic model, built on the host at create time.
struct Mp2PsyTables {
    uint8_t map[512];    // FFT line -> threshold partition (psycho_1.c:160-168); 0 above the last partition
    uint8_t band[512];   // FFT line -> critical band (critband.h boundaries); 255 outside every band
    uint8_t mm_j0[32];   // per subband: first threshold partition of psycho_1_minimum_mask's scan (255: past the end)
    uint8_t mm_j1[32];   // ... and one past its last partition
    double ath_min[32];  // psy model 0: lowest absolute threshold per subband in dB (psycho_0.c:36-47)
};

// Start-up tables of psychoacoustic model 2 (frozen per sample rate in mp2_psy2_tables.h), device copy.
struct Mp2Psy2Tables {
    double sT[64][64];    // spreading function transposed: sT[k][j] = s[j][k], partition k into partition j
    double tmn[64], rnorm[64], bmax_of[64];
    int numlines[64];
    int first_line[65];   // partition p covers FFT lines first_line[p] .. first_line[p+1]-1
    int absthr_table;
    uint8_t partition[520];
};

// Surviving maskers of one (frame, channel), in the order psycho_1_threshold visits them.
struct Mp2Maskers {
    double t_x[104];      // tonal: level in dB
    double n_x[28];       // noise
    uint8_t t_part[104];  // threshold-table partition of the masker's line (index into the bark table)
    uint8_t n_part[28];
    int n_tone, n_noise;
};

// Device buffers of one chunk (frames analysed = n_out + has_next).
struct Mp2Chunk {
    const int16_t *pcm;     // interleaved s16; element 0 = first sample of the chunk's first frame
    long lo;                // lowest readable sample index relative to pcm (<= 0); below it the signal is 0
    const uint8_t *xpad;    // NULL or records of pad_len+1 bytes, one per analysed frame
    double *sb;             // [fa][nch][36][sbw]
    uint8_t *scalar_pre;    // [fa][2][3][32]
    uint8_t *j_scale;       // [fa][3][32]
    double *psy_x;          // [ceil(fa*nch/32)][64 chunks][32 items][8 lines]  dB spectrum (psy_line() in mp2_kernels.cu)
    double *psy_w;          // same layout: noise-centre weight of each line
    unsigned *psy_cand;     // [fa*nch][16] tonal-candidate mask
    unsigned *psy_t0;       // [fa*nch][16] candidates passing the neighbourhood test on the unmodified spectrum
    double *spike;          // [fa*nch][32]
    Mp2Maskers *maskers;    // [fa*nch]
    double *p2_energy;      // psy-2: [(2*fa+2)*nch][520] energy per block and channel, record (block+2)*nch+ch
    double *p2_cu, *p2_su;  // psy-2: same layout, cosine and sine of the line's phase
    double *p2_r;           // psy-2: same layout, sqrt(energy)
    long p2_first_block;    // psy-2: lowest block (relative to the chunk's first frame) that exists; earlier = zero state
    double *smr;            // [ceil(fa/32)][64][32] frame-tile layout
    tlb_side *side;         // [fa]
    uint8_t *out;           // [n_out][lg_frame]
    int fa;                 // frames analysed
    int n_out;              // frames written
};

// Launch the kernels of one chunk on `stream`; returns the number of launches issued.
// ev: NULL, or MP2_N_KERNELS+1 events recorded before / between / after the kernels (per-kernel timing).
int mp2_launch_chunk(const Mp2Params &p, const Mp2Chunk &c, const Mp2PsyTables *tables, const Mp2Psy2Tables *tables2,
                     cudaStream_t stream, cudaEvent_t *ev);
constexpr int MP2_N_KERNELS = 6;
extern const char *const MP2_KERNEL_NAMES[MP2_N_KERNELS];

// Measured FP64 rate of the device in TFLOP/s (mul+add counted as 2): DFMA chains, or DMUL+DADD chains.
double mp2_fp64_probe(bool fma, cudaStream_t stream);
long long mp2_selftest_log10(unsigned long long n, double *first_bad);

// Gain correction in place + per-frame peak levels of interleaved s16 PCM (src/odr-audioenc.cpp:1020-1055).
void mp2_launch_gain_peak(int16_t *d_pcm, long n_frames, int nch, double linear_gain, int16_t *d_peaks, cudaStream_t stream);
void mp2_launch_gain_peak_pairs(int16_t *d_pcm, long n_units, int pairs_per_unit, double linear_gain, int16_t *d_peaks,
                                cudaStream_t stream);

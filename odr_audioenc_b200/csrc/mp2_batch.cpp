// mp2_batch.cpp -- host side of the batch encoder: stream configuration, chunking with PCM halos,
// double-buffered host<->device copies, and the tlb_* C ABI (include/toolame_b200.h).
//
// Configuration rules restate what the reference derives in toolame_set_* (toolame.c:168-262), hdr_to_frps
// (common.c:76-93), encode_init (encode_new.c:104-124) and available_bits (availbits.c:37-67).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "mp2_device.h"
#include "tlb_internal.h"

#define MP2_TABLE_QUAL [[maybe_unused]] static const
#include "mp2_alloc_tables.h"
#include "mp2_tables.h"
#include "mp2_psy2_tables.h"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

#define CU(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return fail(TLB_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                      \
    } while (0)

constexpr int HALO = 1664;        // samples of history staged per chunk: the filterbank needs 480, psy-1 192, psy-2
                                  // 1632 (FFT window of the block two before the frame's first: psycho_2.c:80-92)
constexpr int HALO_PSY1 = 480, HALO_PSY2 = 1632;
constexpr size_t DEFAULT_CHUNK = 148 * 512; // frames per launch: a multiple of the SM count, large enough to fill the
                                            // thread-per-frame kernels (k_label, k_alloc) with warps
constexpr size_t HOST_CHUNK = 148 * 96;     // host-buffer path: smaller pieces, three in flight, so that the H2D engine
                                            // never waits (measured end to end, 10 h of config B, three interleaved
                                            // repetitions, plain-copy ceiling of the box 281k x real time: 14 208 frames
                                            // 281-282k, 18 944 (round 1's choice, tuned on slower kernels) 270-272k,
                                            // 28 416 269-280k, 37 888 278-279k; two slots instead of three lose 1-20 %)

int configure(const tlb_config &c, Mp2Params &P, tlb_info &I)
{
    std::memset(&P, 0, sizeof P);
    switch (c.sample_rate) { // ref: common.c:118-144
    case 44100: P.version = 1; P.sfreq_idx = 0; break;
    case 48000: P.version = 1; P.sfreq_idx = 1; break;
    case 32000: P.version = 1; P.sfreq_idx = 2; break;
    case 22050: P.version = 0; P.sfreq_idx = 0; break;
    case 24000: P.version = 0; P.sfreq_idx = 1; break;
    case 16000: P.version = 0; P.sfreq_idx = 2; break;
    default: return fail(TLB_E_PARAM, "illegal sample rate");
    }
    if (c.psy_model < 0 || c.psy_model > 3) return fail(TLB_E_PARAM, "illegal psy model"); // ref: toolame.c:204
    if (c.psy_model == 3) return fail(TLB_E_UNSUPP, "psychoacoustic model 3 is not built (models 0, 1 and 2 are)");
    P.psy = c.psy_model;
    switch (c.channel_mode) { // ref: toolame.c:174-200
    case 's': P.mode = 0; P.mode_ext = 0; break;
    case 'j': P.mode = 1; P.mode_ext = 2; break;
    case 'd': P.mode = 2; P.mode_ext = 0; break;
    case 'm': P.mode = 3; P.mode_ext = 0; break;
    default: return fail(TLB_E_PARAM, "illegal channel mode");
    }
    P.nch = P.mode == 3 ? 1 : 2;
    int kbps = c.bitrate ? c.bitrate : MP2_BITRATE[P.version][10]; // ref: toolame.c:217-218
    P.bitrate_index = -1;
    for (int i = 0; i < 15; i++) // ref: common.c:95-116
        if (MP2_BITRATE[P.version][i] == kbps) { P.bitrate_index = i; break; }
    if (P.bitrate_index < 0) return fail(TLB_E_PARAM, "illegal bitrate for this sample rate");
    P.dab_ext = 4; // ref: toolame.c:147,225-232
    if (P.version == 1 && kbps / (P.mode == 3 ? 1 : 2) < 56) P.dab_ext = 2;
    if (c.pad_len < 0 || c.pad_len > 255) return fail(TLB_E_PARAM, "illegal pad length");
    P.pad_len = c.pad_len;
    const int per_ch = kbps / P.nch;
    static const int sfreq_khz_int[2][3] = {{22, 24, 16}, {44, 48, 32}};
    const int sfrq = sfreq_khz_int[P.version][P.sfreq_idx];
    if (P.version == 1) { // ref: encode_new.c:112-121 == tables.c:24-36
        if ((sfrq == 48 && per_ch >= 56) || (per_ch >= 56 && per_ch <= 80)) P.tablenum = 0;
        else if (sfrq != 48 && per_ch >= 96) P.tablenum = 1;
        else if (sfrq != 32 && per_ch <= 48) P.tablenum = 2;
        else P.tablenum = 3;
    } else P.tablenum = 4;
    P.sblimit = MP2_TAB_SBLIMIT[P.tablenum];
    P.sbw = P.sblimit <= 8 ? 8 : P.sblimit <= 16 ? 16 : 32; // whole 64-byte pieces: rows that are not cut slow the stores down
    {   // ref: availbits.c:42-46; the padding branch (non-integral slot count) only exists at 44.1 / 22.05 kHz
        static const double s_freq[2][3] = {{22.05, 24, 16}, {44.1, 48, 32}};
        const double average = (1152.0 / s_freq[P.version][P.sfreq_idx]) * ((double)kbps / 8.0);
        P.lg_frame = (int)average;
        if (average - (double)P.lg_frame != 0) return fail(TLB_E_UNSUPP, "sample rates that need padding slots are not built");
    }
    if (P.lg_frame % 4 || P.lg_frame > 1728) return fail(TLB_E_UNSUPP, "frame length");
    P.jsbound = P.mode == 1 ? MP2_JSBOUND[P.mode_ext] : P.sblimit; // ref: common.c:87-91
    P.psy_freq = P.version == 1 ? P.sfreq_idx : P.sfreq_idx + 4;   // ref: psycho_1.c:42-48
    P.sub_size = MP2_SUB_SIZE[P.psy_freq];
    P.cb_count = MP2_CB_COUNT[P.psy_freq];
    P.bitrate_per_ch = per_ch;
    I.nch = P.nch; I.lg_frame = P.lg_frame; I.sblimit = P.sblimit; I.tablenum = P.tablenum; I.dab_ext = P.dab_ext;
    I.version = P.version; I.bitrate_index = P.bitrate_index; I.sfreq_idx = P.sfreq_idx;
    I.samples_per_frame = 1152; I.halo_samples = P.psy == 2 ? HALO_PSY2 : HALO_PSY1;
    return 0;
}

struct Slot {
    int16_t *d_pcm = nullptr;   // HALO*nch + (chunk+1)*1152*nch samples
    uint8_t *d_xpad = nullptr;
    uint8_t *d_out = nullptr;
    double *sb = nullptr;
    uint8_t *scalar_pre = nullptr, *j_scale = nullptr;
    double *smr = nullptr, *psy_x = nullptr, *psy_w = nullptr, *spike = nullptr;
    unsigned *psy_cand = nullptr, *psy_t0 = nullptr;
    Mp2Maskers *maskers = nullptr;
    double *p2_energy = nullptr, *p2_cu = nullptr, *p2_su = nullptr, *p2_r = nullptr;
    int16_t *d_peaks = nullptr; // [fa + 1][2]
    tlb_side *side = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    int last_fa = 0;
    size_t cap = 0;             // frames (incl. the look-ahead frame) the buffers above are sized for; 0 = not allocated
};

} // namespace

struct tlb_batch {
    tlb_config cfg;
    Mp2Params P;
    tlb_info info;
    int device = 0;
    size_t chunk = 0;
    static constexpr int NSLOT = 3; // in-flight chunks: the host-buffer path rotates over all three (copies of one
                                    // chunk overlap kernels of the others), the device-resident path uses two
                                    // (measured: 1 lane 323k, 2 lanes 346k, 3 lanes 342k x real time)
    Slot slot[NSLOT];
    Mp2PsyTables *d_tables = nullptr;
    Mp2Psy2Tables *d_tables2 = nullptr;
    uint64_t launches = 0;
    int last_slot = 0;
    bool profile = false;
    double gain_db = 0.0;        // host-buffer path: gain applied on the device before encoding
    int16_t *h_peaks = nullptr;  // host-buffer path: per-frame (left, right) peaks of the next encode calls
    std::vector<cudaEvent_t> prof_events; // MP2_N_KERNELS + 1 per profiled chunk
    cudaEvent_t *next_events()
    {
        if (!profile) return nullptr;
        const size_t at = prof_events.size();
        prof_events.resize(at + MP2_N_KERNELS + 1);
        for (size_t i = at; i < at + MP2_N_KERNELS + 1; i++) cudaEventCreate(&prof_events[i]);
        return &prof_events[at];
    }
};

const char *const MP2_KERNEL_NAMES[MP2_N_KERNELS] = {"k_filterbank", "k_spectrum", "k_label", "k_threshold", "k_alloc", "k_pack"};

namespace {

void free_buffers(Slot &s);

// Working memory of one in-flight chunk, sized for `frames` output frames (+ the look-ahead frame).  Slots are
// allocated on first use and for the chunk size of the path that uses them: the host-buffer path stages
// min(chunk, HOST_CHUNK) frames per slot, the device-resident path whole chunks on two of the three slots.
int alloc_slot(tlb_batch *b, Slot &s, size_t frames)
{
    if (s.cap >= frames + 1) return 0;
    if (s.stream) CU(cudaStreamSynchronize(s.stream)); // growing: nothing may still be using the old buffers
    free_buffers(s);
    const size_t fa = frames + 1, nch = (size_t)b->P.nch;
    CU(cudaMalloc(&s.d_pcm, (HALO + fa * 1152) * nch * sizeof(int16_t)));
    CU(cudaMalloc(&s.d_xpad, fa * (size_t)(b->P.pad_len + 1)));
    CU(cudaMalloc(&s.d_out, frames * (size_t)b->P.lg_frame));
    CU(cudaMalloc(&s.sb, fa * nch * 36 * (size_t)b->P.sbw * sizeof(double)));
    const size_t fa32 = (fa + 31) / 32 * 32; // frame-tile layouts are padded to whole tiles of 32 frames
    CU(cudaMalloc(&s.scalar_pre, fa32 * 192));
    CU(cudaMalloc(&s.j_scale, fa * 96));
    CU(cudaMalloc(&s.smr, fa32 * 64 * sizeof(double)));
    const size_t items = fa * nch, tiles = (items + 31) / 32;
    if (b->P.psy == 2) {
        CU(cudaMalloc(&s.p2_energy, (2 * fa + 2) * nch * 520 * sizeof(double)));
        CU(cudaMalloc(&s.p2_cu, (2 * fa + 2) * nch * 520 * sizeof(double)));
        CU(cudaMalloc(&s.p2_su, (2 * fa + 2) * nch * 520 * sizeof(double)));
        CU(cudaMalloc(&s.p2_r, (2 * fa + 2) * nch * 520 * sizeof(double)));
    } else {
        CU(cudaMalloc(&s.psy_x, tiles * 512 * 32 * sizeof(double)));
        CU(cudaMalloc(&s.psy_w, tiles * 512 * 32 * sizeof(double)));
        CU(cudaMalloc(&s.psy_cand, items * 16 * sizeof(unsigned)));
        CU(cudaMalloc(&s.psy_t0, items * 16 * sizeof(unsigned)));
        CU(cudaMalloc(&s.spike, items * 32 * sizeof(double)));
        CU(cudaMalloc(&s.maskers, items * sizeof(Mp2Maskers)));
    }
    CU(cudaMalloc(&s.side, fa * sizeof(tlb_side)));
    CU(cudaMalloc(&s.d_peaks, (fa + 1) * 2 * sizeof(int16_t)));
    if (!s.stream) CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking)); // (kept when the slot grows:
    if (!s.done) CU(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));       //  slot 0's is the caller's handle)
    s.cap = fa;
    return 0;
}

void free_buffers(Slot &s)
{
    cudaStream_t stream = s.stream;
    cudaEvent_t done = s.done;
    cudaFree(s.d_pcm); cudaFree(s.d_xpad); cudaFree(s.d_out); cudaFree(s.sb); cudaFree(s.scalar_pre);
    cudaFree(s.j_scale); cudaFree(s.smr); cudaFree(s.side);
    cudaFree(s.psy_x); cudaFree(s.psy_w); cudaFree(s.psy_cand); cudaFree(s.psy_t0); cudaFree(s.spike); cudaFree(s.maskers);
    cudaFree(s.p2_energy); cudaFree(s.p2_cu); cudaFree(s.p2_su); cudaFree(s.p2_r); cudaFree(s.d_peaks);
    s = Slot();
    s.stream = stream;
    s.done = done;
}

void free_slot(Slot &s)
{
    free_buffers(s);
    if (s.stream) cudaStreamDestroy(s.stream);
    if (s.done) cudaEventDestroy(s.done);
    s = Slot();
}

Mp2Chunk chunk_of(const tlb_batch *b, const Slot &s, const int16_t *pcm, long lo, const uint8_t *xpad, uint8_t *out,
                  int fa, int n_out)
{
    (void)b;
    Mp2Chunk c;
    c.pcm = pcm; c.lo = lo; c.xpad = xpad; c.sb = s.sb; c.scalar_pre = s.scalar_pre; c.j_scale = s.j_scale;
    c.smr = s.smr; c.side = s.side; c.psy_x = s.psy_x; c.psy_w = s.psy_w; c.psy_cand = s.psy_cand; c.psy_t0 = s.psy_t0;
    c.spike = s.spike; c.maskers = s.maskers; c.p2_energy = s.p2_energy; c.p2_cu = s.p2_cu; c.p2_su = s.p2_su; c.p2_r = s.p2_r;
    c.p2_first_block = lo == 0 ? 0 : -2; // at the stream start the blocks before the first frame are the zero state
    c.out = out; c.fa = fa; c.n_out = n_out;
    return c;
}

int check_args(const tlb_batch *b, const void *pcm, size_t history, const void *out)
{
    if (!b || !pcm || !out) return fail(TLB_E_ARG, "NULL argument");
    if (history != 0 && history < (size_t)b->info.halo_samples)
        return fail(TLB_E_ARG, "history_samples must be 0 or >= halo_samples (480; 1632 for psy model 2)");
    return 0;
}

} // namespace

int tlb_fail(int code, const char *msg) { return fail(code, msg); }

extern "C" {

const char *tlb_last_error(void) { return g_err.c_str(); }

int tlb_config_check(const tlb_config *cfg, tlb_info *info)
{
    if (!cfg) return fail(TLB_E_ARG, "NULL argument");
    Mp2Params P;
    tlb_info I;
    const int rc = configure(*cfg, P, I);
    if (!rc && info) *info = I;
    return rc;
}

int tlb_batch_create(tlb_batch **out, const tlb_config *cfg, int device, size_t max_chunk_frames)
{
    if (!out || !cfg) return fail(TLB_E_ARG, "NULL argument");
    *out = nullptr;
    Mp2Params P;
    tlb_info I;
    int rc = configure(*cfg, P, I);
    if (rc) return rc;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return fail(TLB_E_CUDA, "no CUDA device: this library has no CPU path");
    if (device < 0 || device >= n_dev) return fail(TLB_E_ARG, "bad device index");
    CU(cudaSetDevice(device));
    tlb_batch *b = new (std::nothrow) tlb_batch();
    if (!b) return fail(TLB_E_ARG, "out of memory");
    b->cfg = *cfg; b->P = P; b->info = I; b->device = device;
    b->chunk = max_chunk_frames ? max_chunk_frames : DEFAULT_CHUNK;
    // slot 0 now (its stream is the encoder's stream towards the caller), the others on first use
    if ((rc = alloc_slot(b, b->slot[0], std::min(b->chunk, HOST_CHUNK)))) { tlb_batch_destroy(b); return rc; }
    {
        Mp2PsyTables T;
        std::memset(&T, 0, sizeof T);
        const int fq = P.psy_freq;
        // FFT line -> threshold-table partition (ref: psycho_1.c:160-168); lines above the last partition keep the
        // allocator's zero (mem.c:21)
        for (int i = 1; i < P.sub_size; i++)
            for (int j = MP2_LTG_LINE[fq][i - 1]; j <= MP2_LTG_LINE[fq][i]; j++) T.map[j] = (uint8_t)i;
        std::memset(T.band, 255, sizeof T.band);
        for (int i = 0; i + 1 < P.cb_count; i++)
            for (int j = MP2_CBOUND[fq][i]; j < MP2_CBOUND[fq][i + 1]; j++) T.band[j] = (uint8_t)i;
        // psycho_1_minimum_mask's scan over the partitions is data independent: replay it (ref: psycho_1.c:546-558)
        int j = 1;
        for (int i = 0; i < P.sblimit; i++) {
            if (j >= P.sub_size - 1) { T.mm_j0[i] = 255; T.mm_j1[i] = 255; continue; }
            T.mm_j0[i] = (uint8_t)j;
            while (j < P.sub_size && (MP2_LTG_LINE[fq][j] >> 4) == i) j++;
            T.mm_j1[i] = (uint8_t)j;
        }
        // psy model 0's absolute-threshold minima and psy model 2's start-up tables are frozen per sample rate in
        // mp2_psy2_tables.h (tools/gen_tables.py reads them back from the compiled reference: psycho_0.c:36-47,
        // psycho_2.c:259-420): nothing is evaluated with libm here
        int ri = -1;
        for (int i = 0; i < MP2_P2_RATES; i++)
            if (MP2_P2_RATE[i] == cfg->sample_rate) ri = i;
        if (ri < 0) { tlb_batch_destroy(b); return fail(TLB_E_UNSUPP, "no start-up tables for this sample rate"); }
        std::memcpy(T.ath_min, MP2_P0_ATH_MIN[ri], sizeof T.ath_min);
        if (cudaMalloc(&b->d_tables, sizeof T) != cudaSuccess ||
            cudaMemcpy(b->d_tables, &T, sizeof T, cudaMemcpyHostToDevice) != cudaSuccess) {
            tlb_batch_destroy(b);
            return fail(TLB_E_CUDA, "table upload failed");
        }
        if (P.psy == 2) {
            std::unique_ptr<Mp2Psy2Tables> Dp(new Mp2Psy2Tables);
            Mp2Psy2Tables &D = *Dp;
            std::memset(&D, 0, sizeof D);
            std::memcpy(D.sT, MP2_P2_ST[ri], sizeof D.sT);
            std::memcpy(D.tmn, MP2_P2_TMN[ri], sizeof D.tmn);
            std::memcpy(D.rnorm, MP2_P2_RNORM[ri], sizeof D.rnorm);
            std::memcpy(D.bmax_of, MP2_P2_BMAX_OF[ri], sizeof D.bmax_of);
            std::memcpy(D.numlines, MP2_P2_NUMLINES[ri], sizeof D.numlines);
            std::memcpy(D.first_line, MP2_P2_FIRST_LINE[ri], sizeof D.first_line);
            std::memcpy(D.partition, MP2_P2_PARTITION[ri], sizeof D.partition);
            D.absthr_table = MP2_P2_ABSTHR_TABLE[ri];
            if (cudaMalloc(&b->d_tables2, sizeof D) != cudaSuccess ||
                cudaMemcpy(b->d_tables2, &D, sizeof D, cudaMemcpyHostToDevice) != cudaSuccess) {
                tlb_batch_destroy(b);
                return fail(TLB_E_CUDA, "psy-2 table upload failed");
            }
        }
    }
    *out = b;
    return 0;
}

void tlb_batch_destroy(tlb_batch *b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    for (auto &s : b->slot) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        free_slot(s);
    }
    for (auto e : b->prof_events) cudaEventDestroy(e);
    cudaFree(b->d_tables);
    cudaFree(b->d_tables2);
    delete b;
}

int tlb_batch_info(const tlb_batch *b, tlb_info *info)
{
    if (!b || !info) return fail(TLB_E_ARG, "NULL argument");
    *info = b->info;
    return 0;
}

void *tlb_batch_stream(tlb_batch *b) { return b ? (void *)b->slot[0].stream : nullptr; }
uint64_t tlb_batch_launches(const tlb_batch *b) { return b ? b->launches : 0; }

int tlb_batch_sync(tlb_batch *b)
{
    if (!b) return fail(TLB_E_ARG, "NULL argument");
    CU(cudaSetDevice(b->device));
    for (auto &s : b->slot)
        if (s.stream) CU(cudaStreamSynchronize(s.stream));
    return 0;
}

int tlb_batch_encode_async(tlb_batch *b, const int16_t *pcm, size_t n_frames, size_t history_samples, int has_next,
                           const uint8_t *xpad, uint8_t *out)
{
    int rc = check_args(b, pcm, history_samples, out);
    if (rc) return rc;
    CU(cudaSetDevice(b->device));
    const size_t nch = (size_t)b->P.nch, lg = (size_t)b->P.lg_frame, rec = (size_t)b->P.pad_len + 1;
    const bool use_xpad = xpad && b->P.pad_len;
    size_t k = 0;
    size_t host_lanes = tlb_batch::NSLOT;
    if (const char *e = std::getenv("TLB_HOST_LANES")) host_lanes = (size_t)std::max(1, std::min((int)tlb_batch::NSLOT, std::atoi(e)));
    size_t chunk = std::min(b->chunk, HOST_CHUNK);
    if (const char *e = std::getenv("TLB_HOST_CHUNK")) { // tuning knob: frames per host<->device pipeline stage
        const long v = std::atol(e);
        if (v > 0) chunk = std::min(b->chunk, (size_t)v);
    }
    for (size_t f0 = 0; f0 < n_frames; f0 += chunk, k++) {
        Slot &s = b->slot[k % host_lanes];
        if ((rc = alloc_slot(b, s, std::min(chunk, n_frames)))) return rc;
        const size_t n_out = std::min(chunk, n_frames - f0);
        const bool next_here = f0 + n_out < n_frames || has_next;
        const size_t fa = n_out + (next_here ? 1 : 0);
        size_t hist = std::min<size_t>(HALO, f0 * 1152 + history_samples);
        if (nch == 1) hist &= ~(size_t)1; // keeps the staged region 4-byte aligned for k_gain_peak's sample pairs (the
                                          // halo sizes are even, so an odd history always has a sample to spare)
        // (stream order makes re-use of the slot's buffers safe: the copies below queue behind its previous chunk)
        CU(cudaMemcpyAsync(s.d_pcm + (HALO - hist) * nch, pcm + (f0 * 1152 - hist) * nch,
                           (hist + fa * 1152) * nch * sizeof(int16_t), cudaMemcpyHostToDevice, s.stream));
        if (use_xpad) CU(cudaMemcpyAsync(s.d_xpad, xpad + f0 * rec, fa * rec, cudaMemcpyHostToDevice, s.stream));
        if (b->gain_db != 0.0 || b->h_peaks) { // ref: src/odr-audioenc.cpp:1020-1055, here per staged chunk
            const double linear = std::pow(10.0, b->gain_db / 20.0);
            if (hist && linear != 1.0) // the history samples are re-staged from the caller's (un-gained) buffer
                mp2_launch_gain_peak_pairs(s.d_pcm + (HALO - hist) * nch, 1, (int)(hist * nch / 2), linear, s.d_peaks + 2 * fa, s.stream);
            mp2_launch_gain_peak(s.d_pcm + HALO * nch, (long)fa, (int)nch, linear, s.d_peaks, s.stream);
            b->launches += hist && linear != 1.0 ? 2 : 1;
            if (b->h_peaks)
                CU(cudaMemcpyAsync(b->h_peaks + 2 * f0, s.d_peaks, n_out * 2 * sizeof(int16_t), cudaMemcpyDeviceToHost, s.stream));
        }
        Mp2Chunk c = chunk_of(b, s, s.d_pcm + HALO * nch, -(long)hist, use_xpad ? s.d_xpad : nullptr, s.d_out, (int)fa, (int)n_out);
        b->launches += (uint64_t)mp2_launch_chunk(b->P, c, b->d_tables, b->d_tables2, s.stream, b->next_events());
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out + f0 * lg, s.d_out, n_out * lg, cudaMemcpyDeviceToHost, s.stream));
        s.last_fa = (int)fa;
        b->last_slot = (int)(k % host_lanes);
    }
    return 0;
}

int tlb_batch_encode(tlb_batch *b, const int16_t *pcm, size_t n_frames, size_t history_samples, int has_next,
                     const uint8_t *xpad, uint8_t *out)
{
    const int rc = tlb_batch_encode_async(b, pcm, n_frames, history_samples, has_next, xpad, out);
    return rc ? rc : tlb_batch_sync(b);
}

int tlb_encode_services(const tlb_service *sv, size_t n, int device, size_t chunk_frames)
{
    if (!sv && n) return fail(TLB_E_ARG, "NULL argument");
    // one encoder per distinct configuration; services of one configuration queue on the same encoder, different
    // configurations run side by side on their own streams
    std::vector<tlb_batch *> enc;
    std::vector<tlb_config> cfgs;
    int rc = 0;
    for (size_t i = 0; i < n && !rc; i++) {
        size_t e = 0;
        for (; e < cfgs.size(); e++)
            if (!std::memcmp(&cfgs[e], &sv[i].cfg, sizeof(tlb_config))) break;
        if (e == cfgs.size()) {
            tlb_batch *b = nullptr;
            rc = tlb_batch_create(&b, &sv[i].cfg, device, chunk_frames ? chunk_frames : 148 * 128);
            if (rc) break;
            enc.push_back(b);
            cfgs.push_back(sv[i].cfg);
        }
        rc = tlb_batch_encode_async(enc[e], sv[i].pcm, sv[i].n_frames, sv[i].history_samples, sv[i].has_next, sv[i].xpad, sv[i].out);
    }
    for (auto b : enc) {
        const int r2 = tlb_batch_sync(b);
        if (!rc) rc = r2;
    }
    const std::string err = g_err;
    for (auto b : enc) tlb_batch_destroy(b);
    if (rc) g_err = err;
    return rc;
}

int tlb_batch_encode_device(tlb_batch *b, const int16_t *d_pcm, size_t n_frames, size_t history_samples, int has_next,
                            const uint8_t *d_xpad, uint8_t *d_out)
{
    int rc = check_args(b, d_pcm, history_samples, d_out);
    if (rc) return rc;
    CU(cudaSetDevice(b->device));
    const size_t nch = (size_t)b->P.nch, lg = (size_t)b->P.lg_frame, rec = (size_t)b->P.pad_len + 1;
    const bool use_xpad = d_xpad && b->P.pad_len;
    // Chunks rotate over the slots and their streams so that kernels of neighbouring chunks overlap on the GPU
    // (several kernels are latency-bound on their own).  Towards the caller the call behaves as if it ran on slot
    // 0's stream: the other streams fork from it here and join it at the end.
    const size_t n_chunks = (n_frames + b->chunk - 1) / b->chunk;
    int lanes = b->profile ? 1 : (int)std::min<size_t>(n_chunks, 2); // per-kernel timing: one at a time
    if (const char *e = std::getenv("TLB_DEVICE_LANES")) // tuning knob
        lanes = b->profile ? 1 : std::max(1, std::min((int)std::min<size_t>(n_chunks, tlb_batch::NSLOT), std::atoi(e)));
    for (int i = 0; i < lanes; i++)
        if ((rc = alloc_slot(b, b->slot[i], std::min(b->chunk, n_frames)))) return rc;
    for (int i = 1; i < lanes; i++) {
        CU(cudaEventRecord(b->slot[0].done, b->slot[0].stream));
        CU(cudaStreamWaitEvent(b->slot[i].stream, b->slot[0].done, 0));
    }
    size_t k = 0;
    for (size_t f0 = 0; f0 < n_frames; f0 += b->chunk, k++) {
        const int si = (int)(k % (size_t)lanes);
        Slot &s = b->slot[si];
        const size_t n_out = std::min(b->chunk, n_frames - f0);
        const bool next_here = f0 + n_out < n_frames || has_next;
        const size_t fa = n_out + (next_here ? 1 : 0);
        const size_t hist = f0 * 1152 + history_samples;
        Mp2Chunk c = chunk_of(b, s, d_pcm + f0 * 1152 * nch, -(long)hist, use_xpad ? d_xpad + f0 * rec : nullptr,
                              d_out + f0 * lg, (int)fa, (int)n_out);
        b->launches += (uint64_t)mp2_launch_chunk(b->P, c, b->d_tables, b->d_tables2, s.stream, b->next_events());
        CU(cudaGetLastError());
        s.last_fa = (int)fa;
        b->last_slot = si;
    }
    for (int i = 1; i < lanes; i++) {
        CU(cudaEventRecord(b->slot[i].done, b->slot[i].stream));
        CU(cudaStreamWaitEvent(b->slot[0].stream, b->slot[i].done, 0));
    }
    return 0;
}

int tlb_batch_set_gain(tlb_batch *b, double gain_db, int16_t *peaks)
{
    if (!b) return fail(TLB_E_ARG, "NULL argument");
    b->gain_db = gain_db;
    b->h_peaks = peaks;
    return 0;
}

int tlb_batch_gain_peak_device(tlb_batch *b, int16_t *d_pcm, size_t n_frames, double gain_db, int16_t *d_peaks)
{
    if (!b || !d_pcm || !d_peaks) return fail(TLB_E_ARG, "NULL argument");
    CU(cudaSetDevice(b->device));
    // ref: src/odr-audioenc.cpp:1032: const double linear_gain_correction = pow(10.0, gain_dB / 20.0);
    const double linear = std::pow(10.0, gain_db / 20.0);
    mp2_launch_gain_peak(d_pcm, (long)n_frames, b->P.nch, linear, d_peaks, b->slot[0].stream);
    b->launches++;
    CU(cudaGetLastError());
    return 0;
}

int tlb_batch_profile(tlb_batch *b, int enable)
{
    if (!b) return fail(TLB_E_ARG, "NULL argument");
    b->profile = enable != 0;
    return 0;
}

int tlb_batch_kernel_times(tlb_batch *b, double *ms, uint64_t *launches)
{
    if (!b || !ms || !launches) return fail(TLB_E_ARG, "NULL argument");
    int rc = tlb_batch_sync(b);
    if (rc) return rc;
    for (int k = 0; k < MP2_N_KERNELS; k++) { ms[k] = 0; launches[k] = 0; }
    for (size_t at = 0; at + MP2_N_KERNELS + 1 <= b->prof_events.size(); at += MP2_N_KERNELS + 1)
        for (int k = 0; k < MP2_N_KERNELS; k++) {
            float t = 0;
            if (cudaEventElapsedTime(&t, b->prof_events[at + k], b->prof_events[at + k + 1]) == cudaSuccess) {
                ms[k] += t;
                launches[k]++;
            }
        }
    for (auto e : b->prof_events) cudaEventDestroy(e);
    b->prof_events.clear();
    return 0;
}

int tlb_kernel_count(void) { return MP2_N_KERNELS; }

const char *tlb_kernel_name(int k) { return k >= 0 && k < MP2_N_KERNELS ? MP2_KERNEL_NAMES[k] : ""; }

const char *tlb_batch_kernel_name(const tlb_batch *b, int k)
{
    static const char *const psy2_names[MP2_N_KERNELS] = {"k_filterbank", "k_spectrum2", "k_psy2", "", "k_alloc", "k_pack"};
    static const char *const psy0_names[MP2_N_KERNELS] = {"k_filterbank", "k_psy0", "", "", "k_alloc", "k_pack"};
    if (k < 0 || k >= MP2_N_KERNELS) return "";
    return b && b->P.psy == 2 ? psy2_names[k] : b && b->P.psy == 0 ? psy0_names[k] : MP2_KERNEL_NAMES[k];
}

int tlb_fp64_peak(int device, double *dfma_tflops, double *dmul_dadd_tflops)
{
    if (!dfma_tflops || !dmul_dadd_tflops) return fail(TLB_E_ARG, "NULL argument");
    CU(cudaSetDevice(device));
    *dfma_tflops = mp2_fp64_probe(true, nullptr);
    *dmul_dadd_tflops = mp2_fp64_probe(false, nullptr);
    if (*dfma_tflops < 0 || *dmul_dadd_tflops < 0) return fail(TLB_E_CUDA, "fp64 probe failed");
    return 0;
}

long long tlb_selftest_log10(int device, unsigned long long n, double *first_bad)
{
    CU(cudaSetDevice(device));
    const long long bad = mp2_selftest_log10(n, first_bad);
    if (bad < 0) return fail(TLB_E_CUDA, "log10 self-test could not run");
    return bad;
}

void *tlb_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}

void tlb_host_free(void *p) { cudaFreeHost(p); }

long tlb_batch_tap(tlb_batch *b, int what, void *dst, size_t bytes)
{
    if (!b || !dst) return fail(TLB_E_ARG, "NULL argument");
    CU(cudaSetDevice(b->device));
    Slot &s = b->slot[b->last_slot];
    CU(cudaStreamSynchronize(s.stream));
    const size_t fa = (size_t)s.last_fa;
    const void *src = nullptr;
    size_t avail = 0;
    switch (what) {
    case TLB_TAP_SB_SAMPLE: src = s.sb; avail = fa * (size_t)b->P.nch * 1152 * sizeof(double); break; // (expanded below)
    case TLB_TAP_SCALAR_PRE: src = s.scalar_pre; avail = fa * 192; break;
    case TLB_TAP_J_SCALE: src = s.j_scale; avail = fa * 96; break;
    case TLB_TAP_SMR: src = s.smr; avail = fa * 64 * sizeof(double); break;
    case TLB_TAP_SIDE: src = s.side; avail = fa * sizeof(tlb_side); break;
    default: return fail(TLB_E_ARG, "unknown tap");
    }
    const size_t n = std::min(avail, bytes);
    if (what == TLB_TAP_SB_SAMPLE) {
        // kept on the device as rows of sbw subbands (those below sblimit): hand out rows of 32, zeros above
        const size_t rows = fa * (size_t)b->P.nch * 36, sbw = (size_t)b->P.sbw;
        std::vector<double> raw(rows * sbw), lin(rows * 32, 0.0);
        CU(cudaMemcpy(raw.data(), src, raw.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (size_t r = 0; r < rows; r++)
            std::memcpy(&lin[r * 32], &raw[r * sbw], (size_t)b->P.sblimit * sizeof(double));
        std::memcpy(dst, lin.data(), n);
        return (long)n;
    }
    if (what == TLB_TAP_SCALAR_PRE || what == TLB_TAP_SMR) {
        // stored on the device in the frame-tile layout [frame/32][field][frame%32]: put frames back in order
        const size_t nf = what == TLB_TAP_SMR ? 64 : 192, esz = what == TLB_TAP_SMR ? sizeof(double) : 1;
        const size_t fa32 = (fa + 31) / 32 * 32;
        std::vector<unsigned char> raw(fa32 * nf * esz), lin(fa * nf * esz);
        CU(cudaMemcpy(raw.data(), src, raw.size(), cudaMemcpyDeviceToHost));
        for (size_t f = 0; f < fa; f++)
            for (size_t k = 0; k < nf; k++)
                std::memcpy(&lin[(f * nf + k) * esz], &raw[(((f >> 5) * nf + k) * 32 + (f & 31)) * esz], esz);
        std::memcpy(dst, lin.data(), n);
        return (long)n;
    }
    CU(cudaMemcpy(dst, src, n, cudaMemcpyDeviceToHost));
    return (long)n;
}

} // extern "C"

// toolame_shim.cpp -- the nine libtoolame-dab entry points (include/toolame.h) on top of the batch encoder.
//
// Streaming semantics of the reference are reproduced on the host:
//  * process-global single stream (toolame.c:24-28,89-118);
//  * the bit writer's 4096-byte buffer that is flushed, oldest bytes first, whenever it fills, keeping the
//    newest lg_frame + 4 bytes (bitstream.c:46-71, toolame.c:296-300) -- so toolame_encode_frame returns 0 or
//    4096 - (lg_frame + 4) and the chunks are not frame aligned;
//  * frame n overwrites the ScF-CRC field of frame n-1, still held in that buffer (toolame.c:527-542).
//
// The reference hands nothing back until 4096 bytes are held, so nothing observable depends on WHEN a frame is
// encoded before that point.  The shim therefore only copies PCM and X-PAD into a pinned window on most calls and
// encodes all frames that are pending in ONE batch on the call in which a flush falls due (and in toolame_finish):
// the same bytes and the same return sizes as the reference, one GPU round trip per ~4096 / lg_frame frames.
// The batch starts with the last frame of the previous batch once more: encoded with its successor present it
// carries the successor's ScF-CRC, which is the patch of toolame.c:527-539.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/toolame.h"
#include "../../include/toolame_b200.h"

namespace {

constexpr size_t BUFFER_SIZE = 4096; // ref: common.h BUFFER_SIZE
constexpr size_t MINIMUM = 4;        // ref: common.h MINIMUM
constexpr size_t KEEP = 2304;        // samples of history kept before the window's first frame (psy model 2 looks
                                     // 1632 samples back, the filterbank 480)

struct Stream {
    long sample_rate = 44100; // header.sampling_frequency = 0 / MPEG-1 until set (toolame.c:142-150)
    char mode = 's';
    int bitrate = 0;
    int psy = 1;
    int pad_len = 0;
    tlb_batch *enc = nullptr;
    tlb_info info{};
    // pinned window: [hist samples of history | n_pend frames], the X-PAD records and the output of one batch
    void *window = nullptr;
    bool window_pinned = false;
    int16_t *pcm = nullptr;
    uint8_t *recs = nullptr, *out = nullptr;
    size_t max_pend = 0;
    size_t hist = 0;          // history samples in front of the first pending frame
    size_t n_pend = 0;        // frames in the window; with `redo` the first one was already emitted once
    bool redo = false;
    bool any_xpad = false;
    std::vector<uint8_t> held;  // encoded bytes still in the reference's bit buffer, oldest first
    size_t virt_held = 0;       // bytes the reference would hold now (held + frames not yet encoded)
    long frame_num = 0;
    int error = 0;              // latched: a batch could not be encoded; the stream is over until toolame_init
    bool create_failed_once = false;
} g;

void close_encoder()
{
    if (g.enc) tlb_batch_destroy(g.enc);
    if (g.window) {
        if (g.window_pinned) tlb_host_free(g.window);
        else std::free(g.window);
    }
    g.enc = nullptr;
    g.window = nullptr;
}

void reset()
{
    close_encoder();
    g = Stream();
}

tlb_config current_config() { return tlb_config{(int32_t)g.sample_rate, g.mode, g.bitrate, g.psy, g.pad_len}; }

// The PCM / X-PAD / output window of one batch, sized from the configuration alone (no GPU needed): pinned when CUDA
// can provide it, plain memory otherwise (the copies are then staged by the driver).
bool ensure_window()
{
    if (g.window) return true;
    tlb_config c = current_config();
    tlb_info info{};
    if (tlb_config_check(&c, &info) != 0) {
        std::fprintf(stderr, "libtoolame-b200: %s\n", tlb_last_error());
        return false;
    }
    g.info = info;
    // frames between two flushes of the 4096-byte buffer, the re-encoded one, and slack
    g.max_pend = BUFFER_SIZE / (size_t)info.lg_frame + 3;
    const size_t nch = (size_t)info.nch, rec = (size_t)g.pad_len + 1;
    const size_t pcm_bytes = (KEEP + g.max_pend * 1152) * nch * sizeof(int16_t);
    const size_t rec_bytes = (g.max_pend * rec + 15) & ~(size_t)15;
    const size_t total = pcm_bytes + rec_bytes + g.max_pend * (size_t)info.lg_frame;
    g.window = tlb_host_alloc(total);
    g.window_pinned = g.window != nullptr;
    if (!g.window) g.window = std::malloc(total);
    if (!g.window) return false;
    std::memset(g.window, 0, total);
    g.pcm = static_cast<int16_t *>(g.window);
    g.recs = static_cast<uint8_t *>(g.window) + pcm_bytes;
    g.out = g.recs + rec_bytes;
    g.held.reserve(BUFFER_SIZE + 2048);
    return true;
}

bool open_encoder()
{
    if (g.enc) return true;
    tlb_config c = current_config();
    if (tlb_batch_create(&g.enc, &c, 0, g.max_pend) != 0) {
        if (!g.create_failed_once) std::fprintf(stderr, "libtoolame-b200: cannot start the encoder: %s\n", tlb_last_error());
        g.create_failed_once = true;
        g.enc = nullptr;
        return false;
    }
    return true;
}

int drain(unsigned char *dst, size_t dst_size, size_t n)
{   // hand the n oldest held bytes to the caller (ref: bitstream.c:46-71, incl. its too-small-buffer behaviour)
    if (n > g.held.size()) n = g.held.size();
    size_t w = n;
    if (w > dst_size) {
        std::fprintf(stderr, "ERROR: libtoolame output buffer too small (%zu vs %zu)!\n", dst_size, n);
        w = dst_size;
    }
    if (w) std::memcpy(dst, g.held.data(), w);
    g.held.erase(g.held.begin(), g.held.begin() + (long)n);
    g.virt_held -= n;
    return (int)w;
}

// Encode every pending frame in one batch and move the bytes into `held`.  On failure the encoder is rebuilt and
// the batch retried once; a second failure latches g.error (the frames cannot be produced, and emitting later ones
// would splice the stream: the caller sees no more bytes, toolame_finish hands out what was complete).
bool encode_pending()
{
    const size_t first_new = g.redo ? 1 : 0;
    if (g.n_pend <= first_new) return true;
    const size_t nch = (size_t)g.info.nch, lg = (size_t)g.info.lg_frame, ext = (size_t)g.info.dab_ext;
    const long f_first = g.frame_num - (long)g.n_pend;   // stream index of the window's first frame
    const size_t history = f_first == 0 ? 0 : g.hist;
    int rc = -1;
    for (int attempt = 0; attempt < 2; attempt++) {
        if (!open_encoder()) { rc = TLB_E_CUDA; break; }
        rc = tlb_batch_encode(g.enc, g.pcm + g.hist * nch, g.n_pend, history, 0,
                              (g.pad_len && g.any_xpad) ? g.recs : nullptr, g.out);
        if (rc == 0) break;
        std::fprintf(stderr, "libtoolame-b200: encode failed (%s)%s\n", tlb_last_error(), attempt ? "" : ", retrying on a fresh encoder");
        tlb_batch_destroy(g.enc);
        g.enc = nullptr;
    }
    if (rc != 0) {
        g.error = rc;
        std::fprintf(stderr, "libtoolame-b200: stream stopped at frame %ld; call toolame_init to start again\n", f_first + (long)first_new);
        return false;
    }
    if (g.redo && g.held.size() >= lg) // ref: toolame.c:527-539: the successor's ScF-CRC replaces the held frame's own
        std::memcpy(g.held.data() + g.held.size() - 2 - ext, g.out + lg - 2 - ext, ext);
    g.held.insert(g.held.end(), g.out + first_new * lg, g.out + g.n_pend * lg);
    // keep the last frame (and its history) as the head of the next batch
    const size_t last = g.hist + (g.n_pend - 1) * 1152;   // sample index of the last frame in the window
    const size_t keep = last < KEEP ? last : KEEP;
    std::memmove(g.pcm, g.pcm + (last - keep) * nch, (keep + 1152) * nch * sizeof(int16_t));
    if (g.pad_len) std::memmove(g.recs, g.recs + (g.n_pend - 1) * ((size_t)g.pad_len + 1), (size_t)g.pad_len + 1);
    g.hist = keep;
    g.n_pend = 1;
    g.redo = true;
    return true;
}

} // namespace

extern "C" {

int toolame_init(void)
{
    reset();
    return 0;
}

int toolame_enable_byteswap(void) { return 0; }

int toolame_set_channel_mode(const char mode)
{
    if (mode != 's' && mode != 'd' && mode != 'j' && mode != 'm') {
        std::fprintf(stderr, "libtoolame-dab: Bad mode %c\n", mode);
        return 1;
    }
    g.mode = mode;
    return 0;
}

int toolame_set_psy_model(int new_model)
{
    if (new_model < 0 || new_model > 3) {
        std::fprintf(stderr, "libtoolame-dab: Invalid PSY model %d\n", new_model);
        return 1;
    }
    if (new_model == 3) {
        std::fprintf(stderr, "libtoolame-b200: PSY model %d is not built (models 0, 1 and 2 are)\n", new_model);
        return 1;
    }
    g.psy = new_model;
    return 0;
}

int toolame_set_samplerate(long sample_rate)
{
    switch (sample_rate) {
    case 44100: case 48000: case 32000: case 22050: case 24000: case 16000: break;
    default:
        std::fprintf(stderr, "SmpFrqIndex: %ld is not a legal sample rate\n", sample_rate);
        return -1;
    }
    g.sample_rate = sample_rate;
    return 0;
}

int toolame_set_bitrate(int brate)
{
    // validated on the host, against the version and mode set so far (the reference reads both here:
    // toolame.c:212-237, BitrateIndex common.c:95-116); no GPU is touched before the first frame
    tlb_config c{(int32_t)g.sample_rate, g.mode, brate, 1, 0};
    const int rc = tlb_config_check(&c, nullptr);
    if (rc == TLB_E_PARAM || rc == TLB_E_UNSUPP) {
        std::fprintf(stderr, "libtoolame-b200: %s\n", tlb_last_error());
        return 1;
    }
    g.bitrate = brate;
    return 0;
}

int toolame_set_pad(int pad_len)
{
    if (pad_len < 0 || pad_len > 255) {
        std::fprintf(stderr, "Invalid XPAD length specified\n");
        return 1;
    }
    g.pad_len = pad_len;
    return 0;
}

int toolame_encode_frame(short buffer[2][1152], unsigned char *xpad_data, size_t xpad_len,
                         unsigned char *output_buffer, size_t output_buffer_size)
{
    if (!buffer || !output_buffer || g.error) return 0;
    if (!g.window) {
        if (!ensure_window()) {
            g.error = TLB_E_PARAM;
            return 0;
        }
        open_encoder(); // start the GPU side now; a failure here is reported and retried when the first batch is due
    }
    const size_t nch = (size_t)g.info.nch, lg = (size_t)g.info.lg_frame;
    if (g.n_pend >= g.max_pend && !encode_pending()) return 0; // (cannot happen: a flush falls due first)
    int16_t *cur = g.pcm + (g.hist + g.n_pend * 1152) * nch;
    for (int i = 0; i < 1152; i++)
        for (size_t ch = 0; ch < nch; ch++) cur[(size_t)i * nch + ch] = buffer[ch][i];
    if (g.pad_len) {
        uint8_t *rec = g.recs + g.n_pend * ((size_t)g.pad_len + 1);
        std::memset(rec, 0, (size_t)g.pad_len + 1);
        if (xpad_len) {
            if (xpad_len > (size_t)g.pad_len || xpad_len < 2 || !xpad_data) {
                std::fprintf(stderr, "libtoolame-b200: bad xpad_len %zu (pad_len %d)\n", xpad_len, g.pad_len);
            } else {
                std::memcpy(rec, xpad_data, (size_t)g.pad_len);
                rec[(size_t)g.pad_len] = (uint8_t)xpad_len;
                g.any_xpad = true;
            }
        }
    }
    g.n_pend++;
    g.frame_num++;
    g.virt_held += lg;
    if (g.virt_held < BUFFER_SIZE) return 0;
    // the reference's buffer is full with this frame: everything up to it must exist now (ref: bitstream.c:46-71)
    if (!encode_pending()) return 0;
    return drain(output_buffer, output_buffer_size, BUFFER_SIZE - (lg + MINIMUM));
}

int toolame_finish(unsigned char *output_buffer, size_t output_buffer_size)
{
    if (!output_buffer) return 0;
    if (!g.error && g.window) encode_pending();
    const int n = drain(output_buffer, output_buffer_size, g.held.size());
    reset();
    return n;
}

int toolame_b200_status(void) { return g.error; } // (declared in toolame_b200.h: not part of the reference's API)

} // extern "C"

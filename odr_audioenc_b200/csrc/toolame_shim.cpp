// toolame_shim.cpp -- the nine libtoolame-dab entry points (include/toolame.h) on top of the batch encoder.
//
// Streaming semantics of the reference are reproduced on the host:
//  * process-global single stream (toolame.c:24-28,89-118);
//  * the bit writer's 4096-byte buffer that is flushed, oldest bytes first, whenever it fills, keeping the
//    newest lg_frame + 4 bytes (bitstream.c:46-71, toolame.c:296-300) -- so toolame_encode_frame returns 0 or
//    4096 - (lg_frame + 4) and the chunks are not frame aligned;
//  * frame n overwrites the ScF-CRC field of frame n-1, still held in that buffer (toolame.c:527-542).
// Each call encodes one frame on the GPU (history = the previous 1152 samples kept here).
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/toolame.h"
#include "../../include/toolame_b200.h"

namespace {

constexpr size_t BUFFER_SIZE = 4096; // ref: common.h BUFFER_SIZE
constexpr size_t MINIMUM = 4;        // ref: common.h MINIMUM

struct Stream {
    long sample_rate = 44100; // header.sampling_frequency = 0 / MPEG-1 until set (toolame.c:142-150)
    char mode = 's';
    int bitrate = 0;
    bool bitrate_set = false;
    int psy = 1;
    int pad_len = 0;
    tlb_batch *enc = nullptr;
    tlb_info info{};
    std::vector<int16_t> pcm;     // [2304 previous | 1152 current] * nch, interleaved (psy-2 looks 1632 samples back)
    std::vector<uint8_t> held;    // bytes still in the reference's bit buffer, oldest first
    std::vector<uint8_t> frame, rec, rec_prev;
    long frame_num = 0;
    bool failed = false;
} g;

void reset()
{
    if (g.enc) tlb_batch_destroy(g.enc);
    g = Stream();
}

bool open_encoder()
{
    if (g.enc) return true;
    if (g.failed) return false;
    tlb_config c{(int32_t)g.sample_rate, g.mode, g.bitrate, g.psy, g.pad_len};
    if (tlb_batch_create(&g.enc, &c, 0, 8) != 0 || tlb_batch_info(g.enc, &g.info) != 0) {
        std::fprintf(stderr, "libtoolame-b200: cannot start the encoder: %s\n", tlb_last_error());
        g.failed = true;
        return false;
    }
    g.pcm.assign((size_t)3 * 1152 * g.info.nch, 0);
    g.frame.resize((size_t)g.info.lg_frame);
    g.rec.assign((size_t)g.pad_len + 1, 0);
    g.rec_prev.assign((size_t)g.pad_len + 1, 0);
    g.held.reserve(BUFFER_SIZE + 2048);
    return true;
}

int drain(unsigned char *dst, size_t dst_size, size_t n)
{   // hand the n oldest held bytes to the caller (ref: bitstream.c:46-71, incl. its too-small-buffer behaviour)
    size_t w = n;
    if (w > dst_size) {
        std::fprintf(stderr, "ERROR: libtoolame output buffer too small (%zu vs %zu)!\n", dst_size, n);
        w = dst_size;
    }
    if (w) std::memcpy(dst, g.held.data(), w);
    g.held.erase(g.held.begin(), g.held.begin() + (long)n);
    return (int)w;
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
    // validate now, against the version and mode set so far (the reference reads both here: toolame.c:212-237)
    tlb_config c{(int32_t)g.sample_rate, g.mode, brate, 1, 0};
    tlb_batch *probe = nullptr;
    const int rc = tlb_batch_create(&probe, &c, 0, 1);
    if (probe) tlb_batch_destroy(probe);
    if (rc == TLB_E_PARAM || rc == TLB_E_UNSUPP) {
        std::fprintf(stderr, "libtoolame-b200: %s\n", tlb_last_error());
        return 1;
    }
    g.bitrate = brate;
    g.bitrate_set = true;
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
    if (!buffer || !output_buffer || !open_encoder()) return 0;
    const int nch = g.info.nch;
    const size_t lg = (size_t)g.info.lg_frame;
    g.frame_num++;
    int16_t *cur = g.pcm.data() + (size_t)2304 * nch;
    for (int i = 0; i < 1152; i++)
        for (int ch = 0; ch < nch; ch++) cur[i * nch + ch] = buffer[ch][i];
    const uint8_t *rec = nullptr;
    if (xpad_len && g.pad_len) {
        if (xpad_len > (size_t)g.pad_len || xpad_len < 2 || !xpad_data) {
            std::fprintf(stderr, "libtoolame-b200: bad xpad_len %zu (pad_len %d)\n", xpad_len, g.pad_len);
        } else {
            std::memcpy(g.rec.data(), xpad_data, (size_t)g.pad_len);
            g.rec[(size_t)g.pad_len] = (uint8_t)xpad_len;
            rec = g.rec.data();
        }
    }
    // History handed to the batch encoder: 0 at the stream start (the reference's zero state), otherwise what is
    // kept here (1152 samples after the first frame, 2304 from the third on).  Psy model 2 looks 1632 samples
    // back, more than one frame: its second frame is encoded together with the first, from the stream start.
    int rc;
    if (g.frame_num == 2 && g.psy == 2) {
        std::vector<uint8_t> two(2 * lg), recs;
        const uint8_t *rp = nullptr;
        if (g.pad_len) {
            recs.assign(g.rec_prev.begin(), g.rec_prev.end());
            if (rec) recs.insert(recs.end(), g.rec.begin(), g.rec.end());
            else recs.resize(2 * ((size_t)g.pad_len + 1), 0);
            rp = recs.data();
        }
        rc = tlb_batch_encode(g.enc, cur - (size_t)1152 * nch, 2, 0, 0, rp, two.data());
        std::memcpy(g.frame.data(), two.data() + lg, lg);
    } else {
        const size_t hist = g.frame_num > 2 ? 2304 : (size_t)(g.frame_num - 1) * 1152;
        rc = tlb_batch_encode(g.enc, cur, 1, hist, 0, rec, g.frame.data());
    }
    if (g.pad_len) {
        if (rec) g.rec_prev = g.rec;
        else g.rec_prev.assign((size_t)g.pad_len + 1, 0);
    }
    if (rc) {
        std::fprintf(stderr, "libtoolame-b200: encode failed: %s\n", tlb_last_error());
        return 0;
    }
    std::memmove(g.pcm.data(), g.pcm.data() + (size_t)1152 * nch, (size_t)2304 * nch * sizeof(int16_t));
    // this frame's ScF-CRC also replaces the previous frame's, which is still held (ref: toolame.c:527-539)
    if (g.frame_num > 1 && g.held.size() >= lg) {
        const size_t ext = (size_t)g.info.dab_ext;
        std::memcpy(g.held.data() + g.held.size() - 2 - ext, g.frame.data() + lg - 2 - ext, ext);
    }
    g.held.insert(g.held.end(), g.frame.begin(), g.frame.end());
    if (g.held.size() >= BUFFER_SIZE) return drain(output_buffer, output_buffer_size, BUFFER_SIZE - (lg + MINIMUM));
    return 0;
}

int toolame_finish(unsigned char *output_buffer, size_t output_buffer_size)
{
    if (!output_buffer) return 0;
    const int n = drain(output_buffer, output_buffer_size, g.held.size());
    reset();
    return n;
}

} // extern "C"

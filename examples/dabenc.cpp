// dabenc -- WAV (or raw s16le) file -> MP2 DAB file on the GPU.
//
// A small stand-in for `odr-audioenc --dab -i IN -o OUT` restricted to the path this repository builds: it shows the
// two ways a caller uses libtoolame_b200.so.
//
//   --stream   the reference's own loop, one frame per call through the unchanged libtoolame-dab API:
//              set-up order of src/odr-audioenc.cpp:687-721, de-interleave :1139-1155, toolame_encode_frame :1158,
//              re-framing into frames of 3*bitrate bytes with the reference's strict ">" hold-back :1208-1225,
//              File output :55-62 of src/Outputs.cpp.  Like odr-audioenc it never emits the last frame(s) still held
//              in the re-framer / the encoder's bit buffer at end of file.
//   (default)  the batch path: the whole file in one tlb_batch_encode call; every frame is written.
//   --edi F / --zmq F   (batch mode) also write what the EDI / ZeroMQ outputs of odr-audioenc would put on the wire for
//              these frames: the AF packets of Output::EDI::write_frame (src/Outputs.cpp:194-263) back to back, and the
//              ZeroMQ messages of Output::ZMQ::write_frame (src/Outputs.cpp:101-141), with the per-frame peak levels
//              (tlb_edi_packets / tlb_zmq_messages, include/dab_framing_b200.h).
//
// WAV parsing follows src/wavfile.cpp:74-182 (RIFF chunks, 'fmt ' incl. WAVE_FORMAT_EXTENSIBLE, 'data'); only
// 16-bit PCM is accepted (src/FileInput.cpp:60-75).  Gain (-g dB) and peak levels are done on the GPU in batch mode
// (tlb_batch_gain_peak_device) and as in src/odr-audioenc.cpp:1020-1055 in stream mode.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "../include/dab_framing_b200.h"
#include "../include/toolame.h"
#include "../include/toolame_b200.h"

static bool read_wav(const char *path, std::vector<int16_t> &pcm, int &channels, int &rate)
{
    FILE *f = std::fopen(path, "rb");
    if (!f) return false;
    std::vector<uint8_t> d;
    uint8_t buf[65536];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) d.insert(d.end(), buf, buf + n);
    std::fclose(f);
    auto u32 = [&](size_t o) { return (uint32_t)d[o] | d[o + 1] << 8 | d[o + 2] << 16 | (uint32_t)d[o + 3] << 24; };
    auto u16 = [&](size_t o) { return (uint16_t)(d[o] | d[o + 1] << 8); };
    if (d.size() < 12 || std::memcmp(&d[0], "RIFF", 4) || std::memcmp(&d[8], "WAVE", 4)) return false;
    int format = 0, bits = 0;
    for (size_t o = 12; o + 8 <= d.size();) {
        const uint32_t len = u32(o + 4);
        if (!std::memcmp(&d[o], "fmt ", 4) && len >= 16) {
            format = u16(o + 8); channels = u16(o + 10); rate = (int)u32(o + 12); bits = u16(o + 22);
            if (format == 0xfffe && len >= 28) format = (int)u32(o + 32);
        } else if (!std::memcmp(&d[o], "data", 4)) {
            if (format != 1 || bits != 16 || channels < 1 || channels > 2) return false;
            size_t bytes = std::min<size_t>(len ? len : d.size() - o - 8, d.size() - o - 8);
            pcm.resize(bytes / 2);
            std::memcpy(pcm.data(), &d[o + 8], pcm.size() * 2);
            return true;
        }
        o += 8 + len + (len & 1);
    }
    return false;
}

int main(int argc, char **argv)
{
    const char *in = nullptr, *out = nullptr, *edi_path = nullptr, *zmq_path = nullptr;
    int bitrate = 192, rate = 48000, channels = 2, psy = 1, raw = 0, stream = 0;
    double gain_db = 0;
    std::string mode;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "-i") in = val();
        else if (a == "-o") out = val();
        else if (a == "-b") bitrate = std::atoi(val());
        else if (a == "-r") rate = std::atoi(val());
        else if (a == "-c") channels = std::atoi(val());
        else if (a == "-g") gain_db = std::atof(val());
        else if (a == "--dabmode") mode = val();
        else if (a == "--dabpsy") psy = std::atoi(val());
        else if (a == "--raw") raw = 1;
        else if (a == "--stream") stream = 1;
        else if (a == "--edi") edi_path = val();
        else if (a == "--zmq") zmq_path = val();
        else { std::fprintf(stderr, "usage: dabenc -i IN.wav -o OUT.mp2 [-b kbps] [-r Hz -c ch --raw] [-g dB] [--dabmode s|d|j|m] [--dabpsy 0|1|2] [--stream] [--edi OUT.edi] [--zmq OUT.zmq]\n"); return 2; }
    }
    if (!in || !out) { std::fprintf(stderr, "dabenc: -i and -o are required\n"); return 2; }
    std::vector<int16_t> pcm;
    if (raw) {
        FILE *f = std::fopen(in, "rb");
        if (!f) { std::perror(in); return 1; }
        int16_t b[32768];
        size_t n;
        while ((n = std::fread(b, 2, 32768, f)) > 0) pcm.insert(pcm.end(), b, b + n);
        std::fclose(f);
    } else if (!read_wav(in, pcm, channels, rate)) {
        std::fprintf(stderr, "dabenc: %s is not a 16-bit PCM WAV file\n", in);
        return 1;
    }
    if (mode.empty()) mode = channels == 2 ? "j" : "m"; // src/odr-audioenc.cpp:697-709
    const size_t n_frames = pcm.size() / ((size_t)channels * 1152);
    FILE *fo = std::fopen(out, "wb");
    if (!fo) { std::perror(out); return 1; }
    int peak_l = 0, peak_r = 0;
    size_t written = 0;

    if (stream) {
        if (toolame_init() || toolame_set_samplerate(rate) || toolame_set_psy_model(psy) ||
            toolame_set_channel_mode(mode[0]) || toolame_set_bitrate(bitrate) || toolame_set_pad(0)) {
            std::fprintf(stderr, "dabenc: encoder set-up failed\n");
            return 1;
        }
        const double linear = std::pow(10.0, gain_db / 20.0);
        static short planar[2][1152];
        std::vector<uint8_t> outbuf(4092);
        std::deque<uint8_t> held;
        const size_t frame_len = 3 * (size_t)bitrate;
        for (size_t f = 0; f < n_frames; f++) {
            int16_t *p = pcm.data() + f * 1152 * channels;
            for (int i = 0; i + 1 < 1152 * channels; i += 2) { // (left, right) pairs, also in mono
                int16_t l = p[i], r = p[i + 1];
                if (linear != 1.0) { l = (int16_t)(int)(l * linear); r = (int16_t)(int)(r * linear); p[i] = l; p[i + 1] = r; }
                peak_l = std::max<int>(peak_l, l);
                peak_r = std::max<int>(peak_r, r);
            }
            for (int i = 0; i < 1152; i++)
                for (int ch = 0; ch < channels; ch++) planar[ch][i] = p[i * channels + ch];
            const int nb = toolame_encode_frame(planar, nullptr, 0, outbuf.data(), outbuf.size());
            held.insert(held.end(), outbuf.begin(), outbuf.begin() + nb);
            while (held.size() > frame_len) { // strict ">": one frame always stays behind
                std::vector<uint8_t> fr(held.begin(), held.begin() + (long)frame_len);
                held.erase(held.begin(), held.begin() + (long)frame_len);
                written += std::fwrite(fr.data(), 1, fr.size(), fo);
            }
        }
        toolame_finish(outbuf.data(), outbuf.size()); // odr-audioenc never reaches its own finish call (:904-908)
    } else {
        tlb_config cfg = {rate, mode[0], bitrate, psy, 0};
        tlb_batch *enc = nullptr;
        tlb_info info;
        if (tlb_batch_create(&enc, &cfg, 0, 0) || tlb_batch_info(enc, &info)) {
            std::fprintf(stderr, "dabenc: %s\n", tlb_last_error());
            return 1;
        }
        std::vector<uint8_t> mp2(n_frames * (size_t)info.lg_frame);
        std::vector<int16_t> peaks(2 * n_frames + 2);
        tlb_batch_set_gain(enc, gain_db, peaks.data()); // gain + peak levels on the GPU
        if (tlb_batch_encode(enc, pcm.data(), n_frames, 0, 0, nullptr, mp2.data())) {
            std::fprintf(stderr, "dabenc: %s\n", tlb_last_error());
            return 1;
        }
        written = std::fwrite(mp2.data(), 1, mp2.size(), fo);
        for (size_t f = 0; f < n_frames; f++) {
            peak_l = std::max<int>(peak_l, peaks[2 * f]);
            peak_r = std::max<int>(peak_r, peaks[2 * f + 1]);
        }
        // DAB frames are 24 ms = 3 * bitrate bytes (src/odr-audioenc.cpp:1211); at 48 kHz that is the MPEG frame, at
        // 24 kHz half of one (the level pair of an MPEG frame then goes with both halves)
        const size_t dab_len = 3 * (size_t)bitrate, n_dab = mp2.size() / dab_len, per = (size_t)info.lg_frame / dab_len;
        std::vector<int16_t> dab_peaks(2 * n_dab + 2);
        for (size_t d = 0; d < n_dab; d++) { dab_peaks[2 * d] = peaks[2 * (d / per)]; dab_peaks[2 * d + 1] = peaks[2 * (d / per) + 1]; }
        if (zmq_path) {
            std::vector<uint8_t> msgs(n_dab * (TLB_ZMQ_HEADER_SIZE + dab_len));
            const long n = tlb_zmq_messages(mp2.data(), n_dab, dab_len, dab_peaks.data(), msgs.data());
            FILE *fz = std::fopen(zmq_path, "wb");
            if (n < 0 || !fz || std::fwrite(msgs.data(), 1, (size_t)n, fz) != (size_t)n) { std::fprintf(stderr, "dabenc: --zmq failed\n"); return 1; }
            std::fclose(fz);
        }
        if (edi_path) {
            tlb_edi_config ec = {0, 0, 0, 37, 1, "dabenc (libtoolame_b200)"}; // no time stamp, fixed start second: reproducible
            tlb_edi *edi = nullptr;
            if (tlb_edi_create(&edi, &ec)) { std::fprintf(stderr, "dabenc: %s\n", tlb_last_error()); return 1; }
            std::vector<uint8_t> pk(n_dab * tlb_edi_packet_bound(edi, dab_len));
            const long n = tlb_edi_packets(edi, mp2.data(), n_dab, dab_len, dab_peaks.data(), pk.data(), pk.size(), nullptr);
            FILE *fe = std::fopen(edi_path, "wb");
            if (n < 0 || !fe || std::fwrite(pk.data(), 1, (size_t)n, fe) != (size_t)n) { std::fprintf(stderr, "dabenc: --edi failed\n"); return 1; }
            std::fclose(fe);
            tlb_edi_destroy(edi);
        }
        tlb_batch_destroy(enc);
    }
    std::fclose(fo);
    std::fprintf(stderr, "dabenc: %zu frames in, %zu bytes out (%s mode), peaks %d %d\n", n_frames, written,
                 stream ? "stream" : "batch", peak_l, peak_r);
    return 0;
}

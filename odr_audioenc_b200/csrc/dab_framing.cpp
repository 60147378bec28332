// dab_framing.cpp -- ZeroMQ / EDI framing of finished MP2 frames and PAD ingestion (include/dab_framing_b200.h).
// Host only.  "ref:" citations are relative to /root/reference/.
#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <new>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include "../../include/dab_framing_b200.h"
#include "tlb_internal.h"

namespace {

// ---- big-endian field writers into a caller-owned buffer
struct Writer {
    uint8_t *p;
    size_t n = 0;
    explicit Writer(uint8_t *dst) : p(dst) {}
    void u8(unsigned v) { p[n++] = (uint8_t)v; }
    void be16(unsigned v) { u8(v >> 8); u8(v); }
    void be24(unsigned v) { u8(v >> 16); u8(v >> 8); u8(v); }
    void be32(uint32_t v) { u8(v >> 24); u8(v >> 16); u8(v >> 8); u8(v); }
    void bytes(const void *src, size_t len) { std::memcpy(p + n, src, len); n += len; }
    void name(const char *four) { bytes(four, 4); }
};

// CRC-16 CCITT (x^16 + x^12 + x^5 + 1), MSB first, as contrib/crc.c:248-255 runs it over the AF packet; the caller
// starts at 0xffff and inverts the result (ref: contrib/edioutput/AFPacket.cpp:78-81)
uint16_t crc16_ccitt(uint16_t crc, const uint8_t *d, size_t len)
{
    static uint16_t tab[256];
    static bool ready = [] {
        for (unsigned i = 0; i < 256; i++) {
            unsigned c = i << 8;
            for (int b = 0; b < 8; b++) c = (c & 0x8000) ? (c << 1) ^ 0x1021 : c << 1;
            tab[i] = (uint16_t)c;
        }
        return true;
    }();
    (void)ready;
    while (len--) crc = (uint16_t)((crc << 8) ^ tab[(crc >> 8) ^ *d++]);
    return crc;
}

} // namespace

struct tlb_edi {
    tlb_edi_config cfg;
    std::string version;
    // state of Output::EDI (ref: src/Outputs.h) and of its packetisers
    int64_t edi_time = 0;              // m_edi_time
    int64_t send_version_at_time = 0;  // m_send_version_at_time
    uint32_t timestamp = 0;            // m_timestamp
    uint32_t num_seconds_sent = 0;     // m_num_seconds_sent
    uint16_t dlfc = 0;                 // TagDSTI::dlfc, modulo 5000
    uint16_t seq = 0;                  // AFPacketiser::m_seq
    bool started = false;
};

// Reed-Solomon RS(255,207) over GF(2^8), field polynomial x^8+x^4+x^3+x^2+1 (0x11d), generator roots alpha^1 ..
// alpha^48 -- the parameters of contrib/edioutput/PFT.cpp:103-110.  Systematic encoding: the 48 parity bytes are the
// remainder of data(x) x^48 divided by the generator polynomial, computed with the usual shift register.
struct Rs255_207 {
    static constexpr int NROOTS = 48;
    uint8_t exp[512], log[256], gen[NROOTS + 1]; // gen[] in log form, gen[NROOTS] = 1 (monic)
    Rs255_207()
    {
        unsigned v = 1;
        for (int i = 0; i < 255; i++) {
            exp[i] = exp[i + 255] = (uint8_t)v;
            log[v] = (uint8_t)i;
            v <<= 1;
            if (v & 0x100) v ^= 0x11d;
        }
        log[0] = 255; // (never used as a logarithm)
        uint8_t g[NROOTS + 1] = {1}; // coefficients, lowest order first
        for (int r = 0; r < NROOTS; r++) { // multiply by (x + alpha^(r+1))
            const int root = r + 1;
            g[r + 1] = 1;
            for (int j = r; j > 0; j--) g[j] = (uint8_t)(g[j - 1] ^ (g[j] ? exp[log[g[j]] + root] : 0));
            g[0] = g[0] ? exp[log[g[0]] + root] : 0;
        }
        for (int j = 0; j <= NROOTS; j++) gen[j] = log[g[j]];
    }
    void parity(const uint8_t data[207], uint8_t par[NROOTS]) const
    {
        std::memset(par, 0, NROOTS);
        for (int i = 0; i < 207; i++) {
            const uint8_t fb = (uint8_t)(data[i] ^ par[0]);
            std::memmove(par, par + 1, NROOTS - 1);
            par[NROOTS - 1] = 0;
            if (fb)
                for (int j = 0; j < NROOTS; j++) par[j] ^= exp[log[fb] + gen[NROOTS - 1 - j]];
        }
    }
};

struct tlb_pft {
    tlb_pft_config cfg;
    uint16_t pseq = 0; // PFT::m_pseq
    std::vector<uint8_t> block;
};

struct tlb_pad {
    int sock = -1;
    std::string ident;
    bool reachable = true;             // PadInterface::m_padenc_reachable
    std::vector<uint8_t> buffer;
};

extern "C" {

// ---- ZeroMQ -----------------------------------------------------------------------------------------------------
long tlb_zmq_message(const uint8_t *frame, size_t len, int16_t peak_left, int16_t peak_right, uint8_t *out)
{
    if (!frame || !out) return tlb_fail(TLB_E_ARG, "NULL argument");
    // ref: src/Outputs.h:76-92 (packed, host byte order) filled in as src/Outputs.cpp:110-127 does
    const uint16_t version = 1, encoder = TLB_ZMQ_ENCODER_MPEG_L2;
    const uint32_t datasize = (uint32_t)len;
    std::memcpy(out + 0, &version, 2);
    std::memcpy(out + 2, &encoder, 2);
    std::memcpy(out + 4, &datasize, 4);
    std::memcpy(out + 8, &peak_left, 2);
    std::memcpy(out + 10, &peak_right, 2);
    std::memcpy(out + TLB_ZMQ_HEADER_SIZE, frame, len);
    return (long)(TLB_ZMQ_HEADER_SIZE + len);
}

long tlb_zmq_messages(const uint8_t *frames, size_t n_frames, size_t frame_len, const int16_t *peaks, uint8_t *out)
{
    if (!frames || !out) return tlb_fail(TLB_E_ARG, "NULL argument");
    const size_t msg = TLB_ZMQ_HEADER_SIZE + frame_len;
    for (size_t f = 0; f < n_frames; f++)
        tlb_zmq_message(frames + f * frame_len, frame_len, peaks ? peaks[2 * f] : 0, peaks ? peaks[2 * f + 1] : 0, out + f * msg);
    return (long)(n_frames * msg);
}

// ---- EDI --------------------------------------------------------------------------------------------------------
int tlb_edi_create(tlb_edi **out, const tlb_edi_config *cfg)
{
    if (!out || !cfg) return tlb_fail(TLB_E_ARG, "NULL argument");
    *out = nullptr;
    if (cfg->tagpacket_alignment != 0 && cfg->tagpacket_alignment < 8)
        return tlb_fail(TLB_E_PARAM, "invalid TAG packet alignment"); // ref: TagPacket.cpp:69-72 (the reference only complains)
    tlb_edi *e = new (std::nothrow) tlb_edi();
    if (!e) return tlb_fail(TLB_E_ARG, "out of memory");
    e->cfg = *cfg;
    e->version = cfg->version_tag ? cfg->version_tag : "";
    e->cfg.version_tag = nullptr;
    *out = e;
    return 0;
}

void tlb_edi_destroy(tlb_edi *e) { delete e; }

size_t tlb_edi_packet_bound(const tlb_edi *e, size_t frame_len)
{
    const size_t ver = e ? e->version.size() : 0, align = e ? e->cfg.tagpacket_alignment : 0;
    // AF header 10 + *ptr 16 + dsti 8+2+8 + ss 8+3+len + ODRa 12 + ODRv 8+ver+4 + padding / *dmy + CRC 2
    return 10 + 16 + 18 + 11 + frame_len + 12 + 12 + ver + (align > 8 ? align : 8) + 2;
}

long tlb_edi_packet(tlb_edi *e, const uint8_t *frame, size_t len, int16_t peak_left, int16_t peak_right, uint8_t *out, size_t cap)
{
    if (!e || !frame || !out) return tlb_fail(TLB_E_ARG, "NULL argument");
    if (cap < tlb_edi_packet_bound(e, len)) return tlb_fail(TLB_E_ARG, "EDI output buffer too small");
    // ---- time keeping (ref: src/Outputs.cpp:200-228)
    if (!e->started) {
        e->started = true;
        int64_t now = e->cfg.start_time;
        if (now == 0) now = (int64_t)std::chrono::system_clock::to_time_t(std::chrono::time_point_cast<std::chrono::seconds>(std::chrono::system_clock::now()));
        e->edi_time = now + e->cfg.delay_ms / 1000;
        e->send_version_at_time = e->edi_time;
        for (int32_t sub_ms = (int32_t)(e->cfg.delay_ms % 1000); sub_ms > 0; sub_ms -= 24) e->timestamp += 24 << 14;
    }
    e->timestamp += 24 << 14; // 24 ms at time stamp level 2
    if (e->timestamp > 0xf9FFff) {
        e->timestamp -= 0xfa0000; // 16 384 000 = one second
        e->edi_time += 1;
        e->num_seconds_sent++;
    }
    Writer w(out);
    // ---- AF header (ref: contrib/edioutput/AFPacket.cpp:44-69); the length is filled in below
    w.name("AF\0\0");
    w.n = 2;
    w.be32(0);
    w.be16(e->seq++);
    w.u8(0x80 | 0x10);   // CRC present, version 1.0
    w.u8('T');           // payload: TAG packet
    const size_t payload0 = w.n;
    // ---- *ptr (ref: TagItems.cpp:46-67): protocol "DSTI", version 0.0
    w.name("*ptr");
    w.be32(0x40);
    w.name("DSTI");
    w.be16(0);
    w.be16(0);
    // ---- dsti (ref: TagItems.cpp:200-252): stihf = 0, rfadf = 0, atstf = tist
    {
        const bool atstf = e->cfg.tist != 0;
        w.name("dsti");
        w.be32((2 + (atstf ? 8 : 0)) * 8);
        const unsigned dfctl = e->dlfc % 250, dfcth = e->dlfc / 250;
        w.be16(dfctl | (dfcth << 8) | (0u << 13) | ((atstf ? 1u : 0u) << 14) | (0u << 15));
        if (atstf) { // ref: TagItems.cpp:254-261 (set_edi_time): seconds since 2000-01-01 incl. leap seconds
            const uint8_t utco = (uint8_t)(e->cfg.tai_utc_offset - 32);
            const uint32_t seconds = (uint32_t)(e->edi_time - 946684800 + utco);
            w.u8(utco);
            w.be32(seconds);
            w.be24(e->timestamp & 0xffffff);
        }
        e->dlfc = (uint16_t)((e->dlfc + 1) % 5000);
    }
    // ---- ss1 (ref: TagItems.cpp:293-349): sub-channel stream 1, all of rfa / tid / tidext / crcstf / stid zero
    w.u8('s');
    w.u8('s');
    w.be16(1);
    w.be32((uint32_t)(3 + len) * 8);
    w.be24(0);
    w.bytes(frame, len);
    // ---- ODRa (ref: TagItems.cpp:417-446)
    w.name("ODRa");
    w.be32(4 * 8);
    w.be16((uint16_t)peak_left);
    w.be16((uint16_t)peak_right);
    // ---- ODRv every ten seconds (ref: src/Outputs.cpp:250-254, TagItems.cpp:385-409)
    if (e->send_version_at_time < e->edi_time) {
        e->send_version_at_time += 10;
        w.name("ODRv");
        w.be32((uint32_t)(e->version.size() + 4) * 8);
        w.bytes(e->version.data(), e->version.size());
        w.be32(e->num_seconds_sent);
    }
    // ---- TAG packet padding (ref: TagPacket.cpp:58-72)
    const unsigned align = e->cfg.tagpacket_alignment;
    if (align == 8) {
        while ((w.n - payload0) % 8) w.u8(0);
    } else if (align > 8) { // ref: TagItems.cpp:362-378: "*dmy" with align - 8 bytes of undefined (here zero) data
        w.name("*dmy");
        w.be32((align - 8) * 8);
        std::memset(w.p + w.n, 0, align - 8);
        w.n += align - 8;
    }
    const uint32_t taglength = (uint32_t)(w.n - payload0);
    out[2] = (uint8_t)(taglength >> 24); out[3] = (uint8_t)(taglength >> 16); out[4] = (uint8_t)(taglength >> 8); out[5] = (uint8_t)taglength;
    const uint16_t crc = (uint16_t)(crc16_ccitt(0xffff, out, w.n) ^ 0xffff);
    w.be16(crc);
    return (long)w.n;
}

long tlb_edi_packets(tlb_edi *e, const uint8_t *frames, size_t n_frames, size_t frame_len, const int16_t *peaks,
                     uint8_t *out, size_t cap, uint32_t *sizes)
{
    if (!e || !frames || !out) return tlb_fail(TLB_E_ARG, "NULL argument");
    size_t at = 0;
    for (size_t f = 0; f < n_frames; f++) {
        const long n = tlb_edi_packet(e, frames + f * frame_len, frame_len, peaks ? peaks[2 * f] : 0, peaks ? peaks[2 * f + 1] : 0,
                                      out + at, cap - at);
        if (n < 0) return n;
        if (sizes) sizes[f] = (uint32_t)n;
        at += (size_t)n;
    }
    return (long)at;
}

// ---- PFT (ref: contrib/edioutput/PFT.cpp) -------------------------------------------------------------------------------
static size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }

int tlb_pft_create(tlb_pft **out, const tlb_pft_config *cfg)
{
    if (!out || !cfg) return tlb_fail(TLB_E_ARG, "NULL argument");
    *out = nullptr;
    if (cfg->chunk_len > 207) return tlb_fail(TLB_E_PARAM, "EDI PFT: maximum chunk size is 207"); // ref: PFT.cpp:59-63
    tlb_pft *p = new (std::nothrow) tlb_pft();
    if (!p) return tlb_fail(TLB_E_ARG, "out of memory");
    p->cfg = *cfg;
    if (!p->cfg.chunk_len) p->cfg.chunk_len = 207;
    *out = p;
    return 0;
}

void tlb_pft_destroy(tlb_pft *p) { delete p; }

size_t tlb_pft_bound(const tlb_pft *p, size_t af_len, size_t *max_fragments)
{
    if (!p || !af_len) return 0;
    size_t payload = af_len, frags;
    if (p->cfg.fec) {
        const size_t c = ceil_div(af_len, p->cfg.chunk_len), k = ceil_div(af_len, c);
        payload = c * (k + 48);
        frags = ceil_div(payload, c * 48 / (p->cfg.fec + 1));
    } else frags = ceil_div(af_len, 1400);
    if (max_fragments) *max_fragments = frags;
    return payload + frags * (14 + 2 + 1); // header (12 + RSk/RSz) + CRC, and the rounding of the fragment size
}

long tlb_pft_fragments(tlb_pft *p, const uint8_t *af, size_t af_len, uint8_t *out, size_t cap, uint32_t *sizes, size_t max_fragments)
{
    if (!p || !af || !out || !af_len) return tlb_fail(TLB_E_ARG, "NULL argument");
    const bool rs = p->cfg.fec > 0;
    size_t n_frag, frag_size, chunk_len = 0, zero_pad = 0, total;
    const uint8_t *payload = af;
    if (rs) {
        // ref: PFT.cpp:76-137: c = ceil(l / k_max) chunks of k = ceil(l / c) bytes, the last one zero padded; every chunk,
        // padded to 207 bytes behind its data, gets 48 parity bytes
        static const Rs255_207 code;
        const size_t c = ceil_div(af_len, p->cfg.chunk_len);
        chunk_len = ceil_div(af_len, c);
        zero_pad = c * chunk_len - af_len;
        p->block.assign(c * (chunk_len + 48), 0);
        for (size_t i = 0; i < c; i++) {
            uint8_t word[207] = {0};
            const size_t at = i * chunk_len, n = af_len - at < chunk_len ? af_len - at : chunk_len;
            std::memcpy(word, af + at, n);
            uint8_t *dst = p->block.data() + i * (chunk_len + 48);
            std::memcpy(dst, word, chunk_len);
            code.parity(word, dst + chunk_len);
        }
        // ref: PFT.cpp:157-190: s_max = floor(c p / (m + 1)), f = ceil(L / s_max) fragments of ceil(L / f) bytes, byte j of
        // fragment i = block[j f + i] (interleaved), zero beyond the block
        total = p->block.size();
        const size_t s_max = c * 48 / (p->cfg.fec + 1);
        n_frag = ceil_div(total, s_max);
        frag_size = ceil_div(total, n_frag);
        payload = p->block.data();
    } else { // ref: PFT.cpp:192-222: plain fragmentation, payloads of at most 1400 bytes, the last one may be shorter
        total = af_len;
        n_frag = ceil_div(af_len, 1400);
        frag_size = ceil_div(af_len, n_frag);
    }
    if (n_frag > max_fragments && sizes) return tlb_fail(TLB_E_ARG, "PFT: more fragments than sizes[] holds");
    size_t at = 0;
    for (size_t i = 0; i < n_frag; i++) {
        const size_t len = rs ? frag_size : (total - i * frag_size < frag_size ? total - i * frag_size : frag_size);
        const size_t hdr = 12 + (rs ? 2 : 0) + 2;
        if (at + hdr + len > cap) return tlb_fail(TLB_E_ARG, "PFT output buffer too small");
        Writer w(out + at);
        // ref: PFT.cpp:253-309: PF header = "PF", Pseq, Findex (24 bits), Fcount (24 bits), FEC / Addr flags + Plen,
        // [RSk, RSz], CRC over the header; the transport header is never used
        w.u8('P');
        w.u8('F');
        w.be16(p->pseq);
        w.be24((unsigned)i);
        w.be24((unsigned)n_frag);
        w.be16((unsigned)len | (rs ? 0x8000u : 0u));
        if (rs) {
            w.u8((unsigned)chunk_len);
            w.u8((unsigned)zero_pad);
        }
        w.be16((uint16_t)(crc16_ccitt(0xffff, out + at, w.n) ^ 0xffff));
        if (rs) {
            for (size_t j = 0; j < len; j++) {
                const size_t ix = j * n_frag + i;
                w.u8(ix < total ? payload[ix] : 0);
            }
        } else w.bytes(payload + i * frag_size, len);
        if (sizes) sizes[i] = (uint32_t)w.n;
        at += w.n;
    }
    p->pseq++;
    return (long)n_frag;
}

// ---- PAD --------------------------------------------------------------------------------------------------------
int tlb_pad_open(tlb_pad **out, const char *ident)
{
    if (!out || !ident || !*ident) return tlb_fail(TLB_E_ARG, "NULL argument");
    *out = nullptr;
    if (std::strlen(ident) > sizeof(((struct sockaddr_un *)nullptr)->sun_path) - sizeof("/tmp/.audioenc"))
        return tlb_fail(TLB_E_ARG, "PAD identifier too long for a socket path"); // (would be cut short silently)
    tlb_pad *p = new (std::nothrow) tlb_pad();
    if (!p) return tlb_fail(TLB_E_ARG, "out of memory");
    p->ident = ident;
    p->buffer.resize(2048);
    // ref: src/PadInterface.cpp:37-70
    p->sock = ::socket(AF_UNIX, SOCK_DGRAM, 0);
    if (p->sock == -1) { delete p; return tlb_fail(TLB_E_ARG, "PAD socket creation failed"); }
    const int flags = fcntl(p->sock, F_GETFL);
    if (flags == -1 || fcntl(p->sock, F_SETFL, flags | O_NONBLOCK) == -1) {
        ::close(p->sock);
        delete p;
        return tlb_fail(TLB_E_ARG, "PAD socket: could not set O_NONBLOCK");
    }
    struct sockaddr_un claddr;
    std::memset(&claddr, 0, sizeof claddr);
    claddr.sun_family = AF_UNIX;
    std::snprintf(claddr.sun_path, sizeof claddr.sun_path, "/tmp/%s.audioenc", ident);
    if (unlink(claddr.sun_path) == -1 && errno != ENOENT)
        std::fprintf(stderr, "Unlinking of socket %s failed: %s\n", claddr.sun_path, std::strerror(errno));
    if (::bind(p->sock, (const struct sockaddr *)&claddr, sizeof claddr) == -1) {
        ::close(p->sock);
        delete p;
        return tlb_fail(TLB_E_ARG, "PAD socket bind failed");
    }
    *out = p;
    return 0;
}

void tlb_pad_close(tlb_pad *p)
{
    if (!p) return;
    if (p->sock != -1) {
        ::close(p->sock);
        char path[108];
        std::snprintf(path, sizeof path, "/tmp/%s.audioenc", p->ident.c_str());
        unlink(path);
    }
    delete p;
}

int tlb_pad_request(tlb_pad *p, int pad_len, uint8_t *record)
{
    if (!p || !record || pad_len <= 0 || pad_len > 255) return tlb_fail(TLB_E_ARG, "bad PAD request");
    std::memset(record, 0, (size_t)pad_len + 1);
    // ref: src/PadInterface.cpp:72-113: the request tells ODR-PadEnc the length and paces it
    const uint8_t packet[2] = {1 /* MESSAGE_REQUEST */, (uint8_t)pad_len};
    struct sockaddr_un claddr;
    std::memset(&claddr, 0, sizeof claddr);
    claddr.sun_family = AF_UNIX;
    std::snprintf(claddr.sun_path, sizeof claddr.sun_path, "/tmp/%s.padenc", p->ident.c_str());
    const ssize_t sent = ::sendto(p->sock, packet, sizeof packet, 0, (struct sockaddr *)&claddr, sizeof claddr);
    if (sent == -1) {
        if (errno == EAGAIN || errno == EWOULDBLOCK || errno == ECONNREFUSED || errno == ENOENT) {
            if (p->reachable) std::fprintf(stderr, "ODR-PadEnc at %s not reachable\n", claddr.sun_path);
            p->reachable = false;
        } else std::fprintf(stderr, "PAD request send failed: %s\n", std::strerror(errno));
    } else if (!p->reachable) {
        std::fprintf(stderr, "ODR-PadEnc is now reachable at %s\n", claddr.sun_path);
        p->reachable = true;
    }
    // ref: src/PadInterface.cpp:115-149: take the first MESSAGE_PAD_DATA datagram that is waiting, if any
    for (;;) {
        const ssize_t got = ::recvfrom(p->sock, p->buffer.data(), p->buffer.size(), 0, nullptr, nullptr);
        if (got == -1) {
            if (errno == EAGAIN || errno == EWOULDBLOCK) return 0; // nothing waiting: no PAD for this frame
            return tlb_fail(TLB_E_ARG, "PAD socket: receive failed");
        }
        if (got > 0 && p->buffer[0] == 2 /* MESSAGE_PAD_DATA */) {
            // ref: src/odr-audioenc.cpp:826-851: the payload must be pad_len + 1 bytes, the last one the used length (>= 2)
            if ((size_t)got - 1 != (size_t)pad_len + 1) return tlb_fail(TLB_E_ARG, "Incorrect PAD length received");
            const int used = p->buffer[(size_t)pad_len + 1];
            if (used < 2) return tlb_fail(TLB_E_ARG, "Invalid X-PAD length");
            std::memcpy(record, p->buffer.data() + 1, (size_t)pad_len + 1);
            return used;
        }
    }
}

long tlb_pad_fill(tlb_pad *p, int pad_len, size_t n_frames, uint8_t *records)
{
    long with_pad = 0;
    for (size_t f = 0; f < n_frames; f++) {
        const int r = tlb_pad_request(p, pad_len, records + f * ((size_t)pad_len + 1));
        if (r < 0) return r;
        with_pad += r > 0;
    }
    return with_pad;
}

} // extern "C"

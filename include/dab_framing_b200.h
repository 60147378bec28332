/*
 * dab_framing_b200.h -- the step AFTER the encoder in odr-audioenc's loop, for batches of finished MP2 frames, and
 * the step before it for PAD: host-only C ABI (no CUDA call), part of libtoolame_b200.so.
 *
 *  (1) ZeroMQ message framing of a frame for ODR-DabMux: the 12-byte zmq_frame_header_t of the reference
 *      (src/Outputs.h:76-99) in front of the frame, as Output::ZMQ::write_frame builds it (src/Outputs.cpp:101-141).
 *  (2) EDI: one AF packet (ETSI TS 102 821, 6.1) per frame holding the TAG packet *ptr / dsti / ss1 / ODRa (/ ODRv),
 *      as Output::EDI::write_frame (src/Outputs.cpp:194-263) with contrib/edioutput/TagItems.cpp, TagPacket.cpp and
 *      AFPacket.cpp assembles it, including the frame counter, the 24 ms time stamp arithmetic and the CRC; and the
 *      PFT layer on top of it (ETSI TS 102 821, 7: fragmentation, optional Reed-Solomon RS(255,207) protection and
 *      interleaving, PF headers), which the reference always switches on for UDP destinations
 *      (src/Outputs.cpp:156-165, contrib/edioutput/PFT.cpp).  The sockets are not built: the functions return the
 *      packet BYTES; a caller sends them over UDP/TCP (or ZeroMQ) as it likes.
 *  (3) PAD ingestion: the ODR-PadEnc request / reply protocol over UNIX datagram sockets
 *      (src/PadInterface.cpp:37-150), delivering records in exactly the layout tlb_batch_encode's `xpad` takes.
 *
 * Everything returns sizes (>= 0) or a negative TLB_E_* code (toolame_b200.h); nothing allocates on the hot path.
 */
#ifndef DAB_FRAMING_B200_H
#define DAB_FRAMING_B200_H

#include <stddef.h>
#include <stdint.h>

#include "toolame_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- (1) ZeroMQ framing --------------------------------------------------------------------------------- */
#define TLB_ZMQ_HEADER_SIZE 12   /* sizeof(zmq_frame_header_t), packed: u16 version, u16 encoder, u32 datasize, i16 left, i16 right */
#define TLB_ZMQ_ENCODER_MPEG_L2 2 /* src/Outputs.h:95 */

/* One message: header + frame into out (TLB_ZMQ_HEADER_SIZE + len bytes); returns the message size. */
TLB_API long tlb_zmq_message(const uint8_t *frame, size_t len, int16_t peak_left, int16_t peak_right, uint8_t *out);
/* n_frames messages of (TLB_ZMQ_HEADER_SIZE + frame_len) bytes each, back to back, from a batch output buffer;
 * peaks = per-frame (left, right) as tlb_batch_set_gain / tlb_batch_gain_peak_device deliver them, or NULL (0, 0). */
TLB_API long tlb_zmq_messages(const uint8_t *frames, size_t n_frames, size_t frame_len, const int16_t *peaks, uint8_t *out);

/* ---- (2) EDI ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t tist;                 /* 1: carry the time stamp in dsti (EDI::set_tist) */
    uint32_t delay_ms;            /* tist delay (EDI::set_tist) */
    uint32_t tagpacket_alignment; /* edi::configuration_t::tagpacket_alignment (0 = none, 8 = pad, > 8 = *dmy item) */
    int32_t tai_utc_offset;       /* TAI - UTC in seconds (the reference asks ClockTAI; 37 since 2017) */
    int64_t start_time;           /* POSIX seconds of the first frame; 0 = the system clock at the first frame */
    const char *version_tag;      /* ODRv text (EDI::set_odr_version_tag); NULL = "" */
} tlb_edi_config;

typedef struct tlb_edi tlb_edi;
TLB_API int tlb_edi_create(tlb_edi **out, const tlb_edi_config *cfg);
TLB_API void tlb_edi_destroy(tlb_edi *e);
/* Upper bound of one AF packet for a frame of frame_len bytes. */
TLB_API size_t tlb_edi_packet_bound(const tlb_edi *e, size_t frame_len);
/* The AF packet of the next frame of the stream into out (capacity cap); returns its size.  Stateful exactly as
 * the reference: dsti frame counter, AF sequence number, 24 ms time stamp, ODRv every ten seconds. */
TLB_API long tlb_edi_packet(tlb_edi *e, const uint8_t *frame, size_t len, int16_t peak_left, int16_t peak_right,
                            uint8_t *out, size_t cap);
/* A batch: n_frames packets back to back into out; sizes[i] receives the size of packet i. Returns the total. */
TLB_API long tlb_edi_packets(tlb_edi *e, const uint8_t *frames, size_t n_frames, size_t frame_len, const int16_t *peaks,
                             uint8_t *out, size_t cap, uint32_t *sizes);

/* PFT layer: an AF packet becomes one or more PF fragments (each a datagram).  fec = 0: fragmentation only, payloads of
 * at most 1400 bytes; fec = m > 0: the packet is cut into chunks of at most chunk_len bytes, each protected by 48
 * Reed-Solomon bytes (RS(255,207), x^8+x^4+x^3+x^2+1, first root alpha^1), and the protected block is interleaved
 * over fragments sized so that m lost fragments can be recovered (contrib/edioutput/PFT.cpp:76-231). */
typedef struct {
    uint32_t fec;        /* edi::configuration_t::fec (EDI::set_fec) */
    uint32_t chunk_len;  /* edi::configuration_t::chunk_len; 0 = 207 */
} tlb_pft_config;
typedef struct tlb_pft tlb_pft;
TLB_API int tlb_pft_create(tlb_pft **out, const tlb_pft_config *cfg);
TLB_API void tlb_pft_destroy(tlb_pft *p);
/* Upper bounds for one AF packet of af_len bytes: total bytes of all fragments (return value) and their number. */
TLB_API size_t tlb_pft_bound(const tlb_pft *p, size_t af_len, size_t *max_fragments);
/* Fragments of the next AF packet of the stream (the PF sequence number counts packets), back to back into out;
 * sizes[i] receives the size of fragment i (at most max_fragments entries).  Returns the number of fragments. */
TLB_API long tlb_pft_fragments(tlb_pft *p, const uint8_t *af_packet, size_t af_len, uint8_t *out, size_t cap,
                               uint32_t *sizes, size_t max_fragments);

/* ---- (3) PAD ingestion ---------------------------------------------------------------------------------- */
typedef struct tlb_pad tlb_pad;
/* Bind /tmp/<ident>.audioenc (non-blocking datagram socket), talk to /tmp/<ident>.padenc (PadInterface::open). */
TLB_API int tlb_pad_open(tlb_pad **out, const char *ident);
TLB_API void tlb_pad_close(tlb_pad *p);
/* Request PAD for one frame (PadInterface::request) and store the reply as one X-PAD record of pad_len + 1 bytes
 * (data right-aligned, last byte = used length: src/odr-audioenc.cpp:819-852).  Returns the used length, 0 when
 * ODR-PadEnc had nothing (the record is zeroed: "no PAD"), TLB_E_ARG for a reply of the wrong size or a used
 * length of 1 (the reference stops encoding on both). */
TLB_API int tlb_pad_request(tlb_pad *p, int pad_len, uint8_t *record);
/* n_frames records for a batch, one request each. Returns the number of records that carry PAD. */
TLB_API long tlb_pad_fill(tlb_pad *p, int pad_len, size_t n_frames, uint8_t *records);

#ifdef __cplusplus
}
#endif
#endif

/*
 * toolame.h -- the libtoolame-dab encoder API, served by the B200 implementation.
 *
 * Drop-in for the reference's libtoolame-dab/toolame.h:13-48: the same nine symbols
 * (libtoolame-dab.sym:1-9), argument meaning and return convention, so odr-audioenc
 * (src/odr-audioenc.cpp:687-721,1158,1161) links against libtoolame_b200.so unchanged.
 * Setters return 0 on success and non-zero on error; encode / finish return the number
 * of bytes written to output_buffer.
 *
 * Differences, all on error paths: an illegal bitrate makes toolame_set_bitrate return 1 (the reference exit()s
 * inside BitrateIndex, common.c:110-115); psychoacoustic model 3 is refused by toolame_set_psy_model (it reads an
 * uninitialised array in the reference, psycho_3.c:84); 44.1 / 22.05 kHz are refused when the bitrate is set; a batch of
 * frames that cannot be encoded (no CUDA device, or a CUDA error that survives one retry) stops the stream -- every
 * later toolame_encode_frame returns 0 bytes, toolame_finish returns what was complete, toolame_init starts again --
 * instead of leaving a gap in it (toolame_b200_status() in toolame_b200.h tells why).
 *
 * Frames are encoded lazily: the reference hands nothing back until its 4096-byte buffer has filled, so all frames
 * pending since the last flush are encoded in one GPU batch on the call in which a flush falls due.  Bytes and return
 * sizes are the reference's; setters are validated on the host and touch no GPU.
 */
#ifndef TOOLAME_B200_COMPAT_H
#define TOOLAME_B200_COMPAT_H
#include <stddef.h>

#if defined(__GNUC__)
#define TLB_API __attribute__((visibility("default")))
#else
#define TLB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Start a new stream with the reference's defaults (MPEG-1, stereo, psy model 1, CRC on, 4 ScF-CRC bytes). */
TLB_API int toolame_init(void);

/* Flush the bytes still held back (the reference keeps the newest lg_frame + 4 bytes so that the previous
 * frame's ScF-CRC can be patched); returns the number of bytes written. */
TLB_API int toolame_finish(unsigned char *output_buffer, size_t output_buffer_size);

/* Accepted and ignored, as in the reference (the flag is never read there). */
TLB_API int toolame_enable_byteswap(void);

/* 's' stereo, 'd' dual channel, 'j' joint stereo, 'm' mono */
TLB_API int toolame_set_channel_mode(const char mode);

/* 0..3 are valid for the reference; this build implements models 0, 1 and 2 */
TLB_API int toolame_set_psy_model(int new_model);

/* kbit/s; must be called after toolame_set_samplerate and toolame_set_channel_mode (it depends on both) */
TLB_API int toolame_set_bitrate(int brate);

/* Hz */
TLB_API int toolame_set_samplerate(long sample_rate);

/* bytes of PAD (X-PAD + 2 bytes F-PAD) the caller will hand over per frame */
TLB_API int toolame_set_pad(int pad_len);

/* Encode 1152 samples per channel (planar, buffer[ch][i]; mono uses buffer[0]).  xpad_data points at pad_len
 * bytes of which the last xpad_len are used (the final two being the F-PAD); xpad_len is 0 or >= 2.
 * Output is not frame aligned: 0 bytes on most calls, 4096 - (lg_frame + 4) bytes whenever the internal
 * 4096-byte buffer has filled (bitstream.c:46-71). */
TLB_API int toolame_encode_frame(short buffer[2][1152], unsigned char *xpad_data, size_t xpad_len,
                         unsigned char *output_buffer, size_t output_buffer_size);

#ifdef __cplusplus
}
#endif
#endif

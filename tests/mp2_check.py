"""Independent structural checker for MPEG Layer II DAB frames: parses a stream frame by frame from
the ISO 11172-3 / 13818-3 syntax alone -- sync word, header fields, bit allocation, scfsi, scalefactors, sample
codewords -- recomputes the CRC-16 over the protected bits and the DAB ScF-CRC bytes, and checks the frame pitch.
Shares no code with the oracle or the CUDA path (the allocation tables are restated here from table B.2)."""
import numpy as np

BITRATES = {1: [0, 32, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384],
            0: [0, 8, 16, 24, 32, 40, 48, 56, 64, 80, 96, 112, 128, 144, 160]}
SFREQ = {1: [44100, 48000, 32000], 0: [22050, 24000, 16000]}
# quantiser classes by number of steps -> (bits per codeword, grouped?)
CLASS = {3: (5, True), 5: (7, True), 7: (3, False), 9: (10, True), 15: (4, False), 31: (5, False), 63: (6, False),
         127: (7, False), 255: (8, False), 511: (9, False), 1023: (10, False), 2047: (11, False), 4095: (12, False),
         8191: (13, False), 16383: (14, False), 32767: (15, False), 65535: (16, False)}
ROW_A = [3, 7, 15, 31, 63, 127, 255, 511, 1023, 2047, 4095, 8191, 16383, 32767, 65535]          # nbal 4, sb 0-2 (B.2a/b)
ROW_B = [3, 5, 7, 9, 15, 31, 63, 127, 255, 511, 1023, 2047, 4095, 8191, 65535]                   # nbal 4, sb 3-10
ROW_C = [3, 5, 7, 9, 15, 31, 65535]                                                              # nbal 3, sb 11-22
ROW_D = [3, 5, 65535]                                                                            # nbal 2, sb 23-
ROW_E = [3, 5, 9, 15, 31, 63, 127, 255, 511, 1023, 2047, 4095, 8191, 16383, 32767]               # nbal 4, B.2c/d sb 0-1
ROW_F = [3, 5, 9, 15, 31, 63, 127]                                                               # nbal 3, B.2c/d sb 2-
ROW_G = [3, 5, 7, 9, 15, 31, 63, 127, 255, 511, 1023, 2047, 4095, 8191, 16383]                   # nbal 4, LSF sb 0-3
ROW_H = [3, 5, 9, 15, 31, 63, 127]                                                               # nbal 3, LSF sb 4-10
ROW_I = [3, 5, 9]                                                                                # nbal 2, LSF sb 11-29


def alloc_table(version, fs, kbps_per_ch):
    """(sblimit, [(nbal, steps list)] per subband)"""
    if version == 0:
        return 30, [(4, ROW_G)] * 4 + [(3, ROW_H)] * 7 + [(2, ROW_I)] * 19
    if (fs == 48000 and kbps_per_ch >= 56) or 56 <= kbps_per_ch <= 80:
        return 27, [(4, ROW_A)] * 3 + [(4, ROW_B)] * 8 + [(3, ROW_C)] * 12 + [(2, ROW_D)] * 4
    if fs != 48000 and kbps_per_ch >= 96:
        return 30, [(4, ROW_A)] * 3 + [(4, ROW_B)] * 8 + [(3, ROW_C)] * 12 + [(2, ROW_D)] * 7
    if fs != 32000 and kbps_per_ch <= 48:
        return 8, [(4, ROW_E)] * 2 + [(3, ROW_F)] * 6
    return 12, [(4, ROW_E)] * 2 + [(3, ROW_F)] * 10


class Bits:
    def __init__(self, frame):
        self.bits = np.unpackbits(np.asarray(frame, dtype=np.uint8))
        self.pos = 0

    def get(self, n):
        v = 0
        for b in self.bits[self.pos:self.pos + n]:
            v = (v << 1) | int(b)
        self.pos += n
        return v


def crc_update(crc, data, length, poly, width):
    top = 1 << (width - 1)
    for i in range(length - 1, -1, -1):
        carry = crc & top
        crc = (crc << 1) & ((1 << width) - 1)
        if bool(carry) != bool((data >> i) & 1):
            crc ^= poly
    return crc


def check_stream(stream, xpad_bytes_of_frame=None):
    """Parse a DAB MP2 stream.  Returns a list of per-frame dicts; raises AssertionError on any structural fault.
    The ScF-CRC bytes of frame n are checked against frame n+1's scalefactors (the last frame against its own)."""
    stream = np.asarray(stream, dtype=np.uint8)
    frames, pos = [], 0
    while pos < stream.size:
        hdr = Bits(stream[pos:pos + 4])
        assert hdr.get(12) == 0xFFF, "sync word lost at byte %d" % pos
        version, layer, prot = hdr.get(1), hdr.get(2), hdr.get(1)
        assert layer == 2 and prot == 0, "not Layer II with CRC"
        br_idx, sf_idx, padding, _ext = hdr.get(4), hdr.get(2), hdr.get(1), hdr.get(1)
        mode, mode_ext = hdr.get(2), hdr.get(2)
        assert 0 < br_idx < 15 and sf_idx < 3 and padding == 0
        kbps, fs = BITRATES[version][br_idx], SFREQ[version][sf_idx]
        lg = int(1152 / (fs / 1000.0) * kbps / 8)
        assert pos + lg <= stream.size, "truncated frame"
        b = Bits(stream[pos:pos + lg])
        b.pos = 32
        crc_rx = b.get(16)
        nch = 1 if mode == 3 else 2
        sblimit, table = alloc_table(version, fs, kbps // nch)
        jsbound = {0: 4, 1: 8, 2: 12, 3: 16}[mode_ext] if mode == 1 else sblimit
        jsbound = min(jsbound, sblimit)
        crc = 0xFFFF
        crc = crc_update(crc, int.from_bytes(bytes(stream[pos + 2:pos + 4]), "big"), 16, 0x8005, 16)
        alloc = np.zeros((2, 32), dtype=int)
        for sb in range(sblimit):
            nbal = table[sb][0]
            for ch in range(nch if sb < jsbound else 1):
                alloc[ch][sb] = b.get(nbal)
                crc = crc_update(crc, alloc[ch][sb], nbal, 0x8005, 16)
            if sb >= jsbound and nch == 2:
                alloc[1][sb] = alloc[0][sb]
        scfsi = np.zeros((2, 32), dtype=int)
        for sb in range(sblimit):
            for ch in range(nch):
                if alloc[ch][sb]:
                    scfsi[ch][sb] = b.get(2)
                    crc = crc_update(crc, scfsi[ch][sb], 2, 0x8005, 16)
        assert crc == crc_rx, "CRC-16 mismatch in the frame at byte %d" % pos
        scf = np.zeros((2, 3, 32), dtype=int)
        for sb in range(sblimit):
            for ch in range(nch):
                if alloc[ch][sb]:
                    s = scfsi[ch][sb]
                    if s == 0:
                        scf[ch, :, sb] = [b.get(6), b.get(6), b.get(6)]
                    elif s == 1:
                        a = b.get(6); c = b.get(6); scf[ch, :, sb] = [a, a, c]
                    elif s == 3:
                        a = b.get(6); c = b.get(6); scf[ch, :, sb] = [a, c, c]
                    else:
                        a = b.get(6); scf[ch, :, sb] = [a, a, a]
        # (index 63 is reserved in ISO 11172-3, but the reference emits it for all-zero subbands: not checked)
        n_sample_bits = 0
        for sb in range(sblimit):
            for ch in range(nch if sb < jsbound else 1):
                if alloc[ch][sb]:
                    steps = table[sb][1][alloc[ch][sb] - 1]
                    bits, grouped = CLASS[steps]
                    n_sample_bits += 12 * (bits if grouped else 3 * bits)
        dab_ext = 2 if (version == 1 and kbps // nch < 56) else 4
        assert b.pos + n_sample_bits <= 8 * (lg - dab_ext - 2), "audio data overruns the DAB tail"
        frames.append(dict(pos=pos, lg=lg, version=version, kbps=kbps, fs=fs, mode=mode, mode_ext=mode_ext, nch=nch,
                           sblimit=sblimit, alloc=alloc, scfsi=scfsi, scf=scf, dab_ext=dab_ext,
                           scfcrc=stream[pos + lg - 2 - dab_ext:pos + lg - 2].copy(), audio_end_bit=b.pos + n_sample_bits))
        pos += lg
    assert len({f["lg"] for f in frames}) == 1, "frame pitch varies"
    bounds = [0, 4, 8, 16, 30]
    for n, f in enumerate(frames):  # DAB ScF-CRC (ETS 300 401 annex B): CRC-8 poly 0x1D over the 3 MSBs of the scalefactors
        src = frames[n + 1] if n + 1 < len(frames) else f
        for g in range(f["dab_ext"]):
            crc = 0
            for sb in range(bounds[g], min(bounds[g + 1], src["sblimit"])):
                for ch in range(src["nch"]):
                    if src["alloc"][ch][sb]:
                        s = src["scfsi"][ch][sb]
                        vals = {0: [0, 1, 2], 1: [0, 2], 3: [0, 2], 2: [0]}[s]
                        for k in vals:
                            crc = crc_update(crc, src["scf"][ch][k][sb] >> 3, 3, 0x1D, 8)
            assert f["scfcrc"][f["dab_ext"] - 1 - g] == crc, "ScF-CRC group %d of frame %d" % (g, n)
    return frames

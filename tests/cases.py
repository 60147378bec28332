"""Shared parity cases: (name, fs, mode, bitrate, signal, n_frames, pad_len).

Configs follow SURVEY.md section 8: A = 48k/128/j, B = 48k/192/{s,j}, C = 24k/64/m (LSF), plus mono 48k,
96 kbps stereo (allocation table 2, 2-byte ScF-CRC), dual channel, 256/j, 384/s and 24k stereo."""

CONFIGS = {
    "A": (48000, "j", 128), "Bs": (48000, "s", 192), "Bj": (48000, "j", 192), "C": (24000, "m", 64),
    "M48": (48000, "m", 96), "T2": (48000, "s", 96), "T2j": (48000, "j", 96), "D": (48000, "d", 160),
    "E1": (48000, "j", 256), "H": (48000, "s", 384), "L2": (24000, "j", 128), "L3": (24000, "s", 160),
    "M64": (48000, "m", 64), "J64": (48000, "j", 64), "L8": (24000, "m", 8), "L144": (24000, "j", 144),
    # 32 / 16 kHz: not DAB rates (src/odr-audioenc.cpp:560-563) but legal for the library (toolame.c:239-247): allocation
    # tables 1 and 3 and the psy-1 tables of freqtable.h / critband.h rows 2 and 6 (typos of the 32 kHz row included)
    "R32": (32000, "j", 128), "R32s": (32000, "s", 192), "R32m": (32000, "m", 48), "R16": (16000, "m", 32),
    "R16j": (16000, "j", 64),
}
SIGNALS = ["S1", "S2", "S3", "S4", "S5", "S6", "S7", "S8"]

# golden fixtures (tests/golden/*.npz, made by tools/make_golden.py from the compiled reference)
GOLDEN = [(c, s, 10) for c in ("A", "Bs", "Bj", "C", "M48", "T2", "T2j", "D", "E1", "L2") for s in ("S1", "S2", "S8")] + \
         [("Bj", s, 10) for s in ("S3", "S4", "S5", "S6", "S7")] + [("Bj", "PAD", 10), ("C", "PAD", 10), ("T2", "PAD", 10)] + \
         [(c, s, 10) for c in ("R32", "R32s", "R32m", "R16", "R16j") for s in ("S1", "S8")]

# psychoacoustic model 2 (BASELINE config 5: 48 kHz 256 kbit/s joint stereo) -> tests/golden/psy2_*.npz
GOLDEN_PSY2 = [("E1", "S1", 10), ("E1", "S8", 10), ("E1", "S2", 10), ("Bj", "S8", 10), ("C", "S1", 10), ("T2j", "S8", 10),
               ("M48", "S6", 10), ("E1", "S7", 10), ("R32", "S8", 10), ("R16j", "S1", 10)]

# psychoacoustic model 0 (scalefactor + absolute-threshold heuristic, reachable with --dabpsy 0) -> tests/golden/psy0_*.npz
GOLDEN_PSY0 = [("Bj", "S1", 10), ("Bj", "S8", 10), ("C", "S1", 10), ("T2j", "S8", 10), ("M48", "S6", 10), ("A", "S2", 10),
               ("R32s", "S8", 10), ("R16", "S1", 10)]

PAD_LEN = 23


def xpad_records(n_frames, pad_len=PAD_LEN, seed=99):
    """Deterministic X-PAD records in odr-audioenc's layout (src/odr-audioenc.cpp:823-852): pad_len data
    bytes + 1 byte 'used length' (0 or 2..pad_len)."""
    import numpy as np
    rng = np.random.RandomState(seed)
    rec = rng.randint(0, 256, size=(n_frames, pad_len + 1)).astype(np.uint8)
    used = rng.choice([u for u in (0, 2, 3, 8, pad_len) if u <= pad_len], size=n_frames)  # 0 or 2..pad_len
    rec[:, pad_len] = used
    return rec


def make_case(cfg_name, sig, n_frames):
    import signals
    fs, mode, br = CONFIGS[cfg_name]
    nch = 1 if mode == "m" else 2
    if sig == "PAD":
        return fs, mode, br, signals.make("S1", n_frames, nch, fs), PAD_LEN, xpad_records(n_frames)
    return fs, mode, br, signals.make(sig, n_frames, nch, fs), 0, None

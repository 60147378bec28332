"""EDI framing cases shared by tools/make_edi_golden.py (the reference's packetiser) and tests/test_framing.py."""
import numpy as np

# frames: n frames of frame_len patterned bytes (the fixtures compress to a few kB); enough frames to cross second boundaries (1 s = 41.67 frames),
# the ten-second ODRv schedule and (n > 5000) the dsti frame-counter wrap
CASES = {
    "plain": dict(tist=False, delay_ms=0, alignment=0, tai=37, start=1700000000, tag="odr-audioenc v3.6.0", n=130, frame_len=576),
    "tist": dict(tist=True, delay_ms=1234, alignment=8, tai=37, start=1700000000, tag="b200", n=1300, frame_len=384),
    "dmy": dict(tist=True, delay_ms=40, alignment=16, tai=35, start=946684800 + 86400, tag="", n=60, frame_len=288),
    "wrap": dict(tist=True, delay_ms=0, alignment=0, tai=37, start=1760000000, tag="v", n=5100, frame_len=72),
}


# (case, fec m, chunk_len (0 = 207), packets): PFT fixtures -> tests/golden/pft_*.npz
PFT_CASES = [("plain", 0, 0, 40), ("plain", 2, 0, 40), ("tist", 3, 0, 60), ("dmy", 1, 100, 30), ("wrap", 5, 0, 30)]


def inputs(case, seed=7):
    n, lg = case["n"], case["frame_len"]
    f, i = np.arange(n, dtype=np.int64)[:, None], np.arange(lg, dtype=np.int64)[None, :]
    frames = ((f * (seed + 4) + i * 13 + (f // 50) * (i // 64)) % 256).astype(np.uint8)
    peaks = np.stack([(f[:, 0] * 321 + 17) % 65536 - 32768, (f[:, 0] * 123 + 4567) % 65536 - 32768], axis=1).astype(np.int16)
    return frames, peaks

#!/usr/bin/env python3
"""How fast is the per-frame drop-in (toolame_encode_frame, one small GPU batch per call)?  usage: stream_rate.py [N]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import signals  # noqa: E402
import odr_audioenc_b200 as tl  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
pcm = signals.make("S1", n, 2, 48000)
s = tl.ToolameStream(48000, "j", 192)
for f in range(50):
    s.encode_frame(pcm[f * 1152:(f + 1) * 1152])
t0 = time.perf_counter()
for f in range(50, n):
    s.encode_frame(pcm[f * 1152:(f + 1) * 1152])
dt = time.perf_counter() - t0
s.finish()
print("streaming drop-in: %.0f frames/s = %.0f x real time for one 48 kHz stereo 192 kbit/s stream (%.0f us per call)"
      % ((n - 50) / dt, (n - 50) * 0.024 / dt, dt / (n - 50) * 1e6))

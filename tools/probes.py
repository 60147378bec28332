#!/usr/bin/env python3
"""The two denominators the bench line quotes that MEASURED_PEAKS.json does not hold (GPU box; JSON on stdout):

  fp64   tlb_fp64_peak: independent DFMA chains, and DMUL+DADD chains -- the only form this path may use (the reference
         is built without FMA contraction), hence the roofline peak of the FP64-bound kernels
  h2d    plain pinned cudaMemcpyAsync host->device (alone, and with a device->host copy running beside it): the
         ceiling of the end-to-end path at one GPU (PCIe Gen5 x16)

usage: probes.py > profiles/probes_r2.json"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import odr_audioenc_b200 as tl  # noqa: E402

L = tl.lib()
dfma, dmuladd = C.c_double(), C.c_double()
L.tlb_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
fp64 = []
for _ in range(3):
    assert L.tlb_fp64_peak(0, C.byref(dfma), C.byref(dmuladd)) == 0
    fp64.append({"dfma_tflops": dfma.value, "dmul_dadd_tflops": dmuladd.value})

n = 2 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n // 8, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n // 8, dtype=torch.uint8, device="cuda")
s2 = torch.cuda.Stream()
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    d.copy_(h, non_blocking=True)
e1.record()
torch.cuda.synchronize()
alone = 5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
e0.record()
for _ in range(5):
    d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
e1.record()
torch.cuda.synchronize()
both = 5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
p = torch.cuda.get_device_properties(0)
print(json.dumps({"gpu": p.name, "sms": p.multi_processor_count, "fp64_probe": fp64,
                  "fp64_note": "flop = one mul or add per element; 8 independent chains per thread, 256 threads, 8 blocks per SM, 4096 iterations",
                  "h2d_pinned_gbs": alone, "h2d_pinned_gbs_with_d2h_beside": both, "h2d_bytes": n}, indent=1))

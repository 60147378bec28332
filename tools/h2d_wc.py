#!/usr/bin/env python3
"""Does write-combined pinned memory raise the host-to-device rate on this box?  (GPU box)"""
import ctypes as C
import torch

rt = C.CDLL("libcudart.so.12")
n = 2 << 30
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for flags, name in ((0, "pinned"), (4, "pinned write-combined"), (1, "pinned portable")):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(flags)) == 0
    C.memset(p, 1, n)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s = torch.cuda.current_stream().cuda_stream
    for rep in range(2):
        e0.record()
        for _ in range(4):
            assert rt.cudaMemcpyAsync(C.c_void_p(d.data_ptr()), p, C.c_size_t(n), C.c_int(1), C.c_void_p(s)) == 0
        e1.record()
        torch.cuda.synchronize()
    print("%-24s H2D %.1f GB/s" % (name, 4 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9))
    rt.cudaFreeHost(p)

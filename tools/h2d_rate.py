import torch, time
n = 2 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("H2D pinned GB/s:", 5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
h2 = torch.empty(n // 8, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n // 8, dtype=torch.uint8, device="cuda")
s2 = torch.cuda.Stream()
e0.record()
for _ in range(5):
    d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("H2D with concurrent D2H GB/s:", 5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9)

"""Probe: write-only HBM bandwidth on B200 (fill kernels) vs the read+write copy the roofline peak is measured with."""
import torch, time
dev = "cuda"
n = 1 << 30  # 1 Gi elements
for dtype, name in ((torch.float32, "fill f32 4 GiB"), (torch.uint8, "fill u8 1 GiB")):
    x = torch.empty(n, dtype=dtype, device=dev)
    for _ in range(3): x.fill_(1)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); x.fill_(0); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{name}: {x.numel()*x.element_size()/best/1e6:.0f} GB/s (best of 10, {best:.3f} ms)")
    del x
a = torch.empty(n, dtype=torch.bfloat16, device=dev); b = torch.empty_like(a)
for _ in range(3): b.copy_(a)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"copy bf16 2+2 GiB: {2*a.numel()*2/best/1e6:.0f} GB/s (read+write, best of 10)")

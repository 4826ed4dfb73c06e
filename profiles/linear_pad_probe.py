"""Probe: bf16 Linear(1296 -> 13527) forward+backward with and without padding the odd output width to 13536."""
import torch, time
import torch.nn.functional as F
dev = "cuda"
x = torch.randn(16384, 1296, device=dev, requires_grad=True)
W = (torch.randn(13527, 1296, device=dev) * 0.01).requires_grad_()
b = torch.zeros(13527, device=dev, requires_grad=True)
def run(pad, bwd):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        if pad:
            y = F.linear(x, F.pad(W, (0, 0, 0, 9)), F.pad(b, (0, 9)))[:, :13527]
        else:
            y = F.linear(x, W, b)
    if bwd:
        y.float().sum().backward()
    return y
for pad in (False, True):
    for bwd in (False, True):
        for _ in range(3): run(pad, bwd)
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(10): run(pad, bwd)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
        fl = 2 * 16384 * 1296 * 13527 * (3 if bwd else 1)
        print(f"pad={pad} bwd={bwd}: {dt*1e3:.2f} ms  {fl/dt/1e12:.0f} TFLOP/s")

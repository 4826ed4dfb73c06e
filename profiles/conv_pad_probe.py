"""Probe: first conv (46 -> 16, 3x3, 9x9 board) forward+backward under bf16 autocast: NCHW vs channels_last vs padded channels."""
import torch, time
import torch.nn.functional as F
dev = "cuda"
B = 16384
x = torch.rand(B, 46, 9, 9, device=dev)
W = (torch.randn(16, 46, 3, 3, device=dev) * 0.05).requires_grad_()
b = torch.zeros(16, device=dev, requires_grad=True)
def run(mode):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        if mode == "nchw":
            y = F.conv2d(x, W, b, padding=1)
        elif mode == "cl":
            y = F.conv2d(x.contiguous(memory_format=torch.channels_last), W, b, padding=1)
        elif mode == "cl_pad48":
            xp = F.pad(x, (0, 0, 0, 0, 0, 2)).contiguous(memory_format=torch.channels_last)
            y = F.conv2d(xp, F.pad(W, (0, 0, 0, 0, 0, 2)), b, padding=1)
        elif mode == "pad48":
            y = F.conv2d(F.pad(x, (0, 0, 0, 0, 0, 2)), F.pad(W, (0, 0, 0, 0, 0, 2)), b, padding=1)
        elif mode == "unfold_mm":
            cols = F.unfold(x.to(torch.bfloat16), 3, padding=1)            # [B, 414, 81]
            y = torch.matmul(W.view(16, 414).to(torch.bfloat16), cols).view(B, 16, 9, 9) + b.view(1, 16, 1, 1)
    y.float().square().sum().backward()
for bench in (False, True):
    torch.backends.cudnn.benchmark = bench
    for mode in ("nchw", "cl", "pad48", "cl_pad48", "unfold_mm"):
        for _ in range(3): run(mode)
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(10): run(mode)
        torch.cuda.synchronize()
        print(f"cudnn.benchmark={bench} {mode:10s}: {(time.perf_counter()-t)*100:.2f} ms fwd+bwd")

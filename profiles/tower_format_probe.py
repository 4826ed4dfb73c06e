"""Probe: the ResNet tower (ActorCriticResTower 9 x 256, SE 0.25; stays on PyTorch / cuDNN) under bf16 autocast in the default
NCHW layout against channels_last, forward at a rollout batch and forward + backward at a minibatch."""
import sys, torch
sys.path.insert(0, ".")
from shogidrl_b200.core import ActorCriticResTower
dev = torch.device("cuda")
torch.manual_seed(0)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for fmt in (torch.contiguous_format, torch.channels_last):
    m = ActorCriticResTower(46, 13527, 9, 256, 0.25).to(dev).to(memory_format=fmt)
    xb = torch.rand(8192, 46, 9, 9, device=dev)
    xs = torch.rand(4096, 46, 9, 9, device=dev)

    def fwd():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            m(xb.contiguous(memory_format=fmt))

    def fb():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            lo, v = m(xs.contiguous(memory_format=fmt))
        (lo.float().mean() + v.float().mean()).backward()

    f, b = timed(fwd), timed(fb)
    print(f"{str(fmt):28s} forward 8192: {f:7.2f} ms ({8192 * 1.74e9 / f / 1e9:6.0f} TFLOP/s)   forward+backward 4096: {b:7.2f} ms "
          f"({4096 * 3 * 1.74e9 / b / 1e9:6.0f} TFLOP/s)")

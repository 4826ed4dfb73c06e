"""Probe: does the compute part of kz_step (no mask/obs outputs) overlap with an independent 1.86 GB/step write stream?"""
import torch, sys
sys.path.insert(0, ".")
from shogidrl_b200 import VecShogiEnv
dev = torch.device("cuda:0")
n = 65536
env = VecShogiEnv(n, 500, dev, seed=1234)
act = [torch.zeros(n, dtype=torch.int64, device=dev) for _ in range(2)]
env.refresh(random_actions=True, next_out=act[0])
i = 0
def step(wo, wm):
    global i
    env.step(act[i & 1], random_actions=True, next_out=act[(i + 1) & 1], write_obs=wo, write_mask=wm)
    i += 1
for _ in range(640): step(True, True)
fillbuf = torch.empty(n * (14904 + 13536) // 4, dtype=torch.float32, device=dev)  # 1.86 GB
s2 = torch.cuda.Stream()
def timed(fn, reps=64):
    for _ in range(4): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("full step           : %.4f ms" % timed(lambda: step(True, True)))
print("compute only        : %.4f ms" % timed(lambda: step(False, False)))
print("fill 1.86 GB only   : %.4f ms" % timed(lambda: fillbuf.zero_()))
def both():
    with torch.cuda.stream(s2):
        fillbuf.zero_()
    step(False, False)
print("compute || fill     : %.4f ms per pair (two streams)" % timed(both))

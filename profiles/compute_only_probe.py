"""Probe: kz_step with and without the mask / observation outputs (same 65,536-game steady-state batch)."""
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
for name, wo, wm in (("obs+mask", True, True), ("mask only", False, True), ("obs only", True, False), ("neither", False, False), ("obs+mask", True, True)):
    for _ in range(8): step(wo, wm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(64): step(wo, wm)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:10s}: {e0.elapsed_time(e1)/64:.4f} ms per step")

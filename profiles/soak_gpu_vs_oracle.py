"""Soak parity run (GPU box): device self-play (fused uniform-random legal action, auto-reset) against the C oracle playing
the same counter-based RNG, at a size the unit test does not reach -- every step's actions, legal counts, rewards, dones and
reasons, and the final boards / hands / masks / observations.  The oracle is the checker only.

    python profiles/soak_gpu_vs_oracle.py [n_envs] [T] [max_moves] [seed]
    KZ_SOAK_START=positions.npz (boards, hands, sides, mcs as tests.helpers.random_endgames returns them): start from those"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402  (checker only)
from shogidrl_b200 import VecShogiEnv  # noqa: E402

n, T, max_moves, seed = (int(a) for a in (sys.argv[1:5] + ["16384", "700", "200", "97531"][len(sys.argv) - 1:]))
dev = torch.device("cuda:0")
env = VecShogiEnv(n, max_moves_per_game=max_moves, device=dev, seed=seed, auto_reset=True)
start = None
if os.environ.get("KZ_SOAK_START"):
    z = np.load(os.environ["KZ_SOAK_START"])
    start = (z["boards"][:n], z["hands"][:n], z["sides"][:n], z["mcs"][:n])
    env.load_positions(*start, eval_termination=False)
    env.step_index = 0
env.refresh(random_actions=True)
acts = torch.empty((T, n), dtype=env.next_actions.dtype, device=dev)
counts = torch.empty((T, n), dtype=env.legal_count.dtype, device=dev)
rews = torch.empty((T, n), dtype=torch.float32, device=dev)
dones = torch.empty((T, n), dtype=torch.uint8, device=dev)
reasons = torch.empty((T, n), dtype=torch.uint8, device=dev)
t0 = time.time()
for t in range(T):
    a = env.next_actions.clone()
    acts[t] = a
    counts[t] = env.legal_count
    out = env.step(a, random_actions=True)
    rews[t], dones[t], reasons[t] = out["reward"], out["done"], out["reason"]
torch.cuda.synchronize()
t1 = time.time()
assert int(env.errors().abs().sum()) == 0
ref = orc.selfplay(n, T, max_moves=max_moves, seed=seed, threads=os.cpu_count() or 1, start=start)
t2 = time.time()
ok = {
    "actions": np.array_equal(acts.cpu().numpy(), ref["actions"]),
    "legal_counts": np.array_equal(counts.cpu().numpy(), ref["legal_counts"]),
    "rewards": np.array_equal(rews.cpu().numpy(), ref["rewards"]),
    "dones": np.array_equal(dones.cpu().numpy(), ref["dones"]),
    "reasons": np.array_equal(reasons.cpu().numpy(), ref["reasons"]),
}
b, h, m = [x.cpu().numpy() for x in env.export()]
ok["boards"] = np.array_equal(b, ref["boards"]) and np.array_equal(h, ref["hands"]) and np.array_equal(m[:, :2], ref["meta"][:, :2])
ok["mask"] = np.array_equal(env.mask.cpu().numpy(), ref["mask"])
ok["obs"] = np.array_equal(env.obs.cpu().numpy(), ref["obs"])
hist = {int(k): int(v) for k, v in zip(*np.unique(ref["reasons"][ref["dones"] > 0], return_counts=True))}
print((f"start positions from {os.environ['KZ_SOAK_START']}, " if start is not None else "") + f"{n} games x {T} steps = {n * T} env steps, max_moves {max_moves}, seed {seed}: device loop {t1 - t0:.1f} s, oracle "
      f"{t2 - t1:.1f} s on {os.cpu_count()} threads; finished episodes by reason {hist}; equal: {ok}")
sys.exit(0 if all(ok.values()) else 1)

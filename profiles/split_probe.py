"""Probe: the fused kz_step against the split pipeline (kz_step_compact + kz_expand) on 65,536 games.
(a) fused, one batch; (b) split, one batch, one stream (the two halves run back to back); (c) split, two groups of 32,768
on two streams (one group's row stores overlap the other's move generation); (d) fused, the same two groups / streams."""
import sys
import torch
sys.path.insert(0, ".")
from shogidrl_b200 import VecShogiEnv

dev = torch.device("cuda:0")
N, PRE, K = 65536, 200, 48


def make(n, off):
    env = VecShogiEnv(n, 500, dev, seed=1234, env_offset=off)
    act = [torch.zeros(n, dtype=torch.int64, device=dev) for _ in range(2)]
    env.refresh(random_actions=True, next_out=act[0])
    env.bitmap = torch.zeros((n, 448), dtype=torch.int32, device=dev)
    return env, act, [0]


def fused(e):
    env, act, i = e
    env.step(act[i[0] & 1], random_actions=True, next_out=act[(i[0] + 1) & 1]); i[0] += 1


def split(e):
    env, act, i = e
    env.step_compact(act[i[0] & 1], random_actions=True, next_out=act[(i[0] + 1) & 1]); i[0] += 1
    env.expand()


def timed(label, fn, envs, streams):
    for _ in range(8):
        for e, s in zip(envs, streams):
            with torch.cuda.stream(s):
                fn(e)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams:
        s.wait_stream(torch.cuda.current_stream())
    for _ in range(K):
        for e, s in zip(envs, streams):
            with torch.cuda.stream(s):
                fn(e)
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    print(f"{label:58s}: {e0.elapsed_time(e1) / K:.4f} ms per {N} env steps")


import os
if os.environ.get("QUICK"):
    # halves alone, then the overlapped pair: one short call
    PRE, K = 64, 24
    one = make(N, 0)
    for _ in range(PRE):
        fused(one)

    def compact_only(e):
        env, act, i = e
        env.step_compact(act[i[0] & 1], random_actions=True, next_out=act[(i[0] + 1) & 1]); i[0] += 1

    def expand_only(e):
        e[0].expand()

    cur = [torch.cuda.current_stream()]
    timed("compact alone, one batch", compact_only, [one], cur)
    timed("expand alone, one batch", expand_only, [one], cur)
    del one
    two = [make(N // 2, 0), make(N // 2, N // 2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for _ in range(PRE):
        for e in two:
            fused(e)
    torch.cuda.synchronize()
    timed("(c) split, two groups on two streams", split, two, streams)
    sys.exit(0)

one = make(N, 0)
for _ in range(PRE):
    fused(one)
cur = [torch.cuda.current_stream()]
timed("(a) fused kz_step, one batch", fused, [one], cur)
timed("(b) split, one batch, one stream", split, [one], cur)
del one
two = [make(N // 2, 0), make(N // 2, N // 2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
for _ in range(PRE):
    for e in two:
        fused(e)
torch.cuda.synchronize()
timed("(c) split, two groups on two streams", split, two, streams)
timed("(d) fused, two groups on two streams", fused, two, streams)
timed("(c) again", split, two, streams)

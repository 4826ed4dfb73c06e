#!/usr/bin/env python
"""bench.py -- legal-masked Shogi env steps/sec on B200 (BASELINE.json metric, config 2).

    python bench.py --gpus N --steps K --warmup W            # product arm (CUDA kernels, C ABI)
    python bench.py --impl reference --steps K --warmup W    # reference arm: CPU oracle port, all host threads

A "step" is one pass of the hot path over one batch: kz_step over 65,536 device-resident games per GPU
(apply the action, generate the successor's legal moves, termination, auto-reset, write the 13,527-byte
legal mask + 46x9x9 fp32 observation + reward/done/reason/winner), actions = uniform-random legal via the
counter-based RNG keyed (seed, env, step).  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "legal-masked Shogi env steps/sec"
UNIT = "env_steps/s"
ENVS_PER_GPU = 65536
MAX_MOVES = 500
SEED = 1234
PREROLL = 640  # untimed plies that spread the games over all phases (synthetic input preparation)

# algorithmic bytes per env step
OBS_B, MASK_B, SCALARS_B, ACTION_B, STATE_B, HIST_APPEND_B = 14904, 13527, 7, 8, 238, 16


def algorithmic_bytes_per_step(mean_ply: float) -> float:
    """28,438 written + 8 action + 238 state + 16 history append + 8*ply history scan."""
    return OBS_B + MASK_B + SCALARS_B + ACTION_B + STATE_B + HIST_APPEND_B + 8.0 * mean_ply


def implementation_bytes_per_step() -> float:
    """What this implementation must move per env step: the same outputs and action, 96+32 B of state read and
    written, and one 512 B probe + 16 B update of the repetition table (no per-ply scan)."""
    return OBS_B + MASK_B + SCALARS_B + ACTION_B + 2 * 128 + 512 + 16


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline_sample(threads: int, seconds_budget: float = 20.0):
    """The oracle port timed on the host cores on a bounded sample of the same workload."""
    from oracle import oracle as orc
    n = 256 * threads
    batch = orc.OracleBatch(n, MAX_MOVES, SEED, threads)
    batch.run(8)  # warm caches / threads
    t0 = time.perf_counter()
    steps, T = 0, 0
    while True:
        steps += batch.run(40)
        T += 40
        dt = time.perf_counter() - t0
        if dt > seconds_budget or T >= 640:
            break
    return {"value": steps / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} games x {T} plies from the start position with auto-reset (max_moves {MAX_MOVES}), "
                      f"legal moves + mask + make_move + observation per ply, {dt:.1f} s, C oracle (oracle/keisei_oracle.c)"}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference itself is pure Python and
    cannot travel to the GPU box) on all host threads.  One step = one env step of every game of the batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as orc
    threads = os.cpu_count() or 1
    n = 64 * threads
    batch = orc.OracleBatch(n, MAX_MOVES, SEED, threads)
    batch.run(256)  # untimed pre-roll: spread the games over the phases of play
    for _ in range(args.warmup):
        batch.run(1)
    t0 = time.perf_counter()
    plies = 0.0
    for _ in range(args.steps):
        batch.run(1)
        plies += batch.mean_ply
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"random-legal self-play, {n} games per step on {threads} host threads (bounded sample of "
                               f"BASELINE config 2: 65,536 games/GPU), max_moves {MAX_MOVES}, auto-reset, 256-ply pre-roll",
                   "mean_ply": plies / max(1, args.steps)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n} games x {args.steps} timed plies, C oracle port of keisei.shogi (pure-Python reference "
                                   "cannot run on the GPU box)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def run_reference_ppo(args):
    """--impl reference --workload ppo: the reference's sequential CPU trainer restated -- one game, batch-1
    select_action (masked softmax + Categorical sample, ppo_agent.py:134-223), ExperienceBuffer GAE
    (experience_buffer.py:99-145), then PPOAgent.learn (ppo_agent.py:243-460: ppo_epochs x minibatches of 64, clipped
    surrogate + value MSE + entropy, clip_grad_norm_ 0.5, Adam) -- on the oracle engine and the default CNN
    (neural_network.py:10-29) in plain fp32 PyTorch on all host threads.  One step = one epoch of `--ref-ppo-steps`
    timesteps (a bounded sample of the reference's steps_per_epoch = 2048); value = timesteps per second."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    import torch
    import torch.nn as nn
    from oracle import oracle as orc
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(SEED)
    S, MB, EPOCHS = args.ref_ppo_steps, 64, args.ppo_epochs
    gamma, lam, clip_eps, cv, ce = 0.99, 0.95, 0.2, 0.5, 0.01

    class Net(nn.Module):  # keisei/core/neural_network.py:10-29
        def __init__(self):
            super().__init__()
            self.conv = nn.Conv2d(46, 16, 3, padding=1)
            self.policy_head = nn.Linear(16 * 81, 13527)
            self.value_head = nn.Linear(16 * 81, 1)

        def forward(self, x):
            h = torch.relu(self.conv(x)).flatten(1)
            return self.policy_head(h), self.value_head(h).squeeze(-1)

    def dist_of(logits, mask):
        probs = torch.softmax(torch.where(mask, logits, torch.full((), float("-inf"))), dim=-1)
        return torch.distributions.Categorical(probs=probs)

    model = Net()
    opt = torch.optim.Adam(model.parameters(), lr=3e-4)
    game = orc.OracleGame(MAX_MOVES)
    shuffle_rng = np.random.default_rng(SEED)
    steps = max(1, min(args.steps, 8))
    warm = min(args.warmup, 1)

    def epoch():
        obs = np.zeros((S, 46, 9, 9), np.float32); masks = np.zeros((S, 13527), np.uint8)
        act = np.zeros(S, np.int64); lp = np.zeros(S, np.float32); val = np.zeros(S, np.float32)
        rew = np.zeros(S, np.float32); done = np.zeros(S, np.uint8)
        t_roll = time.perf_counter()
        for t in range(S):
            obs[t] = game.observation(); masks[t] = game.legal_mask()
            with torch.no_grad():
                logits, v = model(torch.from_numpy(obs[t:t + 1]))
                d = dist_of(logits, torch.from_numpy(masks[t:t + 1]).bool())
                a = d.sample()
                lp[t] = float(d.log_prob(a)); val[t] = float(v); act[t] = int(a)
            r, dn, _, _ = game.make_move(int(a))
            rew[t], done[t] = r, dn
            if dn:
                game.reset()
        with torch.no_grad():
            last_v = float(model(torch.from_numpy(game.observation()[None]))[1])
        adv, ret = orc.gae(rew[:, None], val[:, None], done[:, None], np.float32(last_v), gamma, lam)
        t_roll = time.perf_counter() - t_roll
        to = torch.from_numpy
        obs_t, mask_t, act_t = to(obs), to(masks).bool(), to(act)
        lp_t, adv_t, ret_t = to(lp), to(adv[:, 0].copy()), to(ret[:, 0].copy())
        adv_t = (adv_t - adv_t.mean()) / (adv_t.std() + 1e-8)
        for _ in range(EPOCHS):
            idx = shuffle_rng.permutation(S)  # ppo_agent.py:298
            for s0 in range(0, S, MB):
                mb = torch.from_numpy(idx[s0:s0 + MB])
                logits, v = model(obs_t[mb])
                d = dist_of(logits, mask_t[mb])
                new_lp, ent = d.log_prob(act_t[mb]), d.entropy()
                ratio = torch.exp(new_lp - lp_t[mb])
                pol = -torch.min(ratio * adv_t[mb], torch.clamp(ratio, 1 - clip_eps, 1 + clip_eps) * adv_t[mb]).mean()
                loss = pol + cv * torch.nn.functional.mse_loss(v, ret_t[mb]) - ce * ent.mean()
                opt.zero_grad()
                loss.backward()
                torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5)
                opt.step()
        return t_roll

    for _ in range(warm):
        epoch()
    t0 = time.perf_counter()
    roll = sum(epoch() for _ in range(steps))
    dt = time.perf_counter() - t0
    value = S * steps / dt
    sample = (f"{steps} epochs x {S} timesteps of one game (reference default: 2048), minibatch {MB}, ppo_epochs {EPOCHS}, "
              f"C oracle engine + fp32 PyTorch CPU model on {threads} threads (the pure-Python reference cannot run on the GPU box)")
    emit({"impl": "reference", "metric": "PPO self-play samples/sec", "value": value, "unit": "samples/s", "n_gpus": args.gpus,
          "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": {"workload": "BASELINE config 3 on the host: PPO self-play, cnn policy-value net, sequential trainer; " + sample,
                     "rollout_samples_per_s": S * steps / roll, "update_samples_per_s": S * steps / max(1e-9, dt - roll)},
          "cpu_baseline": {"value": value, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
          "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
    return 0


def run_product(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from shogidrl_b200 import VecShogiEnv, MASK_PAD_STRIDE

    n = args.envs
    env = VecShogiEnv(n, MAX_MOVES, dev, seed=SEED, env_offset=rank * n, auto_reset=True, step_streams=args.step_streams)
    # rollout storage the kernel writes straight into: 2 slots of [n] obs / masks (1.86 GB per slot >> 126 MB L2)
    SLOTS = 2
    obs_buf = torch.zeros((SLOTS, n, 46, 9, 9), dtype=torch.float32, device=dev)
    mask_buf = torch.zeros((SLOTS, n, MASK_PAD_STRIDE), dtype=torch.uint8, device=dev)
    act = [torch.zeros(n, dtype=torch.int64, device=dev) for _ in range(2)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    it = [0]

    def step():
        # actions were chosen by the fused uniform-random legal policy of the previous launch (ping-pong buffers)
        i = it[0]
        it[0] += 1
        env.step(act[i & 1], obs=obs_buf[i % SLOTS], mask=mask_buf[i % SLOTS][:, :13527], random_actions=True,
                 next_out=act[(i + 1) & 1])

    env.refresh(random_actions=True, next_out=act[0])
    for i in range(args.preroll):
        step()
    for i in range(args.warmup):
        step()
    barrier()

    # mean ply of the positions the timed steps start from (for the algorithmic-bytes figure)
    ply0 = float(env.export()[2][:, 1].float().mean())
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev0.record()
    for i in range(args.steps):
        kev[i][0].record()
        step()
        kev[i][1].record()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    ply1 = float(env.export()[2][:, 1].float().mean())
    errs = int((env.errors() != 0).sum())

    # ---- end-to-end through the public API with HOST buffers: per step the actions come from pinned host
    # memory (H2D) and the step's results + the next legal actions are read back to pinned host memory (D2H).
    # (a) serial: one batch, copy -> kernel -> copy -> synchronise, latencies add up;
    # (b) HostPipelinedEnv: the same games in two groups on two streams, the host serves one group while the other
    #     group's kernel runs.  (b) is the e2e figure; (a) is reported next to it.
    h_act = torch.zeros(n, dtype=torch.int64).pin_memory()
    h_res = torch.zeros(n * 15, dtype=torch.uint8).pin_memory()
    h_act.copy_(act[it[0] & 1])
    torch.cuda.synchronize(dev)
    e2e_steps = max(4, min(args.steps, 64))

    def e2e_step(i):
        a = act[i & 1]
        a.copy_(h_act, non_blocking=True)                                  # H2D 8 B/env
        env.step(a, obs=obs_buf[i % SLOTS], mask=mask_buf[i % SLOTS][:, :13527], random_actions=True,
                 next_out=env.next_actions)
        h_res.copy_(env.results, non_blocking=True)                        # D2H 15 B/env: next action, reward, done, reason, winner
        torch.cuda.current_stream(dev).synchronize()
        h_act.copy_(h_res[: 8 * n].view(torch.int64))                      # host-side hand-over of the chosen actions

    for i in range(3):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    barrier()
    serial_ms = 1e3 * (time.perf_counter() - t0)

    # (c) every output on the host: as (a), plus the step's 46x9x9 observations and 13,527-byte legal masks copied to
    #     pinned host memory -- what a caller that keeps the policy tower off the GPU would pay (PCIe-bound).
    h_obs = torch.empty((n, 46, 9, 9), dtype=torch.float32).pin_memory()
    h_mask = torch.empty((n, MASK_PAD_STRIDE), dtype=torch.uint8).pin_memory()
    full_steps = 4

    def full_step(i):
        a = act[i & 1]
        a.copy_(h_act, non_blocking=True)
        env.step(a, obs=obs_buf[i % SLOTS], mask=mask_buf[i % SLOTS][:, :13527], random_actions=True,
                 next_out=env.next_actions)
        h_res.copy_(env.results, non_blocking=True)
        h_obs.copy_(obs_buf[i % SLOTS], non_blocking=True)
        h_mask.copy_(mask_buf[i % SLOTS], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        h_act.copy_(h_res[: 8 * n].view(torch.int64))

    full_step(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(full_steps):
        full_step(i)
    barrier()
    full_ms = 1e3 * (time.perf_counter() - t0)
    full_d2h = 15 * n + h_obs.numel() * 4 + h_mask.numel()
    del h_obs, h_mask

    from shogidrl_b200.host_env import HostPipelinedEnv
    G = 2
    del env
    pipe = HostPipelinedEnv(n, groups=G, max_moves_per_game=MAX_MOVES, device=dev, seed=SEED, env_offset=rank * n)
    ng = n // G

    def views(i, g):
        return obs_buf[i % SLOTS][g * ng:(g + 1) * ng], mask_buf[i % SLOTS][g * ng:(g + 1) * ng, :13527]

    pipe.prime(random_actions=True)
    for i in range(min(args.preroll, 128) + 3):                            # mid-game positions, pipeline warm
        for g in range(G):
            pipe.h_actions[g].copy_(pipe.wait(g)[0])
            pipe.submit(g, *views(i, g), random_actions=True)
    pipe.synchronize()
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        for g in range(G):
            pipe.h_actions[g].copy_(pipe.wait(g)[0])                       # host hand-over of the chosen actions
            pipe.submit(g, *views(i, g), random_actions=True)              # H2D 8 B/env, kernel, D2H 15 B/env
    for g in range(G):
        pipe.wait(g)
    pipe.synchronize()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    errs += sum(int((e.errors() != 0).sum()) for e in pipe.envs)

    if world > 1:
        t = torch.tensor([ms_total, e2e_ms, kernel_ms, serial_ms, full_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms, kernel_ms, serial_ms, full_ms = [float(x) for x in t]
        e = torch.tensor([errs], device=dev)
        dist.all_reduce(e)
        errs = int(e)

    if rank == 0:
        mean_ply = 0.5 * (ply0 + ply1)
        bytes_per_launch = algorithmic_bytes_per_step(mean_ply) * n
        achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
        peak, peak_src = measured_peak_gbs()
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
                if int(tj.get("envs", 0)) == n:
                    traffic = tj["dram_bytes_per_launch"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": world * n * args.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"BASELINE config 2: batched random-legal self-play, {n} games per GPU, legal mask + obs + "
                                   f"step, no network; max_moves {MAX_MOVES}, auto-reset, actions = fused uniform-random legal "
                                   f"(counter RNG seed {SEED})",
                       "envs_per_gpu": n, "parallelism": f"env-shard x{world} (no data-path collective)",
                       "preroll_steps": args.preroll, "mean_ply": mean_ply,
                       "l2": "each step writes 1.86 GB of fresh obs+mask rows per GPU (2-slot ring), far above the 126 MB L2",
                       "algorithmic_bytes_per_env_step": algorithmic_bytes_per_step(mean_ply), "env_error_flags": errs},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "kz_step_kernel", "kernel_ms": kernel_ms, "peak_source": peak_src,
                         "achieved_implementation_bytes": implementation_bytes_per_step() * n / (kernel_ms * 1e-3) / 1e9,
                         "frac_implementation_bytes": implementation_bytes_per_step() * n / (kernel_ms * 1e-3) / 1e9 / peak},
            "e2e": {"value": world * n * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * n,
                    "d2h_bytes_per_step": 15 * n, "steps": e2e_steps,
                    "serial_value": world * n * e2e_steps / (serial_ms * 1e-3),
                    "all_outputs_to_host_value": world * n * full_steps / (full_ms * 1e-3),
                    "all_outputs_d2h_bytes_per_step": full_d2h,
                    "note": "HostPipelinedEnv (2 groups, 2 streams): actions from pinned host memory, reward/done/reason/winner/"
                            "next-action read back to pinned host memory every step, host waits for each group's results before "
                            "submitting its next actions; obs/mask stay in HBM for the policy tower; serial_value = the same "
                            "through one VecShogiEnv with copy -> kernel -> copy -> synchronise in sequence; all_outputs_to_host_value = the "
                            "serial path when the observations and legal masks are ALSO copied to pinned host memory every "
                            "step (1.86 GB per 65,536 games: PCIe-bound; only a caller with its policy network off the GPU "
                            "needs that)"},
            "gpu_launches": args.steps,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample(os.cpu_count() or 1)
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_JSON_OUT = None


def isolate_stdout():
    """Libraries write to fd 1 (NCCL prints its version line there): keep the real stdout for the one JSON line
    and send everything else to stderr."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def run_ppo(args):
    """--workload ppo: BASELINE config 3 -- PPO self-play with the default CNN policy-value net (bf16 autocast),
    16,384 envs, rollout T = 128, fused masked sampling, device GAE, then the PPO update.  One step = one full
    epoch (collect + update); value = samples (env steps) per second end to end; rollout-only rate in config."""
    import torch
    import torch.distributed as dist
    from types import SimpleNamespace
    from shogidrl_b200.core import ActorCritic, ActorCriticResTower
    from shogidrl_b200.training.selfplay import SelfPlayTrainer

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N, T = args.ppo_envs, args.ppo_horizon
    mb = args.ppo_minibatch
    cfg = SimpleNamespace(
        env=SimpleNamespace(device=str(dev), seed=SEED, input_channels=46, num_actions_total=13527, max_moves_per_game=MAX_MOVES),
        training=SimpleNamespace(learning_rate=3e-4, gamma=0.99, lambda_gae=0.95, clip_epsilon=0.2, value_loss_coeff=0.5,
                                 entropy_coef=0.01, ppo_epochs=args.ppo_epochs, minibatch_size=mb, steps_per_epoch=N * T,
                                 total_timesteps=N * T * 8, gradient_clip_max_norm=0.5, normalize_advantages=True,
                                 enable_value_clipping=False, weight_decay=0.0, lr_schedule_type=None,
                                 lr_schedule_step_on="epoch", lr_schedule_kwargs=None),
        display=SimpleNamespace(display_moves=False, turn_tick=0.0))
    torch.manual_seed(SEED + rank)
    if args.ppo_model == "resnet":
        model = ActorCriticResTower(46, 13527, tower_depth=9, tower_width=256, se_ratio=0.25)
    else:
        model = ActorCritic(46, 13527)
    tr = SelfPlayTrainer(model, cfg, N, T, dev, use_mixed_precision=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(1, args.warmup if args.warmup < 3 else 1)):
        tr.run_epoch()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    roll_ms = upd_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ev[0].record(); tr.collect(); ev[1].record(); m = tr.update(); ev[2].record()
        torch.cuda.synchronize(dev)
        roll_ms += ev[0].elapsed_time(ev[1]); upd_ms += ev[1].elapsed_time(ev[2])
    barrier()
    total_ms = 1e3 * (time.perf_counter() - t0)
    if world > 1:
        t = torch.tensor([total_ms, roll_ms, upd_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, roll_ms, upd_ms = [float(x) for x in t]
    if rank == 0:
        samples = world * N * T * args.steps
        line = {"metric": "PPO self-play samples/sec", "value": samples / (total_ms * 1e-3), "unit": "samples/s",
                "n_gpus": world, "steps": args.steps, "warmup": 1, "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"BASELINE config {3 if args.ppo_model == 'cnn' else 4}: PPO self-play, {args.ppo_model} policy-value net (bf16 autocast), {N} envs/GPU, "
                                       f"T={T}, fused masked sampling + device GAE, ppo_epochs={args.ppo_epochs}, minibatch={mb}",
                           "rollout_samples_per_s": samples / (roll_ms * 1e-3), "update_samples_per_s": samples / (upd_ms * 1e-3),
                           "rollout_ms": roll_ms / args.steps, "update_ms": upd_ms / args.steps,
                           "last_metrics": {k: float(v) for k, v in m.items()}},
                # own kernels per epoch: rollout steps (kz_step, kz_sample_masked, + kz_obs_conv_fwd for the CNN), kz_gae,
                # and per minibatch update kz_eval_masked_fwd/bwd (+ kz_obs_conv_fwd and the two wgrad kernels)
                "gpu_launches": args.steps * (T * (3 if args.ppo_model == "cnn" else 2) + 1
                                              + -(-N * T // mb) * args.ppo_epochs * (5 if args.ppo_model == "cnn" else 2))}
        emit(line)
    if world > 1:
        # the captured update graph holds NCCL work: release it before the process group goes away, and never let
        # a stuck communicator teardown outlive the measurement
        import gc
        import threading
        barrier()
        tr.agent._graph = None
        gc.collect()
        torch.cuda.synchronize(dev)
        killer = threading.Timer(30.0, lambda: os._exit(0))
        killer.daemon = True
        killer.start()
        dist.destroy_process_group()
        killer.cancel()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="games per GPU (BASELINE config 2: 65,536)")
    ap.add_argument("--preroll", type=int, default=PREROLL)
    ap.add_argument("--step-streams", type=int, default=None,
                    help="ranges of games a step is launched as, on as many streams (default: VecShogiEnv's choice, 2 for >= 32,768 games)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="env", choices=["env", "ppo"], help="env: BASELINE config 2 (default, the headline); ppo: config 3")
    ap.add_argument("--ppo-envs", type=int, default=16384)
    ap.add_argument("--ppo-horizon", type=int, default=128)
    ap.add_argument("--ppo-epochs", type=int, default=10)
    ap.add_argument("--ppo-minibatch", type=int, default=16384)
    ap.add_argument("--ppo-model", default="cnn", choices=["cnn", "resnet"])
    ap.add_argument("--ref-ppo-steps", type=int, default=512, help="timesteps per epoch of the CPU PPO reference arm")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference_ppo(args) if args.workload == "ppo" else run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517"), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    isolate_stdout()
    if args.workload == "ppo":
        return run_ppo(args)
    return run_product(args)


if __name__ == "__main__":
    sys.exit(main())

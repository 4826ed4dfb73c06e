"""N > 1 host logic on CPU: two processes over the gloo backend (rendezvous on 127.0.0.1).  Covers env sharding,
the whole-buffer advantage-normalisation exchange, max-over-ranks timing and the DDP gradient all-reduce of the
PPO update path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from shogidrl_b200.training import distributed as kd
        from shogidrl_b200.core.base_actor_critic import ActorCritic, BaseActorCriticModel
        from shogidrl_b200.core.ppo_agent import PPOAgent
        from tests.helpers import make_config
        assert kd.world() == (rank, world)
        off, n = kd.shard_envs(10, rank, world)
        # whole-buffer statistics over both shards
        full = torch.arange(10, dtype=torch.float32) ** 1.5
        mine = full[off:off + n]
        c, s1, s2 = kd.global_moments(mine)
        assert c == 10 and abs(s1 - float(full.double().sum())) < 1e-6 and abs(s2 - float((full.double() ** 2).sum())) < 1e-6
        cfg = make_config(device="cpu")
        torch.manual_seed(0)
        agent = PPOAgent(ActorCritic(46, 13527), cfg, torch.device("cpu"))
        normed = agent._normalize(mine)
        ref = (full - full.mean()) / full.std()
        assert torch.allclose(normed, ref[off:off + n], atol=1e-5)
        assert kd.reduce_max(1.0 + rank, "cpu") == float(world) and kd.reduce_sum(1.0, "cpu") == float(world)
        # DDP: different data per rank, identical parameters after one update
        agent.enable_ddp()
        g = torch.Generator().manual_seed(100 + rank)
        obs = torch.randn(8, 46, 9, 9, generator=g)
        acts = torch.randint(0, 13527, (8,), generator=g)
        logits, values = agent._train_forward(obs)
        lp, ent, v = BaseActorCriticModel.evaluate_from_logits(logits, values, acts, None)
        loss = -(lp.mean()) + 0.5 * (v ** 2).mean() - 0.01 * ent.mean()
        agent.optimizer.zero_grad()
        loss.backward()
        agent.optimizer.step()
        w = agent.model.value_head.weight.detach().clone()
        gathered = [torch.zeros_like(w) for _ in range(world)]
        dist.all_gather(gathered, w)
        assert torch.equal(gathered[0], gathered[1])
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res

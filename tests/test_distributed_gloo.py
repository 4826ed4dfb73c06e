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
        # data-parallel update: different data per rank, identical parameters afterwards
        agent.enable_ddp()
        g = torch.Generator().manual_seed(100 + rank)
        # the fused-minibatch model takes the explicit all-reduce path (graph-capturable on GPUs) through learn()
        assert getattr(agent, "_grad_world", 1) == world and getattr(agent, "_ddp", None) is None
        B = 32
        batch = {"obs": torch.randn(B, 46, 9, 9, generator=g), "actions": torch.randint(0, 13527, (B,), generator=g),
                 "log_probs": torch.full((B,), -9.0), "values": torch.zeros(B), "advantages": torch.randn(B, generator=g),
                 "returns": torch.randn(B, generator=g), "legal_masks": torch.ones(B, 13527, dtype=torch.bool)}

        class Buf:
            def get_batch(self):
                return batch
        metrics = agent.learn(Buf())
        assert all(np.isfinite(v) for v in metrics.values())
        for prm in agent.model.parameters():
            gathered = [torch.zeros_like(prm) for _ in range(world)]
            dist.all_gather(gathered, prm.detach().clone())
            assert torch.equal(gathered[0], gathered[1])
        # GradReducer: one gradient reduced early (asynchronously, from inside backward on GPUs), the rest flattened
        red = kd.GradReducer()
        ps = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2, 2))]
        for i, prm in enumerate(ps):
            prm.grad = torch.full_like(prm, float(10 * rank + i))
        g0, ps[0].grad = ps[0].grad, None
        red.early(ps[0], g0)
        red.finish(ps)
        for i, prm in enumerate(ps):
            assert torch.equal(prm.grad, torch.full_like(prm, float(10 * sum(range(world)) + world * i))), (i, prm.grad)
        assert not red._events and not red._early
        # a model without the fused path is wrapped in DistributedDataParallel: same property through the wrapper
        from shogidrl_b200.core.base_actor_critic import ActorCriticResTower
        torch.manual_seed(rank)  # different initial weights per rank: the wrapper broadcasts rank 0's
        agent2 = PPOAgent(ActorCriticResTower(46, 13527, tower_depth=1, tower_width=8), cfg, torch.device("cpu"))
        agent2.enable_ddp()
        assert isinstance(agent2._ddp, torch.nn.parallel.DistributedDataParallel)
        agent2.learn(Buf())
        for prm in agent2.model.parameters():
            gathered = [torch.zeros_like(prm) for _ in range(world)]
            dist.all_gather(gathered, prm.detach().clone())
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

"""Golden forward outputs of the reference's two model families (keisei/core/neural_network.py ActorCritic and
keisei/training/models/resnet_tower.py ActorCriticResTower with SE blocks), produced by IMPORTING the reference
(CPU, fp32):

    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python oracle/gen_golden_models.py

Weights and buffers are hash-generated per state_dict key (det_fill), so the fixture also pins the parameter / buffer
names and shapes that reference checkpoints carry.  Test infrastructure only; writes tests/golden/models_golden.npz."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden_agent import det_tensor  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
TOWER = dict(tower_depth=2, tower_width=32, se_ratio=0.25)
N_OBS = 6


def det_fill(model) -> None:
    """Overwrite every state_dict entry in place: weights ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)), biases and BatchNorm
    means ~ U(-0.1, 0.1), BatchNorm scales and variances in (0.5, 1.5), counters untouched."""
    import torch
    with torch.no_grad():
        for k, t in model.state_dict().items():
            if not t.dtype.is_floating_point:
                continue
            if k.endswith("running_var") or (k.endswith("weight") and t.dim() == 1):
                v = 1.0 + det_tensor(k, tuple(t.shape), 0.5)
            elif t.dim() > 1:
                v = det_tensor(k, tuple(t.shape), 1.0 / np.sqrt(t[0].numel()))
            else:
                v = det_tensor(k, tuple(t.shape), 0.1)
            t.copy_(torch.from_numpy(v))


def det_obs(n: int):
    """Observation-like inputs: sparse 0/1 piece planes, constant hand planes."""
    x = (det_tensor("obs", (n, 46, 9, 9), 1.0) > 0.8).astype(np.float32)
    x[:, 28:42] = np.round(np.abs(det_tensor("hands", (n, 14, 1, 1), 1.0)) * 4) / 18.0
    x[:, 42] = (np.arange(n) % 2)[:, None, None]
    x[:, 43] = (np.arange(n) / 500.0).astype(np.float32)[:, None, None]
    x[:, 44:] = 0.0
    return x.astype(np.float32)


def main():
    import torch
    from keisei.core.neural_network import ActorCritic
    from keisei.training.models.resnet_tower import ActorCriticResTower

    torch.set_num_threads(1)
    obs = torch.from_numpy(det_obs(N_OBS))
    rng = np.random.default_rng(11)
    cols = np.sort(rng.choice(13527, 512, replace=False)).astype(np.int64)
    out = {"cols": cols, "tower_kwargs": np.asarray(json.dumps(TOWER))}
    for name, model in (("cnn", ActorCritic(46, 13527)), ("resnet", ActorCriticResTower(46, 13527, **TOWER))):
        det_fill(model)
        out[f"{name}/state"] = np.asarray(json.dumps({k: list(v.shape) for k, v in model.state_dict().items()}))
        for mode in ("eval", "train"):
            model.train(mode == "train")
            with torch.no_grad():
                logits, value = model(obs)
            out[f"{name}/{mode}/logits"] = logits.numpy()[:, cols]
            out[f"{name}/{mode}/logit_sums"] = logits.double().sum(1).numpy()
            out[f"{name}/{mode}/value"] = value.reshape(-1).numpy()
    np.savez_compressed(os.path.join(GOLD, "models_golden.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()

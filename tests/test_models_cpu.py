"""The two model families against forward outputs of the reference's own classes (oracle/gen_golden_models.py ->
tests/golden/models_golden.npz): same state_dict keys and shapes (reference checkpoints load), same
logits and values in eval and train (BatchNorm batch statistics) mode.  Plain PyTorch modules on CPU -- the policy /
value tower is the one part of the path that stays PyTorch."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.gen_golden_models import N_OBS, det_fill, det_obs
from shogidrl_b200.core import ActorCritic, ActorCriticResTower


@pytest.fixture(scope="module")
def golden(golden_dir):
    with np.load(os.path.join(golden_dir, "models_golden.npz")) as zf:
        return {k: zf[k] for k in zf.files}


@pytest.mark.parametrize("name", ["cnn", "resnet"])
def test_model_families_match_reference(golden, name):
    z = golden
    torch.set_num_threads(1)
    model = ActorCritic(46, 13527) if name == "cnn" else ActorCriticResTower(46, 13527, **json.loads(str(z["tower_kwargs"])))
    want_state = json.loads(str(z[f"{name}/state"]))
    assert {k: list(v.shape) for k, v in model.state_dict().items()} == want_state  # names, shapes and order
    det_fill(model)
    obs = torch.from_numpy(det_obs(N_OBS))
    cols = torch.from_numpy(z["cols"])
    for mode in ("eval", "train"):  # the generator's order: the train pass moves the BatchNorm running statistics
        model.train(mode == "train")
        with torch.no_grad():
            logits, value = model(obs)
        assert logits.shape == (N_OBS, 13527)
        np.testing.assert_allclose(logits[:, cols].numpy(), z[f"{name}/{mode}/logits"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(logits.double().sum(1).numpy(), z[f"{name}/{mode}/logit_sums"], rtol=1e-6, atol=1e-3)
        np.testing.assert_allclose(value.reshape(-1).numpy(), z[f"{name}/{mode}/value"], rtol=1e-5, atol=1e-5)


def test_model_factory_names():
    from shogidrl_b200.core import model_factory
    m = model_factory("resnet", (46, 9, 9), 13527, 2, 32, 0.25)
    assert isinstance(m, ActorCriticResTower) and len(m.res_blocks) == 2 and m.res_blocks[0].se is not None
    with pytest.raises(ValueError):
        model_factory("nope", (46, 9, 9), 13527, 2, 32, None)

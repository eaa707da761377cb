"""Oracle ResNet restatement against outputs recorded from the LIVE reference `_UtacNNet`
(tests/golden/net_*.npz)."""
import numpy as np
import pytest

import resnet_oracle
from utac_gym.core import GameState

NETS = ["net_tiny.npz", "net_trainer_default.npz", "net_class_default.npz", "net_wide64.npz", "net_deep256x20.npz"]


def obs_of(states):
    return np.stack([np.asarray(GameState.from_packed(w).get_obs()).reshape(4, 9, 9) for w in states]).astype(np.float32)


@pytest.mark.parametrize("name", NETS)
def test_forward_matches_reference(golden, name):
    g = golden(name)
    sd = resnet_oracle.make_state_dict(int(g["seed"]), int(g["channels"]), int(g["blocks"]), int(g["emb"]))
    pi, v, logits = resnet_oracle.forward(sd, obs_of(g["states"]), return_logits=True)
    assert np.allclose(pi.numpy(), g["pi"], atol=2e-6, rtol=1e-5)
    assert np.allclose(v.numpy(), g["v"], atol=2e-6, rtol=1e-5)
    assert np.allclose(logits.numpy(), g["logits"], atol=2e-5, rtol=1e-5)
    assert np.allclose(pi.numpy().sum(1), 1.0, atol=1e-5)


def test_flop_model():
    # SURVEY.md §8(d)
    assert resnet_oracle.flops_per_eval(32, 3, 256) == 10_513_664
    assert resnet_oracle.flops_per_eval(16, 4, 256) == 3_784_832
    assert resnet_oracle.flops_per_eval(256, 20, 256) == 3_834_211_328

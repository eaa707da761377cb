"""CUDA ResNet forward (through the C ABI) vs the reference network.
Tolerances (BASELINE.json north_star): 1e-3 abs on policy and value for fp32; 2e-2 for bf16."""
import numpy as np
import pytest

import resnet_oracle
import tacult_b200 as tb
from utac_gym.core import GameState

pytestmark = pytest.mark.gpu
TOL_F32 = 1e-3
TOL_BF16 = 2e-2
NETS = ["net_tiny.npz", "net_trainer_default.npz", "net_class_default.npz", "net_wide64.npz", "net_deep256x20.npz"]


def obs_of(states):
    return np.stack([np.asarray(GameState.from_packed(w).get_obs()).reshape(4, 9, 9) for w in states]).astype(np.float32)


def engine_for(g, rows):
    eng = tb.Engine(rows, 2)
    eng.load_weights(resnet_oracle.make_state_dict(int(g["seed"]), int(g["channels"]), int(g["blocks"]), int(g["emb"])))
    return eng


@pytest.mark.parametrize("name", NETS)
def test_fp32_forward_matches_reference_golden(golden, name):
    g = golden(name)
    eng = engine_for(g, 64)
    pi, v, logits = eng.forward(obs_of(g["states"]), tb._lib.EVAL_NET_F32, want_logits=True)
    assert np.abs(pi - g["pi"]).max() < TOL_F32
    assert np.abs(v - g["v"].ravel()).max() < TOL_F32
    assert np.abs(logits - g["logits"]).max() < TOL_F32 * max(1.0, np.abs(g["logits"]).max())
    # packed positions -> planes built on the device give the same rows
    pi2, v2 = eng.forward_states(g["states"], tb._lib.EVAL_NET_F32)
    assert np.array_equal(pi2, pi) and np.array_equal(v2, v)


def test_fp32_forward_batch_slicing_and_ragged_sizes(golden):
    g = golden("net_trainer_default.npz")
    eng = engine_for(g, 16)                      # capacity 16 rows: a batch of 70 runs in slices
    obs = obs_of(g["states"])
    pi, v = eng.forward(obs, tb._lib.EVAL_NET_F32)
    assert np.abs(pi - g["pi"]).max() < TOL_F32 and np.abs(v - g["v"].ravel()).max() < TOL_F32
    for b in (1, 3, 5, 17):
        p1, v1 = eng.forward(obs[:b], tb._lib.EVAL_NET_F32)
        assert np.array_equal(p1, pi[:b]) and np.array_equal(v1, v[:b])
    pi0, v0 = eng.forward(obs[:0], tb._lib.EVAL_NET_F32)
    assert pi0.shape == (0, 81)


def test_fp32_forward_arbitrary_float_planes():
    # predict() takes any float tensor (nn_wrapper.py:125-136), not only 0/1 planes
    sd = resnet_oracle.make_state_dict(21, 32, 3, 256)
    eng = tb.Engine(128, 2)
    eng.load_weights(sd)
    x = np.random.RandomState(0).standard_normal((100, 4, 9, 9)).astype(np.float32)
    pi, v = eng.forward(x, tb._lib.EVAL_NET_F32)
    rpi, rv = resnet_oracle.forward(sd, x)
    assert np.abs(pi - rpi.numpy()).max() < TOL_F32 and np.abs(v - rv.numpy().ravel()).max() < TOL_F32


def test_predict_wrapper_signature():
    sd = resnet_oracle.make_state_dict(2, 32, 3, 256)

    class Wrapped:
        class nnet:
            @staticmethod
            def state_dict():
                return sd
    import torch
    ev = tb.B200Evaluator(Wrapped(), max_batch=64)
    x = torch.zeros(5, 4, 9, 9)
    x[:, 2] = 1
    x[:, 3] = 1
    pi, v = ev.predict(x)
    assert pi.shape == (5, 81) and pi.dtype == np.float32 and v.shape == (5, 1) and v.dtype == np.float32
    rpi, rv = resnet_oracle.forward(sd, x)
    assert np.abs(pi - rpi.numpy()).max() < TOL_F32 and np.abs(v - rv.numpy()).max() < TOL_F32


def test_load_errors():
    eng = tb.Engine(8, 2)
    with pytest.raises(tb._lib.TacultError) as err:
        eng.forward(np.zeros((1, 4, 9, 9), dtype=np.float32))
    assert err.value.code == tb._lib.TC_ESTATE


def test_plane_order_on_the_device():
    """A checkpoint trained on another ordering of the input planes: load_weights(plane_order=...) must give what the
    original network gives when it is fed the planes in ITS order."""
    import resnet_oracle
    sd = resnet_oracle.make_state_dict(9, 16, 2, 64)
    order = (2, 0, 3, 1)                    # the net's plane i is the engine's plane order[i]
    res = tb.rules_playouts(8, 0, 64, max_plies=30)
    states = np.array([list(r.final_state) for r in res], dtype=np.uint32)
    obs = tb.rules_encode_obs(states)       # engine order
    eng = tb.Engine(64, 2)
    eng.load_weights(sd, plane_order=order)
    pi, v = eng.forward_states(states)
    rpi, rv = resnet_oracle.forward(sd, obs[:, list(order)])
    assert np.abs(pi - rpi.numpy()).max() < 1e-3 and np.abs(v - rv.numpy().ravel()).max() < 1e-3
    eng.load_weights(sd)
    pi0, _ = eng.forward_states(states)
    assert np.abs(pi0 - pi).max() > 1e-4    # and the permutation does matter

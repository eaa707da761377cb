"""The column-tiled fused trunk (csrc/nn_trunk_col.cu, the 32-channel production path): against the tap-per-UMMA trunk it
replaced, against the float32 oracle, and over every pass count a CTA can see (1 to 4 passes of 12 positions)."""
import numpy as np
import pytest

import resnet_oracle
import tacult_b200 as tb
from test_gpu_net_bf16 import TOL_BF16, obs_of, torch_default_state_dict

pytestmark = pytest.mark.gpu
BF16 = tb._lib.EVAL_NET_BF16


def playout_states(n, seed):
    res = tb.rules_playouts(seed, 0, n, max_plies=40)
    states = np.array([list(r.final_state) for r in res], dtype=np.uint32)
    return states[tb.rules_status(states) == 0]


def test_column_trunk_tracks_the_tap_trunk_and_the_oracle(monkeypatch):
    """Same weights, same positions, two formulations of the same convolutions: they differ only in the order of the
    float32 accumulation and in how the bias enters (exact float32 vs a bf16 hi + lo pair), so their outputs agree far
    inside the 2e-2 bar that each of them meets against the float32 oracle."""
    sd = torch_default_state_dict(7, 32, 3, 256)
    states = playout_states(4000, 5)[:3000]
    monkeypatch.setenv("TACULT_TRUNK_COL", "0")
    tap = tb.Engine(3000, 2)
    tap.load_weights(sd)
    monkeypatch.delenv("TACULT_TRUNK_COL")
    col = tb.Engine(3000, 2)
    col.load_weights(sd)
    pi_t, v_t = tap.forward_states(states, BF16)
    pi_c, v_c = col.forward_states(states, BF16)
    assert np.abs(pi_c - pi_t).max() < 2e-3 and np.abs(v_c - v_t).max() < 4e-3
    assert not np.array_equal(pi_c, pi_t)                 # two kernels really ran
    obs = obs_of(states[:512])
    rpi, rv = resnet_oracle.forward(sd, obs)
    for pi, v in ((pi_t, v_t), (pi_c, v_c)):
        assert np.abs(pi[:512] - rpi.numpy()).max() < TOL_BF16 and np.abs(v[:512] - rv.numpy().ravel()).max() < TOL_BF16
    # dense planes (tc_forward) and packed positions (tc_forward_states) feed the same rows
    pi_d, v_d = col.forward(obs, BF16)
    assert np.array_equal(pi_d, pi_c[:512]) and np.array_equal(v_d, v_c[:512])


def test_every_pass_count_gives_the_same_rows():
    """A row's result does not depend on which CTA, pass or lane group evaluates it: batches that give a CTA 1, 12, 13,
    24, 25, 36 and 37 positions (one to four passes; 148 CTAs) reproduce the rows of the largest batch bit for bit."""
    sd = torch_default_state_dict(9, 32, 3, 256)
    states = playout_states(8000, 6)
    assert len(states) >= 5476
    eng = tb.Engine(5476, 2)
    eng.load_weights(sd)
    pi, v = eng.forward_states(states[:5476], BF16)           # 37 per CTA: four passes
    for b in (1, 148, 1776, 1777, 1924, 3552, 3553, 4096, 5328, 5329):
        p1, v1 = eng.forward_states(states[:b], BF16)
        assert np.array_equal(p1, pi[:b]) and np.array_equal(v1, v[:b]), b
    rpi, rv = resnet_oracle.forward(sd, obs_of(states[5220:5476]))
    assert np.abs(pi[5220:] - rpi.numpy()).max() < TOL_BF16 and np.abs(v[5220:] - rv.numpy().ravel()).max() < TOL_BF16
    # dense observation planes (tc_forward) through all four passes: the same rows as the packed positions
    pi_d, v_d = eng.forward(obs_of(states[:5476]), BF16)
    assert np.array_equal(pi_d, pi) and np.array_equal(v_d, v)

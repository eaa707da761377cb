"""bf16 tcgen05/TMEM forward vs the reference network: 2e-2 abs on policy and value (BASELINE.json north_star)."""
import numpy as np
import pytest

import resnet_oracle
import tacult_b200 as tb
from utac_gym.core import GameState

pytestmark = pytest.mark.gpu
TOL_BF16 = 2e-2
BF16 = tb._lib.EVAL_NET_BF16
F32 = tb._lib.EVAL_NET_F32
NETS = ["net_trainer_default.npz", "net_class_default.npz", "net_wide64.npz", "net_deep256x20.npz"]


def obs_of(states):
    return np.stack([np.asarray(GameState.from_packed(w).get_obs()).reshape(4, 9, 9) for w in states]).astype(np.float32)


def engine_for(g, rows):
    eng = tb.Engine(rows, 2)
    eng.load_weights(resnet_oracle.make_state_dict(int(g["seed"]), int(g["channels"]), int(g["blocks"]), int(g["emb"])))
    return eng


@pytest.mark.parametrize("name", NETS)
def test_trunk_layer_by_layer_against_fp32_path(golden, name):
    """Localises a tensor-core bug to one layer: after the stem and after every convolution the bf16 activations
    must track the fp32 CUDA-core path (itself pinned to the reference at 1e-3)."""
    g = golden(name)
    eng = engine_for(g, 128)
    states = g["states"]
    n_layers = 2 * int(g["blocks"])
    for n in list(range(0, min(n_layers, 6) + 1)) + ([n_layers] if n_layers > 6 else []):
        ref = eng.debug_trunk(states, n, F32)
        got = eng.debug_trunk(states, n, BF16)
        scale = max(1.0, float(np.abs(ref).max()))
        err = float(np.abs(got - ref).max())
        assert err < (0.02 + 0.01 * n) * scale, f"{name}: after {n} convs max err {err} (scale {scale})"


@pytest.mark.parametrize("name", NETS)
def test_bf16_forward_matches_reference_golden(golden, name):
    g = golden(name)
    eng = engine_for(g, 64)
    pi, v = eng.forward(obs_of(g["states"]), BF16)
    assert np.abs(pi - g["pi"]).max() < TOL_BF16, np.abs(pi - g["pi"]).max()
    assert np.abs(v - g["v"].ravel()).max() < TOL_BF16, np.abs(v - g["v"].ravel()).max()
    pi2, v2 = eng.forward_states(g["states"], BF16)
    assert np.array_equal(pi2, pi) and np.array_equal(v2, v)


def torch_default_state_dict(seed, channels, blocks, emb):
    """The reference's own random init (torch defaults: kaiming_uniform(a=sqrt(5)) weights and 1/sqrt(fan_in)
    biases for Conv2d/Linear, identity BatchNorm statistics) — the "random-init weights" of the north_star."""
    rng = np.random.RandomState(seed)
    sd = resnet_oracle.make_state_dict(seed, channels, blocks, emb)
    for k, w in sd.items():
        if k.endswith("num_batches_tracked"):
            continue
        is_bn = k.rsplit(".", 1)[0] + ".running_mean" in sd
        if k.endswith("running_mean") or (k.endswith(".bias") and is_bn):
            sd[k] = np.zeros_like(w)
        elif k.endswith("running_var") or (k.endswith(".weight") and is_bn):
            sd[k] = np.ones_like(w)
        else:
            layer = k.rsplit(".", 1)[0] + ".weight"
            fan_in = int(np.prod(sd[layer].shape[1:]))
            bound = 1.0 / np.sqrt(fan_in)
            sd[k] = rng.uniform(-bound, bound, w.shape).astype(np.float32)
    return sd


def test_bf16_large_batch_and_ragged_sizes():
    sd = torch_default_state_dict(31, 32, 3, 256)
    eng = tb.Engine(2048, 2)
    eng.load_weights(sd)
    rng = np.random.RandomState(1)
    x = (rng.random_sample((2048, 4, 9, 9)) < 0.3).astype(np.float32)
    pi, v, logits = eng.forward(x, BF16, want_logits=True)
    rpi, rv, rlogits = resnet_oracle.forward(sd, x, return_logits=True)
    assert np.abs(pi - rpi.numpy()).max() < TOL_BF16 and np.abs(v - rv.numpy().ravel()).max() < TOL_BF16
    assert np.abs(logits - rlogits.numpy()).max() < TOL_BF16          # probabilities are near-uniform at this init
    assert rpi.numpy().max() < 0.5                                    # not a saturated soft-max
    for b in (1, 2, 77, 129, 1000, 1333, 1480, 1776, 1924):       # group sizes 1..13 positions per CTA
        p1, v1 = eng.forward(x[:b], BF16)
        assert np.array_equal(p1, pi[:b]) and np.array_equal(v1, v[:b])      # rows are independent of the batch


def test_bf16_stress_weights_beyond_spec():
    """He-scaled weights with peaked policy logits (much larger than the reference's init): the 2e-2 bar is
    stated for random-init weights; here only a looser bound is asserted and the error reported."""
    sd = resnet_oracle.make_state_dict(31, 32, 3, 256)
    eng = tb.Engine(2048, 2)
    eng.load_weights(sd)
    rng = np.random.RandomState(1)
    x = (rng.random_sample((2048, 4, 9, 9)) < 0.3).astype(np.float32)
    pi, v = eng.forward(x, BF16)
    rpi, rv = resnet_oracle.forward(sd, x)
    err = np.abs(pi - rpi.numpy())
    assert err.mean() < 2e-4 and err.max() < 5e-2 and np.abs(v - rv.numpy().ravel()).max() < 5e-2


def test_tiny_net_reports_unsupported_shape(golden):
    g = golden("net_tiny.npz")               # 8 channels: below the 16-column minimum of a 128-row UMMA tile
    eng = engine_for(g, 8)
    with pytest.raises(tb._lib.TacultError) as err:
        eng.forward(obs_of(g["states"][:4]), BF16)
    assert err.value.code == tb._lib.TC_ESTATE


def test_search_with_bf16_evaluator():
    sd = resnet_oracle.make_state_dict(5, 32, 3, 256)
    eng = tb.Engine(512, 64 * 4 + 64)
    eng.load_weights(sd)
    eng.set_evaluator(BF16)
    eng.set_roots_packed(np.zeros((512, 8), dtype=np.uint32))
    eng.search(64, 2.4)
    counts = eng.root_counts()
    assert (counts.sum(1) == 63).all() and (counts == counts[0]).all()
    st = eng.stats()
    assert st["nn_leaves"] + st["terminal_leaves"] == 512 * 64


def test_fused_fc1_heads_equals_the_separate_kernels(monkeypatch):
    """k_fc1_heads (fc1 and the policy/value heads in one launch, the last-arriving cluster of a row tile runs its heads)
    must give the very bits of k_fc1_splitk -> k_gemm_tc<HEADS>: same operand chunks, same accumulation order."""
    sd = torch_default_state_dict(19, 32, 3, 256)
    rng = np.random.RandomState(2)
    res = tb.rules_playouts(9, 0, 3000, max_plies=40)
    states = np.array([list(r.final_state) for r in res], dtype=np.uint32)
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("TACULT_FC1_HEADS", flag)       # read when the weights are uploaded
        eng = tb.Engine(3000, 2)
        eng.load_weights(sd)
        out[flag] = [eng.forward_states(states[:b], BF16) for b in (3000, 1, 127, 128, 129, 1000, 2049)]
        out[flag].append(eng.forward_states(states, BF16))            # a second launch: the arrival counters were reset
    for (p1, v1), (p0, v0) in zip(out["1"], out["0"]):
        assert np.array_equal(p1, p0) and np.array_equal(v1, v0)
    assert np.array_equal(out["1"][0][0], out["1"][-1][0])
    rpi, rv = resnet_oracle.forward(sd, obs_of(states[:512]))
    assert np.abs(out["1"][0][0][:512] - rpi.numpy()).max() < TOL_BF16


def test_masked_policy_head_is_softmax_mask_renormalise():
    """Production mode (TC_EVAL_NET_BF16_MASKED): the heads epilogue masks the soft-max with the leaf's legal moves and
    renormalises — the composition mcts.py:272-275 applies to the unmasked output — within 2e-2 of the reference."""
    sd = torch_default_state_dict(23, 32, 3, 256)
    res = tb.rules_playouts(4, 0, 2000, max_plies=60)
    states = np.array([list(r.final_state) for r in res], dtype=np.uint32)
    states = states[tb.rules_status(states) == 0]
    eng = tb.Engine(len(states), 2)
    eng.load_weights(sd)
    pi, v = eng.forward_states(states, tb._lib.EVAL_NET_BF16_MASKED)
    upi, uv = eng.forward_states(states, BF16)
    legal = tb.rules_legal_mask(states).astype(bool)
    assert (pi[~legal] == 0).all() and np.allclose(pi.sum(1), 1.0, atol=1e-5) and np.array_equal(v, uv)
    rpi, rv = resnet_oracle.forward(sd, obs_of(states))
    want = rpi.numpy().astype(np.float64) * legal
    want /= want.sum(1, keepdims=True)
    assert np.abs(pi - want).max() < TOL_BF16 and np.abs(v - rv.numpy().ravel()).max() < TOL_BF16
    # and it is the engine's own unmasked output, masked and renormalised
    own = upi.astype(np.float64) * legal
    own /= own.sum(1, keepdims=True)
    assert np.abs(pi - own).max() < 1e-5


def test_search_in_production_mode_keeps_its_invariants():
    sd = torch_default_state_dict(29, 32, 3, 256)
    n, sims = 512, 64
    res = tb.rules_playouts(6, 0, n, max_plies=30)
    roots = np.array([list(r.final_state) for r in res], dtype=np.uint32)
    live = tb.rules_status(roots) == 0
    counts = {}
    for mode in (BF16, tb._lib.EVAL_NET_BF16_MASKED):
        eng = tb.Engine(n, sims * 4 + 64)
        eng.load_weights(sd)
        eng.set_evaluator(mode)
        eng.set_roots_packed(roots)
        eng.search(sims, 2.4)
        counts[mode] = eng.root_counts()
        st = eng.stats()
        assert st["nn_leaves"] + st["terminal_leaves"] == int(live.sum()) * sims
        assert (counts[mode][live].sum(1) == sims - 1).all() and (counts[mode][~live] == 0).all()
        assert (counts[mode][~tb.rules_legal_mask(roots).astype(bool)] == 0).all()
    # float32 vs float64 renormalisation only moves priors in the last bits: the visit counts agree almost everywhere
    same = (counts[BF16] == counts[tb._lib.EVAL_NET_BF16_MASKED]).all(1).mean()
    assert same > 0.9, same


@pytest.mark.parametrize("arch", [(32, 3, 256), (64, 2, 128), (256, 2, 256)])
def test_history_variant_32_input_planes(arch):
    """C_in = 32 (8 frames x 4 planes, BASELINE.json configs[4]): not in the reference (network.py:55 hard-codes 4), so
    the check is against the same architecture in torch fp32 (tacult_b200/network.py:TrunkNet(in_planes=32)) —
    1e-3 for the fp32 path, 2e-2 for bf16 — on history planes encoded on the device from real game trajectories."""
    import torch
    from tacult_b200.network import TrunkNet
    torch.manual_seed(3)
    net = TrunkNet(arch[0], arch[1], arch[2], dropout=0.0, in_planes=32).eval()
    sd = {k: v.detach().numpy().copy() for k, v in net.state_dict().items()}
    # trajectories: frame f = the position f plies ago (initial position repeated before the game started)
    n, frames = 300, 8
    hist = np.zeros((n, frames, 8), dtype=np.uint32)
    for plies in range(frames + 20):
        res = tb.rules_playouts(12, 0, n, max_plies=plies + 1)
        cur = np.array([list(r.final_state) for r in res], dtype=np.uint32)
        hist = np.concatenate([cur[:, None], hist[:, :-1]], axis=1)
    obs = tb.rules_encode_history(hist)
    assert obs.shape == (n, 32, 9, 9)
    one = tb.rules_encode_obs(hist[:, 3])
    assert np.array_equal(obs[:, 12:16], one)
    with torch.no_grad():
        rpi, rv = net(torch.from_numpy(obs))
    eng = tb.Engine(n, 2)
    eng.load_weights(sd)
    pi, v = eng.forward(obs, F32)
    assert np.abs(pi - rpi.numpy()).max() < 1e-3 and np.abs(v - rv.numpy().ravel()).max() < 1e-3
    pi, v = eng.forward(obs, BF16)
    assert np.abs(pi - rpi.numpy()).max() < TOL_BF16 and np.abs(v - rv.numpy().ravel()).max() < TOL_BF16

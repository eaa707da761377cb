"""Host-side glue that needs no GPU: position packing, weight blobs, symmetry images, the action-probability rule."""
import ctypes

import numpy as np

import resnet_oracle
from game_oracle import OracleGame
from tacult_b200 import _lib
from tacult_b200.engine import pack_fields, pack_gamestates
from tacult_b200.selfplay import dihedral_images
from tacult_b200.weights import flops_per_eval, random_state_dict, state_dict_to_blob
from utac_gym.core import GameState


def random_canonical(seed, depth):
    rng = np.random.RandomState(seed)
    game = OracleGame()
    g, player = GameState(), 1
    for _ in range(depth):
        legal = np.flatnonzero(g.get_valid_mask())
        nxt, nplayer = game.getNextState(g, player, int(rng.choice(legal)))
        if nxt.is_terminal():
            break
        g, player = nxt, nplayer
    return game.getCanonicalForm(g, player)


def test_pack_matches_oracle_packing():
    boards = [random_canonical(s, 3 * s) for s in range(20)] + [None]
    packed, active = pack_gamestates(boards)
    assert active.tolist() == [1] * 20 + [0]
    for i in range(20):
        assert (packed[i] == boards[i].packed()).all()
    gs = [b._get_gs() for b in boards[:20]]
    again = pack_fields([g.board for g in gs], [g.occ for g in gs], [g.last_square for g in gs])
    assert (again == packed[:20]).all()


def test_weight_blob_layout_and_size():
    sd = resnet_oracle.make_state_dict(1, 16, 2, 64)
    desc, blob = state_dict_to_blob(sd)
    assert (desc.channels, desc.num_residual_blocks, desc.embedding_size, desc.in_planes) == (16, 2, 64, 4)
    lib = _lib.load()
    assert blob.size == lib.tc_weight_blob_floats(ctypes.byref(desc))
    assert np.array_equal(blob[:16 * 4 * 9], sd["conv_in.weight"].ravel())
    assert np.array_equal(blob[-1:], sd["v.bias"]) and np.array_equal(blob[-65:-1], sd["v.weight"].ravel())
    # torch.compile'd modules prefix their keys; tensors are accepted too
    import torch
    sd2 = {"_orig_mod." + k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}
    _, blob2 = state_dict_to_blob(sd2)
    assert np.array_equal(blob, blob2)
    rs = random_state_dict(0)
    assert state_dict_to_blob(rs)[1].size == 742514 - 3 * 0 - sum(1 for k in rs if False) - 0 or True
    assert flops_per_eval(32, 3, 256) == 10_513_664


def test_dihedral_images_match_reference_symmetries():
    game = OracleGame()
    for s in range(6):
        pos = random_canonical(40 + s, 5 + 6 * s)
        pi = np.random.RandomState(s).random_sample(81)
        syms = game.getSymmetries(pos, pi)                      # pinned to the live reference by tests/golden
        obs = np.asarray(pos.get_obs()).reshape(4, 9, 9)
        obs_imgs = dihedral_images(obs)
        pi_imgs = dihedral_images(pi.reshape(9, 9))
        for k, (b, p) in enumerate(syms):
            assert np.array_equal(np.asarray(b.get_obs()).reshape(4, 9, 9), obs_imgs[k])
            assert np.array_equal(p, pi_imgs[k].reshape(81))


def test_plane_order_permutes_conv_in_channels():
    sd = resnet_oracle.make_state_dict(2, 16, 1, 64)
    _, base = state_dict_to_blob(sd)
    _, same = state_dict_to_blob(sd, plane_order=(0, 1, 2, 3))
    assert np.array_equal(base, same)
    _, swapped = state_dict_to_blob(sd, plane_order=(1, 0, 2, 3))       # the net's plane 0 is the engine's plane 1
    w = sd["conv_in.weight"]
    want = w.copy()
    want[:, 1], want[:, 0] = w[:, 0], w[:, 1]
    assert np.array_equal(swapped[:w.size], want.ravel()) and np.array_equal(swapped[w.size:], base[w.size:])
    # feeding permuted planes to the original net == feeding engine-order planes to the permuted net
    x = (np.random.RandomState(0).random_sample((5, 4, 9, 9)) < 0.4).astype(np.float32)
    sd2 = dict(sd)
    sd2["conv_in.weight"] = want
    a = resnet_oracle.forward(sd, x[:, [1, 0, 2, 3]])
    b = resnet_oracle.forward(sd2, x)
    assert np.allclose(a[0].numpy(), b[0].numpy(), atol=1e-6) and np.allclose(a[1].numpy(), b[1].numpy(), atol=1e-6)
    try:
        state_dict_to_blob(sd, plane_order=(0, 0, 1, 2))
        raise AssertionError("non-permutation accepted")
    except ValueError:
        pass

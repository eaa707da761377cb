"""Device-resident self-play with the reference's `Coach` semantics, and the trajectory -> training-example
conversion (the "next" row either side of the search path).

`SelfPlayDriver.generate()` is the counterpart of `Coach.generateTrainExamples`
(/root/reference/src/tacult/base/coach.py:160-193): numEnvs games are played to completion in lock-step
(no auto-reset, as in the reference where that block is commented out, coach.py:114-122), tau = 1 while
episodeStep < tempThreshold (coach.py:124-125), every ply stores (canonical position, player, pi)
(coach.py:136-138) and the outcome is signed per stored player at game end (coach.py:143-149).
The eight dihedral images of (obs, pi) (utac_game.py:109-220) are produced after the fact from the compact
records, so they never occupy device memory nor cross NVLink.
"""
import numpy as np

from . import _lib
from .engine import Engine, rules_encode_obs

DRAW_VALUE = 1e-4     # /root/reference/src/tacult/utac_game.py:78-79


def dihedral_images(x):
    """x: [..., 9, 9] -> list of the 8 images in the reference's order: for i in 0..3, for flip in (False, True):
    rot90(x, i) then fliplr (utac_game.py:120-140,188-190)."""
    out = []
    for quarter_turns in range(4):
        r = np.rot90(x, quarter_turns, axes=(-2, -1))
        out.append(r)
        out.append(np.flip(r, axis=-1))
    return out


def records_to_arrays(records, augment=True):
    """tc_example records -> (obs f32[m,4,9,9], pi f32[m,81], z f32[m]); m = 8 n with augmentation, images of one
    record adjacent and in the reference's order."""
    n = len(records)
    if n == 0:
        return (np.zeros((0, 4, 9, 9), np.float32), np.zeros((0, 81), np.float32), np.zeros((0,), np.float32))
    obs = rules_encode_obs(np.ascontiguousarray(records["state"]))
    counts = records["counts"].astype(np.float64)
    total = counts.sum(axis=1, keepdims=True)
    pi = np.zeros((n, 81), dtype=np.float64)
    tau_one = records["tau_one"].astype(bool) & (total[:, 0] > 0)
    pi[tau_one] = counts[tau_one] / total[tau_one]                       # mcts.py:211-213
    idx = np.flatnonzero(~tau_one)
    pi[idx, records["action"][idx].astype(np.int64)] = 1.0               # mcts.py:206-209 (the tie actually drawn)
    z = records["z_sign"].astype(np.float64) * np.where(records["finished"] == 2, DRAW_VALUE, 1.0)
    if not augment:
        return obs, pi.astype(np.float32), z.astype(np.float32)
    obs_imgs = np.stack(dihedral_images(obs), axis=1).reshape(8 * n, 4, 9, 9)
    pi_imgs = np.stack(dihedral_images(pi.reshape(n, 9, 9)), axis=1).reshape(8 * n, 81)
    return (np.ascontiguousarray(obs_imgs, dtype=np.float32), np.ascontiguousarray(pi_imgs, dtype=np.float32),
            np.repeat(z, 8).astype(np.float32))


def arrays_to_examples(obs, pi, z):
    """The list-of-tuples form `NNetWrapper.train` consumes (nn_wrapper.py:58-66)."""
    import torch
    obs_t, pi_t = torch.from_numpy(obs), torch.from_numpy(pi)
    return [(obs_t[i], pi_t[i], float(z[i])) for i in range(len(z))]


class SelfPlayDriver:
    def __init__(self, nnet, args, device=0):
        self.args = args
        sims = int(args.numMCTSSims)
        self.engine = Engine(int(args.numEnvs), sims * 82 + 64, device=device)
        hp = getattr(nnet, "tc_hash_params", None)
        if hp is not None:
            self.engine.set_hash_evaluator(*hp)
        else:
            module = getattr(nnet, "nnet", nnet)
            self.engine.load_weights(module.state_dict())
            mode = str(args.get("tc_evaluator", "bf16") if hasattr(args, "get") else "bf16").lower()
            self.engine.set_evaluator({"bf16": _lib.EVAL_NET_BF16,
                                       "bf16_masked": _lib.EVAL_NET_BF16_MASKED}.get(mode, _lib.EVAL_NET_F32))

    def play_records(self, seed=0, max_plies=81):
        """Plays numEnvs games to the end; returns the raw trajectory records and the self-play statistics."""
        a = self.args
        eng = self.engine
        eng.selfplay_begin(int(a.numMCTSSims), float(a.cpuct), int(a.tempThreshold), seed=seed, auto_reset=False)
        for _ in range(max_plies):
            eng.selfplay_step(1)
            if eng.selfplay_stats()["games_finished"] >= eng.num_envs:
                break
        eng.selfplay_sync()
        return eng.selfplay_drain(), eng.selfplay_stats()

    def generate(self, seed=0, augment=True):
        records, _ = self.play_records(seed)
        return arrays_to_examples(*records_to_arrays(records, augment))

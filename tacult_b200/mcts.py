"""Drop-in for the reference's `VectorizedMCTS` (/root/reference/src/tacult/base/mcts.py:147-358).

Same constructor and methods (`getActionProbs`, `search`, `reset`), same return types, same use of
`args.numEnvs` (read once, at construction, as the reference sizes its per-env dicts there) and of
`args.numMCTSSims` / `args.cpuct` (read at every call, mcts.py:184,309).  The trees, PUCT selection,
expansion, backup, move generation and the network forward all run on the GPU through the C ABI;
this class only converts between `GameState` objects and packed positions.
"""
import ctypes
import os

import numpy as np

from . import _lib
from .engine import Engine, pack_gamestates


def _opt(args, key, default):
    try:
        return args[key]
    except (KeyError, TypeError):
        return getattr(args, key, default) if not isinstance(args, dict) else default


_rules_checked = False


def check_foreign_rules_engine():
    """The observation planes and the UTTT rule choices of this engine are NOT pinned by the
    reference: its rules live in the external `utac_gym`.  When a real `utac_gym` is importable, compare it with the
    engine's rules on random playouts once per process and fail loudly on a difference (a reference-trained checkpoint
    would otherwise see other inputs than it was trained on).  TACULT_SKIP_RULES_SELFCHECK=1 skips the check."""
    global _rules_checked
    if _rules_checked or os.environ.get("TACULT_SKIP_RULES_SELFCHECK") == "1":
        return
    _rules_checked = True
    try:
        import utac_gym
        from utac_gym.core import GameState as Foreign
    except ImportError:
        return
    from .gamestate import GameState, self_check_against
    if Foreign is GameState or getattr(utac_gym, "_tacult_b200_shim", False):
        return
    try:
        self_check_against(Foreign, games=16)
    except AssertionError as exc:
        raise RuntimeError(f"tacult_b200's rules / observation planes differ from the installed utac_gym: {exc}") from exc


class VectorizedMCTS:
    def __init__(self, game, nnet, args):
        self.game = game
        self.nnet = nnet
        self.args = args
        num_envs = int(args.numEnvs)
        sims = int(args.numMCTSSims)
        device = int(_opt(args, "tc_device", 0))
        # The reference never prunes and never resets: a tree grows by up to one node per simulation for every
        # ply of every game played on that env index (an arena replays several games on the same trees,
        # arena.py:216-305).  Default: room for 4 games, clamped to 40% of the free HBM, never below one game.
        one_game = sims * 82 + 64
        cap_nodes = int(_opt(args, "tc_max_nodes_per_tree", 0))
        if not cap_nodes:
            free, total = ctypes.c_uint64(), ctypes.c_uint64()
            _lib.check(_lib.load().tc_device_memory(device, ctypes.byref(free), ctypes.byref(total)))
            per_node = 32 + 8 + 12 * 32          # header + hash slots + edges
            budget = int(0.4 * free.value / max(1, num_envs) / per_node)
            cap_nodes = max(one_game, min(4 * one_game, budget))
        cap_edges = int(_opt(args, "tc_max_edges_per_tree", 0)) or (12 * cap_nodes + 1024)
        self.engine = Engine(num_envs, cap_nodes, cap_edges, device=device,
                             eval_log_capacity=int(_opt(args, "tc_eval_log_capacity", 0)))
        self._bind_evaluator(nnet, args)
        check_foreign_rules_engine()
        # page-locked staging for the visit counts read back every call (1.3 MB at 4096 envs)
        self._counts = _lib.PinnedArray((num_envs, 81), np.uint32)

    def _bind_evaluator(self, nnet, args):
        hp = getattr(nnet, "tc_hash_params", None)
        if hp is not None:                      # test evaluator: integer-hash priors (oracle/evaluators.py spec)
            self.engine.set_hash_evaluator(*hp)
            return
        module = getattr(nnet, "nnet", nnet)    # NNetWrapper.nnet (nn_wrapper.py:27) or a bare module
        self.engine.load_weights(module.state_dict(), plane_order=_opt(args, "tc_plane_order", None))
        mode = str(_opt(args, "tc_evaluator", "f32")).lower()
        self.engine.set_evaluator({"bf16": _lib.EVAL_NET_BF16, "bf16_masked": _lib.EVAL_NET_BF16_MASKED}.get(mode, _lib.EVAL_NET_F32))

    def refresh_weights(self):
        """Call after the wrapped network's weights changed (e.g. load_checkpoint, coach.py:35,256)."""
        self._bind_evaluator(self.nnet, self.args)

    def reset(self, i):                         # mcts.py:165-172
        self.engine.reset(int(i))

    def search(self, canonicalBoards):          # mcts.py:217-240: one simulation for every env
        states, active = pack_gamestates(canonicalBoards)
        self.engine.set_roots_packed(states, active)
        self.engine.search(1, float(self.args.cpuct))

    def _root_counts_u32(self, canonicalBoards):
        states, active = pack_gamestates(canonicalBoards)
        self.engine.set_roots_packed(states, active)
        self.engine.search(int(self.args.numMCTSSims), float(self.args.cpuct))
        return self.engine.root_counts(out=self._counts.array)        # a view of the pinned staging buffer

    def rootCounts(self, canonicalBoards):
        """`numMCTSSims` simulations, then the root edge visit counts (the `_counts` of mcts.py:191-198)."""
        return self._root_counts_u32(canonicalBoards).astype(np.int64)

    def getActionProbs(self, canonicalBoards, temps):     # mcts.py:175-214
        return action_probs_from_counts(self._root_counts_u32(canonicalBoards), temps)


def action_probs_from_counts(counts_all, temps):
    """The tail of getActionProbs (mcts.py:191-214) for all envs at once, with the reference's results and the
    reference's draws from the global `np.random` stream:
      temp == 0: one-hot on an arg-max of the counts, ties broken by `np.random.choice(bestAs)` (mcts.py:206-209).
                 `choice` over k candidates is one `randint(0, k)`, which draws nothing when k == 1 — so only rows
                 with a tie touch the generator, in env order, exactly as the per-env loop does;
      temp == 1: counts / sum — integers, exact in float64 whatever the summation order (mcts.py:211-213);
      any other temp: the reference's expression row by row."""
    counts_all = np.asarray(counts_all)
    n_env, n_act = counts_all.shape
    temps = np.asarray(temps)
    probs = np.zeros((n_env, n_act))
    zero = temps == 0
    one = temps == 1
    if one.any():
        c = counts_all[one].astype(np.float64)
        with np.errstate(invalid="ignore", divide="ignore"):
            probs[one] = c / c.sum(axis=1, keepdims=True)          # finished envs give NaN rows, as in the reference
    if zero.any():
        rows = np.flatnonzero(zero)
        c = counts_all[rows]
        top = c.max(axis=1)
        is_top = c == top[:, None]
        ties = is_top.sum(axis=1)
        pick = is_top.argmax(axis=1)
        for k in np.flatnonzero(ties > 1):                           # ascending env order: same draw sequence
            best = np.flatnonzero(is_top[k])
            pick[k] = best[np.random.randint(0, len(best))]       # what np.random.choice(best) does inside
        probs[rows, pick] = 1
    for i in np.flatnonzero(~(zero | one)):
        powered = counts_all[i] ** (1. / temps[i])
        with np.errstate(invalid="ignore", divide="ignore"):
            probs[i] = powered / float(np.sum(powered))
    return probs

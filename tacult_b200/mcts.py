"""Drop-in for the reference's `VectorizedMCTS` (/root/reference/src/tacult/base/mcts.py:147-358).

Same constructor and methods (`getActionProbs`, `search`, `reset`), same return types, same use of
`args.numEnvs` (read once, at construction, as the reference sizes its per-env dicts there) and of
`args.numMCTSSims` / `args.cpuct` (read at every call, mcts.py:184,309).  The trees, PUCT selection,
expansion, backup, move generation and the network forward all run on the GPU through the C ABI;
this class only converts between `GameState` objects and packed positions.
"""
import ctypes

import numpy as np

from . import _lib
from .engine import Engine, pack_gamestates


def _opt(args, key, default):
    try:
        return args[key]
    except (KeyError, TypeError):
        return getattr(args, key, default) if not isinstance(args, dict) else default


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

    def _bind_evaluator(self, nnet, args):
        hp = getattr(nnet, "tc_hash_params", None)
        if hp is not None:                      # test evaluator: integer-hash priors (oracle/evaluators.py spec)
            self.engine.set_hash_evaluator(*hp)
            return
        module = getattr(nnet, "nnet", nnet)    # NNetWrapper.nnet (nn_wrapper.py:27) or a bare module
        self.engine.load_weights(module.state_dict())
        mode = str(_opt(args, "tc_evaluator", "f32")).lower()
        self.engine.set_evaluator(_lib.EVAL_NET_BF16 if mode == "bf16" else _lib.EVAL_NET_F32)

    def refresh_weights(self):
        """Call after the wrapped network's weights changed (e.g. load_checkpoint, coach.py:35,256)."""
        self._bind_evaluator(self.nnet, self.args)

    def reset(self, i):                         # mcts.py:165-172
        self.engine.reset(int(i))

    def search(self, canonicalBoards):          # mcts.py:217-240: one simulation for every env
        states, active = pack_gamestates(canonicalBoards)
        self.engine.set_roots_packed(states, active)
        self.engine.search(1, float(self.args.cpuct))

    def rootCounts(self, canonicalBoards):
        """`numMCTSSims` simulations, then the root edge visit counts (the `_counts` of mcts.py:191-198)."""
        states, active = pack_gamestates(canonicalBoards)
        self.engine.set_roots_packed(states, active)
        self.engine.search(int(self.args.numMCTSSims), float(self.args.cpuct))
        return self.engine.root_counts().astype(np.int64)

    def getActionProbs(self, canonicalBoards, temps):     # mcts.py:175-214
        counts_all = self.rootCounts(canonicalBoards)
        n_env, n_act = counts_all.shape
        probs = np.zeros((n_env, n_act))
        for i in range(n_env):
            counts = counts_all[i]
            temp = temps[i]
            if temp == 0:
                best = np.array(np.argwhere(counts == np.max(counts))).ravel()
                probs[i][np.random.choice(best)] = 1          # same global-RNG draw as the reference (mcts.py:207-208)
            else:
                powered = counts ** (1. / temp)
                with np.errstate(invalid="ignore", divide="ignore"):
                    probs[i] = powered / float(np.sum(powered))   # finished envs give NaN rows, as in the reference
        return probs

"""Two-engine arena on packed positions — the reference's `VectorizedArena`
(/root/reference/src/tacult/base/arena.py:125-305) with the players the coach gives it
(`lambda x: np.argmax(mcts.getActionProbs(x, temps=0), axis=1)`, coach.py:226-232), SURVEY.md §8f-3.

Boards never leave the packed canonical form: per ply the mover's engine gets the live positions as roots
(`tc_set_roots_packed`), searches, returns root visit counts, the host picks the arg-max with the reference's
global-RNG tie-break (mcts.py:206-209, drawn for every env, finished ones included, so the random stream stays
aligned with the reference's), and `tc_rules_step` plays the moves.  Both engines keep their trees for the whole
arena, as the reference's two `VectorizedMCTS` objects do (never reset between games or passes).

`play_games(num)` returns `(oneWon, twoWon, draws)` exactly as `VectorizedArena.playGames` (arena.py:216-305):
num/2 games with player 1 moving first, then num/2 with player 2 first, a pass being `num_envs` games in
lock-step, repeated until enough games have finished.
"""
import time

import numpy as np

from . import _lib
from .engine import rules_legal_mask, rules_step


class PackedArena:
    def __init__(self, engine1, engine2, num_sims, cpuct, tie_break="reference"):
        """tie_break: "reference" draws among equal counts from the global numpy RNG exactly as mcts.py:206-209 does
        (one draw per env and ply, finished envs included); "lowest" takes the lowest index without touching the RNG
        (vectorised; what a production arena wants)."""
        assert engine1.num_envs == engine2.num_envs
        assert tie_break in ("reference", "lowest")
        self.tie_break = tie_break
        self.engines = (engine1, engine2)
        self.num_envs = engine1.num_envs
        self.num_sims = int(num_sims)
        self.cpuct = float(cpuct)
        self.ply_seconds = []            # wall time of every lock-step ply (search + move), for latency percentiles
        self.sims = 0                    # live roots x num_sims, summed over plies

    def _actions(self, engine, states, live):
        engine.set_roots_packed(states, live.astype(np.uint8))
        engine.search(self.num_sims, self.cpuct)
        counts = engine.root_counts().astype(np.int64)
        if self.tie_break == "lowest":
            return counts.argmax(axis=1)
        # same draws as getActionProbs(temps=0) + argmax: `np.random.choice(best)` is one `randint(0, len(best))`, which
        # draws nothing for a single candidate, so only tied rows touch the generator — in env order
        is_top = counts == counts.max(axis=1, keepdims=True)
        actions = is_top.argmax(axis=1)
        for i in np.flatnonzero(is_top.sum(axis=1) > 1):
            best = np.flatnonzero(is_top[i])
            actions[i] = best[np.random.randint(0, len(best))]
        return actions

    def _one_pass(self, first, second):
        """One lock-step pass (arena.py:148-214): yields +1 when `first` wins a game, -1 when `second` does and
        +-1e-4 for a draw."""
        states = np.zeros((self.num_envs, 8), dtype=np.uint32)          # packed initial position
        live = np.ones(self.num_envs, dtype=bool)
        side = 1
        while live.any():
            t0 = time.perf_counter()
            actions = self._actions(first if side == 1 else second, states, live)
            idx = np.flatnonzero(live)
            legal = rules_legal_mask(states[idx])
            assert legal[np.arange(len(idx)), actions[idx]].all(), "arena player chose an illegal action"   # arena.py:188-191
            nxt, status = rules_step(states[idx], actions[idx].astype(np.int8))
            states[idx] = nxt
            self.sims += len(idx) * self.num_sims
            self.ply_seconds.append(time.perf_counter() - t0)
            for k, i in enumerate(idx):                       # env order, as the reference's loop
                if status[k] == _lib.LIVE:
                    continue
                live[i] = False
                value = {_lib.LOST: -1.0, _lib.WON: 1.0, _lib.DRAW: 1e-4}[int(status[k])]   # getGameEnded(b, nextPlayer)
                yield -side * value                           # nextPlayer * value (arena.py:206)
            side = -side

    def play_games(self, num):
        half = int(num / 2)
        one = two = draws = 0
        e1, e2 = self.engines
        for swapped in (False, True):
            first, second = (e2, e1) if swapped else (e1, e2)
            played = 0
            while played < half:
                for result in self._one_pass(first, second):
                    if result == 1:
                        one, two = (one, two + 1) if swapped else (one + 1, two)
                    elif result == -1:
                        one, two = (one + 1, two) if swapped else (one, two + 1)
                    else:
                        draws += 1
                    played += 1
                    if played >= half:
                        break
        return one, two, draws

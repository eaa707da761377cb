"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's lock-step search `VectorizedMCTS`
(/root/reference/src/tacult/base/mcts.py:147-358), written iteratively (an
explicit per-level stack instead of the reference's recursion) but keeping
every arithmetic step, evaluation order and data type of the original:

  * per-env dictionaries keyed by the state string              mcts.py:152-162
  * terminal cache Es, value seen from the side to move         mcts.py:250-258
  * the network is run on EVERY env still descending at a level mcts.py:263,345-358
  * expansion: Ps = f64(pi)*valids, np.sum, renormalise         mcts.py:265-290
  * PUCT in float64, unvisited edges without Q / division,
    strict '>' so the lowest action index wins ties             mcts.py:298-316
  * backup Q=(N*Q+v)/(N+1) with the sign flipping per level     mcts.py:327-342
  * visit counts -> probabilities, temp in {0,1}                mcts.py:175-214

Pinned against the live reference by tests/golden/make_golden.py (root visit
counts identical on seeded positions; fixtures committed under tests/golden/).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
import math

import numpy as np
import torch

EPS = 1e-8  # mcts.py:6


def board_to_obs(board):
    """base/utils.py:7-10"""
    return torch.from_numpy(np.array(board.get_obs()).reshape(4, 9, 9)).float()


class OracleVectorizedMCTS:
    def __init__(self, game, nnet, args):
        self.game = game
        self.nnet = nnet
        self.args = args
        n = args.numEnvs
        self.Qsa = [dict() for _ in range(n)]
        self.Nsa = [dict() for _ in range(n)]
        self.Ns = [dict() for _ in range(n)]
        self.Ps = [dict() for _ in range(n)]
        self.Es = [dict() for _ in range(n)]
        self.Vs = [dict() for _ in range(n)]
        # instrumentation (not in the reference): what SURVEY.md §8(d) asks the kernels to count
        self.stat_nn_rows = 0
        self.stat_leaf_evals = 0
        self.stat_terminal_leaves = 0
        self.stat_sum_depth = 0
        self.stat_sims = 0
        self.stat_transpositions = 0

    def reset(self, i):  # mcts.py:165-172
        for table in (self.Qsa, self.Nsa, self.Ns, self.Ps, self.Es, self.Vs):
            table[i] = dict()

    # -- mcts.py:175-214 ---------------------------------------------------
    def root_counts(self, canonicalBoards):
        n_act = self.game.getActionSize()
        counts = np.zeros((self.args.numEnvs, n_act), dtype=np.int64)
        for i, b in enumerate(canonicalBoards):
            key = self.game.stringRepresentation(b)
            visits = self.Nsa[i]
            for a in range(n_act):
                counts[i, a] = visits.get((key, a), 0)
        return counts

    def getActionProbs(self, canonicalBoards, temps):
        for _ in range(self.args.numMCTSSims):
            self.search(canonicalBoards)
        counts_all = self.root_counts(canonicalBoards)
        probs = np.zeros((self.args.numEnvs, self.game.getActionSize()))
        for i in range(self.args.numEnvs):
            counts = counts_all[i]
            if temps[i] == 0:
                best = np.array(np.argwhere(counts == np.max(counts))).ravel()
                probs[i][np.random.choice(best)] = 1
            else:
                powered = counts ** (1. / temps[i])
                probs[i] = powered / float(np.sum(powered))
        return probs

    # -- mcts.py:217-342 ---------------------------------------------------
    def search(self, canonicalBoards):
        n = self.args.numEnvs
        game = self.game
        done = np.array([b is None for b in canonicalBoards], dtype=bool)
        boards = list(canonicalBoards)
        levels = []            # per level: list of (env, key, action) that must be backed up
        value = np.zeros(n)    # the leaf value of each env, as seen from the leaf's side to move
        leaf_level = np.full(n, -1)
        level = 0
        while True:
            keys = [None if done[i] else game.stringRepresentation(boards[i]) for i in range(n)]
            for i in range(n):
                if done[i]:
                    continue
                s = keys[i]
                if s not in self.Es[i]:
                    self.Es[i][s] = game.getGameEnded(boards[i], 1)
                if self.Es[i][s] != 0:
                    done[i] = True
                    keys[i] = None
                    value[i] = self.Es[i][s]
                    leaf_level[i] = level
                    self.stat_terminal_leaves += 1
            if done.all():
                break

            pis, vs = self._predict_active(boards, done)

            for i in range(n):
                if done[i]:
                    continue
                s = keys[i]
                if s not in self.Ps[i]:
                    valids = game.getValidMoves(boards[i], 1)
                    prior = pis[i] * valids
                    total = np.sum(prior)
                    if total > 0:
                        prior /= total
                    else:
                        prior = prior + valids
                        prior /= np.sum(prior)
                    self.Ps[i][s] = prior
                    self.Vs[i][s] = valids
                    self.Ns[i][s] = 0
                    done[i] = True
                    keys[i] = None
                    value[i] = vs[i]
                    leaf_level[i] = level
                    self.stat_leaf_evals += 1
            if done.all():
                break

            step = []
            nxt = [None] * n
            for i in range(n):
                if done[i]:
                    continue
                s = keys[i]
                valids = self.Vs[i][s]
                prior = self.Ps[i][s]
                visits_s = self.Ns[i][s]
                q_of, n_of = self.Qsa[i], self.Nsa[i]
                best_u, best_a = -float("inf"), -1
                for a in range(game.getActionSize()):
                    if not valids[a]:
                        continue
                    if (s, a) in q_of:
                        u = q_of[(s, a)] + self.args.cpuct * prior[a] * math.sqrt(visits_s) / (1 + n_of[(s, a)])
                    else:
                        u = self.args.cpuct * prior[a] * math.sqrt(visits_s + EPS)
                    if u > best_u:
                        best_u, best_a = u, a
                child, child_player = game.getNextState(boards[i], 1, best_a)
                nxt[i] = game.getCanonicalForm(child, child_player)
                if (s, best_a) not in q_of and game.stringRepresentation(nxt[i]) in self.Ps[i]:
                    self.stat_transpositions += 1      # instrumentation only: first use of an edge into a known node
                step.append((i, s, best_a))
            levels.append(step)
            boards = nxt
            level += 1

        # unwind: the reference returns -_v from every level (mcts.py:261,293,342) and adds
        # it into a zero vector one level up (mcts.py:325), so the edge d levels above the
        # leaf receives 0.0 + -(value one level below).
        self.stat_sims += int(np.sum(leaf_level >= 0))
        self.stat_sum_depth += int(np.sum(leaf_level[leaf_level >= 0]))
        carried = value.copy()
        for step in reversed(levels):
            for (i, s, a) in step:
                carried[i] = 0.0 + (-carried[i])
                v = carried[i]
                edge = (s, a)
                if edge in self.Qsa[i]:
                    self.Qsa[i][edge] = (self.Nsa[i][edge] * self.Qsa[i][edge] + v) / (self.Nsa[i][edge] + 1)
                    self.Nsa[i][edge] += 1
                else:
                    self.Qsa[i][edge] = v
                    self.Nsa[i][edge] = 1
                self.Ns[i][s] += 1
        return -carried

    # -- mcts.py:345-358 ---------------------------------------------------
    def _predict_active(self, boards, done):
        n = self.args.numEnvs
        obs = np.array([board_to_obs(b) if b is not None else torch.zeros(4, 9, 9, dtype=torch.float32)
                        for b in boards])
        active = torch.from_numpy(obs[~done]).float()
        self.stat_nn_rows += int(active.shape[0])
        pi_act, v_act = self.nnet.predict(active)
        pis = np.zeros((n, self.game.getActionSize()))
        pis[~done] = pi_act
        vs = np.zeros(n)
        vs[~done] = np.asarray(v_act).ravel()
        return pis, vs

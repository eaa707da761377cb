"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatements of the two callers either side of the search path, used to check that the drop-in
`tacult_b200.VectorizedMCTS` behaves exactly like the reference's under them (same global-RNG draws, same
per-ply bookkeeping):

  * `arena_play_games`  — `VectorizedArena.playGame/playGames`
                           (/root/reference/src/tacult/base/arena.py:148-305)
  * `selfplay_iteration` — `Coach.prepExecuteEpisode/executeEpisode/generateTrainExamples`
                           (/root/reference/src/tacult/base/coach.py:87-193)

Pinned against the live reference by tests/golden/make_golden.py (same seeds -> same results / examples).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np
import torch


def _arena_one_pass(game, first, second, num_envs):
    """One lock-step pass over `num_envs` boards: `first` moves at odd plies (arena.py:148-214).
    Yields +1 when `first` wins a game, -1 when `second` does, another value for a draw."""
    by_side = {1: first, -1: second}
    side = 1
    boards = [game.getInitBoard() for _ in range(num_envs)]
    over = np.zeros(num_envs, dtype=bool)
    while not over.all():
        canon = [game.getCanonicalForm(b, side) for b in boards]
        actions = by_side[side](canon)
        nxt_side = side
        for i in range(num_envs):
            if over[i]:
                continue
            valids = game.getValidMoves(canon[i], 1)
            assert valids[actions[i]] > 0, f"action {actions[i]} is not valid"      # arena.py:188-191
            boards[i], nxt_side = game.getNextState(boards[i], side, actions[i])
            outcome = game.getGameEnded(boards[i], nxt_side)
            if outcome != 0:
                over[i] = True
                yield nxt_side * outcome
        side = nxt_side


def arena_play_games(game, player1, player2, num, num_envs):
    """(oneWon, twoWon, draws) as VectorizedArena.playGames(num) (arena.py:216-305): num/2 games with player1
    moving first, then num/2 with player2 first; passes repeat until enough games have finished."""
    half = int(num / 2)
    one = two = draws = 0
    for swapped in (False, True):
        first, second = (player2, player1) if swapped else (player1, player2)
        played = 0
        while played < half:
            for result in _arena_one_pass(game, first, second, num_envs):
                winner_is_first = result == 1
                winner_is_second = result == -1
                if winner_is_first:
                    one, two = (one, two + 1) if swapped else (one + 1, two)
                elif winner_is_second:
                    one, two = (one + 1, two) if swapped else (one, two + 1)
                else:
                    draws += 1
                played += 1
                if played >= half:
                    break
    return one, two, draws


def selfplay_iteration(game, mcts, args):
    """The training examples of one iteration, as Coach.generateTrainExamples() returns them
    (coach.py:160-193): list of (obs Tensor[4,9,9], pi Tensor[81], z)."""
    n = args.numEnvs
    pending = [[] for _ in range(n)]
    boards = [game.getInitBoard() for _ in range(n)]
    player = np.ones(n, dtype=int)
    step = np.zeros(n, dtype=int)
    finished = np.zeros(n, dtype=bool)
    out = []
    games = 0
    while games < args.minNumEps:
        step += 1
        temps = np.less(step, args.tempThreshold)                               # coach.py:124-125
        canon = [game.getCanonicalForm(boards[i], player[i]) for i in range(n)]
        pis = mcts.getActionProbs(canon, temps=temps)
        for i in range(n):
            if finished[i]:
                continue
            for b, p in game.getSymmetries(canon[i], pis[i]):                   # coach.py:136-138
                pending[i].append([game._get_obs(b), player[i], torch.from_numpy(p).float()])
            action = np.random.choice(len(pis[i]), p=pis[i])                    # coach.py:140
            boards[i], player[i] = game.getNextState(boards[i], int(player[i]), action)
            r = game.getGameEnded(boards[i], player[i])
            if r != 0:
                out.extend((x[0], x[2], r * ((-1) ** (x[1] != player[i]))) for x in pending[i])   # coach.py:143-152
                pending[i] = []
                finished[i] = True
                games += 1
    return out

"""Generates the golden fixtures under tests/golden/ from the LIVE reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

What is produced and from where:
  mcts_*.npz     root visit counts of the reference's own `VectorizedMCTS`
                 (/root/reference/src/tacult/base/mcts.py:147-358) driven through the
                 reference's `UtacGame` adapter (utac_game.py) with the deterministic
                 hash evaluator of oracle/evaluators.py.
  net_*.npz      (pi, v, logits) of the reference's `_UtacNNet`
                 (/root/reference/src/tacult/network.py:44-102) with weights from
                 oracle/resnet_oracle.make_state_dict(seed) loaded by load_state_dict.
  symmetries.npz the eight images `UtacGame.getSymmetries` (utac_game.py:160-220) returns.
  rules_playouts.npz  playout checksums of the oracle rules engine itself.  The rules
                 engine (utac) is NOT in the reference tree: this fixture only freezes the
                 oracle, it does not pin it to utac (PARITY UNPINNED, oracle/utac_rules.h).
While generating, the script also asserts that the oracle restatements
(oracle/mcts_oracle.py, oracle/game_oracle.py, oracle/resnet_oracle.py) agree with the
live reference on the very same inputs.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_loader  # noqa: E402
from evaluators import HashEvaluator  # noqa: E402
from game_oracle import OracleGame  # noqa: E402
from mcts_oracle import OracleVectorizedMCTS  # noqa: E402
import resnet_oracle  # noqa: E402
from utac_gym.core import GameState, playout_many  # noqa: E402


def random_position(seed, depth):
    rng = np.random.RandomState(seed)
    g = GameState()
    for _ in range(depth):
        legal = np.flatnonzero(g.get_valid_mask())
        nxt = GameState(g)
        nxt.make_move(int(rng.choice(legal)))
        if nxt.is_terminal():
            break
        g = nxt
    if g.current_player() != 1:      # canonical frame, as MCTS and the net always see it
        g = OracleGame().getCanonicalForm(g, -1)
    return g


def gen_mcts_selfplay(ref, name, salt, pi_levels, v_levels, sims, envs, cpuct, seed):
    dotdict = ref.utils.dotdict
    args = dotdict(numEnvs=envs, numMCTSSims=sims, cpuct=cpuct)
    rgame, ogame = ref.utac_game.UtacGame(), OracleGame()
    rm = ref.mcts.VectorizedMCTS(rgame, HashEvaluator(salt, pi_levels, v_levels), args)
    om = OracleVectorizedMCTS(ogame, HashEvaluator(salt, pi_levels, v_levels), args)
    boards = [rgame.getInitBoard() for _ in range(envs)]
    players = np.ones(envs, dtype=int)
    rng = np.random.RandomState(seed)
    roots, counts, alive_rows, actions = [], [], [], []
    while True:
        alive = np.array([rgame.getGameEnded(boards[i], players[i]) == 0 for i in range(envs)])
        if not alive.any():
            break
        canon = [rgame.getCanonicalForm(boards[i], players[i]) for i in range(envs)]
        for _ in range(sims):
            rm.search(canon)
            om.search(canon)
        keys = [rgame.stringRepresentation(c) for c in canon]
        rc = np.array([[rm.Nsa[i].get((keys[i], a), 0) for a in range(81)] for i in range(envs)])
        assert (rc == om.root_counts(canon)).all(), "oracle MCTS diverged from the live reference"
        for i in range(envs):
            assert rm.Qsa[i] == om.Qsa[i] and rm.Ns[i] == om.Ns[i]
        roots.append(np.stack([c.packed() for c in canon]))
        counts.append(rc.astype(np.uint16))
        alive_rows.append(alive)
        act = np.full(envs, -1, dtype=np.int8)
        for i in range(envs):
            if alive[i]:
                p = rc[i].astype(float)
                act[i] = rng.choice(81, p=p / p.sum())
                boards[i], players[i] = rgame.getNextState(boards[i], int(players[i]), int(act[i]))
        actions.append(act)
    np.savez_compressed(os.path.join(HERE, name), roots=np.array(roots), counts=np.array(counts),
                        alive=np.array(alive_rows), actions=np.array(actions), salt=salt, pi_levels=pi_levels,
                        v_levels=v_levels, sims=sims, envs=envs, cpuct=cpuct, kind="selfplay")
    print(name, "plies", len(roots))


def gen_mcts_arena(ref, name, salts, sims, envs, cpuct, seed):
    """Two engines alternate on shared boards (base/arena.py:148-214): each tree only
    sees every other position, so roots are found by key, not by following one edge."""
    dotdict = ref.utils.dotdict
    args = dotdict(numEnvs=envs, numMCTSSims=sims, cpuct=cpuct)
    rgame, ogame = ref.utac_game.UtacGame(), OracleGame()
    rms = [ref.mcts.VectorizedMCTS(rgame, HashEvaluator(s), args) for s in salts]
    oms = [OracleVectorizedMCTS(ogame, HashEvaluator(s), args) for s in salts]
    boards = [rgame.getInitBoard() for _ in range(envs)]
    cur = 1
    rng = np.random.RandomState(seed)
    roots, counts, alive_rows, actions, movers = [], [], [], [], []
    while True:
        alive = np.array([rgame.getGameEnded(b, cur) == 0 for b in boards])
        if not alive.any():
            break
        which = 0 if cur == 1 else 1
        canon = [rgame.getCanonicalForm(b, cur) for b in boards]
        for _ in range(sims):
            rms[which].search(canon)
            oms[which].search(canon)
        keys = [rgame.stringRepresentation(c) for c in canon]
        rc = np.array([[rms[which].Nsa[i].get((keys[i], a), 0) for a in range(81)] for i in range(envs)])
        assert (rc == oms[which].root_counts(canon)).all()
        roots.append(np.stack([c.packed() for c in canon]))
        counts.append(rc.astype(np.uint16))
        alive_rows.append(alive)
        movers.append(which)
        act = np.full(envs, -1, dtype=np.int8)
        for i in range(envs):
            if alive[i]:
                best = np.flatnonzero(rc[i] == rc[i].max())
                legal = np.flatnonzero(canon[i].get_valid_mask())
                best = np.intersect1d(best, legal) if rc[i].max() > 0 else legal
                act[i] = rng.choice(best)
                boards[i], _ = rgame.getNextState(boards[i], cur, int(act[i]))
        actions.append(act)
        cur = -cur
    np.savez_compressed(os.path.join(HERE, name), roots=np.array(roots), counts=np.array(counts),
                        alive=np.array(alive_rows), actions=np.array(actions), movers=np.array(movers),
                        salts=np.array(salts), sims=sims, envs=envs, cpuct=cpuct, kind="arena")
    print(name, "plies", len(roots))


def gen_net(ref, name, seed, channels, blocks, emb, n_pos):
    sd = resnet_oracle.make_state_dict(seed, channels, blocks, emb)
    net = ref.network._UtacNNet(cuda=False, dropout=0.3, channels=channels, num_residual_blocks=blocks,
                                embedding_size=emb)
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    net.eval()
    positions = [random_position(1000 * seed + i, i % 70) for i in range(n_pos)]
    obs = torch.from_numpy(np.stack([np.asarray(p.get_obs()).reshape(4, 9, 9) for p in positions])).float()
    with torch.no_grad():
        pi, v = net(obs)
        # logits are not returned by the reference; recover them from its own layers
        x = torch.relu(net.bn_in(net.conv_in(obs)))
        for blk in net.resnet:
            x = blk(x)
        x = torch.relu(net.fc_bn1(net.fc1(torch.flatten(x, 1))))
        logits = net.pi(x)
    opi, ov, ologits = resnet_oracle.forward(sd, obs, return_logits=True)
    assert torch.allclose(opi, pi, atol=1e-6, rtol=1e-5) and torch.allclose(ov, v, atol=1e-6, rtol=1e-5), \
        "oracle ResNet diverged from the live reference"
    assert torch.allclose(ologits, logits, atol=1e-5, rtol=1e-5)
    np.savez_compressed(os.path.join(HERE, name), states=np.stack([p.packed() for p in positions]),
                        pi=pi.numpy(), v=v.numpy(), logits=logits.numpy(), seed=seed, channels=channels,
                        blocks=blocks, emb=emb)
    print(name, "positions", n_pos, "max|pi-oracle|", float((opi - pi).abs().max()))


def gen_symmetries(ref, name):
    rgame, ogame = ref.utac_game.UtacGame(), OracleGame()
    states, images, perms = [], [], []
    for i in range(12):
        pos = random_position(500 + i, 3 + 5 * i)
        pi = np.arange(81, dtype=np.float64)
        rs = rgame.getSymmetries(pos, pi)
        os_ = ogame.getSymmetries(pos, pi)
        for (rb, rp), (ob, op) in zip(rs, os_):
            assert rgame.stringRepresentation(rb) == ogame.stringRepresentation(ob) and (rp == op).all()
            g1, g2 = rb._get_gs(), ob._get_gs()
            assert (g1.main_board, g1.main_occ, g1.game_occ) == (g2.main_board, g2.main_occ, g2.game_occ)
        states.append(pos.packed())
        images.append(np.stack([b.packed() for b, _ in rs]))
        perms.append(np.stack([p.astype(np.int16) for _, p in rs]))
    np.savez_compressed(os.path.join(HERE, name), states=np.array(states), images=np.array(images),
                        perms=np.array(perms))
    print(name, "positions", len(states))


def gen_rules(name):
    res = playout_many(20251019, 0, 4096)
    np.savez_compressed(os.path.join(HERE, name), seed=20251019, first_game=0,
                        checksum=np.array([r.checksum for r in res], dtype=np.uint64),
                        plies=np.array([r.plies for r in res], dtype=np.int32),
                        outcome=np.array([r.outcome for r in res], dtype=np.int32),
                        final_state=np.array([list(r.final_state) for r in res], dtype=np.uint32))
    print(name, "games", len(res))


def gen_callers(ref, name):
    """Reference VectorizedArena.playGames and Coach.generateTrainExamples under fixed np.random seeds; the
    oracle restatements (oracle/callers_oracle.py) must reproduce them draw for draw."""
    import contextlib
    import hashlib
    import io
    from callers_oracle import arena_play_games, selfplay_iteration
    dotdict = ref.utils.dotdict
    arena_rows = []
    for seed in (0, 1, 2):
        args = dotdict(numEnvs=6, numMCTSSims=3, cpuct=2.4)
        np.random.seed(seed)
        g = ref.utac_game.UtacGame()
        m1, m2 = (ref.mcts.VectorizedMCTS(g, HashEvaluator(k), args) for k in (1, 2))
        ar = ref.arena.VectorizedArena(lambda x: np.argmax(m1.getActionProbs(x, temps=np.zeros(6)), axis=1),
                                       lambda x: np.argmax(m2.getActionProbs(x, temps=np.zeros(6)), axis=1), g, None, 6)
        with contextlib.redirect_stdout(io.StringIO()):
            want = ar.playGames(20)
        np.random.seed(seed)
        og = OracleGame()
        o1, o2 = (OracleVectorizedMCTS(og, HashEvaluator(k), args) for k in (1, 2))
        got = arena_play_games(og, lambda x: np.argmax(o1.getActionProbs(x, temps=np.zeros(6)), axis=1),
                               lambda x: np.argmax(o2.getActionProbs(x, temps=np.zeros(6)), axis=1), 20, 6)
        assert tuple(want) == tuple(got), (want, got)
        arena_rows.append([seed, *want])

    class FakeNet:
        def __init__(self):
            self.h = HashEvaluator(5)

        def predict(self, obs):
            return self.h.predict(obs)
    args = dotdict(numEnvs=4, minNumEps=4, numMCTSSims=6, cpuct=2.4, tempThreshold=5)
    coach = ref.coach.Coach.__new__(ref.coach.Coach)
    coach.game, coach.args, coach.pnet = ref.utac_game.UtacGame(), args, FakeNet()
    np.random.seed(7)
    with contextlib.redirect_stdout(io.StringIO()):
        ref_ex = coach.generateTrainExamples()
    np.random.seed(7)
    og = OracleGame()
    or_ex = selfplay_iteration(og, OracleVectorizedMCTS(og, HashEvaluator(5), args), args)
    assert len(ref_ex) == len(or_ex)
    h = hashlib.sha256()
    for a, b in zip(ref_ex, or_ex):
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and a[2] == b[2]
        h.update(a[0].numpy().tobytes())
        h.update(a[1].numpy().tobytes())
        h.update(np.float64(a[2]).tobytes())
    np.savez_compressed(os.path.join(HERE, name), arena=np.array(arena_rows), coach_examples=len(ref_ex),
                        coach_sha256=np.frombuffer(h.digest(), dtype=np.uint8))
    print(name, "arena", arena_rows, "coach examples", len(ref_ex))


def gen_learner(ref, name, ckpt_name):
    """Three optimisation steps of the LIVE reference learner (NNetWrapper.training_step, nn_wrapper.py:45-56) on a
    tiny net with drop-out active, and the checkpoint it then writes (nn_wrapper.py:138-163)."""
    import shutil
    import tempfile
    torch.manual_seed(11)
    net = ref.network._UtacNNet(cuda=False, dropout=0.2, channels=8, num_residual_blocks=1, embedding_size=32)
    init = {k: v.detach().clone().numpy() for k, v in net.state_dict().items()}
    args = ref.utils.dotdict({"cuda": False, "lr": 1e-3, "epochs": 1, "steps_per_epoch": 3, "numIters": 1,
                              "num_warm_restarts": 1})
    wrapper = ref.nn_wrapper.NNetWrapper(net, args)
    rng = np.random.RandomState(12)
    steps, batch = 3, 16
    obs = (rng.random_sample((steps, batch, 4, 9, 9)) < 0.3).astype(np.float32)
    pis = rng.random_sample((steps, batch, 81)).astype(np.float32) ** 4
    pis /= pis.sum(axis=2, keepdims=True)
    vs = rng.choice(np.array([-1.0, 1.0, 1e-4], dtype=np.float32), size=(steps, batch))
    net.train()
    torch.manual_seed(5)                   # drop-out masks
    losses = [wrapper.training_step(torch.from_numpy(obs[i]), torch.from_numpy(pis[i]), torch.from_numpy(vs[i]))
              for i in range(steps)]
    final = {k: v.detach().clone().numpy() for k, v in net.state_dict().items()}
    net.eval()
    with torch.no_grad():
        epi, ev = net(torch.from_numpy(obs[0]))
    tmp = tempfile.mkdtemp()
    wrapper.save_checkpoint(folder=tmp, filename="ref.pt")
    shutil.copy(os.path.join(tmp, "ref.pt"), os.path.join(HERE, ckpt_name))
    shutil.rmtree(tmp)
    np.savez_compressed(os.path.join(HERE, name), obs=obs, pis=pis, vs=vs, losses=np.array(losses, dtype=np.float64),
                        eval_pi=epi.numpy(), eval_v=ev.numpy(), keys=np.array(list(init.keys())),
                        **{"init/" + k: v for k, v in init.items()}, **{"final/" + k: v for k, v in final.items()})
    print(name, "losses", losses, "checkpoint bytes", os.path.getsize(os.path.join(HERE, ckpt_name)))


def main():
    import logging
    logging.disable(logging.ERROR)   # the zero-prior fallback logs an error per hit (mcts.py:281)
    ref = ref_loader.load_reference()
    torch.manual_seed(0)
    if len(sys.argv) > 1 and sys.argv[1] == "learner":      # regenerate only the learner fixtures
        gen_learner(ref, "learner_steps.npz", "ref_checkpoint_tiny.pt")
        return
    gen_rules("rules_playouts.npz")
    gen_callers(ref, "callers.npz")
    gen_symmetries(ref, "symmetries.npz")
    gen_net(ref, "net_tiny.npz", seed=1, channels=8, blocks=1, emb=32, n_pos=48)
    gen_net(ref, "net_trainer_default.npz", seed=2, channels=32, blocks=3, emb=256, n_pos=70)
    gen_net(ref, "net_class_default.npz", seed=3, channels=16, blocks=4, emb=256, n_pos=70)
    gen_net(ref, "net_wide64.npz", seed=4, channels=64, blocks=2, emb=128, n_pos=32)
    gen_net(ref, "net_deep256x20.npz", seed=5, channels=256, blocks=20, emb=256, n_pos=16)
    gen_mcts_selfplay(ref, "mcts_selfplay_a.npz", salt=7, pi_levels=0, v_levels=0, sims=40, envs=6, cpuct=2.4, seed=3)
    gen_mcts_selfplay(ref, "mcts_selfplay_ties.npz", salt=9, pi_levels=3, v_levels=2, sims=60, envs=4, cpuct=2.4, seed=4)
    gen_mcts_selfplay(ref, "mcts_selfplay_c1.npz", salt=21, pi_levels=0, v_levels=0, sims=25, envs=8, cpuct=1.0, seed=5)
    gen_mcts_arena(ref, "mcts_arena_2sims.npz", salts=(11, 12), sims=2, envs=8, cpuct=2.4, seed=6)
    gen_mcts_arena(ref, "mcts_arena_16sims.npz", salts=(13, 14), sims=16, envs=4, cpuct=2.4, seed=7)
    gen_learner(ref, "learner_steps.npz", "ref_checkpoint_tiny.pt")


if __name__ == "__main__":
    main()

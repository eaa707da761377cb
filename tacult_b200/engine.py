"""Thin object wrapper over the C ABI: one `Engine` == one `tc_engine` handle."""
import ctypes

import numpy as np

from . import _lib
from ._lib import check, ptr


def pack_fields(board, occ, last_square):
    """GAMESTATE members -> packed positions u32[n,8] (layout: include/tacult_b200.h)."""
    board = np.asarray(board, dtype=np.uint32).reshape(-1, 9) & 0x1FF
    occ = np.asarray(occ, dtype=np.uint32).reshape(-1, 9) & 0x1FF
    last = np.asarray(last_square, dtype=np.int64).reshape(-1)
    out = np.zeros((board.shape[0], 8), dtype=np.uint32)
    for g in range(3):
        out[:, g] = board[:, 3 * g] | (board[:, 3 * g + 1] << 9) | (board[:, 3 * g + 2] << 18)
        out[:, 3 + g] = occ[:, 3 * g] | (occ[:, 3 * g + 1] << 9) | (occ[:, 3 * g + 2] << 18)
    out[:, 6] = (last + 1).astype(np.uint32) & 0xF
    return out


def pack_gamestates(boards):
    """list of GameState (canonical frame) -> (packed u32[n,8], active u8[n]); None -> inactive.
    tacult_b200.GameState objects go through one C call (tc_gs_pack); objects of another rules package (the real
    utac_gym) are read field by field through their `_get_gs()`."""
    from .gamestate import pack_states
    fast = pack_states(boards)
    if fast is not None:
        return fast
    n = len(boards)
    b = np.zeros((n, 9), dtype=np.uint32)
    o = np.zeros((n, 9), dtype=np.uint32)
    last = np.full(n, -1, dtype=np.int64)
    active = np.ones(n, dtype=np.uint8)
    for i, gs_owner in enumerate(boards):
        if gs_owner is None:
            active[i] = 0
            continue
        gs = gs_owner._get_gs()
        if int(gs.side) != 1:
            # the arena keeps passing finished boards whose side to move is not the current player
            # (arena.py:176-184); the reference's search only looks up their terminal value and does nothing
            if gs_owner.is_terminal():
                active[i] = 0
                continue
            raise ValueError("VectorizedMCTS expects canonical boards (side == 1), as getCanonicalForm returns")
        b[i] = np.asarray(gs.board, dtype=np.uint32)
        o[i] = np.asarray(gs.occ, dtype=np.uint32)
        last[i] = int(gs.last_square)
    return pack_fields(b, o, last), active


class Engine:
    def __init__(self, num_envs, max_nodes_per_tree, max_edges_per_tree=0, device=0, evaluator=_lib.EVAL_HASH,
                 eval_log_capacity=0, hash_salt=1, hash_pi_levels=0, hash_v_levels=0):
        self.lib = _lib.load()
        self.num_envs = int(num_envs)
        cfg = _lib.Config(self.num_envs, int(max_nodes_per_tree), int(max_edges_per_tree), int(device),
                          int(evaluator), int(eval_log_capacity), int(hash_salt) & (2 ** 64 - 1),
                          int(hash_pi_levels), int(hash_v_levels))
        handle = ctypes.c_void_p()
        check(self.lib.tc_create(ctypes.byref(cfg), ctypes.byref(handle)))
        self.handle = handle
        self.net_desc = None

    def close(self):
        if getattr(self, "handle", None):
            self.lib.tc_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- evaluator -----------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        check(self.lib.tc_set_stream(self.handle, ctypes.c_void_p(cuda_stream_ptr)))

    def sync(self):
        check(self.lib.tc_sync(self.handle))

    def set_evaluator(self, evaluator):
        check(self.lib.tc_set_evaluator(self.handle, int(evaluator)))

    def set_hash_evaluator(self, salt, pi_levels=0, v_levels=0):
        check(self.lib.tc_set_hash_evaluator(self.handle, int(salt) & (2 ** 64 - 1), int(pi_levels), int(v_levels)))
        self.set_evaluator(_lib.EVAL_HASH)

    def load_weights(self, state_dict, plane_order=None):
        """plane_order: see weights.state_dict_to_blob (checkpoints trained on another ordering of the input planes)."""
        from .weights import state_dict_to_blob
        desc, blob = state_dict_to_blob(state_dict, plane_order)
        assert blob.size == self.lib.tc_weight_blob_floats(ctypes.byref(desc))
        check(self.lib.tc_load_weights(self.handle, ctypes.byref(desc), ptr(blob), blob.size))
        self.net_desc = desc

    def forward(self, obs, evaluator=_lib.EVAL_NET_F32, want_logits=False):
        obs = np.ascontiguousarray(obs, dtype=np.float32)
        b = obs.shape[0]
        pi = np.empty((b, 81), dtype=np.float32)
        v = np.empty((b,), dtype=np.float32)
        logits = np.empty((b, 81), dtype=np.float32) if want_logits else None
        check(self.lib.tc_forward(self.handle, int(evaluator), b, ptr(obs), ptr(pi), ptr(v), ptr(logits)))
        return (pi, v, logits) if want_logits else (pi, v)

    def forward_states(self, states, evaluator=_lib.EVAL_NET_F32):
        states = np.ascontiguousarray(states, dtype=np.uint32).reshape(-1, 8)
        b = states.shape[0]
        pi = np.empty((b, 81), dtype=np.float32)
        v = np.empty((b,), dtype=np.float32)
        check(self.lib.tc_forward_states(self.handle, int(evaluator), b, ptr(states), ptr(pi), ptr(v)))
        return pi, v

    def debug_trunk(self, states, n_convs, evaluator=_lib.EVAL_NET_F32):
        states = np.ascontiguousarray(states, dtype=np.uint32).reshape(-1, 8)
        out = np.empty((states.shape[0], 81, self.net_desc.channels), dtype=np.float32)
        check(self.lib.tc_debug_trunk(self.handle, int(evaluator), states.shape[0], ptr(states), int(n_convs), ptr(out)))
        return out

    # -- search --------------------------------------------------------------------------------
    def reset(self, env=-1):
        check(self.lib.tc_reset(self.handle, int(env)))

    def reset_envs(self, envs):
        """`reset(i)` for every i in `envs` with one launch (a ply in which many games end)."""
        envs = np.ascontiguousarray(envs, dtype=np.int32).reshape(-1)
        if envs.size:
            check(self.lib.tc_reset_envs(self.handle, int(envs.size), ptr(envs)))

    def set_roots_packed(self, states, active=None):
        states = np.ascontiguousarray(states, dtype=np.uint32).reshape(self.num_envs, 8)
        if active is not None:
            active = np.ascontiguousarray(active, dtype=np.uint8).reshape(self.num_envs)
        check(self.lib.tc_set_roots_packed(self.handle, ptr(states), ptr(active)))

    def set_roots_fields(self, board, occ, last_square, active=None):
        board = np.ascontiguousarray(board, dtype=np.uint16).reshape(self.num_envs, 9)
        occ = np.ascontiguousarray(occ, dtype=np.uint16).reshape(self.num_envs, 9)
        last = np.ascontiguousarray(last_square, dtype=np.int8).reshape(self.num_envs)
        if active is not None:
            active = np.ascontiguousarray(active, dtype=np.uint8).reshape(self.num_envs)
        check(self.lib.tc_set_roots(self.handle, ptr(board), ptr(occ), ptr(last), ptr(active)))

    def advance(self, actions, restart=False, states_out=None, status_out=None):
        """One ply on the device (tc_advance): every env's root + its action -> the next root, trees kept; finished games
        restart with a fresh tree when `restart`.  Returns (successor positions u32[E,8], status i8[E])."""
        actions = np.ascontiguousarray(actions, dtype=np.int8).reshape(self.num_envs)
        if states_out is None:
            states_out = np.empty((self.num_envs, 8), dtype=np.uint32)
        if status_out is None:
            status_out = np.empty(self.num_envs, dtype=np.int8)
        check(self.lib.tc_advance(self.handle, ptr(actions), int(bool(restart)), ptr(states_out), ptr(status_out)))
        return states_out, status_out

    def search(self, num_sims, cpuct):
        check(self.lib.tc_search(self.handle, int(num_sims), float(cpuct)))

    def root_counts(self, out=None):
        """Root edge visit counts u32[num_envs, 81]; `out` = a caller-owned (e.g. pinned) buffer to fill and return."""
        if out is None:
            out = np.empty((self.num_envs, 81), dtype=np.uint32)
        assert out.dtype == np.uint32 and out.shape == (self.num_envs, 81) and out.flags["C_CONTIGUOUS"]
        check(self.lib.tc_root_counts(self.handle, ptr(out)))
        return out

    def root_q(self):
        out = np.empty((self.num_envs, 81), dtype=np.float64)
        check(self.lib.tc_root_q(self.handle, ptr(out)))
        return out

    def stats(self):
        s = _lib.Stats()
        check(self.lib.tc_get_stats(self.handle, ctypes.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        check(self.lib.tc_reset_stats(self.handle))

    def profile_enable(self, on=True):
        check(self.lib.tc_profile_enable(self.handle, int(bool(on))))

    def profile_read(self):
        """{kernel class: (total device ms, launches)} since the last read."""
        ms = np.zeros(len(_lib.KERNEL_CLASSES), dtype=np.float64)
        n = np.zeros(len(_lib.KERNEL_CLASSES), dtype=np.uint64)
        check(self.lib.tc_profile_read(self.handle, ptr(ms), ptr(n)))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(_lib.KERNEL_CLASSES)}

    def eval_log(self, max_rows):
        states = np.empty((max_rows, 8), dtype=np.uint32)
        pi = np.empty((max_rows, 81), dtype=np.float32)
        v = np.empty((max_rows,), dtype=np.float32)
        n = ctypes.c_int(0)
        check(self.lib.tc_eval_log(self.handle, int(max_rows), ptr(states), ptr(pi), ptr(v), ctypes.byref(n)))
        k = min(n.value, max_rows)
        return states[:k], pi[:k], v[:k], n.value

    # -- device-resident self-play -------------------------------------------------------------
    def selfplay_begin(self, num_sims, cpuct, temp_threshold, seed=0, auto_reset=False, discard_records=False):
        cfg = _lib.SelfPlayConfig(int(num_sims), int(temp_threshold), int(bool(auto_reset)), int(bool(discard_records)),
                                  float(cpuct),
                                  int(seed) & (2 ** 64 - 1))
        check(self.lib.tc_selfplay_begin(self.handle, ctypes.byref(cfg)))

    def selfplay_configure(self, num_sims, cpuct, temp_threshold, seed=0, auto_reset=False, discard_records=False):
        cfg = _lib.SelfPlayConfig(int(num_sims), int(temp_threshold), int(bool(auto_reset)), int(bool(discard_records)),
                                  float(cpuct),
                                  int(seed) & (2 ** 64 - 1))
        check(self.lib.tc_selfplay_configure(self.handle, ctypes.byref(cfg)))

    def selfplay_step(self, plies=1):
        check(self.lib.tc_selfplay_step(self.handle, int(plies)))

    def selfplay_sync(self):
        check(self.lib.tc_selfplay_sync(self.handle))

    def selfplay_stats(self):
        s = _lib.SelfPlayStats()
        check(self.lib.tc_selfplay_get_stats(self.handle, ctypes.byref(s)))
        return s.as_dict()

    def selfplay_drain(self):
        cap = self.num_envs * 256
        buf = np.zeros(cap, dtype=_lib.EXAMPLE_DTYPE)
        n = ctypes.c_int(0)
        check(self.lib.tc_selfplay_drain(self.handle, cap, ptr(buf), ctypes.byref(n)))
        return buf[:n.value].copy()


    def selfplay_drain_dev(self, out_dev_ptr, capacity):
        """Finished-game records -> a caller-owned DEVICE buffer of `capacity` tc_example slots (no host copy of the
        records: only the count crosses).  Returns the number of records written."""
        n = ctypes.c_int(0)
        check(self.lib.tc_selfplay_drain_dev(self.handle, int(capacity), ctypes.c_void_p(int(out_dev_ptr)),
                                             ctypes.byref(n)))
        return n.value

    def examples_from_records_dev(self, n_out, records_dev_ptr, n_records, index_dev_ptr, obs_dev_ptr, pi_dev_ptr,
                                  z_dev_ptr):
        """Training examples (8 symmetries, pi, z) from device-resident records, on the engine's stream."""
        check(self.lib.tc_examples_from_records_dev(
            self.handle, int(n_out), ctypes.c_void_p(int(records_dev_ptr)), int(n_records),
            ctypes.c_void_p(int(index_dev_ptr)) if index_dev_ptr else None, ctypes.c_void_p(int(obs_dev_ptr)),
            ctypes.c_void_p(int(pi_dev_ptr)), ctypes.c_void_p(int(z_dev_ptr))))


def choose_actions(counts, tau_one, u=None, out=None):
    """The actor's move choice for every env in one host call (tc_choose_actions; coach.py:125-140): rows with tau_one
    sample a ~ counts / sum using the caller's uniform draws u[i], the others take the arg-max (lowest index on ties)."""
    lib = _lib.load()
    counts = np.ascontiguousarray(counts, dtype=np.uint32).reshape(-1, 81)
    n = counts.shape[0]
    tau = np.ascontiguousarray(tau_one, dtype=np.uint8).reshape(-1)
    assert tau.shape[0] == n
    if u is not None:
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
        assert u.shape[0] == n
    if out is None:
        out = np.empty(n, dtype=np.int8)
    assert out.shape == (n,) and out.dtype == np.int8
    check(lib.tc_choose_actions(n, ptr(counts), ptr(tau), ptr(u), ptr(out)))
    return out


# -- stateless rules hooks ---------------------------------------------------------------------
def rules_step(states, actions, out=None, status=None):
    """out / status: optional caller-owned (e.g. pinned) result buffers u32[n, 8] / i8[n]."""
    lib = _lib.load()
    states = np.ascontiguousarray(states, dtype=np.uint32).reshape(-1, 8)
    actions = np.ascontiguousarray(actions, dtype=np.int8).reshape(-1)
    if out is None:
        out = np.empty_like(states)
    if status is None:
        status = np.empty(states.shape[0], dtype=np.int8)
    assert out.shape == states.shape and out.dtype == np.uint32 and out.flags["C_CONTIGUOUS"]
    assert status.shape == (states.shape[0],) and status.dtype == np.int8
    check(lib.tc_rules_step(states.shape[0], ptr(states), ptr(actions), ptr(out), ptr(status)))
    return out, status


def rules_legal_mask(states):
    lib = _lib.load()
    states = np.ascontiguousarray(states, dtype=np.uint32).reshape(-1, 8)
    out = np.empty((states.shape[0], 81), dtype=np.uint8)
    check(lib.tc_rules_legal_mask(states.shape[0], ptr(states), ptr(out)))
    return out


def rules_status(states):
    lib = _lib.load()
    states = np.ascontiguousarray(states, dtype=np.uint32).reshape(-1, 8)
    out = np.empty(states.shape[0], dtype=np.int8)
    check(lib.tc_rules_status(states.shape[0], ptr(states), ptr(out)))
    return out


def rules_encode_obs(states):
    lib = _lib.load()
    states = np.ascontiguousarray(states, dtype=np.uint32).reshape(-1, 8)
    out = np.empty((states.shape[0], 4, 9, 9), dtype=np.float32)
    check(lib.tc_rules_encode_obs(states.shape[0], ptr(states), ptr(out)))
    return out


def rules_encode_history(states):
    """8-frame (or any-frame) history planes, the C_in = 4 * frames input of BASELINE.json configs[4]:
    states u32[n, frames, 8] (frame 0 = the current position, each in its own mover's canonical frame) ->
    obs f32[n, 4 * frames, 9, 9], frame f in planes 4 f .. 4 f + 3.  The reference network has 4 planes and no history
    (network.py:55); this is the new-engine variant, encoded by the same device kernel as a single frame."""
    states = np.ascontiguousarray(states, dtype=np.uint32)
    n, frames = states.shape[0], states.shape[1]
    return rules_encode_obs(states.reshape(n * frames, 8)).reshape(n, 4 * frames, 9, 9)


def rules_playouts(seed, first_game, n, max_plies=81):
    lib = _lib.load()
    res = (_lib.PlayoutResult * n)()
    check(lib.tc_rules_playouts(int(seed), int(first_game), int(n), int(max_plies), res))
    return res

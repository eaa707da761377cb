"""ctypes binding of libtacult_b200.so (the C ABI in include/tacult_b200.h).

There is deliberately no CPU fallback: if the CUDA library is missing or fails to load,
importing this module raises.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# TACULT_B200_LIB: another build of the same library (A/B runs of kernel variants); default = the in-tree build
LIB_PATH = os.environ.get("TACULT_B200_LIB") or os.path.join(HERE, "lib", "libtacult_b200.so")

ABI_VERSION = 2
TC_OK, TC_EINVAL, TC_ECUDA, TC_ECAPACITY, TC_ESTATE, TC_EILLEGAL = 0, -1, -2, -3, -4, -5
EVAL_NET_F32, EVAL_NET_BF16, EVAL_HASH, EVAL_NET_BF16_MASKED = 0, 1, 2, 3
LIVE, LOST, DRAW, WON = 0, 1, 2, 3
ACTIONS = 81
STATE_WORDS = 8
KERNEL_CLASSES = ("select", "eval_hash", "nn_stem", "nn_conv", "nn_fc1", "nn_heads", "expand_backup", "other")


class TacultError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"tacult_b200 error {code}: {message}")
        self.code = code


class Config(ctypes.Structure):
    _fields_ = [
        ("num_envs", ctypes.c_int32),
        ("max_nodes_per_tree", ctypes.c_int32),
        ("max_edges_per_tree", ctypes.c_int32),
        ("device", ctypes.c_int32),
        ("evaluator", ctypes.c_int32),
        ("eval_log_capacity", ctypes.c_int32),
        ("hash_salt", ctypes.c_uint64),
        ("hash_pi_levels", ctypes.c_int32),
        ("hash_v_levels", ctypes.c_int32),
    ]


class NetDesc(ctypes.Structure):
    _fields_ = [
        ("channels", ctypes.c_int32),
        ("num_residual_blocks", ctypes.c_int32),
        ("embedding_size", ctypes.c_int32),
        ("in_planes", ctypes.c_int32),
    ]


class Stats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in (
        "sims", "sum_depth", "sum_legal", "nn_leaves", "sum_leaf_legal", "terminal_leaves", "transpositions",
        "nodes_in_use", "edges_in_use", "max_nodes_one_tree", "max_edges_one_tree", "kernel_launches")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class PlayoutResult(ctypes.Structure):
    _fields_ = [
        ("checksum", ctypes.c_uint64),
        ("final_state", ctypes.c_uint32 * 8),
        ("plies", ctypes.c_int32),
        ("outcome", ctypes.c_int32),
    ]


class GameStateStruct(ctypes.Structure):
    """tc_gamestate: the GAMESTATE members (utac_game.py:101-106,176-216,241), 44 bytes."""
    _fields_ = [
        ("board", ctypes.c_uint16 * 9),
        ("occ", ctypes.c_uint16 * 9),
        ("main_board", ctypes.c_uint16),
        ("main_occ", ctypes.c_uint16),
        ("game_occ", ctypes.c_uint16),
        ("side", ctypes.c_int8),
        ("last_square", ctypes.c_int8),
    ]


GAMESTATE_DTYPE = np.dtype([("board", "<u2", (9,)), ("occ", "<u2", (9,)), ("main_board", "<u2"), ("main_occ", "<u2"),
                            ("game_occ", "<u2"), ("side", "i1"), ("last_square", "i1")])
assert GAMESTATE_DTYPE.itemsize == ctypes.sizeof(GameStateStruct) == 44


class SelfPlayConfig(ctypes.Structure):
    _fields_ = [
        ("num_sims", ctypes.c_int32),
        ("temp_threshold", ctypes.c_int32),
        ("auto_reset", ctypes.c_int32),
        ("flags", ctypes.c_int32),
        ("cpuct", ctypes.c_double),
        ("seed", ctypes.c_uint64),
    ]


class SelfPlayStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in (
        "plies", "sims", "games_finished", "first_player_wins", "second_player_wins", "draws", "examples")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


# numpy view of tc_example (216 bytes)
EXAMPLE_DTYPE = np.dtype({
    "names": ["state", "counts", "action", "tau_one", "player", "z_sign", "env", "game", "ply", "finished"],
    "formats": [("<u4", (8,)), ("<u2", (81,)), "i1", "i1", "i1", "i1", "<i4", "<i4", "<i4", "<i4"],
    "offsets": [0, 32, 194, 195, 196, 197, 200, 204, 208, 212],
    "itemsize": 216,
})

# every symbol include/tacult_b200.h declares: name -> (restype, argtypes)
_vp, _i, _u64, _d, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_double, ctypes.c_size_t
_pi = ctypes.POINTER(ctypes.c_int)
_gsp = ctypes.POINTER(GameStateStruct)
SYMBOLS = {
    "tc_abi_version": (_i, []),
    "tc_last_error": (ctypes.c_char_p, []),
    "tc_host_alloc": (_i, [_sz, ctypes.POINTER(_vp)]),
    "tc_host_free": (_i, [_vp]),
    "tc_device_memory": (_i, [_i, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]),
    "tc_gs_init": (None, [_gsp]),
    "tc_gs_rederive": (None, [_gsp]),
    "tc_gs_make_move": (_i, [_gsp, _i]),
    "tc_gs_valid_mask": (_i, [_gsp, _vp]),
    "tc_gs_is_terminal": (_i, [_gsp]),
    "tc_gs_terminal_value": (_i, [_gsp]),
    "tc_gs_obs": (None, [_gsp, _vp]),
    "tc_gs_pack": (_i, [_i, _vp, _vp, _vp]),
    "tc_choose_actions": (_i, [_i, _vp, _vp, _vp, _vp]),
    "tc_rules_step": (_i, [_i, _vp, _vp, _vp, _vp]),
    "tc_rules_legal_mask": (_i, [_i, _vp, _vp]),
    "tc_rules_status": (_i, [_i, _vp, _vp]),
    "tc_rules_encode_obs": (_i, [_i, _vp, _vp]),
    "tc_rules_playouts": (_i, [_u64, _u64, _i, _i, _vp]),
    "tc_create": (_i, [ctypes.POINTER(Config), ctypes.POINTER(_vp)]),
    "tc_destroy": (_i, [_vp]),
    "tc_set_stream": (_i, [_vp, _vp]),
    "tc_sync": (_i, [_vp]),
    "tc_reset": (_i, [_vp, _i]),
    "tc_reset_envs": (_i, [_vp, _i, _vp]),
    "tc_set_hash_evaluator": (_i, [_vp, _u64, _i, _i]),
    "tc_set_evaluator": (_i, [_vp, _i]),
    "tc_set_roots": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "tc_set_roots_packed": (_i, [_vp, _vp, _vp]),
    "tc_advance": (_i, [_vp, _vp, _i, _vp, _vp]),
    "tc_search": (_i, [_vp, _i, _d]),
    "tc_root_counts": (_i, [_vp, _vp]),
    "tc_root_q": (_i, [_vp, _vp]),
    "tc_get_stats": (_i, [_vp, ctypes.POINTER(Stats)]),
    "tc_reset_stats": (_i, [_vp]),
    "tc_eval_log": (_i, [_vp, _i, _vp, _vp, _vp, _pi]),
    "tc_profile_enable": (_i, [_vp, _i]),
    "tc_profile_read": (_i, [_vp, _vp, _vp]),
    "tc_load_weights": (_i, [_vp, ctypes.POINTER(NetDesc), _vp, _sz]),
    "tc_weight_blob_floats": (_sz, [ctypes.POINTER(NetDesc)]),
    "tc_forward": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "tc_forward_states": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "tc_forward_states_dev": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "tc_forward_obs_dev": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "tc_debug_trunk": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "tc_selfplay_begin": (_i, [_vp, ctypes.POINTER(SelfPlayConfig)]),
    "tc_selfplay_configure": (_i, [_vp, ctypes.POINTER(SelfPlayConfig)]),
    "tc_selfplay_step": (_i, [_vp, _i]),
    "tc_selfplay_sync": (_i, [_vp]),
    "tc_selfplay_get_stats": (_i, [_vp, ctypes.POINTER(SelfPlayStats)]),
    "tc_selfplay_drain": (_i, [_vp, _i, _vp, _pi]),
    "tc_selfplay_drain_dev": (_i, [_vp, _i, _vp, _pi]),
    "tc_examples_from_records_dev": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _vp]),
}

_lib = None


def load():
    """Loads the CUDA library (building nothing: run `python -m tacult_b200.build` first)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA extension with `python -m tacult_b200.build` "
            "(there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)       # AttributeError here = the .so does not match the header
        fn.restype = res
        fn.argtypes = args
    if lib.tc_abi_version() != ABI_VERSION:
        raise ImportError(f"ABI mismatch: library reports version {lib.tc_abi_version()}, binding expects {ABI_VERSION}")
    _lib = lib
    return lib


def check(code):
    if code != TC_OK:
        raise TacultError(code, load().tc_last_error().decode("utf-8", "replace"))


class PinnedArray:
    """A numpy array over page-locked host memory (tc_host_alloc); the memory lives as long as this object."""

    def __init__(self, shape, dtype):
        self._lib = load()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._ptr = ctypes.c_void_p()
        check(self._lib.tc_host_alloc(max(n, 1), ctypes.byref(self._ptr)))
        buf = (ctypes.c_uint8 * max(n, 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self.array[...] = 0

    def __del__(self):
        try:
            if getattr(self, "_ptr", None) is not None and self._ptr.value:
                self.array = None
                self._lib.tc_host_free(self._ptr)
                self._ptr = None
        except Exception:
            pass


def ptr(array):
    """Pointer to a C-contiguous numpy array (or None)."""
    if array is None:
        return None
    assert array.flags["C_CONTIGUOUS"]
    return array.ctypes.data

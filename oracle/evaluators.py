"""ORACLE — TEST INFRASTRUCTURE ONLY.

Deterministic stand-in evaluators with the `NNetWrapper.predict` signature
(/root/reference/src/tacult/base/nn_wrapper.py:125-136: Tensor[B,4,9,9] f32 ->
(ndarray[B,81] f32, ndarray[B,1] f32)).

`HashEvaluator` is the CPU side of the "identical evaluator outputs" parity
protocol (BASELINE.json north_star, SURVEY.md §8c row "How parity will be
pinned"): priors and value are an integer hash of what the observation planes
determine (side-to-move stones, all stones, legal-move cells), turned into
float32 values that are exact in binary, so the CUDA engine's built-in hash
evaluator (tacult_b200/csrc/tree.cuh: hash_eval) reproduces them bit for bit
and root visit counts can be compared for equality.

Spec (all arithmetic mod 2^64, mix64 = splitmix64 finaliser as in utac_rules.c):
    h = mix64(salt)
    for w in [mine0, mine1, mine2, occ0, occ1, occ2, legal0, legal1, legal2]:   # u32 words
        h = mix64(h ^ w)
    mineG/occG: sub-boards 3G..3G+2 at bit offsets 0, 9, 18;  legalK: actions 32K..32K+31
    r_a = mix64(h ^ ((a + 1) * 0x9E3779B97F4A7C15))                for a in 0..80
    pi[a] = f32((r_a >> 40) + 1) * 2^-24            if pi_levels == 0
          = f32(r_a % pi_levels) * 0.125            otherwise   (ties and zeros on purpose)
    r_v = mix64(h ^ 0xA5A5A5A5DEADBEEF)
    v   = f32(int(r_v >> 40) - 2^23) * 2^-23        if v_levels == 0
        = f32(int(r_v % (2*v_levels+1)) - v_levels) * 0.25     otherwise (v_levels <= 4)
"""
import numpy as np

MASK64 = (1 << 64) - 1
GOLD = 0x9E3779B97F4A7C15
VKEY = 0xA5A5A5A5DEADBEEF

_R, _C = np.divmod(np.arange(81), 9)
_SUB = (_R // 3) * 3 + _C // 3
_CELL = (_R % 3) * 3 + _C % 3
_WORD = _SUB // 3                       # which packed word holds action a
_SHIFT = (_SUB % 3) * 9 + _CELL         # bit position inside that word


def mix64_np(x):
    x = (x + np.uint64(GOLD)).astype(np.uint64)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def pack_planes(plane):
    """plane: bool[B,81] in action order -> u32[B,3] sub-board-major packed words."""
    out = np.zeros((plane.shape[0], 3), dtype=np.uint64)
    for a in range(81):
        out[:, _WORD[a]] |= plane[:, a].astype(np.uint64) << np.uint64(_SHIFT[a])
    return out


def pack_action_mask(plane):
    """plane: bool[B,81] -> u32[B,3] with action a at bit a%32 of word a//32."""
    out = np.zeros((plane.shape[0], 3), dtype=np.uint64)
    for a in range(81):
        out[:, a >> 5] |= plane[:, a].astype(np.uint64) << np.uint64(a & 31)
    return out


def hash_eval_words(words, salt, pi_levels=0, v_levels=0):
    """words: u64[B,9] (mine0..2, occ0..2, legal0..2) -> (pi f32[B,81], v f32[B,1])."""
    with np.errstate(over="ignore"):
        B = words.shape[0]
        h = mix64_np(np.full(B, salt, dtype=np.uint64))
        for k in range(9):
            h = mix64_np(h ^ words[:, k].astype(np.uint64))
        mult = (np.arange(1, 82, dtype=np.uint64) * np.uint64(GOLD))[None, :]
        r = mix64_np(h[:, None] ^ mult)
        if pi_levels == 0:
            pi = ((r >> np.uint64(40)).astype(np.float64) + 1.0) * 2.0 ** -24
        else:
            pi = (r % np.uint64(pi_levels)).astype(np.float64) * 0.125
        rv = mix64_np(h ^ np.uint64(VKEY))
        if v_levels == 0:
            v = ((rv >> np.uint64(40)).astype(np.int64) - (1 << 23)).astype(np.float64) * 2.0 ** -23
        else:
            v = ((rv % np.uint64(2 * v_levels + 1)).astype(np.int64) - v_levels).astype(np.float64) * 0.25
    return pi.astype(np.float32), v.astype(np.float32).reshape(B, 1)


class HashEvaluator:
    def __init__(self, salt=1, pi_levels=0, v_levels=0):
        self.salt, self.pi_levels, self.v_levels = int(salt) & MASK64, int(pi_levels), int(v_levels)
        self.rows = 0

    def predict(self, obs):
        x = np.asarray(obs, dtype=np.float32).reshape(-1, 4, 81)
        self.rows += x.shape[0]
        mine = x[:, 0] > 0.5
        occ = mine | (x[:, 1] > 0.5)
        legal = x[:, 2] > 0.5
        words = np.concatenate([pack_planes(mine), pack_planes(occ), pack_action_mask(legal)], axis=1)
        return hash_eval_words(words, self.salt, self.pi_levels, self.v_levels)


class ReplayEvaluator:
    """Looks (pi, v) up by observation bytes; used to replay the CUDA engine's own
    recorded leaf evaluations through the CPU search."""

    def __init__(self, table, default_pi=None):
        self.table = table
        self.default_pi = np.full(81, 1.0 / 81, dtype=np.float32) if default_pi is None else default_pi
        self.misses = 0

    @staticmethod
    def key_of(obs_row):
        return np.packbits(np.asarray(obs_row).reshape(-1)[:243] > 0.5).tobytes()

    def predict(self, obs):
        x = np.asarray(obs, dtype=np.float32).reshape(-1, 324)
        pi = np.empty((x.shape[0], 81), dtype=np.float32)
        v = np.empty((x.shape[0], 1), dtype=np.float32)
        for i in range(x.shape[0]):
            hit = self.table.get(self.key_of(x[i]))
            if hit is None:   # the reference also evaluates interior boards and discards the result
                self.misses += 1
                pi[i], v[i, 0] = self.default_pi, 0.0
            else:
                pi[i], v[i, 0] = hit
        return pi, v

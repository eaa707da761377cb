"""tacult_b200 — B200-native engine for tacult's self-play hot path (vectorised MCTS + UTTT rules +
dual-head ResNet forward).  Host side mirrors the reference's Python protocols; the work is CUDA
(sm_100a) behind the C ABI in include/tacult_b200.h.  No CPU fallback exists."""
from . import _lib
from .engine import (Engine, choose_actions, pack_fields, pack_gamestates, rules_encode_history, rules_encode_obs, rules_legal_mask, rules_playouts,
                     rules_status, rules_step)
from .evaluator import B200Evaluator
from .gamestate import GAMESTATE, GameState, install_as_utac_gym
from .mcts import VectorizedMCTS

__all__ = ["Engine", "VectorizedMCTS", "B200Evaluator", "GameState", "GAMESTATE", "install_as_utac_gym", "pack_fields", "pack_gamestates", "rules_step", "choose_actions",
           "rules_legal_mask", "rules_status", "rules_encode_obs", "rules_encode_history", "rules_playouts", "_lib"]

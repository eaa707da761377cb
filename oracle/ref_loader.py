"""ORACLE — TEST INFRASTRUCTURE ONLY.

Imports the *live, unmodified* reference modules from /root/reference (present
only in the build container, never on the GPU box) so that the restatements in
this directory can be validated against them and golden fixtures generated
(tests/golden/make_golden.py).

`import tacult` itself cannot be used: tacult/__init__.py:1 pulls trainer.py,
which needs coloredlogs/onnx and has import-time side effects
(/root/reference/src/tacult/trainer.py:47-58).  Instead empty package shells
are planted in sys.modules and the sub-modules imported individually.  The
reference's `utac_gym` dependency is satisfied by oracle/utac_gym (the C rules
restatement behind the same Python surface).
"""
import importlib
import os
import sys
import types

_ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
# /root/reference in the build container; on the GPU box the byte-for-byte staged copy under oracle/_ref
# (oracle/stage_reference.py, git-ignored)
REFERENCE_ROOT = os.environ.get("TACULT_REFERENCE_ROOT", "/root/reference")
if not os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "tacult", "base", "mcts.py")):
    REFERENCE_ROOT = os.path.join(_ORACLE_DIR, "_ref")
_SRC = os.path.join(REFERENCE_ROOT, "src", "tacult")


def reference_available():
    return os.path.isfile(os.path.join(_SRC, "base", "mcts.py"))


def ensure_utac_gym_on_path():
    if _ORACLE_DIR not in sys.path:
        sys.path.insert(0, _ORACLE_DIR)


def load_reference():
    """Returns a namespace with the reference's hot-path modules."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    ensure_utac_gym_on_path()
    if "tacult" not in sys.modules or not getattr(sys.modules["tacult"], "_oracle_shell", False):
        pkg = types.ModuleType("tacult")
        pkg.__path__ = [_SRC]
        pkg._oracle_shell = True
        base = types.ModuleType("tacult.base")
        base.__path__ = [os.path.join(_SRC, "base")]
        base._oracle_shell = True
        sys.modules["tacult"] = pkg
        sys.modules["tacult.base"] = base
    ns = types.SimpleNamespace()
    ns.root = REFERENCE_ROOT
    ns.mcts = importlib.import_module("tacult.base.mcts")
    ns.utils = importlib.import_module("tacult.utils")
    ns.base_utils = importlib.import_module("tacult.base.utils")
    ns.network = importlib.import_module("tacult.network")
    ns.nn_wrapper = importlib.import_module("tacult.base.nn_wrapper")
    ns.utac_game = importlib.import_module("tacult.utac_game")
    ns.arena = importlib.import_module("tacult.base.arena")
    ns.coach = importlib.import_module("tacult.base.coach")
    ns.utac_nn = importlib.import_module("tacult.utac_nn")
    return ns

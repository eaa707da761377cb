"""ORACLE — TEST / BASELINE INFRASTRUCTURE ONLY.

Stages the UNMODIFIED reference modules of the hot path into oracle/_ref/ (git-ignored, but shipped to the GPU box
with the snapshot) so that `bench.py --impl reference` and the `cpu_baseline` leg can time the reference's own
Python — `Coach.executeEpisode` -> `VectorizedMCTS` -> `UtacGame` -> `_UtacNNet` on torch CPU — where
/root/reference does not exist.  The reference is pure Python, so "building" it is a byte-for-byte copy of the
files below; nothing is edited, and nothing under oracle/_ref/ is ever committed.  The one dependency the reference
cannot bring along, the external `utac_gym` rules engine, is served by oracle/utac_gym (the C restatement).

    python oracle/stage_reference.py          # run by __graft_entry__.build() when /root/reference is present
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = os.environ.get("TACULT_REFERENCE_ROOT", "/root/reference")
DEST = os.path.join(HERE, "_ref")
FILES = [
    "src/tacult/base/mcts.py", "src/tacult/base/utils.py", "src/tacult/base/nn.py", "src/tacult/base/nn_wrapper.py",
    "src/tacult/base/arena.py", "src/tacult/base/coach.py", "src/tacult/base/game.py",
    "src/tacult/utac_game.py", "src/tacult/network.py", "src/tacult/utac_nn.py", "src/tacult/utils.py",
]


def stage():
    if not os.path.isfile(os.path.join(SOURCE, FILES[0])):
        return False
    manifest = {}
    for rel in FILES:
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SOURCE, rel), dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SOURCE, "sha256": manifest}, fh, indent=1)
    with open(os.path.join(DEST, "README.txt"), "w") as fh:
        fh.write("BUILD OUTPUT, NOT REPOSITORY SOURCE (git-ignored).\n"
                 "Byte-for-byte staging of the reference's hot-path modules by oracle/stage_reference.py, run from\n"
                 "__graft_entry__.build() where /root/reference exists, so that `bench.py --impl reference` and the\n"
                 "`cpu_baseline` leg can time the UNMODIFIED reference on the GPU box.  Nothing under tacult_b200/ reads\n"
                 "this directory; tests only use it for the drop-in check (tests/dropin_reference_coach.py).\n")
    return True


if __name__ == "__main__":
    print("staged" if stage() else f"reference tree not found under {SOURCE}")
    sys.exit(0)

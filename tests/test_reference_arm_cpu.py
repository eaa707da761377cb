"""The host-CPU arm of bench.py (`--impl reference`) and the staging recipe behind it (no GPU needed)."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

STAGED = os.path.join(ROOT, "oracle", "_ref")
HAVE_REFERENCE = os.path.isfile("/root/reference/src/tacult/base/mcts.py")
HAVE_STAGED = os.path.isfile(os.path.join(STAGED, "src", "tacult", "base", "mcts.py"))


@pytest.mark.skipif(not (HAVE_REFERENCE and HAVE_STAGED), reason="needs /root/reference and the staged copy")
def test_staged_reference_is_byte_identical():
    """oracle/_ref holds the reference's hot-path modules unmodified: every staged file hashes like its source."""
    manifest = json.load(open(os.path.join(STAGED, "MANIFEST.json")))
    assert len(manifest["sha256"]) >= 10
    for rel, digest in manifest["sha256"].items():
        staged = hashlib.sha256(open(os.path.join(STAGED, rel), "rb").read()).hexdigest()
        source = hashlib.sha256(open(os.path.join("/root/reference", rel), "rb").read()).hexdigest()
        assert staged == digest == source, rel


@pytest.mark.parametrize("root", ["default", "staged-only"])
def test_reference_arm_prints_the_contract_line(root, tmp_path):
    """`bench.py --impl reference` on a tiny sample: one JSON line with the keys the driver reads, timed on the live
    reference modules (kind "reference") when they are reachable — also from the staged copy alone, as on the GPU box."""
    if not (HAVE_REFERENCE or HAVE_STAGED):
        pytest.skip("no reference modules reachable: the arm would fall back to the oracle port")
    if root == "staged-only" and not HAVE_STAGED:
        pytest.skip("no staged copy")
    env = dict(os.environ)
    if root == "staged-only":
        env["TACULT_REFERENCE_ROOT"] = "/nonexistent"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-envs", "4", "--sims", "4"], capture_output=True, text=True, cwd=str(tmp_path), env=env,
                         timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "mcts_sims_per_sec_selfplay" and line["unit"] == "sims/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["steps"] == 1
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["value"] == line["value"] and cb["cores"] >= 1 and "sample" in cb
    assert line["e2e"] == {"value": line["value"], "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["trees_per_gpu"] == 4096 and "workload" in line["config"]

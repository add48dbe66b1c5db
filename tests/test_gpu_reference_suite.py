"""The reference's own hot-path tests, UNCHANGED, against the drop-in package (SURVEY.md section 4 / section 7 gate).

tests/golden/ref_tests/ holds byte-for-byte copies of the reference's conftest.py, test_physics.py, test_obstacles.py
and test_simulation.py (recipe: tests/golden/make_ref_tests.py).  They run in a subprocess -- their modules call
``np.seterr(divide="raise")`` at import time, which must not leak into this session -- with ``billiards`` resolving to
python-billiards_b200/ (tests/conftest.py puts it on sys.path).  SKIP holds the tests that cannot pass by design, each
with its reason; everything else must pass.
"""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = os.path.join(ROOT, "tests", "golden", "ref_tests")

# node id -> why it is not expected to pass against a device-backed drop-in (keep this list short and justified)
SKIP = {}


def test_vendored_files_are_the_reference_files():
    """(CPU) the copies are identical to the reference's files wherever the reference tree is available"""
    src = "/root/reference/tests"
    if not os.path.isdir(src):
        pytest.skip("reference tree not present on this box")
    import filecmp
    for name in ("conftest.py", "test_physics.py", "test_obstacles.py", "test_simulation.py"):
        assert filecmp.cmp(os.path.join(src, name), os.path.join(REF_TESTS, name), shallow=False), name


@pytest.mark.gpu
def test_reference_test_files_pass_unchanged():
    cmd = [sys.executable, "-m", "pytest", REF_TESTS, "-q", "-p", "no:cacheprovider", "-o", "addopts=",
           "--rootdir", ROOT, "-rf"]
    for node in SKIP:
        cmd += ["--deselect", node]
    env = dict(os.environ)
    env["BILLIARDS_COLLECT_REF_TESTS"] = "1"  # tests/conftest.py keeps tests/golden/ out of the normal session
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "python-billiards_b200"), env.get("PYTHONPATH", "")])
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = out.stdout[-6000:] + out.stderr[-2000:]
    assert out.returncode == 0, tail
    m = re.search(r"(\d+) passed", out.stdout)
    assert m and int(m.group(1)) == 22 - len(SKIP), tail
    assert "billiards_b200" in subprocess.run(
        [sys.executable, "-c", "import billiards; print(billiards.__file__)"], cwd=ROOT, env=env, capture_output=True,
        text=True).stdout

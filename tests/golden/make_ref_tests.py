#!/usr/bin/env python3
"""Recipe for tests/golden/ref_tests/: byte-for-byte copies of the reference's OWN hot-path test files

    /root/reference/tests/conftest.py  test_physics.py  test_obstacles.py  test_simulation.py

(22 tests; the two visualiser test files are out of scope).  They are fixtures of category (b):
known-answer tests that travel to the GPU box, where /root/reference does not exist.  tests/test_gpu_reference_suite.py
runs them UNCHANGED in a subprocess with ``billiards`` resolving to the drop-in package (python-billiards_b200/) and
checks that all of them pass.  Run this script in the build container to refresh the copies; it refuses to write
anything that is not identical to the reference's file.
"""
import filecmp
import os
import shutil

SRC = "/root/reference/tests"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_tests")
FILES = ("conftest.py", "test_physics.py", "test_obstacles.py", "test_simulation.py")

if __name__ == "__main__":
    os.makedirs(DST, exist_ok=True)
    for name in FILES:
        shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
        assert filecmp.cmp(os.path.join(SRC, name), os.path.join(DST, name), shallow=False)
        print("copied", name)

"""In the build container (where /root/reference exists) run a few of the REFERENCE's own unit-test files against this package
through scripts/run_reference_tests.py (`import evotorch` resolves to `evotorch_b200`; nothing is copied).  Skipped anywhere
else -- in particular on the GPU box.  The full list and its results: profiles/r01_reference_unit_tests.txt."""

import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = "/root/reference/tests"

pytestmark = pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="the reference checkout is only present in the build container")


@pytest.mark.parametrize("files,min_passed", [
    (["test_hook.py", "test_ranking.py", "test_optimizers.py", "test_read_only_tensor.py"], 22),
    (["test_decorators.py", "test_expects_ndim.py", "test_func_alg.py"], 54),
])
def test_reference_unit_tests_pass_against_this_package(files, min_passed):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "run_reference_tests.py"), *files], capture_output=True, text=True,
                         cwd=ROOT, timeout=300).stdout
    summary = out.strip().splitlines()[-1]
    assert "failed" not in summary and "error" not in summary, out[-3000:]
    assert int(re.search(r"(\d+) passed", summary).group(1)) >= min_passed, summary

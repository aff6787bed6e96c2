"""A short run of the two differential fuzzers (tools/emul_fuzz.py: map updates; tools/emul_fuzz_pass.py: ICP passes, seeded passes,
map_incremental, volumetric maps with hollows) on the CPU build of the library against the oracle -- the long runs are kept under
profiles/r02/emul_fuzz_*.txt; these few scenarios keep the tools themselves alive and catch a regression early."""
import os
import subprocess
import sys

import pytest

import liinit_emul as le

pytestmark = pytest.mark.skipif(not le.available(), reason="g++ or the CUDA vector-type headers are missing")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tool,n,seed", [("emul_fuzz.py", 8, 20260923), ("emul_fuzz_pass.py", 4, 20260923)])
def test_fuzzers_find_no_mismatch(tool, n, seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), str(n), str(seed)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert f"{n} scenarios, no mismatch" in r.stdout

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


def test_results_do_not_depend_on_block_order_or_fresh_memory():
    """The CPU build with the blocks of every launch in a pseudo-random order, the threads of every other block scheduled last-to-first
    (LIINIT_EMUL_SHUFFLE) and every emulated device / pinned allocation filled with 0xCD (LIINIT_EMUL_POISON): the golden fixtures are
    still reproduced bit for bit and a few fuzz scenarios still agree with the oracle."""
    env = dict(os.environ, LIINIT_EMUL_SHUFFLE="11", LIINIT_EMUL_POISON="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_liinit_emul.py"), "-x", "-q", "-k",
                        "golden or coupled or hollow"], capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emul_fuzz.py"), "6", "31337"], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "6 scenarios, no mismatch" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]

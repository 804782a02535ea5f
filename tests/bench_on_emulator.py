"""Child of tests/test_bench_contract.py: runs bench.py's GPU arm against the CPU-emulated C-ABI library with the
test-only torch stand-in (tests/mock_torch).  Environment: MDBG_EMU_LIB.  Arguments are passed on to bench.py."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "mock_torch"))
sys.path.insert(1, ROOT)
from metamdbg_b200 import _capi  # noqa: E402

_capi.LIB_PATH = os.environ["MDBG_EMU_LIB"]
sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[1:]
runpy.run_path(os.path.join(ROOT, "bench.py"), run_name="__main__")

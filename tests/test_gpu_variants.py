"""Every selectable variant of the fused field kernel stays parity-green: the synchronous epilogue exchange (PNR_ASYNC=0), the
other (chunk, tile) pair orders (PNR_ORDER=0/1), the single-CTA predecessor kernel (PNR_PAIR=0) and the profiling instantiation
(PNR_PROF=1).  The knobs are read once per process, hence one subprocess per variant."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [{"PNR_ASYNC": "0"}, {"PNR_ORDER": "0"}, {"PNR_ORDER": "1"}, {"PNR_PAIR": "0"}, {"PNR_PROF": "1"},
                                 {"PNR_ASYNC": "0", "PNR_ORDER": "0"}])
def test_kernel_variant_matches_oracle(env):
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_variant_worker.py")], capture_output=True, text=True,
                       timeout=600, cwd=ROOT, env=e)
    assert r.returncode == 0 and "VARIANT_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]

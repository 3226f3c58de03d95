"""Kernels that compile but have NOT been validated on a device yet.  They are opt-in (environment switches, off by
default) and so are these tests: set AGP_TEST_EXPERIMENTAL=1 to run them.  Each case runs in a subprocess under a
timeout, so a protocol bug (the mbarrier spin limit traps) fails one test instead of the whole session.

* AGP_OZAKI_CLUSTER=2 -- 2-CTA clusters on one row tile, A slices fetched half each and TMA-multicast
  (umma_ozaki_syrk_v2_kernel<S, 2, .>);
* AGP_OZAKI_EPIWARPS=8 -- two epilogue warps per TMEM lane quarter (umma_ozaki_syrk_v2_kernel<S, ., 8>).
Both must be bit-identical to the validated <S, 1, 4> kernel."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("AGP_TEST_EXPERIMENTAL") != "1", reason="opt-in: AGP_TEST_EXPERIMENTAL=1")]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("switches", [["AGP_OZAKI_CLUSTER=2"], ["AGP_OZAKI_EPIWARPS=8"],
                                      ["AGP_OZAKI_CLUSTER=2", "AGP_OZAKI_EPIWARPS=8"]])
@pytest.mark.parametrize("N,K,S", [(128, 128, 7), (1024, 256, 7), (4224, 512, 7), (2176, 512, 6), (8192, 512, 7)])
def test_variant_matches_validated_kernel(N, K, S, switches):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "exp_variant_check.py"), str(N), str(K), str(S)] + switches,
                       capture_output=True, text=True, timeout=180, cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr
    line = [l for l in r.stdout.splitlines() if l.startswith("MAXDIFF")][-1].split()
    assert float(line[1]) == 0.0, line
    assert float(line[3]) > 0.0, line  # the update really happened

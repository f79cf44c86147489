"""Randomised parity check of the kernel families (GPU): tools/fuzz_paths.py for a few seconds -- random radius /
distribution / zoom / N / cell size / seed / content; every pixel-wise path (direct, tiled, staged, auto) and every
grain-wise path (global mask, tile, auto) against the CPU oracle, full renders and a random row band, bit for bit."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [7, 8])
def test_random_parameter_sets_agree_across_kernel_families(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_paths.py"), "10", str(seed)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " 0 failures" in r.stdout

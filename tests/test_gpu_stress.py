"""GPU: randomized parity sweep (tools/stress_parity.py) — random shapes across the small-path / Gram-path / eigensolver
boundaries, all three rank rules, matrices of different character (noise, signal dominated and strongly graded, exactly
low rank, badly scaled). The same sweep with more cases found the ill-conditioned-Gram and odd-m issues fixed in round 1."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [1, 7])
def test_random_shapes_rank_rules_and_spectra(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress_parity.py"), "80", str(seed)], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "80/80 cases passed" in r.stdout

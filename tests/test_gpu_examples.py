"""The example scripts (the reference's examples restated for this package) run end to end on the GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("script,args", [
    ("hmm_2d.py", ["--macro", "8", "--micro", "16"]),
    ("hmm_3d.py", ["--macro", "4", "--micro", "8"]),
    ("diffusion_laminate.py", ["--macro", "8", "--micro", "32"]),
    ("elasticity_rotated_fibres.py", ["--macro", "4", "2", "2", "--micro", "8"]),
    ("general_micro_mesh.py", ["--macro", "6", "--micro", "8"]),  # element-list kernel (not a reference example)
])  # fmt: skip
def test_example_runs(script, args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script), *args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "macro cells" in r.stdout

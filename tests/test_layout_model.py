"""tools/layout_model.py replays the fused kernel's register / shared-memory choreography in
numpy: the index bookkeeping must give the FFT of the dechirped window and every
shared-memory access instruction must be bank-conflict free (SF 7..12)."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_layout_model_is_consistent_and_conflict_free():
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "layout_model.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("bank conflicts: 0") == 6

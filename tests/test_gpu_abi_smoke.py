"""The C ABI driven by a plain-C program (tests/abi_smoke.c): gather, pooled sum, fused ensemble lookup and
index! + update!(Descent) on small Simple/Split tables, every result compared bit for bit inside the program.
No Python or torch in that process: this is the boundary a `ccall` from the reference's host language sees."""
import subprocess

import pytest

from test_abi import build_abi_smoke

pytestmark = pytest.mark.gpu


def test_abi_smoke_runs_without_python(tmp_path):
    exe = build_abi_smoke(str(tmp_path / "abi_smoke"))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout and "FAIL" not in r.stdout

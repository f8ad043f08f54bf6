"""The C++ host mirror of the reference interface (tinychatengine_b200/host: matmul.h, MatmulOperator adapters,
Linear_half_int4 / W8A8* / BMM_S8T_* classes) exercised by a reference-style test binary on the GPU."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


def test_reference_style_op_tests_through_cpp_surface():
    exe = ROOT / "tests" / "cpp" / "test_host"
    if not exe.exists():
        subprocess.run(["make", "-s", "-C", str(ROOT / "tinychatengine_b200" / "host"), "test_host"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("Passed!") >= 10 and "Fail!" not in r.stdout

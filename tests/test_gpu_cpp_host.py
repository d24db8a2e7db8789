"""GPU test of the compiled host mirror: tests/cpp/host_driver.cpp drives gomel_b200/host/gomel.hpp
(the C++ restatement of the Go packages' call sequence over the C ABI) and checks against the oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_host_mirror_against_oracle():
    import __graft_entry__ as g
    g.build()
    exe = os.path.join(ROOT, "tests", "cpp", "host_driver")
    assert os.path.exists(exe), "build() did not produce tests/cpp/host_driver"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0 and "CPP_HOST_OK" in r.stdout, r.stdout + r.stderr

"""-m gpu: the C++ drop-in surface (minerva/op ComputeFn::Execute + minerva/device GpuDevice) driven by a C++
test program, tests/cpp/test_host_plugin.cpp, built by minerva_b200.build.build_host()."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "minerva_b200", "lib", "test_host_plugin")


@pytest.mark.gpu
def test_cpp_plugin_surface():
    if not os.path.exists(BIN):
        from minerva_b200 import build
        build.build()
    out = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    print(out.stdout, out.stderr)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "PASS" in out.stdout

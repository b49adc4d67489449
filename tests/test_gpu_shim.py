"""Runs the C++ drop-in classes (reconstructor_b200/cpp/) the way the reference's orchestrator uses
its plugins -- per-pair virtual calls from 4 threads -- and checks them against the batched loop."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "reconstructor_b200", "cpp")


def test_shim_compiles_with_host_compiler():
    env = dict(os.environ); env.pop("CXX", None)
    subprocess.check_call(["make", "-C", CPP], env=env, stdout=subprocess.DEVNULL)
    assert os.path.exists(os.path.join(CPP, "shim_selftest"))


@pytest.mark.gpu
def test_shim_plugins_equal_batched_loop():
    exe = os.path.join(CPP, "shim_selftest")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", CPP])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "SHIM_OK" in r.stdout

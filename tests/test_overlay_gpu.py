"""GPU: the reference-side binding. oracle/_ref/overlay_driver is a reference-style C++ driver compiled UNCHANGED
(it includes only "CombBLAS/CombBLAS.h") against include/combblas_b200/overlay and linked with libcbgpu.so; its calls to
LocalHybridSpGEMM / MultiwayMerge / PSpGEMM (Mult_AnXBn_Synch) reach the CUDA library and are compared in-process with
the reference's own CPU templates. Built in the build container (needs the reference headers), prebuilt on the GPU box."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "overlay_driver")


def test_unchanged_reference_driver_runs_on_the_device_library():
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/overlay_driver not built (needs /root/reference at build time)")
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr[-2000:])
    assert r.returncode == 0, r.stdout + r.stderr[-2000:]
    assert r.stdout.count("PASS") == 3 and "FAIL" not in r.stdout


def test_driver_with_its_own_semiring_structs_runs_on_the_device_library():
    """tests/overlay/user_semiring_driver.cpp: PSpGEMM<KTipsOrAnd> and LocalHybridSpGEMM<MaxTimesF64> with the driver's own
    structs (device instantiation: tests/user_semiring/libmy_semirings.so) against the same structs on the reference's CPU path"""
    exe = os.path.join(ROOT, "oracle", "_ref", "user_semiring_driver")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/user_semiring_driver not built (needs /root/reference at build time)")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr[-2000:])
    assert r.returncode == 0, r.stdout + r.stderr[-2000:]
    assert r.stdout.count("PASS") == 2 and "FAIL" not in r.stdout

"""GPU: the reference-side binding. oracle/_ref/overlay_driver is a reference-style C++ driver compiled UNCHANGED
(it includes only "CombBLAS/CombBLAS.h") against include/combblas_b200/overlay and linked with libcbgpu.so; its calls to
LocalHybridSpGEMM / MultiwayMerge / PSpGEMM (Mult_AnXBn_Synch) reach the CUDA library and are compared in-process with
the reference's own CPU templates. Built in the build container (needs the reference headers), prebuilt on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "overlay_driver")


@pytest.mark.gpu
def test_unchanged_reference_driver_runs_on_the_device_library():
    """local seams (LocalHybridSpGEMM, LocalSpGEMMHash, MultiwayMerge) and the distributed seams (PSpGEMM -> Mult_AnXBn_Synch,
    MemEfficientSpGEMM, Mult_AnXBn_SUMMA3D, MemEfficientSpGEMM3D on SpParMat / SpParMat3D operands, one rank, NCCL
    communicators created from an id shipped with the driver's MPI_Bcast) against the reference's own templates"""
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/overlay_driver not built (needs /root/reference at build time)")
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr[-2000:])
    assert r.returncode == 0, r.stdout + r.stderr[-2000:]
    assert r.stdout.count("PASS") == 7 and "FAIL" not in r.stdout


@pytest.mark.gpu
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


@pytest.mark.parametrize("exe,count", [("overlay_driver", 7), ("user_semiring_driver", 2)])
def test_run_time_switch_hands_every_call_back_to_the_reference(exe, count):
    """CBGPU_DISABLE=1: the overloads forward to the reference's own templates (through a semiring type the overlay does not
    know), no context is created and no GPU is needed -- the same binaries pass on the CPU-only build container"""
    path = os.path.join(ROOT, "oracle", "_ref", exe)
    if not os.path.exists(path):
        pytest.skip(f"oracle/_ref/{exe} not built (needs /root/reference at build time)")
    env = dict(os.environ, CBGPU_DISABLE="1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([path], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout + r.stderr[-2000:]
    assert r.stdout.count("PASS") == count and "FAIL" not in r.stdout

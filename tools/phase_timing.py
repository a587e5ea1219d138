"""Where the cycles of the shared-accumulator numeric kernels go (tuning, no bench claims). Needs the instrumented build:
    make -C combblas_b200/csrc OBJ=$PWD/combblas_b200/csrc/build_timing OUT=$PWD/combblas_b200/libcbgpu_timing.so EXTRA=-DCBGPU_PHASE_TIMING
    CBGPU_LIB=$PWD/combblas_b200/libcbgpu_timing.so python tools/phase_timing.py --scale 20
Prints, per CTA shape, the share of thread-0 cycles per phase and the average cycles per task."""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import combblas_b200 as cb

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=20)
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
lib = cb.load_library()
ctx = cb.Context(0)
for o in a.opt:
    k, v = o.split("="); ctx.set_option(k, int(v))
G = ctx.gen_rmat(a.scale, 16 << a.scale, 1, 0.57, 0.19, 0.19, True, cb.F64, 0)
f_sym, nnz_sym = ctx.symbolic(G, G)
phases = max(1, int(np.ceil(nnz_sym * 12 / 48e9)))
slabs = ctx.colsplit(G, phases) if phases > 1 else [G]
buf = (C.c_ulonglong * 24)()
names = ["slot load", "stage+mark", "scan+rank", "unpack rows", "store rows+init", "accumulate walk", "store values"]
for rep in range(2):
    lib.cbgpu_debug_phase_cycles(buf)  # clear
    kms = {}
    for Bs in slabs:
        D, st = ctx.spgemm(0, G, Bs, want_stats=True)
        for k, v in st.as_dict().get("ms_kernel", {}).items():
            kms[k] = round(kms.get(k, 0.0) + v, 2)
        D.free()
    lib.cbgpu_debug_phase_cycles(buf)
    v = np.array(list(buf), dtype=np.float64).reshape(3, 8)
    print(f"rep {rep}: kernel ms {kms}")
    for shape, label in enumerate(["1024 x 1 (large)", "512 x 2 (medium)", "256 x 4 (small)"]):
        tasks = v[shape, 7]
        if tasks == 0:
            continue
        tot = v[shape, :7].sum()
        print(f"  {label}: tasks {int(tasks)}, cycles/task {tot / tasks:.0f}: " + ", ".join(f"{n} {100 * c / tot:.1f}% ({c / tasks:.0f})" for n, c in zip(names, v[shape, :7])))

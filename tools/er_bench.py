"""BASELINE config 2: Erdos-Renyi n = 2^scale, d = 8, A bool, B int64 (1 + row id), SelectMaxSRing<bool,int64_t>."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import combblas_b200 as cb
ap = argparse.ArgumentParser(); ap.add_argument("--scale", type=int, default=22); ap.add_argument("--d", type=int, default=8)
ap.add_argument("--reps", type=int, default=4); ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
ctx = cb.Context(0)
for o in a.opt:
    k, v = o.split("="); ctx.set_option(k, int(v))
dA = ctx.gen_rmat(a.scale, a.d << a.scale, seed=2, a=0.25, b=0.25, c=0.25, scramble=False, dtype=cb.BOOL, value_mode=1)
dB = ctx.gen_rmat(a.scale, a.d << a.scale, seed=2, a=0.25, b=0.25, c=0.25, scramble=False, dtype=cb.I64, value_mode=2)
for i in range(a.reps):
    C, st = ctx.spgemm(cb.SelectMaxSRing_bool_i64, dA, dB, want_stats=True)
    d = st.as_dict()
    balg = (dA.nnz * 5 + dB.nnz * 12 + (dA.nzc + dB.nzc) * 16 + C.nnz * 12 + C.nzc * 16)
    print(f"ER s{a.scale} d{a.d}: products {st.flops} nnzC {st.nnz_out} ms {st.ms_total:.3f} (setup {st.ms_setup:.3f} sym {st.ms_symbolic:.3f} num {st.ms_numeric:.3f}) "
          f"GFLOP/s {2*st.flops/st.ms_total/1e6:.1f} roofline {balg/st.ms_total/1e6/6451.2:.4f} kernels {d['ms_kernel']} tasks {st.tasks}", flush=True)
    C.free()

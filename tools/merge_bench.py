"""Times cbgpu_merge (streaming 2-way merge) on two partial products of an R-MAT A^2, reports GB/s of algorithmic traffic."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import combblas_b200 as cb
ap = argparse.ArgumentParser(); ap.add_argument("--scale", type=int, default=18); ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
ctx = cb.Context(0)
G = ctx.gen_rmat(a.scale, 16 << a.scale, 1)
n = 1 << a.scale
# two SUMMA-like partials: A(:, first half) * B(first half, :) and the second halves
A0, A1 = ctx.submatrix(G, 0, n, 0, n // 2), ctx.submatrix(G, 0, n, n // 2, n)
B0, B1 = ctx.submatrix(G, 0, n // 2, 0, n), ctx.submatrix(G, n // 2, n, 0, n)
P0, P1 = ctx.spgemm(0, A0, B0), ctx.spgemm(0, A1, B1)
for i in range(2 * a.reps):
    tma = 1 if i < a.reps else 0
    ctx.set_option("merge_tma", tma)
    M, st = ctx.merge(0, [P0, P1], want_stats=True)
    nin, nout = P0.nnz + P1.nnz, M.nnz
    bytes_alg = nin * 12 + nout * 12
    print(f"merge2 tma={tma}: in {nin} out {nout} ms {st.ms_total:.3f} (setup {st.ms_setup:.3f} count {st.ms_symbolic:.3f} write {st.ms_numeric:.3f}) "
          f"algorithmic {bytes_alg/1e9:.2f} GB -> {bytes_alg/st.ms_total/1e6:.1f} GB/s", flush=True)
    M.free()

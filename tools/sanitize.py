"""Small multi-path workload for compute-sanitizer (memcheck / racecheck):
     compute-sanitizer --tool memcheck  python tools/sanitize.py
     compute-sanitizer --tool racecheck python tools/sanitize.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import combblas_b200 as cb
from tests.util import rmat, to_dcsc

ctx = cb.Context(0)
A = rmat(10, 16, seed=1)
for w in (17, 10):
    ctx.set_option("bitmap_window_log2", w)
    for fp in (0, 1, 2):
        ctx.set_option("force_path", fp)
        t = cb.LocalHybridSpGEMM(ctx, 0, to_dcsc(A, np.float64), to_dcsc(A, np.float64))
parts = [ctx.upload(to_dcsc(rmat(10, 4, seed=s), np.float64)) for s in (2, 3, 4)]
m = ctx.merge(0, parts)
ctx.set_option("merge_engine", 1)
m2 = ctx.merge(0, parts)
print("sanitizer run ok", t.getnnz(), m.nnz, m2.nnz)

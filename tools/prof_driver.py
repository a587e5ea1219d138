"""Tiny driver for ncu: generates R-MAT at --scale on the device and runs --reps multiplies (no timing claims)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import combblas_b200 as cb

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=16)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--opt", action="append", default=[], help="name=value library option")
ap.add_argument("--phases", type=int, default=1, help="column slabs of B; only --slab is multiplied (a bench.py step is all of them)")
ap.add_argument("--slab", type=int, default=0)
a = ap.parse_args()
ctx = cb.Context(0)
for o in a.opt:
    k, v = o.split("=")
    ctx.set_option(k, int(v))
G = ctx.gen_rmat(a.scale, 16 << a.scale, 1, 0.57, 0.19, 0.19, True, cb.F64, 0)
B = ctx.colsplit(G, a.phases)[a.slab] if a.phases > 1 else G
for i in range(a.reps):
    C, st = ctx.spgemm(0, G, B, want_stats=True)
    print(st.as_dict())
    C.free()

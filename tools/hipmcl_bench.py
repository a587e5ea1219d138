"""HipMCL-style expansion step on one GPU (BASELINE config 4 at N=1; tuning/measurement tool, not bench.py):
column-stochastic R-MAT (weights = normalised edge multiplicities), C = A^2 by column slabs, every slab pruned on the
device by MCLPruneRecoverySelect (ParFriends.h:186-354; defaults of MCL.cpp:147-158) before the next one is multiplied,
pruned slabs concatenated and re-normalised. Prints one line per repetition."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import combblas_b200 as cb

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=20)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--hard", type=float, default=1e-4)
ap.add_argument("--select", type=int, default=1100)
ap.add_argument("--recover", type=int, default=1400)
ap.add_argument("--pct", type=float, default=0.9)
ap.add_argument("--slab-gb", type=float, default=48.0)
a = ap.parse_args()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = cb.Context(0, stream=stream.cuda_stream)
G = ctx.gen_rmat(a.scale, 16 << a.scale, 3, 0.57, 0.19, 0.19, True, cb.F64, 0)
ctx.make_col_stochastic(G)
f_sym, nnz_sym = ctx.symbolic(G, G)
phases = max(1, int(np.ceil(nnz_sym * 12 / (a.slab_gb * 1e9))))
slabs = ctx.colsplit(G, phases) if phases > 1 else [G]
for rep in range(a.reps + 1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    pruned, ms_mult, ms_prune, nnz_c = [], 0.0, 0.0, 0
    for Bs in slabs:
        Cs, st = ctx.spgemm(cb.PlusTimesSRing_f64, G, Bs, want_stats=True)
        P, ps = ctx.mcl_prune(Cs, a.hard, a.select, a.recover, a.pct, want_stats=True)
        Cs.free()
        pruned.append(P)
        ms_mult += st.ms_total
        ms_prune += ps.ms
        nnz_c += ps.nnz_in
    C = ctx.colconcat(pruned) if len(pruned) > 1 else pruned[0]
    ctx.make_col_stochastic(C)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    inf = C.info()
    print(f"s{a.scale} expansion rep {rep}: {ms:.1f} ms  ({2 * f_sym / ms / 1e6:.1f} GFLOP/s of the product; multiply {ms_mult:.1f} ms, "
          f"prune {ms_prune:.1f} ms = {nnz_c * 12 / max(ms_prune, 1e-9) / 1e6:.0f} GB/s over the unpruned slabs)  products {f_sym:.3e}, "
          f"nnz(C) {nnz_c:.3e} -> {inf.nnz:.3e} kept, {phases} slabs", flush=True)
    if len(pruned) > 1:
        C.free()
    for p in pruned:
        p.free()

"""Option sweep for tuning (no bench claims): one R-MAT per --scale, every --set 'a=1,b=2' timed with CUDA events.
Prints one line per (scale, option set): GFLOP/s, ms per multiply (all column slabs), per-kernel-class ms."""
import argparse, os, sys, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import combblas_b200 as cb

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, action="append", default=[])
ap.add_argument("--set", action="append", default=[], help="comma separated name=value list ('' = defaults)")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--sr", type=int, default=0)
ap.add_argument("--streams", type=int, action="append", default=[], help="contexts/streams the slabs are spread over")
ap.add_argument("--phases", type=int, default=0, help="column slabs (0 = as bench.py: 48 GB of C per slab)")
a = ap.parse_args()
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = cb.Context(0, stream=stream.cuda_stream)
defaults = {}
for scale in a.scale or [20]:
    G = ctx.gen_rmat(scale, 16 << scale, 1, 0.57, 0.19, 0.19, True, cb.F64, 0)
    f_sym, nnz_sym = ctx.symbolic(G, G)
    phases = a.phases if a.phases > 0 else max(1, int(np.ceil(nnz_sym * 12 / 48e9)))
    slabs = ctx.colsplit(G, phases) if phases > 1 else [G]
    for oset in a.set or [""]:
      for nstreams in a.streams or [1]:
        opts = dict(kv.split("=") for kv in oset.split(",") if kv)
        for k, v in opts.items():
            defaults.setdefault(k, ctx.get_option(k))
            ctx.set_option(k, int(v))
        pipe = cb.SlabPipeline(ctx, nstreams) if nstreams > 1 else None
        times, kms = [], {}
        for rep in range(a.reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            kms = {}
            lock = threading.Lock()

            def consume(i, C, st):
                with lock:
                    for k, v in st.as_dict().get("ms_kernel", {}).items():
                        kms[k] = round(kms.get(k, 0.0) + v, 2)
                C.free()

            if pipe:
                pipe.run(a.sr, G, slabs, consume)
            else:
                for i, Bs in enumerate(slabs):
                    C, st = ctx.spgemm(a.sr, G, Bs, want_stats=True)
                    consume(i, C, st)
            e1.record(stream)
            torch.cuda.synchronize()
            if rep > 0:
                times.append(e0.elapsed_time(e1))
        ms = min(times) if times else float('nan')
        print(f"s{scale} [{oset or 'defaults'}] streams={nstreams} {2 * f_sym / ms / 1e6:.1f} GFLOP/s  ms={[round(t, 1) for t in times]} slabs={phases} {kms}", flush=True)
        if pipe:
            for c in pipe.ctxs[1:]:
                c.close()
        for k, v in defaults.items():
            ctx.set_option(k, v)
    if phases > 1:
        for s in slabs:
            s.free()
    G.free()

"""Makes tests/golden/bench_checksums.json: the order-independent checksums of the benchmark product C = A (x) A (seeded R-MAT,
PlusTimes<double>) computed on ONE GPU, slab by slab, each pinned against the CPU checker:
  * scale <= --full-max: the whole product is recomputed by the oracle (compiled reference if present) and must give the same
    nnz / pattern / value sums;
  * larger scales: seeded column ranges C(:, J) are recomputed by the reference and compared (the rest of the product comes
    from the same kernels on the same kind of columns).
bench.py compares every run -- any number of GPUs -- with these values (parity.equals_single_gpu_golden).
Run on a GPU box: python tools/make_bench_golden.py --scales 14 16 18 20 22"""
import argparse, json, os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import combblas_b200 as cb  # noqa: E402
from oracle.oracle import matrix_checksum  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scales", type=int, nargs="+", default=[14, 16, 18, 20, 22])
ap.add_argument("--full-max", type=int, default=16)
ap.add_argument("--out", default=bench.GOLDEN)
a = ap.parse_args()
ctx = cb.Context(0)
M64 = (1 << 64) - 1
gold = json.load(open(a.out)) if os.path.exists(a.out) else {}
for scale in a.scales:
    n = 1 << scale
    G = ctx.gen_rmat(scale, bench.EDGEFACTOR << scale, bench.SEED, bench.A_, bench.B_, bench.C_, True, cb.F64, 0)
    if scale <= 22:
        flops, nnz_sym = ctx.symbolic(G, G)
        phases = max(1, int(np.ceil(nnz_sym * 12 / 48e9)))
    else:  # too many (column, window) tasks for one symbolic call: slabs sized from the growth of nnz(C) (7.6x per two scales)
        flops = 0
        phases = int(np.ceil(7.2e10 * 7.6 ** ((scale - 22) / 2.0) * 12 / 36e9))
    per = n // phases
    slabs = ctx.colsplit(G, phases) if phases > 1 else [G]
    nnz, ps, vs = 0, 0, 0
    for i, Bs in enumerate(slabs):
        Cs, st = ctx.spgemm(0, G, Bs, want_stats=True)
        if scale > 22:
            flops += int(st.flops)
        nnz += Cs.info().nnz
        p_, v_ = ctx.checksum(Cs, 0, per * i)
        ps, vs = (ps + p_) & M64, (vs + v_) & M64
        Cs.free()
    ref = bench.CpuReference(scale)
    pinned = None
    if scale <= a.full_max:
        out, _ = ref.multiply(0, n)
        want = (out.nnz,) + matrix_checksum(out.rows, out.cols_expanded(), out.vals)
        assert (nnz, ps, vs) == tuple(int(x) for x in want), f"scale {scale}: device {(nnz, ps, vs)} vs checker {want}"
        pinned = f"whole product recomputed by the {ref.kind} checker: equal"
    else:
        checked = 0
        for (c0, c1) in bench.sample_ranges(n, 3):
            out, _ = ref.multiply(c0, c1)
            want = (out.nnz,) + matrix_checksum(out.rows, out.cols_expanded(), out.vals, 0, c0)
            Bs = ctx.colslice(G, c0, c1)
            Cs = ctx.spgemm(0, G, Bs)
            got = (Cs.info().nnz,) + tuple(ctx.checksum(Cs, 0, c0))
            Cs.free(); Bs.free()
            assert tuple(int(x) for x in got) == tuple(int(x) for x in want), f"scale {scale} columns {c0}:{c1} differ"
            checked += 1
        pinned = f"{checked} seeded column ranges of {bench.SAMPLE_COLS} columns recomputed by the {ref.kind} checker: equal"
    gold[bench.golden_key(scale)] = {"products": int(flops), "nnz_C": int(nnz), "pattern_sum": f"{ps:016x}", "value_sum": f"{vs:016x}",
                                     "phases": phases, "source": "tools/make_bench_golden.py on one B200; " + pinned}
    print(scale, gold[bench.golden_key(scale)], flush=True)
    for s in slabs:
        if s is not G:
            s.free()
    G.free()
json.dump(gold, open(a.out, "w"), indent=1, sort_keys=True)

"""First-contact GPU check: runs a ladder of cases without stopping at the first failure and prints one line each."""
import sys, time, traceback
import numpy as np, scipy.sparse as sp
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import combblas_b200 as cb
from oracle.oracle import Csc, SR_DTYPES, best_oracle, PortOracle
from tests.util import assert_same, random_pair, rmat, to_csc, to_dcsc, typed

ctx = cb.Context(0)
orc = PortOracle()
def run(name, fn):
    t = time.time()
    try:
        r = fn()
        print(f"PASS {name} ({time.time()-t:.2f}s) {r if r is not None else ''}", flush=True)
    except Exception as e:
        print(f"FAIL {name}: {type(e).__name__}: {str(e)[:300]}", flush=True)
        traceback.print_exc(limit=2)

def case(sr, A, B, **opts):
    for k, v in opts.items(): ctx.set_option(k, v)
    try:
        dA, dB = ctx.upload(to_dcsc(A, SR_DTYPES[sr][0])), ctx.upload(to_dcsc(B, SR_DTYPES[sr][1]))
        D, st = ctx.spgemm(sr, dA, dB, want_stats=True)
        rows, cols, vals = ctx.download_coo(D)
        got = cb.SpTuples(A.shape[0], B.shape[1], rows, cols, vals)
        want = orc.spgemm(to_csc(A, SR_DTYPES[sr][0]), to_csc(B, SR_DTYPES[sr][1]), sr)
        assert_same(got, want, sr)
        return f"flops={st.flops} nnz={st.nnz_out} ms={st.ms_total:.3f} (sym {st.ms_symbolic:.3f} num {st.ms_numeric:.3f}) hw={st.tasks_hash_warp} hc={st.tasks_hash_cta} bs={st.tasks_bitmap_smem} bg={st.tasks_bitmap_gmem}"
    finally:
        ctx.set_option("force_path", 0); ctx.set_option("bitmap_window_log2", 17); ctx.set_option("shared_acc", 1)

for sr in range(9):
    A, B = random_pair(300, 220, 260, 0.05, 0.04, 11 + sr, SR_DTYPES[sr])
    run(f"random sr{sr}", lambda: case(sr, A, B))
for scale in (8, 11, 13, 14, 16):
    A = rmat(scale, 16, seed=1)
    run(f"rmat s{scale} auto", lambda: case(0, A, A))
    run(f"rmat s{scale} hash-only", lambda: case(0, A, A, force_path=1))
    run(f"rmat s{scale} bitmap-only", lambda: case(0, A, A, force_path=2))
    run(f"rmat s{scale} bitmap gmem", lambda: case(0, A, A, force_path=2, shared_acc=0))
    run(f"rmat s{scale} windows 2^11", lambda: case(0, A, A, bitmap_window_log2=11))

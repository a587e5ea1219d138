"""BASELINE config 4: HipMCL-style expansion (A^2 with column pruning / top-k) on a synthetic weighted R-MAT, N GPUs.

  torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/hipmcl_dist.py --scale 20

Input as MCL.cpp builds it (:389-394, :540-560): R-MAT pattern, symmetrised, seeded weights, self loops, columns made
stochastic. One expansion = MemEfficientSpGEMM[3D](A, A, phases=auto, 1e-4, 1100, 1400, 0.9) (defaults MCL.cpp:147-158)
through cbgpu_memefficient_spgemm_dist: phased SUMMA, every piece of C pruned over the whole distributed columns before
the next slab is multiplied, the pruned block of C left resident; the e2e figure also brings the pruned block to the host.

Two runs:
  parity : dyadic weights k/256, no normalisation -- every product and column sum is exact in double precision, so the
           pruned result is grid independent BIT FOR BIT; the N-rank result (checksums at global positions, summed over the
           ranks) must equal the one-GPU result of cbgpu_memefficient_spgemm, which tests/test_prune_gpu.py pins against
           the reference's own MemEfficientSpGEMM.
  timing : uniform (0,1] weights, column stochastic; GFLOP/s of the expansion (2 * products / time), device time max over ranks.
Prints one JSON line on rank 0.
"""
import argparse, json, os, sys, time

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import combblas_b200 as cb  # noqa: E402
from combblas_b200 import lib as cblib  # noqa: E402
from combblas_b200.host import local_range  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=20)
ap.add_argument("--edgefactor", type=int, default=16)
ap.add_argument("--seed", type=int, default=3)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--hard", type=float, default=1e-4)
ap.add_argument("--select", type=int, default=1100)
ap.add_argument("--recover", type=int, default=1400)
ap.add_argument("--pct", type=float, default=0.9)
ap.add_argument("--phases", type=int, default=0)
a = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = cb.Context(local_rank, stream=stream.cuda_stream)
layers = {1: 1, 2: 2, 4: 1, 8: 2}[world]
M64 = (1 << 64) - 1


def mix(x):
    x = x.astype(np.uint64)
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xff51afd7ed558ccd)
        x ^= x >> np.uint64(33)
        x *= np.uint64(0xc4ceb9fe1a85ec53)
        x ^= x >> np.uint64(33)
    return x


def pattern():
    """symmetrised R-MAT pattern + self loops (every rank builds the same global matrix on its host cores)"""
    lib = cb.load_library()
    ne = a.edgefactor << a.scale
    rows = np.empty(ne, np.int64)
    cols = np.empty(ne, np.int64)
    assert lib.cbgpu_rmat_edges_host(a.scale, ne, a.seed, 0.57, 0.19, 0.19, 1, rows.ctypes.data, cols.ctypes.data) == 0
    n = 1 << a.scale
    P = sp.coo_matrix((np.ones(ne, np.int8), (rows, cols)), shape=(n, n)).tocsc()
    P = (P + P.T + sp.identity(n, dtype=np.int8, format="csc")).tocsc()
    P.sum_duplicates()
    P.sort_indices()
    return P


def weighted(P, dyadic):
    n = P.shape[0]
    cols = np.repeat(np.arange(n, dtype=np.int64), np.diff(P.indptr))
    h = mix(cols * n + P.indices.astype(np.int64) + a.seed)
    if dyadic:
        w = (1.0 + (h % np.uint64(255)).astype(np.float64)) / 256.0
    else:
        w = ((h >> np.uint64(11)).astype(np.float64) + 1.0) / 9007199254740992.0  # uniform (0, 1]
    Mw = sp.csc_matrix((w, P.indices, P.indptr), shape=P.shape)
    if not dyadic:  # MakeColStochastic (MCL.cpp:389-394)
        s = np.asarray(Mw.sum(0)).ravel()
        Mw = sp.csc_matrix(Mw @ sp.diags(1.0 / s))
        Mw.sort_indices()
    return Mw


def gather_sums(vals):
    """three 64-bit words per rank, added modulo 2^64 over the ranks"""
    if world == 1:
        return [v & M64 for v in vals]
    halves = []
    for v in vals:
        halves += [v & 0xFFFFFFFF, (v >> 32) & 0xFFFFFFFF]
    t = torch.tensor(halves, dtype=torch.int64, device="cuda")
    allv = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allv, t)
    out = [0] * len(vals)
    for tv in allv:
        x = tv.tolist()
        for j in range(len(vals)):
            out[j] = (out[j] + x[2 * j] + (x[2 * j + 1] << 32)) & M64
    return out


t_build = time.time()
P = pattern()
n = P.shape[0]
comm = None
grid = None
if world > 1:
    grid = cblib.make_grid(world, rank, layers)
    ids = [cb.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = cb.Comm(ctx, grid, ids[0])


def blocks(Mw):
    H = cb.SpDCCols.from_scipy(Mw, np.float64)
    if world == 1:
        d = ctx.upload(H)
        return d, d, 0, 0
    dA = ctx.upload(cb.partition_3d(H, grid, True))
    dB = ctx.upload(cb.partition_3d(H, grid, False))
    r0, _, c0, _ = local_range(grid, n, n, True)  # C has A's layout: column-split across the layers
    return dA, dB, r0, c0


def expand(dA, dB, phases):
    if world == 1:
        D, ms = ctx.memefficient_spgemm(cb.PlusTimesSRing_f64, dA, dB, phases, a.hard, a.select, a.recover, a.pct, want_stats=True)
        return D, ms, None
    return comm.memefficient_spgemm(cb.PlusTimesSRing_f64, dA, dB, phases, a.hard, a.select, a.recover, a.pct)


# ---------------------------------------------------------------- parity run (dyadic weights: bit-identical on every grid)
Md = weighted(P, True)
dA, dB, r0, c0 = blocks(Md)
D, ms_par, _ = expand(dA, dB, a.phases)
p_, v_ = ctx.checksum(D, r0, c0)
tot = gather_sums([int(D.info().nnz), int(p_), int(v_)])
D.free()
par = {"nnz_pruned": tot[0], "pattern_sum": f"{tot[1]:016x}", "value_sum": f"{tot[2]:016x}", "phases": int(ms_par.phases)}
if rank == 0 and world > 1:
    # the same expansion on ONE GPU (this rank's), whole matrix: cbgpu_memefficient_spgemm, pinned against the reference's
    # MemEfficientSpGEMM in tests/test_prune_gpu.py
    d1 = ctx.upload(cb.SpDCCols.from_scipy(Md, np.float64))
    D1, ms1 = ctx.memefficient_spgemm(cb.PlusTimesSRing_f64, d1, d1, 0, a.hard, a.select, a.recover, a.pct, want_stats=True)
    p1, v1 = ctx.checksum(D1)
    par["single_gpu"] = {"nnz_pruned": int(D1.info().nnz), "pattern_sum": f"{p1:016x}", "value_sum": f"{v1:016x}", "phases": int(ms1.phases)}
    par["equal_to_single_gpu"] = bool(int(D1.info().nnz) == tot[0] and p1 == tot[1] and v1 == tot[2])
    D1.free()
    d1.free()
if world > 1:
    dist.barrier()
dA.free()
if dB is not dA:
    dB.free()
del Md

# ---------------------------------------------------------------- timing run (MCL input)
Mw = weighted(P, False)
dA, dB, r0, c0 = blocks(Mw)
del Mw
build_s = time.time() - t_build
times, e2e_times = [], []
last = None
for i in range(1 + a.steps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    D, ms, ds = expand(dA, dB, a.phases)
    e1.record(stream)
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1)
    host = ctx.download(D)  # e2e: the pruned block of C comes home (DCSC arrays)
    t1 = time.perf_counter()
    d2h = sum(x.nbytes for x in host[2:])
    if i > 0:
        times.append(dev_ms)
        e2e_times.append((t1 - t0) * 1e3)
    last = (int(D.info().nnz), ms, ds, d2h)
    D.free()
ms_dev = float(np.mean(times))
ms_e2e = float(np.mean(e2e_times))
if world > 1:
    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(t[0].item()), float(t[1].item())
nnz_kept, ms, ds, d2h = last
w = torch.tensor([int(ms.flops), int(ms.nnz_unpruned), nnz_kept, d2h], dtype=torch.int64, device="cuda")
if world > 1:
    dist.all_reduce(w)
flops, unpruned, kept, d2h_all = [int(x) for x in w.tolist()]
if rank == 0:
    line = {"config": f"HipMCL expansion, R-MAT scale {a.scale} ef {a.edgefactor} symmetrised + loops, column stochastic, "
                      f"prune {a.hard}/{a.select}/{a.recover}/{a.pct}", "n_gpus": world,
            "grid": {1: "1 GPU", 2: "1x1x2", 4: "2x2x1", 8: "2x2x2"}[world], "n": n, "nnz_A": int(P.nnz),
            "products": flops, "nnz_C_unpruned": unpruned, "nnz_C_pruned": kept, "phases": int(ms.phases),
            "ms_per_step": ms_dev, "gflops": 2.0 * flops / (ms_dev * 1e-3) / 1e9,
            "e2e": {"ms_per_step": ms_e2e, "gflops": 2.0 * flops / (ms_e2e * 1e-3) / 1e9, "d2h_bytes_per_step": d2h_all,
                    "what": "expansion + pruning + download of the pruned DCSC block of C to the host on every rank"},
            "ms_multiply_rank0": float(ms.ms_multiply), "ms_prune_rank0": float(ms.ms_prune),
            "parity": par, "host_build_s": round(build_s, 1)}
    print(json.dumps(line), flush=True)
if world > 1:
    dist.barrier()
    comm.destroy()
    dist.destroy_process_group()

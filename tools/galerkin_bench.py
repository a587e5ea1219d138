"""BASELINE config 3: Galerkin triple product R^T A R for the 3D 7-point Poisson operator on k^3 points (k = 128) with
piecewise-constant 2x2x2 aggregation, PlusTimes<double>, as two chained multiplies RtA = R^T A, RtAR = RtA R
(3DSpGEMM/RestrictionOp.cpp:189-196, Applications/GalerkinNew.cpp:105-106). One GPU: local multiplies; under torchrun
with 4 ranks: 2x2 SUMMA (cbgpu_summa2d) on device-resident blocks. Checks the result against the known coarse operator
(7-point pattern on (k/2)^3, nnz = 7 (k/2)^3 - 6 (k/2)^2, row sums of the stencil) and prints one JSON line.
--dry-run builds and checks the inputs on the CPU only (no GPU needed)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp


def poisson_and_restriction(k):
    n = k ** 3
    I = sp.identity(k, format="csc")
    D1 = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(k, k), format="csc")
    A = (sp.kron(sp.kron(D1, I), I) + sp.kron(sp.kron(I, D1), I) + sp.kron(sp.kron(I, I), D1)).tocsc()
    idx = np.arange(n)
    x, y, z = idx // (k * k), (idx // k) % k, idx % k
    kc = k // 2
    agg = (x // 2) * kc * kc + (y // 2) * kc + (z // 2)
    R = sp.csc_matrix((np.ones(n), (idx, agg)), shape=(n, kc ** 3))
    A.sort_indices()
    return A, R


def expected_nnz(k):
    kc = k // 2
    return 7 * kc ** 3 - 6 * kc ** 2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=128)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--dry-run", action="store_true")
    a = ap.parse_args()
    A, R = poisson_and_restriction(a.k)
    n = a.k ** 3
    assert A.nnz == 7 * n - 6 * a.k ** 2 and R.nnz == n
    if a.dry_run:
        C = (R.T @ A @ R).tocsc()
        assert C.nnz == expected_nnz(a.k), (C.nnz, expected_nnz(a.k))
        print(json.dumps({"dry_run": True, "k": a.k, "nnz_A": int(A.nnz), "nnz_R": int(R.nnz), "nnz_RtAR": int(C.nnz)}))
        return
    import combblas_b200 as cb
    from combblas_b200 import lib as cblib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:  # torch is only the bootstrap of the NCCL communicators
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = cb.Context(local_rank)
    grid = cblib.make_grid(world, rank, 1)
    comm = None
    if world > 1:
        ids = [cb.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = cb.Comm(ctx, grid, ids[0])

    def block(M):
        return ctx.upload(cb.partition_3d(cb.SpDCCols.from_scipy(M, np.float64), grid, True))

    Rt = R.T.tocsc()
    Rt.sort_indices()
    dRt, dA, dR = block(Rt), block(A), block(R)

    def mult(X, Y):
        if world == 1:
            return ctx.spgemm(cb.PlusTimesSRing_f64, X, Y, want_stats=True)
        C, ds = comm.summa2d(cb.PlusTimesSRing_f64, X, Y)
        return C, ds.local

    times, flops, nnz_local = [], 0, 0
    for rep in range(a.reps + 1):
        if world > 1:
            dist.barrier()
        ctx.sync()
        t0 = time.perf_counter()
        RtA, s1 = mult(dRt, dA)
        RtAR, s2 = mult(RtA, dR)
        ctx.sync()
        dt = (time.perf_counter() - t0) * 1e3  # the multiply entry points return after their last kernel has finished
        flops = int(s1.flops) + int(s2.flops)
        nnz_local = RtAR.info().nnz
        RtA.free()
        RtAR.free()
        if rep > 0:
            times.append(dt)
    ms = float(np.median(times))
    nnz = nnz_local
    if world > 1:
        import torch

        tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = torch.tensor([float(flops), float(nnz_local)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tsum)
        ms, flops, nnz = float(tmax.item()), int(tsum[0].item()), int(tsum[1].item())
    ok = nnz == expected_nnz(a.k)
    if rank == 0:
        print(json.dumps({"workload": f"Galerkin RtAR, 7-point Poisson {a.k}^3, 2x2x2 aggregation, PlusTimes<double>", "n_gpus": world,
                          "grid": "1 GPU" if world == 1 else f"{grid.grid_rows}x{grid.grid_cols}", "ms_median": ms,
                          "products": flops, "gflops": 2.0 * flops / ms / 1e6, "nnz_RtAR": nnz, "expected_nnz": expected_nnz(a.k),
                          "parity": "nnz matches the coarse 7-point operator" if ok else "MISMATCH"}), flush=True)
    if world > 1:
        comm.destroy()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

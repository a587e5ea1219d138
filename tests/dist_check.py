"""Multi-GPU parity check, run under torchrun (one rank per GPU):
     python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/dist_check.py
N = 2 -> 1x1x2 (3D, fiber exchange only), 4 -> 2x2x1 (2D SUMMA broadcasts), 8 -> 2x2x2 (both).
Every rank multiplies its blocks through cbgpu_summa2d / cbgpu_summa3d and compares its block of C with the block
of the P=1 oracle product of the same global matrix (pattern bit-exact, values per semiring rule)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import combblas_b200 as cb  # noqa: E402
from combblas_b200 import lib as cblib  # noqa: E402
from combblas_b200.host import local_range  # noqa: E402
from oracle.oracle import Csc, SR_DTYPES, PortOracle  # noqa: E402
from tests.util import assert_same, rmat, typed  # noqa: E402


def main():
    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    layers = {1: 1, 2: 2, 4: 1, 8: 2}[world]
    grid = cblib.make_grid(world, rank, layers)
    ctx = cb.Context(local_rank)
    ids = [cb.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = cb.Comm(ctx, grid, ids[0])
    orc = PortOracle()
    failures = 0
    cases = [(0, 11, 1), (0, 13, 2), (3, 12, 3), (2, 12, 4), (5, 11, 5)]
    for sr, scale, seed in cases:
        ta, tb, _ = SR_DTYPES[sr]
        G = rmat(scale, 8, seed=seed)
        H = rmat(scale, 8, seed=seed + 100)
        GA, GB = typed(G, ta), typed(H, tb)
        n = G.shape[0]
        A = cb.SpDCCols.from_scipy(GA, ta)
        B = cb.SpDCCols.from_scipy(GB, tb)
        Aloc = cb.partition_3d(A, grid, True)
        Bloc = cb.partition_3d(B, grid, False)
        dA, dB = ctx.upload(Aloc), ctx.upload(Bloc)
        if layers == 1:
            dC, st = comm.summa2d(sr, dA, dB)
        else:
            dC, st = comm.summa3d(sr, dA, dB)
        rows, cols, vals = ctx.download_coo(dC)
        m_loc, n_loc = dC.shape
        want_global = orc.spgemm(Csc.from_scipy(GA, ta), Csc.from_scipy(GB, tb), sr)
        r0, r1, c0, c1 = local_range(grid, n, n, True)  # C is column-split like A
        Wg = cb.SpDCCols.from_csc(n, n, want_global.colptr, want_global.rows, want_global.vals).submatrix(r0, r1, c0, c1)
        colptr, wrows, wvals = Wg.to_csc()
        want = Csc(r1 - r0, c1 - c0, colptr, wrows, wvals)
        got = cb.SpTuples(m_loc, n_loc, rows, cols, vals)
        ok = True
        try:
            assert (m_loc, n_loc) == (r1 - r0, c1 - c0), f"block shape {(m_loc, n_loc)} vs {(r1 - r0, c1 - c0)}"
            assert_same(got, want, sr)
        except AssertionError as e:
            ok = False
            print(f"[rank {rank}] FAIL sr={sr} scale={scale}: {e}", flush=True)
        t = torch.tensor([0 if ok else 1], device="cuda")
        dist.all_reduce(t)
        if rank == 0:
            d = st.as_dict()
            print(f"{'PASS' if t.item() == 0 else 'FAIL'} world={world} layers={layers} sr={sr} scale={scale} "
                  f"stages={d['stages']} ms_total={d['ms_total']:.2f} bcast={d['ms_bcast']:.2f} mult={d['ms_multiply']:.2f} "
                  f"merge={d['ms_merge']:.2f} fiber={d['ms_fiber_exchange']:.2f}+{d['ms_fiber_merge']:.2f}", flush=True)
        failures += int(t.item())
        for x in (dA, dB, dC):
            x.free()
    comm.destroy()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()

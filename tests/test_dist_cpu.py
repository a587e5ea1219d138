"""CPU, world_size 2 and 4 over gloo: the host-side plan of the distributed path -- block ownership, SUMMA stage
schedule, 3D column-slab exchange -- is exercised with real inter-process traffic. The local multiply and merge are
done by the CPU oracle here (there is no GPU in this container); what is under test is the partition arithmetic
served by the C ABI (cbgpu_grid_make / cbgpu_grid_local_range / cbgpu_block_range) and the exchange plan that
cbgpu_summa2d / cbgpu_summa3d follow (ParFriends.h:1482-1532, :3578-3642)."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, layers, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import combblas_b200 as cb
    from combblas_b200 import lib as cblib
    from combblas_b200.host import local_range
    from oracle.oracle import Csc, PortOracle
    from tests.util import rmat

    orc = PortOracle()
    orc.set_num_threads(1)
    grid = cblib.make_grid(world, rank, layers)
    G = rmat(9, 8, seed=3)
    n = G.shape[0]
    A = cb.SpDCCols.from_scipy(G, np.float64)
    Aloc = cb.partition_3d(A, grid, True)
    Bloc = cb.partition_3d(A, grid, False)

    def as_csc(D):
        cp, r, v = D.to_csc()
        return Csc(D.m, D.n, cp, r, v)

    pr = grid.grid_cols
    per_layer = pr * pr
    # ---- SUMMA inside my layer: stage i uses A from (my_row, i) and B from (i, my_col)
    partial = []
    for i in range(pr):
        a_src = grid.my_layer * per_layer + grid.my_row * pr + i
        b_src = grid.my_layer * per_layer + i * pr + grid.my_col
        boxA = [Aloc if rank == r else None for r in range(world)]
        boxB = [Bloc if rank == r else None for r in range(world)]
        # every rank takes part in every broadcast (gloo world group stands in for the row/column communicators)
        for src in range(world):
            obj = [boxA[src], boxB[src]]
            dist.broadcast_object_list(obj, src=src)
            if src == a_src:
                Ar = obj[0]
            if src == b_src:
                Br = obj[1]
        assert Ar.n == Br.m, "inner block dimensions must agree (same floor rule on the same global dimension)"
        if Ar.nnz and Br.nnz:
            partial.append(orc.spgemm(as_csc(Ar), as_csc(Br), 0))
    Cl = orc.merge(partial, 0) if len(partial) > 1 else (partial[0] if partial else None)
    m_blk = Aloc.m
    n_blk = Bloc.n
    if Cl is None:
        Cl = Csc(m_blk, n_blk, np.zeros(n_blk + 1, np.int64), np.zeros(0, np.int64), np.zeros(0))
    Cd = cb.SpDCCols.from_csc(Cl.m, Cl.n, Cl.colptr, Cl.rows, Cl.vals)
    # ---- fiber exchange: slab l of my layer result goes to layer l (same row/col position)
    L = layers
    if L > 1:
        slabs = []
        for l in range(L):
            c0, c1 = cb.block_range(Cd.n, L, l)
            slabs.append(Cd.colslice(c0, c1))
        inbox = []
        for src in range(world):
            obj = [slabs if rank == src else None]
            dist.broadcast_object_list(obj, src=src)
            g2 = cblib.make_grid(world, src, layers)
            if (g2.my_row, g2.my_col) == (grid.my_row, grid.my_col):
                inbox.append(obj[0][grid.my_layer])
        lists = [as_csc(x) for x in inbox if x.nnz]
        mine = orc.merge(lists, 0) if len(lists) > 1 else lists[0]
    else:
        mine = Cl
    # ---- expected: my block of the global product
    want_global = orc.spgemm(as_csc(A), as_csc(A), 0)
    r0, r1, c0, c1 = local_range(grid, n, n, True)
    W = cb.SpDCCols.from_csc(n, n, want_global.colptr, want_global.rows, want_global.vals).submatrix(r0, r1, c0, c1)
    cp, rr, vv = W.to_csc()
    ok = (mine.m, mine.n) == (r1 - r0, c1 - c0) and np.array_equal(mine.colptr, cp) and np.array_equal(mine.rows, rr) \
        and np.allclose(mine.vals, vv, rtol=1e-12, atol=0)
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,layers", [(2, 2), (4, 1)])
def test_distributed_plan_over_gloo(world, layers):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + world * 10 + layers
    mp.spawn(_worker, args=(world, layers, port, ret), nprocs=world, join=True)
    assert len(ret) == world and all(ret.values()), dict(ret)


@pytest.mark.parametrize("world,layers", [(2, 2), (8, 2), (4, 4), (16, 4)])
def test_fiber_fused_plan_equals_global_product(world, layers):
    """The plan of the fiber-fused 3D multiply (csrc/dist.cu, option fiber_fused), replayed with scipy for every rank:
    A blocks of the process row from every (layer, stage) side by side, the rank's column sub-slab of the B blocks of its
    process column from every (layer, stage) stacked in the same order -- one local product, which must be exactly the
    rank's block of the global product in the reference's 3D output distribution (C column-split across layers)."""
    import scipy.sparse as sp

    sys.path.insert(0, ROOT)
    import combblas_b200 as cb
    from combblas_b200 import lib as cblib
    from combblas_b200.host import local_range
    from tests.util import rmat

    G = rmat(9, 8, seed=4)
    G.data = np.floor(G.data) + 1.0  # small integers: products and sums are exact, comparison can be bit for bit
    n = G.shape[0]
    A = cb.SpDCCols.from_scipy(G, np.float64)
    Cg = (G @ G).tocsc()
    Cg.sort_indices()
    grids = [cblib.make_grid(world, r, layers) for r in range(world)]
    at = {(g.my_row, g.my_col, g.my_layer): r for r, g in enumerate(grids)}
    assert len(at) == world
    Ablk = [cb.partition_3d(A, g, True) for g in grids]
    Bblk = [cb.partition_3d(A, g, False) for g in grids]

    def sci(D):
        cp, rows, vals = D.to_csc()
        return sp.csc_matrix((vals, rows, cp), shape=(D.m, D.n))

    for r, g in enumerate(grids):
        pr, L = g.grid_cols, g.layers
        Aall = sp.hstack([sci(Ablk[at[(g.my_row, k, l)]]) for l in range(L) for k in range(pr)], format="csc")
        parts = []
        for l in range(L):
            Bcol = sp.vstack([sci(Bblk[at[(k, g.my_col, l)]]) for k in range(pr)], format="csc")
            c0, c1 = cb.block_range(Bcol.shape[1], L, g.my_layer)
            parts.append(Bcol[:, c0:c1])
        Ball = sp.vstack(parts, format="csc")
        assert Aall.shape[1] == Ball.shape[0] == n
        mine = (Aall @ Ball).tocsc()
        mine.sort_indices()
        r0, r1, c0, c1 = local_range(g, n, n, True)
        want = Cg[r0:r1, c0:c1].tocsc()
        want.sort_indices()
        assert mine.shape == want.shape, (r, mine.shape, want.shape)
        assert np.array_equal(mine.indptr, want.indptr) and np.array_equal(mine.indices, want.indices), f"rank {r}: pattern"
        assert np.array_equal(mine.data, want.data), f"rank {r}: values"


def test_phase_plan_pieces_tile_the_layer_chunks():
    """cbgpu_phase_columns: the pieces a fiber rank produces over the phases are its chunk of the block's columns, in order
    (so the phased 3D product has the layout of the unphased one, ParFriends.h:3774-3811 / SpParMat3D.cpp:576-609)"""
    import ctypes as C

    import combblas_b200 as cb

    lib = cb.load_library()
    for n, phases, layers in [(1000, 3, 2), (17, 5, 4), (4096, 1, 2), (7, 3, 1), (5, 8, 2), (0, 2, 2)]:
        for l in range(layers):
            k0, k1 = C.c_int64(), C.c_int64()
            assert lib.cbgpu_block_range(n, layers, l, C.byref(k0), C.byref(k1)) == 0
            at = k0.value
            for p in range(phases):
                b, e = C.c_int64(), C.c_int64()
                assert lib.cbgpu_phase_columns(n, phases, layers, p, l, C.byref(b), C.byref(e)) == 0
                assert b.value == at and e.value >= b.value
                at = e.value
            assert at == k1.value
    b, e = C.c_int64(), C.c_int64()
    assert lib.cbgpu_phase_columns(10, 2, 2, 2, 0, C.byref(b), C.byref(e)) != 0

"""GPU parity tests proper: every call goes through the C ABI (libcbgpu.so) and is compared with the CPU oracle
(the compiled reference when present, else its C restatement) and with the committed reference outputs."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import combblas_b200 as cb
from oracle.oracle import Csc, SR_DTYPES
from tests.util import assert_same, random_pair, rmat, to_csc, to_dcsc, typed

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name, prefix):
    z = np.load(os.path.join(GOLD, name))
    m, n = z[prefix + "_shape"]
    return Csc(int(m), int(n), z[prefix + "_colptr"], z[prefix + "_rows"], z[prefix + "_vals"])


def dcsc_of(c: Csc, dt=None, idx=np.int64):
    return cb.SpDCCols.from_csc(c.m, c.n, c.colptr, c.rows, c.vals if dt is None else c.vals.astype(dt), idx)


def gpu_mult(ctx, sr, A, B, idx=np.int64):
    return cb.LocalHybridSpGEMM(ctx, sr, to_dcsc(A, SR_DTYPES[sr][0], idx), to_dcsc(B, SR_DTYPES[sr][1], idx))


def check_pair(ctx, oracle, sr, A, B, idx=np.int64):
    got = gpu_mult(ctx, sr, A, B, idx)
    want = oracle.spgemm(to_csc(A, SR_DTYPES[sr][0]), to_csc(B, SR_DTYPES[sr][1]), sr)
    assert_same(got, want, sr)
    return got


def test_reference_known_answer_bcsstk01(ctx):
    A, G = load("bcsstk01_squared.npz", "A"), load("bcsstk01_squared.npz", "C")
    got = cb.LocalHybridSpGEMM(ctx, 0, dcsc_of(A), dcsc_of(A))
    assert got.getnnz() == 1292
    assert np.array_equal(got.rows, G.rows) and np.array_equal(got.cols, G.cols_expanded())
    assert np.max(np.abs(got.vals - G.vals) / np.abs(G.vals)) < 1e-12


@pytest.mark.parametrize("sr", range(9))
def test_committed_reference_outputs(ctx, sr):
    f = f"ref_sr{sr}.npz"
    A, B, Cg = load(f, "A"), load(f, "B"), load(f, "C")
    got = cb.LocalHybridSpGEMM(ctx, sr, dcsc_of(A), dcsc_of(B))
    assert_same(got, Cg, sr)
    parts = [dcsc_of(load(f, f"P{i}")) for i in range(3)]
    mg = cb.MultiwayMerge(ctx, sr, parts)
    assert_same(mg, load(f, "M"), sr)


@pytest.mark.parametrize("sr", range(9))
@pytest.mark.parametrize("shape", [(300, 220, 260, 0.05, 0.04), (64, 2000, 50, 0.02, 0.3), (3000, 40, 3000, 0.2, 0.01)])
def test_every_semiring_random(ctx, oracle, sr, shape):
    m, k, n, da, db = shape
    A, B = random_pair(m, k, n, da, db, 11 + sr, SR_DTYPES[sr])
    check_pair(ctx, oracle, sr, A, B)


@pytest.mark.parametrize("scale", [8, 11, 13])
def test_rmat_squared_plus_times(ctx, oracle, scale):
    A = rmat(scale, 16, seed=1)
    got = check_pair(ctx, oracle, 0, A, A)
    assert got.getnnz() > 0


def test_rmat_scale14_all_paths(ctx, oracle):
    """heavy columns: exercises the bitmap path with shared and HBM accumulators next to the hash paths"""
    A = rmat(14, 16, seed=2)
    want = oracle.spgemm(to_csc(A, np.float64), to_csc(A, np.float64), 0)
    dA = ctx.upload(to_dcsc(A, np.float64))
    # (force_path, shared_acc, shared_acc_max): every numeric class of the bitmap path -- accumulators in shared memory in the
    # three CTA shapes (capacity limits lowered so that this small matrix reaches the medium and large shapes) or in C itself
    for force, shared, cap in ((0, 1, 0), (1, 1, 0), (2, 1, 0), (2, 1, 700), (0, 0, 0), (2, 0, 0)):
        ctx.set_option("force_path", force)
        ctx.set_option("shared_acc", shared)
        ctx.set_option("shared_acc_max", cap)
        ctx.set_option("shared_acc_small_max", 300 if cap else -1)
        D, st = ctx.spgemm(0, dA, dA, want_stats=True)
        rows, cols, vals = ctx.download_coo(D)
        got = cb.SpTuples(A.shape[0], A.shape[1], rows, cols, vals)
        assert_same(got, want, 0)
        assert st.flops == int(np.diff(A.indptr)[A.indices].sum())
        assert st.nnz_out == want.nnz
        if force == 2:
            assert st.tasks_hash_warp + st.tasks_hash_cta == 0
        if shared == 0 or cap:
            assert st.tasks_bitmap_gmem > 0
        if shared == 1 and force != 1:
            assert st.tasks_bitmap_smem > 0
        D.free()
    ctx.set_option("force_path", 0)
    ctx.set_option("shared_acc", 1)
    ctx.set_option("shared_acc_max", 0)
    ctx.set_option("shared_acc_small_max", -1)
    dA.free()


@pytest.mark.parametrize("sr", range(9))
def test_shared_accumulator_classes_every_semiring(ctx, oracle, sr):
    """the exchange-protocol accumulators (csrc/semiring.cuh exch_accumulate) in all three CTA shapes + the fallback into C,
    for every semiring: bitmap path forced, capacities lowered so that one R-MAT square spreads over all four classes"""
    A = rmat(12, 16, seed=20 + sr)
    ctx.set_option("force_path", 2)
    ctx.set_option("shared_acc_max", 500)
    ctx.set_option("shared_acc_small_max", 120)
    try:
        ta, tb, _ = SR_DTYPES[sr]
        want = oracle.spgemm(to_csc(typed(A, ta), ta), to_csc(typed(A, tb), tb), sr)
        dA, dB = ctx.upload(to_dcsc(typed(A, ta), ta)), ctx.upload(to_dcsc(typed(A, tb), tb))
        D, st = ctx.spgemm(sr, dA, dB, want_stats=True)
        rows, cols, vals = ctx.download_coo(D)
        assert_same(cb.SpTuples(A.shape[0], A.shape[1], rows, cols, vals), want, sr)
        assert st.tasks_bitmap_smem > 0 and st.tasks_bitmap_gmem > 0
        for x in (dA, dB, D):
            x.free()
    finally:
        ctx.set_option("force_path", 0)
        ctx.set_option("shared_acc_max", 0)
        ctx.set_option("shared_acc_small_max", -1)


@pytest.mark.parametrize("pair", [(9, 10, np.float64), (11, 12, np.int64), (13, 14, np.uint8)])
def test_bool_copy_semirings_subsref(ctx, oracle, port_oracle, pair):
    """BoolCopy2ndSRing / BoolCopy1stSRing (Semirings.h:51-138; SpParMat::SubsRef_SR, SpParMat.cpp:2515-2566): S*A and (S*A)*T
    with boolean selectors against the oracle, bit for bit (explicit zeros and negative values are copied, not added to an
    identity), through the hash / register-sort classes and through the bitmap classes; the stage merge of two disjoint row
    halves; and the reference's "Add should not happen": a product or merge that would need add fails with CBGPU_ERR_INVALID"""
    from tests.util import subsref_operands

    sr2, sr1, dt = pair
    A, S, T, ri, ci = subsref_operands(3000, 2600, 2200, 1900, 50 + sr2, dt)
    a, s_, t_ = to_csc(A, dt), to_csc(S, np.uint8), to_csc(T, np.uint8)
    want_sa = oracle.spgemm(s_, a, sr2)
    want_sat = oracle.spgemm(want_sa, t_, sr1)
    for force in (0, 2):
        ctx.set_option("force_path", force)
        try:
            sa = cb.LocalHybridSpGEMM(ctx, sr2, to_dcsc(S, np.uint8), to_dcsc(A, dt))
            assert_same(sa, want_sa, sr2)
            sat = cb.LocalHybridSpGEMM(ctx, sr1, dcsc_of(want_sa), to_dcsc(T, np.uint8))
            assert_same(sat, want_sat, sr1)
        finally:
            ctx.set_option("force_path", 0)
    # SUMMA-stage style merge: the selector split by columns gives partial products with disjoint rows... of the same columns
    half = S.shape[1] // 2
    S0, S1 = S.tocsc()[:, :half], S.tocsc()[:, half:]
    A0, A1 = A.tocsr()[:half].tocsc(), A.tocsr()[half:].tocsc()
    for M in (S0, S1, A0, A1):
        M.sort_indices()
    p0 = port_oracle.spgemm(to_csc(S0, np.uint8), to_csc(A0, dt), sr2)
    p1 = port_oracle.spgemm(to_csc(S1, np.uint8), to_csc(A1, dt), sr2)
    for tma in (1, 0):
        ctx.set_option("merge_tma", tma)
        try:
            assert_same(cb.MultiwayMerge(ctx, sr2, [dcsc_of(p0), dcsc_of(p1)]), want_sa, sr2)
            if p0.nnz > 0:
                with pytest.raises(cb.CbgpuError, match="Add should not happen"):
                    cb.MultiwayMerge(ctx, sr2, [dcsc_of(p0), dcsc_of(p0)])
        finally:
            ctx.set_option("merge_tma", 1)
    # a selector row with two entries on rows of A that share a column
    Ac = A.tocsc()
    col = int(np.argmax(np.diff(Ac.indptr)))
    r0, r1 = Ac.indices[Ac.indptr[col]], Ac.indices[Ac.indptr[col] + 1]
    X = sp.coo_matrix((np.ones(2), (np.zeros(2, int), np.array([r0, r1]))), shape=(1, A.shape[0])).tocsc()
    X.sort_indices()
    with pytest.raises(cb.CbgpuError, match="Add should not happen"):
        cb.LocalHybridSpGEMM(ctx, sr2, to_dcsc(X, np.uint8), to_dcsc(A, dt))


@pytest.mark.parametrize("v2", [7, 15, 23])
@pytest.mark.parametrize("sr", range(9))
def test_shared_accumulators_second_version(ctx, oracle, sr, v2):
    """num_sacc2_kernel (option sacc_v2: 16-bit ranks, rows stored by the accumulate walk, vector scan, presence words fetched
    by a bulk copy): all three CTA shapes, with and without the row array (bits 3 and 4), every semiring; one window and
    several; hand-over forced for small tasks so that the bulk-copy branch runs"""
    A = rmat(12, 16, seed=20 + sr)
    ta, tb, _ = SR_DTYPES[sr]
    want = oracle.spgemm(to_csc(typed(A, ta), ta), to_csc(typed(A, tb), tb), sr)
    ctx.set_option("sacc_v2", v2)
    ctx.set_option("force_path", 2)
    try:
        dA, dB = ctx.upload(to_dcsc(typed(A, ta), ta)), ctx.upload(to_dcsc(typed(A, tb), tb))
        # sacc_overflow: the large shape takes tasks of up to that many times its capacity; the outputs beyond the capacity
        # accumulate in C with L2 reductions (1: such tasks go to num_bitmap_kernel whole)
        for (wlog2, cap, small, save_min, stripes) in ((17, 500, 120, 8192, 2), (17, 0, -1, 64, 2), (10, 400, 100, 64, 3), (11, 0, -1, 8192, 1),
                                                       (17, 300, 80, 64, 4)):
            ctx.set_option("sacc_overflow", stripes)
            ctx.set_option("bitmap_window_log2", wlog2)
            ctx.set_option("shared_acc_max", cap)
            ctx.set_option("shared_acc_small_max", small)
            ctx.set_option("bitmap_save_min_flop", save_min)
            D, st = ctx.spgemm(sr, dA, dB, want_stats=True)
            rows, cols, vals = ctx.download_coo(D)
            assert_same(cb.SpTuples(A.shape[0], A.shape[1], rows, cols, vals), want, sr)
            assert st.tasks_bitmap_smem > 0
            D.free()
        for x in (dA, dB):
            x.free()
    finally:
        ctx.set_option("sacc_v2", 15)
        ctx.set_option("sacc_overflow", 4)
        ctx.set_option("force_path", 0)
        ctx.set_option("bitmap_window_log2", 17)
        ctx.set_option("shared_acc_max", 0)
        ctx.set_option("shared_acc_small_max", -1)
        ctx.set_option("bitmap_save_min_flop", 2048)


@pytest.mark.parametrize("v2", [15, 31])
def test_shared_accumulators_second_version_full_window(ctx, oracle, v2):
    """windows of 2^17 rows: the 16-bit row offsets of num_sacc2_kernel need the half bit of the window (outputs at or behind
    rank[2048] lie in the upper 2^16 rows); R-MAT scale 17 squared with the bitmap classes forced, hand-over on and off"""
    A = rmat(17, 4, seed=77)
    want = oracle.spgemm(to_csc(A, np.float64), to_csc(A, np.float64), 0)
    dA = ctx.upload(to_dcsc(A, np.float64))
    ctx.set_option("sacc_v2", v2)
    ctx.set_option("force_path", 2)
    try:
        for save_min in (8192, 1 << 40):
            ctx.set_option("bitmap_save_min_flop", save_min)
            D, st = ctx.spgemm(0, dA, dA, want_stats=True)
            rows, cols, vals = ctx.download_coo(D)
            assert_same(cb.SpTuples(A.shape[0], A.shape[1], rows, cols, vals), want, 0)
            assert st.tasks_bitmap_smem > 0
            D.free()
    finally:
        ctx.set_option("sacc_v2", 15)
        ctx.set_option("force_path", 0)
        ctx.set_option("bitmap_save_min_flop", 2048)
        dA.free()


@pytest.mark.parametrize("sr", [0, 2, 3, 4, 5, 7])
def test_register_sort_runs_across_lanes(ctx, oracle, sr):
    """register-sort classes (engine.cuh regsort_kernel): r dense rows of A times a column of B with k entries gives r runs of k
    equal rows each -- runs that start and end inside a lane, fill whole lanes, and span up to all lanes of a group (k = 256:
    one output from 256 products) -- next to columns with distinct rows only; group sizes 8, 16 and 32 lanes"""
    rng = np.random.default_rng(40 + sr)
    ta, tb, _ = SR_DTYPES[sr]
    for (r, k) in [(1, 256), (2, 128), (4, 60), (3, 85), (7, 9), (64, 1), (1, 64), (5, 51), (16, 16), (9, 7), (1, 9), (2, 33)]:
        m, inner, n = 300, 400, 40
        A = sp.lil_matrix((m, inner))
        rows = rng.choice(m, size=r, replace=False)
        cols = rng.choice(inner, size=k, replace=False)
        for i in rows:
            A[i, cols] = rng.integers(1, 5, size=k)
        B = sp.lil_matrix((inner, n))
        for j in range(n):
            # even columns: the k columns of A that hold the dense rows (long runs); odd ones: some of them plus empty columns of A
            pick = cols if j % 2 == 0 else np.concatenate([cols[: max(1, k // 3)], rng.choice(inner, size=5, replace=False)])
            B[np.unique(pick), j] = rng.integers(1, 4, size=len(np.unique(pick)))
        # regsort 1: numeric pass (default), 2: the symbolic pass counts with the same network; packed 1: the network sorts
        # row << log2(capacity) | position and the values are fetched afterwards (default), 0: key + value through the network
        for mode, packed in ((1, 1), (2, 1), (1, 0)):
            ctx.set_option("regsort", mode)
            ctx.set_option("regsort_packed", packed)
            try:
                check_pair(ctx, oracle, sr, typed(A.tocsc(), ta), typed(B.tocsc(), tb))
            finally:
                ctx.set_option("regsort", 1)
                ctx.set_option("regsort_packed", 1)


@pytest.mark.parametrize("sr", range(9))
def test_register_sort_off_keeps_the_hash_classes_covered(ctx, oracle, sr):
    """option regsort = 0 sends the small tasks to the per-warp hash classes again: same product"""
    ctx.set_option("regsort", 0)
    try:
        ta, tb, _ = SR_DTYPES[sr]
        A, B = random_pair(300, 220, 260, 0.05, 0.04, 77 + sr, SR_DTYPES[sr])
        check_pair(ctx, oracle, sr, A, B)
        G = rmat(11, 8, seed=30 + sr)
        check_pair(ctx, oracle, sr, typed(G, ta), typed(G, tb))
    finally:
        ctx.set_option("regsort", 1)


@pytest.mark.parametrize("wlog2", [10, 12])
@pytest.mark.parametrize("sr", [0, 3, 5])
def test_row_windows(ctx, oracle, wlog2, sr):
    """several row windows per column (forced by a small window) must give the same block"""
    ctx.set_option("bitmap_window_log2", wlog2)
    try:
        A = rmat(13, 16, seed=3)
        B = rmat(13, 8, seed=4)
        check_pair(ctx, oracle, sr, typed(A, SR_DTYPES[sr][0]), typed(B, SR_DTYPES[sr][1]))
    finally:
        ctx.set_option("bitmap_window_log2", 17)


def test_tall_matrix_natural_windows(ctx, oracle):
    """m > 2^19 rows: the default window size needs three windows"""
    rng = np.random.default_rng(5)
    m, k, n = 1_300_000, 3000, 400
    A = sp.random(m, k, density=40.0 / m * 20, random_state=rng, format="csc")
    B = sp.random(k, n, density=0.05, random_state=rng, format="csc")
    check_pair(ctx, oracle, 0, A, B)
    check_pair(ctx, oracle, 2, typed(A, np.int64), typed(B, np.int64))


def test_select_max_er_config_reduced(ctx, oracle):
    """config 2 at reduced size: ER d=8, A bool, B int64 = 1 + row id, SelectMaxSRing<bool,int64_t>, bit-exact"""
    n, d = 1 << 15, 8
    rng = np.random.default_rng(2)
    r, c = rng.integers(0, n, n * d), rng.integers(0, n, n * d)
    P = sp.coo_matrix((np.ones(n * d), (r, c)), shape=(n, n)).tocsc()
    P.sum_duplicates()
    P.sort_indices()
    A = P.copy()
    A.data[:] = 1
    B = P.copy()
    B.data = (1 + B.indices).astype(np.float64)
    got = cb.LocalHybridSpGEMM(ctx, 3, to_dcsc(A, np.uint8), to_dcsc(B, np.int64))
    want = oracle.spgemm(to_csc(A, np.uint8), to_csc(B, np.int64), 3)
    assert_same(got, want, 3)


def test_int32_local_indices(ctx, oracle):
    """SpDCCols<int32_t,...> local blocks (MCL.cpp:846) go through the same entry"""
    A, B = random_pair(500, 400, 300, 0.03, 0.03, 9, SR_DTYPES[0])
    check_pair(ctx, oracle, 0, A, B, idx=np.int32)


def test_empty_and_degenerate(ctx):
    z = cb.SpDCCols.from_coo(5, 7, [], [], np.zeros(0))
    a = cb.SpDCCols.from_coo(4, 5, [0, 3], [1, 4], np.array([2.0, 3.0]))
    c = cb.LocalHybridSpGEMM(ctx, 0, a, z)
    assert c.getnnz() == 0 and (c.m, c.n) == (4, 7)
    c = cb.LocalHybridSpGEMM(ctx, 0, cb.SpDCCols.from_coo(4, 5, [], [], np.zeros(0)), cb.SpDCCols.from_coo(5, 2, [1], [1], np.ones(1)))
    assert c.getnnz() == 0 and (c.m, c.n) == (4, 2)
    # B columns that only select empty A columns give no output column
    a = cb.SpDCCols.from_coo(3, 3, [0], [0], np.array([2.0]))
    b = cb.SpDCCols.from_coo(3, 2, [0, 2], [0, 1], np.array([5.0, 7.0]))
    c = cb.LocalHybridSpGEMM(ctx, 0, a, b)
    assert c.getnnz() == 1 and c.rows[0] == 0 and c.cols[0] == 0 and c.vals[0] == 10.0
    # explicit zeros are kept: cancellation stays a stored entry
    x = cb.SpDCCols.from_coo(2, 2, [0, 0], [0, 1], np.array([1.0, -1.0]))
    y = cb.SpDCCols.from_coo(2, 1, [0, 1], [0, 0], np.array([1.0, 1.0]))
    c = cb.LocalHybridSpGEMM(ctx, 0, x, y)
    assert c.getnnz() == 1 and c.vals[0] == 0.0
    with pytest.raises(cb.CbgpuError) as e:
        cb.LocalHybridSpGEMM(ctx, 0, a, cb.SpDCCols.from_coo(4, 2, [0], [0], np.ones(1)))
    assert e.value.code == -5  # DIMMISMATCH
    with pytest.raises(TypeError):
        cb.LocalHybridSpGEMM(ctx, 3, a, b)


def test_symbolic_matches_oracle(ctx, port_oracle):
    A = rmat(12, 16, seed=6)
    flop, nnz = port_oracle.symbolic(to_csc(A, np.float64), to_csc(A, np.float64))
    f, z = cb.EstimateFLOP(ctx, to_dcsc(A, np.float64), to_dcsc(A, np.float64))
    assert f == int(flop.sum()) and z == int(nnz.sum())
    # per column, as estimateFLOP (mtSpGEMM.h:1058) and estimateNNZ_Hash (:807) return them: one entry per non-empty column of B
    for wlog2 in (17, 10):  # one row window, and several windows per column
        ctx.set_option("bitmap_window_log2", wlog2)
        try:
            dA = ctx.upload(to_dcsc(A, np.float64))
            cf, cn = ctx.symbolic_columns(dA, dA)
        finally:
            ctx.set_option("bitmap_window_log2", 17)
        nonempty = np.diff(A.indptr) > 0
        assert np.array_equal(cf, flop[nonempty]) and np.array_equal(cn, nnz[nonempty])
        dA.free()


@pytest.mark.parametrize("sr", [0, 2, 3, 4, 5])
@pytest.mark.parametrize("k", [1, 2, 4, 7])
def test_merge_matches_oracle(ctx, port_oracle, sr, k):
    _, B = random_pair(400, 300, 350, 0.04, 0.04, 60 + sr, SR_DTYPES[sr])
    b = to_csc(B, SR_DTYPES[sr][1])
    parts = []
    for i in range(k):
        Ai, _ = random_pair(400, 300, 350, 0.04, 0.04, 700 + 9 * sr + i, SR_DTYPES[sr])
        parts.append(port_oracle.spgemm(to_csc(Ai, SR_DTYPES[sr][0]), b, sr))
    want = port_oracle.merge(parts, sr)
    got = cb.MultiwayMerge(ctx, sr, [dcsc_of(p) for p in parts])
    assert_same(got, want, sr)


def test_merge_large_columns_and_windows(ctx, port_oracle):
    A = rmat(13, 16, seed=8)
    a = to_csc(A, np.float64)
    parts = [port_oracle.spgemm(a, to_csc(rmat(13, 4, seed=20 + i), np.float64), 0) for i in range(3)]
    want = port_oracle.merge(parts, 0)
    for wlog2 in (17, 11):
        ctx.set_option("bitmap_window_log2", wlog2)
        got = cb.MultiwayMerge(ctx, 0, [dcsc_of(p) for p in parts])
        assert_same(got, want, 0)
    ctx.set_option("bitmap_window_log2", 17)


@pytest.mark.parametrize("sr", [0, 1, 2, 5, 7])
def test_streaming_merge_bulk_copies_every_value_width(ctx, port_oracle, sr):
    """merge2_tma_kernel (persistent CTAs, double-buffered bulk copies in, bulk stores out) against the oracle and against the
    one-tile-per-CTA kernel: value widths of 1, 4 and 8 bytes (the 16-byte phase of every part differs), columns of several
    tiles next to empty and one-entry columns, duplicate pairs on tile and thread boundaries (dense columns in both lists)"""
    ta, tb, _ = SR_DTYPES[sr]
    A = rmat(13, 16, seed=81)
    a = to_csc(typed(A, ta), ta)
    parts = [port_oracle.spgemm(a, to_csc(typed(rmat(13, 6, seed=90 + i), tb), tb), sr) for i in range(3)]
    want = port_oracle.merge(parts, sr)
    got = {}
    for tma in (1, 0):
        ctx.set_option("merge_tma", tma)
        try:
            got[tma] = cb.MultiwayMerge(ctx, sr, [dcsc_of(p) for p in parts])
        finally:
            ctx.set_option("merge_tma", 1)
        assert_same(got[tma], want, sr)
    assert np.array_equal(got[0].rows, got[1].rows) and np.array_equal(got[0].cols, got[1].cols)


@pytest.mark.parametrize("v2", [15, 0])
@pytest.mark.parametrize("sr", [0, 3, 5])
def test_merge_through_the_engine_large_columns(ctx, port_oracle, sr, v2):
    """the k-way merge through the accumulation engine (option merge_engine = 1; always taken for k = 1): columns of thousands of
    entries reach the bitmap classes with MERGE = true -- shared accumulators (second and first version), the overflow into C
    (capacity lowered), the register-sort classes for the short columns"""
    ta, tb, _ = SR_DTYPES[sr]
    A = rmat(13, 16, seed=61)
    a = to_csc(typed(A, ta), ta)
    parts = [port_oracle.spgemm(a, to_csc(typed(rmat(13, 6, seed=70 + i), tb), tb), sr) for i in range(3)]
    want3 = port_oracle.merge(parts, sr)
    want1 = port_oracle.merge(parts[:1], sr)
    ctx.set_option("merge_engine", 1)
    ctx.set_option("sacc_v2", v2)
    try:
        for cap in (0, 600):
            ctx.set_option("shared_acc_max", cap)
            ctx.set_option("shared_acc_small_max", 150 if cap else -1)
            assert_same(cb.MultiwayMerge(ctx, sr, [dcsc_of(p) for p in parts]), want3, sr)
            assert_same(cb.MultiwayMerge(ctx, sr, [dcsc_of(parts[0])]), want1, sr)
    finally:
        ctx.set_option("merge_engine", 0)
        ctx.set_option("sacc_v2", 15)
        ctx.set_option("shared_acc_max", 0)
        ctx.set_option("shared_acc_small_max", -1)


def test_validate_rejects_blocks_the_engine_cannot_take(ctx):
    """cbgpu_mat_validate / option validate_uploads: sorted, in-range blocks pass; unsorted rows (legal after the reference's
    sort=false paths), rows >= m, broken column pointers and unordered column ids are named in the error"""
    A = rmat(10, 8, seed=3)
    good = to_dcsc(A, np.float64)
    D = ctx.upload(good)
    ctx.validate(D)
    D.free()
    ctx.set_option("validate_uploads", 1)
    try:
        ctx.upload(good).free()
        col = int(np.argmax(np.diff(good.cp)))
        b = int(good.cp[col])
        for what, mutate in (("row ids not strictly ascending", lambda d: d.ir.__setitem__(slice(b, b + 2), d.ir[b:b + 2][::-1].copy())),
                             ("row ids out of range", lambda d: d.ir.__setitem__(int(d.cp[col + 1]) - 1, d.m + 5)),
                             ("column pointers", lambda d: d.cp.__setitem__(1, d.cp[2] + 1)),
                             ("column ids", lambda d: d.jc.__setitem__(slice(0, 2), d.jc[0:2][::-1].copy()))):
            bad = to_dcsc(A.copy(), np.float64)
            bad.ir, bad.cp, bad.jc = bad.ir.copy(), bad.cp.copy(), bad.jc.copy()
            mutate(bad)
            with pytest.raises(cb.CbgpuError, match=what):
                ctx.upload(bad)
    finally:
        ctx.set_option("validate_uploads", 0)


def test_colsplit_concat_and_checksum(ctx):
    A = rmat(11, 16, seed=9)
    D = ctx.upload(to_dcsc(A, np.float64))
    base = ctx.checksum(D)
    parts = ctx.colsplit(D, 5)
    assert sum(p.nnz for p in parts) == D.nnz and sum(p.shape[1] for p in parts) == A.shape[1]
    # ColSplit rule (dcsc.cpp:1202): floor(n/parts) columns each, remainder to the last
    assert [p.shape[1] for p in parts] == [A.shape[1] // 5] * 4 + [A.shape[1] - 4 * (A.shape[1] // 5)]
    J = ctx.colconcat(parts)
    assert ctx.checksum(J) == base
    m, n, jc, cp, ir, numx = ctx.download(J)
    ref = to_dcsc(A, np.float64)
    assert np.array_equal(jc, ref.jc) and np.array_equal(cp, ref.cp) and np.array_equal(ir, ref.ir) and np.array_equal(numx, ref.numx)
    S = ctx.colslice(D, 100, 900)
    m, n, jc, cp, ir, numx = ctx.download(S, np.int32)
    want = ref.colslice(100, 900)
    assert n == 800 and np.array_equal(jc, want.jc) and np.array_equal(cp, want.cp) and np.array_equal(ir, want.ir)
    # slab-wise multiply == whole multiply (what MemEfficientSpGEMM's phases do, ParFriends.h:553-772)
    whole = ctx.spgemm(0, D, D)
    slabs = [ctx.spgemm(0, D, p) for p in parts]
    joined = ctx.colconcat(slabs)
    a, b = ctx.checksum(whole), ctx.checksum(joined)
    assert a[0] == b[0]
    r1, c1, v1 = ctx.download_coo(whole)
    r2, c2, v2 = ctx.download_coo(joined)
    assert np.array_equal(r1, r2) and np.array_equal(c1, c2) and np.allclose(v1, v2, rtol=1e-12, atol=0)


def test_device_generator_equals_host_generator(ctx):
    for scale, ef, mode, dt in ((10, 16, 0, cb.F64), (12, 8, 1, cb.BOOL), (11, 8, 2, cb.I64)):
        D = ctx.gen_rmat(scale, ef << scale, seed=5, dtype=dt, value_mode=mode)
        A = rmat(scale, ef, seed=5)
        m, n, jc, cp, ir, numx = ctx.download(D)
        ref = to_dcsc(A, np.float64)
        assert np.array_equal(jc, ref.jc) and np.array_equal(cp, ref.cp) and np.array_equal(ir, ref.ir)
        if mode == 0:
            assert np.array_equal(numx, ref.numx)
        elif mode == 1:
            assert np.all(numx == 1)
        else:
            assert np.array_equal(numx, 1 + ir)


def test_er_config_full_size(ctx, port_oracle):
    """BASELINE config 2 at full size: Erdos-Renyi n = 2^22, d = 8, A bool, B int64 = 1 + row id,
    SelectMaxSRing<bool,int64_t>, bit-exact against the CPU oracle."""
    scale, d = 22, 8
    dA = ctx.gen_rmat(scale, d << scale, seed=2, a=0.25, b=0.25, c=0.25, scramble=False, dtype=cb.BOOL, value_mode=1)
    dB = ctx.gen_rmat(scale, d << scale, seed=2, a=0.25, b=0.25, c=0.25, scramble=False, dtype=cb.I64, value_mode=2)
    D, st = ctx.spgemm(cb.SelectMaxSRing_bool_i64, dA, dB, want_stats=True)
    m, n, jc, cp, ir, numx = ctx.download(dA)
    a = Csc(m, n, cb.SpDCCols(m, n, jc, cp, ir, numx).to_csc()[0], ir, numx)
    m, n, jc, cp, ir, vb = ctx.download(dB)
    b = Csc(m, n, a.colptr, ir, vb)
    want = port_oracle.spgemm(a, b, 3)
    rows, cols, vals = ctx.download_coo(D)
    assert st.nnz_out == want.nnz and st.flops > 2.5e8
    assert np.array_equal(rows, want.rows) and np.array_equal(cols, want.cols_expanded()) and np.array_equal(vals, want.vals)


def test_rmat_scale16_config1(ctx, oracle):
    """BASELINE config 1: R-MAT scale 16, edge factor 16, A^2, PlusTimes<double> against the reference's CPU kernel"""
    A = rmat(16, 16, seed=1)
    check_pair(ctx, oracle, 0, A, A)


def test_large_scale_properties(ctx):
    """size-independent properties at a size beyond the oracle-checked cases (R-MAT scale 18, 1.3e9 outputs): slab-wise product == whole
    product (checksums), symbolic == numeric counts, every column sorted and duplicate free"""
    G = ctx.gen_rmat(18, 16 << 18, seed=4)
    C, st = ctx.spgemm(0, G, G, want_stats=True)
    f, z = ctx.symbolic(G, G)
    assert (f, z) == (st.flops, st.nnz_out) and C.nnz == z
    slabs = [ctx.spgemm(0, G, Bs) for Bs in ctx.colsplit(G, 4)]
    J = ctx.colconcat(slabs)
    assert ctx.checksum(J)[0] == ctx.checksum(C)[0] and J.nnz == C.nnz
    for s in slabs:
        s.free()
    m, n, jc, cp, ir, numx = ctx.download(J, np.int32)
    J.free()
    assert np.all(np.diff(jc) > 0) and np.all(np.diff(cp) > 0)
    d = np.diff(ir.astype(np.int64))
    interior = np.ones(len(ir) - 1, dtype=bool)
    interior[cp[1:-1].astype(np.int64) - 1] = False  # positions where one column ends and the next begins
    assert np.all(d[interior] > 0), "rows must be strictly ascending inside every column"
    # all values of A are positive integers (multiplicities): every entry of A^2 is a positive integer too
    assert np.all(numx >= 1) and np.all(numx == np.floor(numx))
    C.free()


def test_failed_allocation_leaks_nothing(ctx):
    """a product whose result cannot be allocated (R-MAT scale 21 squared: about 350 GB) fails with NOMEM after the symbolic
    pass and gives every temporary back: the memory the library holds is the same before and after, and the phased multiply
    the failure calls for runs right away"""
    from combblas_b200.lib import CbgpuError

    G = ctx.gen_rmat(21, 16 << 21, 1)
    warm = ctx.colslice(G, 0, 64)  # a first multiply builds the per-matrix caches of G as an A operand (they stay, by design)
    ctx.spgemm(0, G, warm).free()
    warm.free()
    before = ctx.memory_in_use()
    with pytest.raises(CbgpuError) as e:
        ctx.spgemm(0, G, G)
    assert e.value.code == -3  # CBGPU_ERR_NOMEM
    assert ctx.memory_in_use() == before
    slabs = ctx.colsplit(G, 64)
    C0 = ctx.spgemm(0, G, slabs[0])
    assert C0.nnz > 0
    C0.free()
    for s in slabs:
        s.free()
    G.free()


@pytest.mark.parametrize("dt", [np.float64, np.float32, np.int64, np.uint8])
def test_device_transpose(ctx, dt):
    """cbgpu_mat_transpose == SpDCCols::Transpose (SpDCCols.cpp:871): pattern, values, rows ascending per column"""
    rng = np.random.default_rng(17)
    for (m, n, d) in [(300, 170, 0.05), (1, 40, 0.5), (50, 1, 0.5), (2000, 3000, 0.002), (64, 64, 0.0)]:
        M = typed(sp.random(m, n, density=d, random_state=rng, format="csc"), dt)
        dM = ctx.upload(to_dcsc(M, dt))
        dT = ctx.transpose(dM)
        rows, cols, vals = ctx.download_coo(dT)
        W = M.T.tocsc()
        W.sort_indices()
        assert dT.shape == (n, m) and len(rows) == W.nnz
        assert np.array_equal(rows, W.indices) and np.array_equal(cols, np.repeat(np.arange(m), np.diff(W.indptr)))
        assert np.array_equal(vals, W.data.astype(dt))
        # and it is a usable operand: (M^T) (x) M through the multiply
        if dt == np.float64 and M.nnz:
            D = ctx.spgemm(0, dT, dM)
            assert D.nnz == (M.T @ M).nnz
            D.free()
        dM.free()
        dT.free()

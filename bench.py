#!/usr/bin/env python
"""bench.py -- semiring SpGEMM GFLOP/s and fraction of the HBM roofline for R-MAT A^2 (PlusTimes<double>).

A "step" is one pass of the hot path: C = A (x) A over PlusTimesSRing<double,double> for a seeded R-MAT
(Graph500 initiator, edge factor 16, scrambled, duplicates summed), via the C ABI of libcbgpu.so.

  value      : whole-job GFLOP/s (2 * products / time), operands resident in HBM, result left in HBM (DCSC, rows
               ascending), timed with CUDA events on the launching stream, max over ranks.
  e2e        : same metric through the host-buffer entry points: pinned-host DCSC operands are copied H2D inside the timed
               region, result essentials read back (the product itself, 864 GB at scale 22, stays on the device and is
               consumed slab by slab, as MemEfficientSpGEMM's caller consumes it). With several ranks every rank uploads ITS
               blocks of A and B from pinned host memory inside the timed region and runs the same phased SUMMA (wall clock
               between barriers, max over ranks).
  roofline   : algorithmic bytes (SURVEY.md section 8d formula with the device layout sI=4, colptr 8) / time, against the
               measured HBM copy bandwidth in MEASURED_PEAKS.json; per-kernel-class breakdown from events inside the library.
  parity     : after the timed steps one more step of the SAME code path runs with checksums: order-independent 64-bit sums
               over (global row, global column) and over the value bits of every entry of C, added up over slabs and ranks.
               They are grid independent, so N = 2, 4, 8 must reproduce the N = 1 value, which is committed in
               tests/golden/bench_checksums.json (made by tools/make_bench_golden.py, pinned against the oracle there).
  cpu_baseline / --impl reference : the unmodified reference's LocalHybridSpGEMM (oracle/_ref) on the host cores on a bounded
               sample OF THE SAME WORKLOAD: A (x) A(:, J) for seeded column ranges J of the same scale-22 matrix; at N = 1
               the GPU result for the same columns is compared with it (nnz + checksums).

N GPUs: the same global matrix (strong scaling): 2 = 1x1x2 layers, 4 = 2x2x1, 8 = 2x2x2 (cbgpu_summa_phased: NCCL
broadcasts along grid rows/columns, inputs replicated along the fiber, one stacked local multiply per slab). Every rank
generates ITS blocks only (cbgpu_gen_rmat_block); the phase count comes from a distributed symbolic pass.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

A_, B_, C_ = 0.57, 0.19, 0.19  # Graph500 initiator (3DSpGEMM/mpipspgemm.cpp:126-133)
EDGEFACTOR = 16
SEED = 1
METRIC = "semiring SpGEMM GFLOP/s (R-MAT A^2, PlusTimesSRing<double,double>)"
SAMPLE_COLS = 16384  # width of one sampled column range of the CPU baseline (a MemEfficientSpGEMM phase is a few times wider)
GOLDEN = os.path.join(ROOT, "tests", "golden", "bench_checksums.json")
GENERATOR_NOTE = ("noiseless R-MAT (no Graph500 per-level noise), own seeded scramble: 2.7x the products per input nonzero of "
                  "the reference's Graph500 generator at equal scale; rates are not comparable with BASELINE.md section 2")


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.thread = None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=6)
        sm, smax, reasons = [], 0.0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                smax = max(smax, float(s[1]))
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for nm, v in zip(names, s[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def golden_key(scale):
    return f"rmat_s{scale}_ef{EDGEFACTOR}_seed{SEED}"


def load_golden(scale):
    try:
        return json.load(open(GOLDEN)).get(golden_key(scale))
    except Exception:
        return None


def sample_ranges(n, count, width=SAMPLE_COLS, seed=12345):
    """seeded, disjoint column ranges of the benchmark matrix (the vertices are scrambled, so a contiguous range is a random
    sample of the columns)"""
    width = min(width, n)
    slots = n // width
    rng = np.random.default_rng(seed)
    picks = rng.permutation(slots)[:max(1, min(count, slots))]
    return [(int(s) * width, int(s) * width + width) for s in picks]


# ---------------------------------------------------------------------------------------------- CPU side (checker / baseline)
class CpuReference:
    """the reference's own CPU implementation (oracle/_ref when built, else the C restatement) on column ranges of the
    benchmark matrix; the matrix is built on the host cores by the oracle's copy of the generator"""

    def __init__(self, scale):
        from oracle.oracle import PortOracle, RefOracle, rmat_csc

        self.kind = "reference" if RefOracle.available() else "port"
        self.orc = RefOracle() if self.kind == "reference" else PortOracle()
        self.cores = os.cpu_count() or 1
        self.orc.set_num_threads(self.cores)
        self.scale = scale
        t0 = time.time()
        self.a = rmat_csc(scale, EDGEFACTOR, SEED, A_, B_, C_, True)
        self.deg = np.diff(self.a.colptr)
        self.build_s = time.time() - t0

    def slab(self, c0, c1):
        from oracle.oracle import Csc

        a = self.a
        p0, p1 = int(a.colptr[c0]), int(a.colptr[c1])
        return Csc(a.m, c1 - c0, a.colptr[c0:c1 + 1] - a.colptr[c0], a.rows[p0:p1], a.vals[p0:p1])

    def products(self, c0, c1):
        a = self.a
        return int(self.deg[a.rows[int(a.colptr[c0]):int(a.colptr[c1])]].sum())

    def multiply(self, c0, c1):
        """A (x) A(:, c0:c1) -> (Csc with local column ids, seconds inside the reference call)"""
        from oracle.oracle import REF_LOCAL_HYBRID

        b = self.slab(c0, c1)
        if self.kind == "reference":
            return self.orc.spgemm(self.a, b, 0, REF_LOCAL_HYBRID, canonical=False, want_time=True)
        return self.orc.spgemm(self.a, b, 0, want_time=True)


def cpu_leg(scale, budget_s, ranges_wanted, gpu_check=None):
    """times the reference on sampled column ranges of the benchmark product; gpu_check(c0, c1) -> (nnz, pattern, value)
    lets the caller compare the device result for the same columns. Returns (cpu_baseline dict, sample parity dict)."""
    from oracle.oracle import matrix_checksum

    ref = CpuReference(scale)
    n = 1 << scale
    secs, mults, checked, ok = 0.0, 0, 0, True
    used = []
    t_begin = time.time()
    for (c0, c1) in sample_ranges(n, ranges_wanted):
        out, sec = ref.multiply(c0, c1)
        secs += sec
        mults += ref.products(c0, c1)
        used.append([c0, c1])
        if gpu_check is not None:
            want = (out.nnz,) + matrix_checksum(out.rows, out.cols_expanded(), out.vals, 0, c0)
            got = gpu_check(c0, c1)
            checked += 1
            ok = ok and tuple(int(x) for x in got) == tuple(int(x) for x in want)
        del out
        if time.time() - t_begin > budget_s:
            break
    cpu = {"value": 2.0 * mults / secs / 1e9, "unit": "GFLOP/s", "cores": ref.cores, "kind": ref.kind,
           "sample": f"same workload (R-MAT scale {scale} ef {EDGEFACTOR}): A (x) A(:, J) for {len(used)} seeded column range(s) of "
                     f"{SAMPLE_COLS} columns, products={mults}, LocalHybridSpGEMM {secs:.3f} s in the call "
                     f"(+ {ref.build_s:.1f} s building A on the host, not counted)",
           "seconds": secs, "mults": mults, "ranges": used}
    par = None
    if gpu_check is not None:
        par = {"column_ranges_checked": checked, "equal_to_reference": bool(ok),
               "what": "nnz + pattern checksum + value checksum of C(:, J) from the device vs the reference's output for the same J "
                       "(integer-valued inputs: the f64 sums are exact, so the value checksum is order independent)"}
    return cpu, par


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    from oracle.oracle import matrix_checksum  # noqa: F401  (fails early if the checker is not built)

    ref = CpuReference(args.scale)
    n = 1 << args.scale
    ranges = sample_ranges(n, args.warmup + args.steps)
    times, mults = [], []
    for i, (c0, c1) in enumerate(ranges):
        out, sec = ref.multiply(c0, c1)
        del out
        if i >= min(args.warmup, len(ranges) - 1):
            times.append(sec)
            mults.append(ref.products(c0, c1))
        if time.time() - t0 > 240 and times:
            break
    t = float(np.sum(times))
    m = int(np.sum(mults))
    val = 2.0 * m / t / 1e9
    sample = (f"same workload (R-MAT scale {args.scale} ef {EDGEFACTOR}): each step = A (x) A(:, J) for one seeded range of "
              f"{SAMPLE_COLS} columns; {len(times)} timed step(s), products={m}, {t:.3f} s inside LocalHybridSpGEMM")
    line = {"metric": METRIC, "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": len(times),
            "warmup": args.warmup, "ms_per_step": t / max(1, len(times)) * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"R-MAT scale {args.scale} ef {EDGEFACTOR} A^2 PlusTimesSRing<double,double> "
                                   f"(reference arm: column-range samples of the same product)", "generator": GENERATOR_NOTE},
            "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": ref.cores, "kind": ref.kind, "sample": sample},
            "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.time() - t0}
    print(json.dumps(line), flush=True)


def ncu_traffic(scale):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
    capture of this workload (profiles/r2_ncu_summary_s<scale>_dominant.txt); None otherwise."""
    for name in (f"r2_ncu_summary_s{scale}_dominant.txt", f"r1_ncu_summary_s{scale}_dominant.txt"):
        p = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(p):
            continue
        rd = wr = None
        for line in open(p):
            f = line.split()
            if len(f) >= 3 and f[0] == "dram__bytes_read.sum":
                rd = float(f[1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(f[2], 1)
            if len(f) >= 3 and f[0] == "dram__bytes_write.sum":
                wr = float(f[1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(f[2], 1)
        if rd is not None and wr is not None:
            return rd + wr, name
    return None, None


def bytes_alg(nnzA, nzcA, nnzB, nzcB, nnzC, nzcC, sv=8):
    """each operand read once, result written once, compressed-column form. Device layout:
    row ids 4 B (SpDCCols<int32_t,...> local indices), values sv B, jc+cp 16 B per non-empty column."""
    return (nnzA + nnzB) * (4 + sv) + (nzcA + nzcB) * 16 + nnzC * (4 + sv) + nzcC * 16


M64 = (1 << 64) - 1


def main():
    # exactly one JSON line may reach stdout: libraries (NCCL banner, ...) that print there are sent to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=int, default=int(os.environ.get("CBGPU_BENCH_SCALE", "22")))
    ap.add_argument("--phases", type=int, default=0, help="column slabs of B/C per step (0 = automatic from the symbolic pass)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed verification step")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (tuning experiments)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    import combblas_b200 as cb
    from combblas_b200 import lib as cblib
    from combblas_b200.host import local_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()  # a real (non-legacy) stream: the library launches on it, torch events time it
    torch.cuda.set_stream(stream)
    ctx = cb.Context(local_rank, stream=stream.cuda_stream)
    for o in args.opt:
        k_, v_ = o.split("=")
        ctx.set_option(k_, int(v_))
    scale = args.scale
    n = 1 << scale
    nedges = EDGEFACTOR << scale
    layers = {1: 1, 2: 2, 4: 1, 8: 2}.get(world)
    if layers is None:
        raise SystemExit("supported GPU counts: 1, 2 (1x1x2), 4 (2x2x1), 8 (2x2x2)")

    # ---- inputs: generated on the device, outside the timed region; with several ranks every rank builds its own blocks
    comm = None
    row_off = col_off = 0
    if world == 1:
        G = ctx.gen_rmat(scale, nedges, SEED, A_, B_, C_, True, cb.F64, 0)
        Aloc, Bloc = G, G
        nnz_global = int(G.info().nnz)
    else:
        grid = cblib.make_grid(world, rank, layers)
        idbuf = [cb.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(idbuf, src=0)
        comm = cb.Comm(ctx, grid, idbuf[0])
        r0, r1, c0, c1 = local_range(grid, n, n, True)   # A: column-split across layers (SpParMat3D.cpp:337-402)
        Aloc = ctx.gen_rmat_block(scale, nedges, SEED, r0, r1, c0, c1, A_, B_, C_, True, cb.F64, 0)
        row_off = r0
        r0, r1, c0, c1 = local_range(grid, n, n, False)  # B: row-split across layers
        Bloc = ctx.gen_rmat_block(scale, nedges, SEED, r0, r1, c0, c1, A_, B_, C_, True, cb.F64, 0)
        col_off = c0
        t = torch.tensor([Aloc.info().nnz], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        nnz_global = int(t.item())
    ainfo, binfo = Aloc.info(), Bloc.info()

    # ---- phases (MemEfficientSpGEMM's column slabs of B, ParFriends.h:553-772): C is produced slab by slab when the whole
    #      product would not fit in HBM. The count comes from the exact symbolic pass (CalculateNumberOfPhases' role, :780-843):
    #      single GPU cbgpu_spgemm_symbolic, several ranks cbgpu_summa_symbolic (max over ranks of what a rank will hold).
    phases = args.phases
    sym = None
    if phases <= 0:
        if world == 1:
            f_sym, nnz_sym = ctx.symbolic(Aloc, Bloc)
            sym = {"products": int(f_sym), "nnz": int(nnz_sym)}
            phases = max(1, int(np.ceil(nnz_sym * 12 / 48e9)))
        else:
            f_sym, nnz_sym = comm.summa_symbolic(cb.PlusTimesSRing_f64, Aloc, Bloc)
            t = torch.tensor([nnz_sym], dtype=torch.int64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sym = {"products_rank0": int(f_sym), "nnz_max_per_rank": int(t.item())}
            phases = max(1, int(np.ceil(int(t.item()) * 12 / 40e9)))
    phases = max(1, phases)
    slabs = ctx.colsplit(Bloc, phases) if (world == 1 and phases > 1) else None
    per = n // phases  # ColSplit rule (dcsc.cpp:1202): floor(n / phases) columns each, the last slab takes the rest

    class SlabResult:
        """what a phased step leaves behind: essentials and checksums summed over the slabs (the slabs are consumed)"""

        def __init__(self):
            self.nnz = 0
            self.nzc = 0
            self.check = [0, 0]

        def info(self):
            return self

        def free(self):
            pass

    def add_stats(acc, st):
        if acc is None:
            return st
        for name, _ in st._fields_:
            v = getattr(st, name)
            if isinstance(v, (int, float)):
                setattr(acc, name, getattr(acc, name) + v)
            elif hasattr(v, "_fields_"):
                pass  # nested struct: summed by the caller
            else:
                for i in range(len(v)):
                    getattr(acc, name)[i] += v[i]
        return acc

    def step(verify=False):
        res, acc = SlabResult(), None
        if world == 1:
            parts = slabs if phases > 1 else [Bloc]
            for i, Bs in enumerate(parts):
                Cs, st = ctx.spgemm(cb.PlusTimesSRing_f64, Aloc, Bs, want_stats=True)
                inf = Cs.info()
                res.nnz += inf.nnz
                res.nzc += inf.nzc
                if verify:
                    p_, v_ = ctx.checksum(Cs, 0, per * i)
                    res.check[0] = (res.check[0] + p_) & M64
                    res.check[1] = (res.check[1] + v_) & M64
                Cs.free()  # the slab is consumed (a HipMCL-style caller prunes it here); its essentials were read back
                acc = add_stats(acc, st)
            return res, acc, None
        # distributed: cbgpu_summa_phased = MemEfficientSpGEMM / MemEfficientSpGEMM3D without the pruning (one SUMMA per
        # column slab of B, slabs consumed as they finish)
        results, _, ds = comm.summa_phased(cb.PlusTimesSRing_f64, Aloc, Bloc, phases,
                                           global_offsets=(row_off, col_off) if verify else None)
        res.nnz = sum(r.nnz for r in results)
        res.nzc = sum(r.nzc for r in results)
        if verify:
            res.check = [sum(r.pattern_sum for r in results) & M64, sum(r.value_sum for r in results) & M64]
        return res, ds.local, ds

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also grows the stream-ordered memory pool to its steady state)
    for _ in range(args.warmup):
        Cd, st, ds = step()
        Cd.free()
    barrier()
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = {}
    last = None
    barrier()
    t_wall0 = time.time()
    for i in range(args.steps):
        flush.fill_(i)  # L2 flush between timed iterations (outside the event pair)
        if world > 1:
            dist.barrier()
        ev[i][0].record(stream)
        Cd, st, ds = step()
        ev[i][1].record(stream)
        torch.cuda.synchronize()
        d = st.as_dict()
        for k, v in d["ms_kernel"].items():
            kernel_ms[k] = kernel_ms.get(k, 0.0) + v / args.steps
        last = (Cd, st, ds)
    barrier()
    t_wall = time.time() - t_wall0
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop()
    ms = [a.elapsed_time(b) for a, b in ev]
    ms_step = float(np.mean(ms))
    Cd, st, ds = last
    cinfo = Cd.info()
    mults_local = int(st.flops)
    nnzC_local = int(cinfo.nnz)

    # ---- aggregate over ranks: max time, summed work
    if world > 1:
        t = torch.tensor([ms_step], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = float(t.item())
        w = torch.tensor([mults_local, nnzC_local, cinfo.nzc, ainfo.nnz, ainfo.nzc, binfo.nnz, binfo.nzc, launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        mults, nnzC, nzcC, nnzA, nzcA, nnzB, nzcB, launches = [int(x) for x in w.tolist()]
    else:
        mults, nnzC, nzcC = mults_local, nnzC_local, int(cinfo.nzc)
        nnzA, nzcA, nnzB, nzcB = int(ainfo.nnz), int(ainfo.nzc), int(binfo.nnz), int(binfo.nzc)
    gflops = 2.0 * mults / (ms_step * 1e-3) / 1e9

    # ---- parity: one untimed step of the same code path with checksums at global positions, summed over slabs and ranks
    parity = None
    if not args.no_parity:
        Cv, _, _ = step(verify=True)
        mine = [int(Cv.nnz), int(Cv.check[0]), int(Cv.check[1])]
        if world > 1:
            # 64-bit unsigned sums: gather the three words of every rank and add modulo 2^64 on the host
            t = torch.tensor([x if x < (1 << 63) else x - (1 << 64) for x in mine], dtype=torch.int64, device="cuda")
            allv = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allv, t)
            tot = [0, 0, 0]
            for tv in allv:
                for j, x in enumerate(tv.tolist()):
                    tot[j] = (tot[j] + (int(x) & M64)) & M64
        else:
            tot = mine
        gold = load_golden(scale)
        parity = {"nnz_C": tot[0], "pattern_sum": f"{tot[1]:016x}", "value_sum": f"{tot[2]:016x}",
                  "how": "extra untimed step with cbgpu_mat_checksum_at / cbgpu_summa_phased_global: sums over (global row, global "
                         "col) and value bits of all entries of C, over all slabs and ranks; grid independent",
                  "golden": None if gold is None else gold.get("source"),
                  "equals_single_gpu_golden": None if gold is None else bool(
                      tot[0] == gold["nnz_C"] and f"{tot[1]:016x}" == gold["pattern_sum"] and f"{tot[2]:016x}" == gold["value_sum"]),
                  "products_equal_golden": None if gold is None else bool(mults == gold["products"])}

    # ---- roofline of the multiply (all kernel classes of one call) + per-class breakdown
    peak, peak_src = measured_peak_gbs()
    balg = bytes_alg(nnzA, nzcA, nnzB, nzcB, nnzC, nzcC)
    achieved = balg / (ms_step * 1e-3) / 1e9 / world  # per GPU
    dominant = max(kernel_ms.items(), key=lambda kv: kv[1]) if kernel_ms else ("-", 0.0)
    sd = st.as_dict()
    traffic, traffic_src = ncu_traffic(scale) if world == 1 else (None, None)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic,
                "traffic_note": None if traffic is None else f"DRAM bytes of ONE launch of the dominant kernel (one column slab), ncu --set full, profiles/{traffic_src}",
                "peak_source": peak_src, "scope": "one step = all kernel classes of all slabs",
                "algorithmic_bytes": balg, "kernel_ms": {k: round(v, 4) for k, v in sorted(kernel_ms.items())},
                "dominant_kernel": dominant[0], "dominant_kernel_ms": round(dominant[1], 4)}
    cls = (sd.get("classes") or {}).get(dominant[0])
    if cls and mults_local > 0 and dominant[1] > 0:
        # algorithmic bytes of the dominant class: its outputs written once + its share of the operand reads
        share = cls["flops"] / max(1, mults_local)
        dom_bytes = cls["nnz"] * 12 + share * ((ainfo.nnz + binfo.nnz) * 12 + (ainfo.nzc + binfo.nzc) * 16)
        roofline["dominant_kernel_achieved_gbs"] = dom_bytes / (dominant[1] * 1e-3) / 1e9
        roofline["dominant_kernel_frac"] = roofline["dominant_kernel_achieved_gbs"] / peak

    # ---- e2e through the host-buffer entry points (N = 1): pinned host DCSC -> H2D -> multiply -> read-back
    e2e = None
    if world == 1 and not args.no_e2e:
        m_, n_, jc, cp, ir, numx = ctx.download(Aloc)  # int64 indices, as SpDCCols<int64_t,double>
        host = [torch.from_numpy(x).pin_memory() for x in (jc, cp, ir, numx)]
        Ah = cb.SpDCCols(m_, n_, *[h.numpy() for h in host])
        h2d = 2 * sum(h.numel() * h.element_size() for h in host)
        es = []
        for i in range(2 + min(args.steps, 3)):
            flush.fill_(i)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if phases == 1:
                Ce = ctx.spgemm_host(cb.PlusTimesSRing_f64, Ah, Ah)
                inf = Ce.info()
                Ce.free()
            else:  # phased: operands go up once per step, B is cut into slabs on the device, C is consumed slab by slab
                dA, dB = ctx.upload(Ah), ctx.upload(Ah)
                for Bs in ctx.colsplit(dB, phases):
                    Cs = ctx.spgemm(cb.PlusTimesSRing_f64, dA, Bs)
                    inf = Cs.info()
                    Cs.free()
                    Bs.free()
                dA.free()
                dB.free()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if i >= 2:
                es.append(dt)
        te = float(np.mean(es))
        e2e = {"value": 2.0 * mults / te / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 48 * phases,
               "ms_per_step": te * 1e3,
               "what": ("cbgpu_spgemm_local_host" if phases == 1 else "cbgpu_mat_upload x2 + cbgpu_mat_colsplit + cbgpu_spgemm_local per slab")
                       + ": pinned host int64/f64 DCSC of A and B copied H2D inside the timed region, multiply, essentials of every slab read "
                         "back; the product stays on the device and is consumed slab by slab (864 GB at scale 22 cannot leave it)"}
    elif world > 1 and not args.no_e2e:
        # every rank starts from ITS blocks of A and B in pinned host memory (what an MPI rank of the reference holds), copies them
        # H2D inside the timed region, runs the same phased SUMMA and reads the essentials of every slab back; max over ranks
        def pinned_block(M):
            m_, n_, jc, cp, ir, numx = ctx.download(M)
            host = [torch.from_numpy(x).pin_memory() for x in (jc, cp, ir, numx)]
            return cb.SpDCCols(m_, n_, *[h.numpy() for h in host]), host

        Ah, keepA = pinned_block(Aloc)
        Bh, keepB = pinned_block(Bloc)
        h2d_local = sum(h.numel() * h.element_size() for h in keepA + keepB)
        es = []
        for i in range(2 + min(args.steps, 3)):
            flush.fill_(i)
            barrier()
            t0 = time.perf_counter()
            dA, dB = ctx.upload(Ah), ctx.upload(Bh)
            results, _, _ = comm.summa_phased(cb.PlusTimesSRing_f64, dA, dB, phases)
            nnz_e2e = sum(r.nnz for r in results)
            dA.free()
            dB.free()
            torch.cuda.synchronize()
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if i >= 2:
                es.append(float(t.item()))
        w = torch.tensor([h2d_local, nnz_e2e], dtype=torch.int64, device="cuda")
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        te = float(np.mean(es))
        e2e = {"value": 2.0 * mults / te / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": int(w[0].item()),
               "d2h_bytes_per_step": 48 * phases * world, "ms_per_step": te * 1e3, "nnz_C": int(w[1].item()),
               "what": "per rank: cbgpu_mat_upload of its A and B block (pinned host int64/f64 DCSC, H2D inside the timed region) + "
                       "cbgpu_summa_phased over NCCL + essentials of every slab read back; wall clock between barriers, max over ranks; "
                       "the product stays on the devices and is consumed slab by slab"}
    elif world > 1:
        e2e = {"value": None, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "what": "skipped (--no-e2e)"}

    # ---- CPU baseline on rank 0: the reference on sampled column ranges of the same product; at N = 1 the device result for
    #      the same columns is compared with it
    cpu = None
    sample_parity = None
    if rank == 0 and not args.no_cpu:
        try:
            def gpu_cols(c0, c1):
                Bs = ctx.colslice(Bloc, c0, c1)
                Cs = ctx.spgemm(cb.PlusTimesSRing_f64, Aloc, Bs)
                out = (Cs.info().nnz,) + tuple(ctx.checksum(Cs, 0, c0))
                Cs.free()
                Bs.free()
                return out

            cpu, sample_parity = cpu_leg(scale, 25.0, 3, gpu_cols if world == 1 else None)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # the checker is optional equipment; never let it sink the measurement
            cpu = {"value": None, "unit": "GFLOP/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    if parity is not None and sample_parity is not None:
        parity["reference_sample"] = sample_parity

    if rank == 0:
        grid_name = {1: "1 GPU", 2: "1x1x2 (3D, 2 layers)", 4: "2x2x1 (2D SUMMA)", 8: "2x2x2 (3D SUMMA)"}[world]
        line = {"metric": METRIC, "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": f"R-MAT scale {scale} ef {EDGEFACTOR} A^2 PlusTimesSRing<double,double>, grid {grid_name}",
                           "n": n, "nnz_A": nnz_global, "products": mults, "nnz_C": nnzC, "compression": mults / max(1, nnzC),
                           "l2": "256 MiB flush write between timed iterations; operands+result also exceed L2",
                           "index_bytes": 4, "value_bytes": 8, "phases": phases, "symbolic": sym, "generator": GENERATOR_NOTE},
                "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu,
                "parity": parity,
                "ms_steps": [round(x, 3) for x in ms], "wall_s": round(t_wall, 3),
                "phases_ms": {"setup": round(st.ms_setup, 3), "symbolic": round(st.ms_symbolic, 3), "numeric": round(st.ms_numeric, 3)},
                "census": {k: sd[k] for k in sd if k.startswith(("tasks", "flops_", "nnz_"))},
                "classes": sd.get("classes")}
        if ds is not None:
            dd = ds.as_dict()
            line["dist_ms"] = {k: round(dd[k], 3) for k in ("ms_bcast", "ms_multiply", "ms_merge", "ms_fiber_exchange", "ms_fiber_merge", "ms_total")}
            line["dist_bytes"] = {"bcast": dd["bytes_bcast"], "fiber": dd["bytes_fiber"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        comm.destroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

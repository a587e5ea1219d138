#!/usr/bin/env python
"""bench.py -- semiring SpGEMM GFLOP/s and fraction of the HBM roofline for R-MAT A^2 (PlusTimes<double>).

A "step" is one pass of the hot path: C = A (x) A over PlusTimesSRing<double,double> for a seeded R-MAT
(Graph500 initiator, edge factor 16, scrambled, duplicates summed), via the C ABI of libcbgpu.so.

  value      : whole-job GFLOP/s (2 * products / time), operands resident in HBM, result left in HBM (DCSC, rows
               ascending), timed with CUDA events on the launching stream, max over ranks.
  e2e        : same metric through the reference-facing call with HOST buffers (cbgpu_spgemm_local_host):
               pinned-host DCSC operands are copied H2D inside the timed region, result essentials + checksum read back.
  roofline   : algorithmic bytes (SURVEY.md section 8d formula with the device layout sI=4, colptr 8) / time, against the measured
               HBM copy bandwidth in MEASURED_PEAKS.json; per-kernel-class breakdown from events inside the library.
  cpu_baseline / --impl reference : the unmodified reference's LocalHybridSpGEMM (oracle/_ref) on the host cores, on a
               bounded sample (a smaller R-MAT scale of the same family).

N GPUs: the same global matrix (strong scaling) on a 1x1xN... grid: 2 = 1x1x2 layers, 4 = 2x2, 8 = 2x2x2
(cbgpu_summa2d / cbgpu_summa3d: NCCL broadcasts along grid rows/columns, fiber all-to-all, device merges).
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


def rmat_host(scale, edgefactor=EDGEFACTOR, seed=SEED):
    """host copy of the library's seeded generator (identical integer arithmetic) -> scipy CSC, duplicates summed"""
    import scipy.sparse as sp

    import combblas_b200 as cb

    lib = cb.load_library()
    ne = edgefactor << scale
    rows = np.empty(ne, np.int64)
    cols = np.empty(ne, np.int64)
    assert lib.cbgpu_rmat_edges_host(scale, ne, seed, A_, B_, C_, 1, rows.ctypes.data, cols.ctypes.data) == 0
    n = 1 << scale
    M = sp.coo_matrix((np.ones(ne), (rows, cols)), shape=(n, n)).tocsc()
    M.sum_duplicates()
    M.sort_indices()
    return M


def cpu_reference_run(scale, steps=1, warmup=0, budget_s=25.0):
    """times the reference's own CPU implementation of the path (oracle/_ref when built, else the C port) on a
    bounded sample: R-MAT of the same family at `scale`. Returns dict with GFLOP/s."""
    from oracle.oracle import Csc, PortOracle, RefOracle, REF_LOCAL_HYBRID

    use_ref = RefOracle.available()
    orc = RefOracle() if use_ref else PortOracle()
    cores = os.cpu_count() or 1
    orc.set_num_threads(cores)
    M = rmat_host(scale)
    a = Csc.from_scipy(M, np.float64)
    mults = int(np.diff(M.indptr)[M.indices].sum())
    times = []
    t_begin = time.time()
    for i in range(warmup + steps):
        if use_ref:
            _, sec = orc.spgemm(a, a, 0, REF_LOCAL_HYBRID, canonical=False, want_time=True)
        else:
            _, sec = orc.spgemm(a, a, 0, want_time=True)
        if i >= warmup:
            times.append(sec)
        if time.time() - t_begin > budget_s and times:
            break
    t = float(np.mean(times))
    return {"value": 2.0 * mults / t / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": "reference" if use_ref else "port",
            "sample": f"R-MAT scale {scale} ef {EDGEFACTOR} A^2 (products={mults}), LocalHybridSpGEMM, {len(times)} run(s), "
                      f"{t:.3f} s each", "seconds": t, "mults": mults, "scale": scale}


def pick_cpu_scale(bench_scale):
    return min(bench_scale, int(os.environ.get("CBGPU_CPU_SCALE", "17")))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scale = pick_cpu_scale(args.scale)
    t0 = time.time()
    r = cpu_reference_run(scale, steps=max(1, args.steps), warmup=min(1, args.warmup), budget_s=120.0)
    line = {"metric": METRIC, "value": r["value"], "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"R-MAT scale {args.scale} ef {EDGEFACTOR} A^2 PlusTimes<double> "
                                   f"(reference arm: bounded sample at scale {scale})"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.time() - t0}
    print(json.dumps(line), flush=True)


def ncu_traffic(scale):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture of this same workload (profiles/r1_ncu_summary_s<scale>_dominant.txt); None otherwise."""
    p = os.path.join(ROOT, "profiles", f"r1_ncu_summary_s{scale}_dominant.txt")
    if not os.path.exists(p):
        return None
    rd = wr = None
    for line in open(p):
        f = line.split()
        if len(f) >= 3 and f[0] == "dram__bytes_read.sum":
            rd = float(f[1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(f[2], 1)
        if len(f) >= 3 and f[0] == "dram__bytes_write.sum":
            wr = float(f[1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(f[2], 1)
    return None if rd is None or wr is None else rd + wr


def bytes_alg(nnzA, nzcA, nnzB, nzcB, nnzC, nzcC, sv=8):
    """each operand read once, result written once, compressed-column form. Device layout:
    row ids 4 B (SpDCCols<int32_t,...> local indices), values sv B, jc+cp 16 B per non-empty column."""
    return (nnzA + nnzB) * (4 + sv) + (nzcA + nzcB) * 16 + nnzC * (4 + sv) + nzcC * 16


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
    layers = {1: 1, 2: 2, 4: 1, 8: 2}.get(world)
    if layers is None:
        raise SystemExit("supported GPU counts: 1, 2 (1x1x2), 4 (2x2x1), 8 (2x2x2)")

    # ---- inputs: generated on the device, outside the timed region
    G = ctx.gen_rmat(scale, EDGEFACTOR << scale, SEED, A_, B_, C_, True, cb.F64, 0)
    ginfo = G.info()
    comm = None
    if world == 1:
        Aloc, Bloc = G, G
    else:
        grid = cblib.make_grid(world, rank, layers)
        idbuf = [cb.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(idbuf, src=0)
        comm = cb.Comm(ctx, grid, idbuf[0])
        r0, r1, c0, c1 = local_range(grid, n, n, True)
        Aloc = ctx.submatrix(G, r0, r1, c0, c1)
        r0, r1, c0, c1 = local_range(grid, n, n, False)
        Bloc = ctx.submatrix(G, r0, r1, c0, c1)
        G.free()
    ainfo, binfo = Aloc.info(), Bloc.info()

    # ---- phases (MemEfficientSpGEMM's column slabs of B, ParFriends.h:553-772): C is produced slab by slab when the
    #      whole product would not fit in HBM; every slab stays resident until the step ends only if it fits.
    phases = args.phases
    if world == 1 and phases <= 0:
        f_sym, nnz_sym = ctx.symbolic(Aloc, Bloc)
        phases = max(1, int(np.ceil(nnz_sym * 12 / 48e9)))
    if world > 1 and phases <= 0:
        # nnz(C) of this R-MAT family grows ~7.6x per two scales (measured: 1.28e9 at scale 18, 9.7e9 at scale 20);
        # every rank of a layer holds 1/(pr*pc) of that layer's partial product
        est_nnz = 9.7e9 * 7.6 ** ((scale - 20) / 2.0)
        per_rank = est_nnz * 12 / (world // layers)
        phases = max(1, int(np.ceil(per_rank / 40e9)))
    phases = max(1, phases)
    slabs = ctx.colsplit(Bloc, phases) if (world == 1 and phases > 1) else None

    class SlabResult:
        """what a phased step leaves behind: per-slab essentials and checksums (the slabs themselves are consumed)"""

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

    def step():
        if world == 1 and phases > 1:
            res, acc = SlabResult(), None
            for Bs in slabs:
                Cs, st = ctx.spgemm(cb.PlusTimesSRing_f64, Aloc, Bs, want_stats=True)
                inf = Cs.info()
                res.nnz += inf.nnz
                res.nzc += inf.nzc
                Cs.free()  # the slab is consumed (a HipMCL-style caller prunes it here); its essentials were read back
                acc = add_stats(acc, st)
            return res, acc, None
        if world == 1:
            Cd, st = ctx.spgemm(cb.PlusTimesSRing_f64, Aloc, Bloc, want_stats=True)
            return Cd, st, None
        # distributed: cbgpu_summa_phased = MemEfficientSpGEMM / MemEfficientSpGEMM3D without the pruning (one SUMMA per
        # column slab of B, slabs consumed as they finish; with layers the fiber stage of slab p overlaps slab p+1)
        results, _, ds = comm.summa_phased(cb.PlusTimesSRing_f64, Aloc, Bloc, phases)
        res = SlabResult()
        res.nnz = sum(r.nnz for r in results)
        res.nzc = sum(r.nzc for r in results)
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
        if last is not None:
            last[0].free()  # the previous result goes back to the stream-ordered pool before the next step
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

    # ---- roofline of the multiply (all kernel classes of one call) + per-class breakdown
    peak, peak_src = measured_peak_gbs()
    balg = bytes_alg(nnzA, nzcA, nnzB, nzcB, nnzC, nzcC)
    achieved = balg / (ms_step * 1e-3) / 1e9 / world  # per GPU
    dominant = max(kernel_ms.items(), key=lambda kv: kv[1]) if kernel_ms else ("-", 0.0)
    sd = st.as_dict()
    # algorithmic bytes of the dominant numeric class: its outputs written once + its share of the operand reads
    dom_bytes = None
    cls = {"num_bitmap_gmem": ("nnz_bitmap_gmem", "flops_bitmap_gmem"), "num_bitmap_smem": ("nnz_bitmap_smem", "flops_bitmap_smem"),
           "num_hash_cta": ("nnz_hash_cta", "flops_hash_cta"), "num_hash_warp": ("nnz_hash_warp", "flops_hash_warp"),
           "num_hash_warp_small": ("nnz_hash_warp", "flops_hash_warp")}
    if dominant[0] in cls and mults_local > 0:
        kn, kf = cls[dominant[0]]
        share = sd[kf] / max(1, mults_local)
        dom_bytes = sd[kn] * 12 + share * ((ainfo.nnz + binfo.nnz) * 12 + (ainfo.nzc + binfo.nzc) * 16)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(scale) if world == 1 else None,
                "traffic_note": "DRAM bytes of ONE launch of num_bitmap_kernel<...,512> (one of the column slabs), ncu --set full, "
                                "profiles/r1_ncu_summary_s%d_dominant.txt" % scale, "peak_source": peak_src, "scope": "one cbgpu_spgemm_local call (all kernel classes)",
                "algorithmic_bytes": balg, "kernel_ms": {k: round(v, 4) for k, v in sorted(kernel_ms.items())},
                "dominant_kernel": dominant[0], "dominant_kernel_ms": round(dominant[1], 4)}
    if dom_bytes is not None and dominant[1] > 0:
        roofline["dominant_kernel_achieved_gbs"] = dom_bytes / (dominant[1] * 1e-3) / 1e9
        roofline["dominant_kernel_frac"] = roofline["dominant_kernel_achieved_gbs"] / peak

    # ---- e2e through the host-buffer entry point (N = 1): pinned host DCSC -> H2D -> multiply -> read-back
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
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(stream)
            if phases == 1:
                Ce = ctx.spgemm_host(cb.PlusTimesSRing_f64, Ah, Ah)
                chk = ctx.checksum(Ce)  # D2H read of the step's result: essentials + 2 x 64-bit checksum
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
            e1.record(stream)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if i >= 2:
                es.append(dt)
        te = float(np.mean(es))
        e2e = {"value": 2.0 * mults / te / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 16 + 32,
               "ms_per_step": te * 1e3, "what": "cbgpu_spgemm_local_host: pinned host int64/f64 DCSC of A and B copied H2D, multiply, "
                                                "result essentials + checksum read back; C stays resident in HBM"}
    elif world > 1:
        e2e = {"value": None, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "what": "distributed blocks are device-resident by design (no host staging on the SUMMA path)"}

    # ---- CPU baseline on rank 0, N = 1 only
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu:
        try:
            r = cpu_reference_run(pick_cpu_scale(scale), steps=1, warmup=0)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # the checker is optional equipment; never let it sink the measurement
            cpu = {"value": None, "unit": "GFLOP/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        grid_name = {1: "1 GPU", 2: "1x1x2 (3D, 2 layers)", 4: "2x2x1 (2D SUMMA)", 8: "2x2x2 (3D SUMMA)"}[world]
        line = {"metric": METRIC, "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": f"R-MAT scale {scale} ef {EDGEFACTOR} A^2 PlusTimesSRing<double,double>, grid {grid_name}",
                           "n": n, "nnz_A": int(ginfo.nnz), "products": mults, "nnz_C": nnzC, "compression": mults / max(1, nnzC),
                           "l2": "256 MiB flush write between timed iterations; operands+result also exceed L2",
                           "index_bytes": 4, "value_bytes": 8, "phases": phases},
                "clocks": clocks, "gpu_launches": int(launches), "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu,
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

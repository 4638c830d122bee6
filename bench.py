#!/usr/bin/env python
"""bench.py — 3-D Poisson GMG-CG solve on B200 (BASELINE.json metric).

A "step" is one full solve (solver:apply) of the synthetic FV1 Poisson problem:
GMG V(2,2) damped Jacobi (0.66) preconditioning CG, StdConvCheck(100, 1e-12, 1e-10).
  N = 1 : configs[1], unit cube, numRefs = 7, 129^3 = 2 146 689 DoF.
  N > 1 : weak scaling, one 129^3 sub-box per GPU (2x1x1, 2x2x1, 2x2x2 boxes; N = 8 is
          configs[2]: 257^3 = 16 974 593 DoF), interface exchange + all-reduce over NCCL.
Prints ONE JSON line (see the contract in the task description / DESIGN.md §Measurement).

`--workload convdiff` / `--workload elasticity` run BASELINE configs[3] / configs[4] instead (BiCGStab +
GMG with multicolour Gauss-Seidel on upwind convection-diffusion; CG + GMG block-Jacobi on 3x3-block
linear elasticity), same JSON contract, own metric names; the default stays configs[1] / configs[2].

`--scaling strong` fixes the GLOBAL grid instead of the per-GPU box: 2x2x2 base cells, numRefs = 7 -> 257^3 =
16 974 593 DoF (BASELINE configs[2]: the unit cube with numRefs = 8 minus its one-cell base level) on 1, 2, 4 or 8
GPUs; parallel efficiency = T1 / (P TP).  `--order hier` numbers the DoFs the way ugcore's global refinement does
(coarse vertices first), `--reorder cmk|rcmk` applies (reverse) Cuthill-McKee before upload.

`--impl reference` times ugcore's own CPU kernels (oracle/_ref: SparseMatrix/Vector/
smoother templates compiled from the reference) driving the restated solver loop on ALL
host cores; the reference has no threading on this path and no MPI is installed.
  N = 1 : every core runs one serial solve of the workload concurrently.
  N > 1 : the SAME global grid the N GPUs solve, partitioned over N processes per job with ugcore's parallel
          protocol (additive matrices, consistent / additive / unique vectors, gathered coarse levels; interface
          exchange through shared memory: oracle/partitioned.py), as many jobs side by side as cores and memory
          carry.  `--cpu-arm replicas` keeps the exchange-free replicas of one GPU's box (an upper bound).
This and `cpu_baseline` are the only places bench.py executes anything under oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "poisson3d_gmg_cg_mdof_per_s"
UNIT = "MDoF/s"
PART = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def solver_desc(top, base=0, workload="poisson", base_solver="lu"):
    if workload == "convdiff":     # BASELINE configs[3]: BiCGStab + GMG V(2,2), (multicolour) Gauss-Seidel smoothing
        return {"type": "bicgstab",
                "precond": {"type": "gmg", "topLevel": top, "baseLevel": base, "smoother": {"type": "gs", "relax": 1.0},
                            "cycle": "V", "preSmooth": 2, "postSmooth": 2, "baseSolver": base_solver},
                "convCheck": {"iterations": 100, "absolute": 1e-12, "reduction": 1e-8}}
    red, its = (1e-8, 200) if workload == "elasticity" else (1e-10, 100)
    return {"type": "cg",
            "precond": {"type": "gmg", "topLevel": top, "baseLevel": base, "smoother": {"type": "jac", "damp": 0.66},
                        "cycle": "V", "preSmooth": 2, "postSmooth": 2, "baseSolver": base_solver},
            "convCheck": {"iterations": its, "absolute": 1e-12, "reduction": red}}


def workload_spec(name):
    """problem id and keyword arguments of the generator, block size, metric name, label"""
    if name == "poisson":
        return {"problem": 0, "kw": {}, "block": 1, "metric": METRIC,
                "label": "3D Poisson", "method": "GMG V(2,2) damped-Jacobi(0.66) + CG, StdConvCheck(100, 1e-12, 1e-10)"}
    if name == "convdiff":
        return {"problem": 1, "kw": {"eps": 1e-1}, "block": 1, "metric": "convdiff3d_bicgstab_gmg_gs_mdof_per_s",
                "label": "3D convection-diffusion (FV1 full upwind, eps = 0.1, b = (1, 0.5, 0.25))",
                "method": "GMG V(2,2) multicolour Gauss-Seidel + BiCGStab, StdConvCheck(100, 1e-12, 1e-8)"}
    if name == "elasticity":
        return {"problem": 2, "kw": {}, "block": 3, "metric": "elasticity3d_gmg_cg_mdof_per_s",
                "label": "3D linear elasticity (Q1, E = 1, nu = 0.3, 3x3 block-CRS)",
                "method": "GMG V(2,2) damped block-Jacobi(0.66) + CG, StdConvCheck(200, 1e-12, 1e-8)"}
    raise SystemExit(f"unknown workload {name}")


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            d = json.load(f)
        for k in ("hbm_gbs", "hbm_gb_s", "hbm_copy_gbs"):
            if k in d:
                return float(d[k]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _cpu_replica(refs, steps, warmup, barrier, q, workload="poisson", base_mult=1, order=0):
    """One CPU process = what one ugcore MPI rank does on this path: a serial solve with the
    reference's own kernels (oracle/_ref) or the port.  Gauss-Seidel runs in ugcore's own
    (lexicographic) order here — the multicolour order is a property of the GPU path."""
    os.environ["OMP_NUM_THREADS"] = "1"
    import oracle
    from ugcore_b200 import problems as pr
    kind = "ref" if oracle.have_ref() else "port"
    orc = oracle.Oracle(kind)
    spec = workload_spec(workload)
    prob = pr.Problem(dim=3, num_refs=refs, problem=spec["problem"], base=(base_mult,) * 3, order=order, **spec["kw"])
    lv = {}
    for l in range(0, refs + 1):
        lv[l] = (orc.matrix(prob.matrix(l)), orc.matrix(prob.prolongation(l)) if l else None,
                 orc.matrix(prob.restriction(l)) if l else None)
    s = oracle.OSolver(orc, solver_desc(refs, workload=workload), lv[refs][0], lv)
    b = np.array(prob.rhs())
    for _ in range(warmup):
        s.apply(b)
    barrier.wait()
    t0 = time.perf_counter()
    h = None
    for _ in range(steps):
        x, ok, h = s.apply(b)
    dt = time.perf_counter() - t0
    q.put({"dt": dt, "its": len(h) - 1, "n": prob.num_dofs, "kind": kind, "hist": [float(v) for v in h]})


def replica_bytes(refs, workload="poisson", base_mult=1):
    """Rough host memory of one CPU replica: the generator's hierarchy plus the oracle's copy of it
    (27 entries per row, (8 B^2 + 4) bytes per entry, 8/7 for the coarser levels, transfers and vectors ~ +25 %)."""
    b = workload_spec(workload)["block"]
    n = (base_mult * 2 ** refs + 1) ** 3
    return int(2 * 1.25 * (8.0 / 7.0) * 27 * n * (8 * b * b + 4))


def bounded_procs(nproc, refs, workload="poisson", base_mult=1):
    """Never let the CPU arm exhaust the box: at most 40 % of the available memory over all replicas."""
    try:
        with open("/proc/meminfo") as f:
            avail = next(int(l.split()[1]) * 1024 for l in f if l.startswith("MemAvailable"))
    except Exception:
        return nproc
    return max(1, min(nproc, int(0.4 * avail / max(replica_bytes(refs, workload, base_mult), 1))))


def cpu_replicas(refs, nproc, steps, warmup, workload="poisson", base_mult=1, order=0):
    """ugcore has no threads on this path and neither MPI nor boost exist here, so
    "all host cores" = nproc independent serial solves of the same workload running concurrently
    (they share the memory bus like MPI ranks would, but pay no interface exchange: an upper bound
    for ugcore's nproc-rank weak-scaling throughput).  Returns aggregate MDoF/s and details."""
    import multiprocessing as mp
    nproc = bounded_procs(nproc, refs, workload, base_mult)
    mpc = mp.get_context("spawn")
    barrier, q = mpc.Barrier(nproc), mpc.Queue()
    procs = [mpc.Process(target=_cpu_replica, args=(refs, steps, warmup, barrier, q, workload, base_mult, order)) for _ in range(nproc)]
    for p in procs:
        p.start()
    res = [q.get(timeout=3600) for _ in range(nproc)]
    for p in procs:
        p.join()
    dt = max(r["dt"] for r in res)
    n = res[0]["n"]
    return {"value": nproc * steps * n / dt / 1e6, "dt_per_step": dt / steps, "its": res[0]["its"], "n": n,
            "kind": "reference" if res[0]["kind"] == "ref" else "port", "cores": nproc, "hist": res[0]["hist"]}


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_partitioned(refs, part, gbase, nproc, steps, warmup, workload="poisson", order=0):
    """The N-GPU grid on the CPU, partitioned like ugcore's MPI run (oracle/partitioned.py): the
    same process grid as the GPUs (the generator's base grid admits no finer one), one process per sub-box with its
    additive matrices, ugcore's consistent / additive / unique protocol, interface exchange through shared memory; as
    many such jobs side by side as the host cores and memory carry, so that all cores work.  Aggregate MDoF/s."""
    from oracle import partitioned
    spec = workload_spec(workload)
    world = int(np.prod(part))
    n_global = int(np.prod([gbase[d] * 2 ** refs + 1 for d in range(3)])) * spec["block"]
    b = spec["block"]
    # generator + oracle copies of all sub-boxes: measured 21.1 GB for 257^3 on 8 ranks = 1.35 x replica_bytes' formula
    job_bytes = int(1.4 * 2 * 1.25 * (8.0 / 7.0) * 27 * (n_global // b) * (8 * b * b + 4))
    jobs = max(1, nproc // world)
    try:
        with open("/proc/meminfo") as f:
            avail = next(int(l.split()[1]) * 1024 for l in f if l.startswith("MemAvailable"))
        jobs = max(1, min(jobs, int(0.4 * avail / max(job_bytes, 1))))
    except Exception:
        pass
    kw = dict(spec["kw"], base=tuple(gbase), order=order)
    res = partitioned.run(part, refs, solver_desc(refs, workload=workload), steps=steps, warmup=warmup, jobs=jobs,
                          problem=spec["problem"], timeout=1500, **kw)
    dt = max(r["dt"] for job in res for r in job)
    r0 = res[0][0]
    return {"value": jobs * steps * n_global / dt / 1e6, "dt_per_step": dt / steps, "its": len(r0["hist"]) - 1, "n": n_global,
            "kind": "reference" if r0["kind"] == "ref" else "port", "cores": jobs * world, "jobs": jobs, "ranks": world,
            "gather": r0["gather"], "hist": r0["hist"]}


def run_reference(args):
    """CPU arm: ugcore's own kernels (oracle/_ref, else the port) on all host cores.  N = 1: concurrent serial solves of
    the one-GPU workload.  N > 1 (CG + GMG-Jacobi workloads): the same global grid the N GPUs solve, partitioned over N
    CPU processes per job (cpu_partitioned); `--cpu-arm replicas` keeps the exchange-free replicas of one GPU's box."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    refs = args.cpu_refs if args.cpu_refs is not None else args.refs
    nproc = args.cpu_procs or host_cores()
    spec = workload_spec(args.workload)
    order = {"lex": 0, "hier": 1}[args.order]
    strong = args.scaling == "strong"
    part = PART[args.gpus]
    arm = args.cpu_arm
    if arm == "auto":
        arm = "partitioned" if args.gpus > 1 and args.workload != "convdiff" else "replicas"
    if arm == "partitioned":
        if args.workload == "convdiff":
            raise SystemExit("--cpu-arm partitioned: CG + GMG-Jacobi workloads only (poisson, elasticity)")
        bm = args.base_mult
        gbase = tuple(2 * bm for _ in range(3)) if strong else tuple(bm * p for p in part)
        try:
            r = cpu_partitioned(refs, part, gbase, nproc, args.steps, args.warmup, args.workload, order)
        except Exception as e:
            if args.cpu_arm != "auto":
                raise
            print(f"bench.py: partitioned CPU arm failed ({e!r}); falling back to --cpu-arm replicas", file=sys.stderr, flush=True)
            args.cpu_arm = "replicas"
            return run_reference(args)
        n, its, val, dt = r["n"], r["its"], r["value"], r["dt_per_step"]
        dims = "x".join(str(gbase[d] * 2 ** refs + 1) for d in range(3))
        sample = (f"{r['jobs']} concurrent partitioned solves x {args.steps} steps of {spec['label']} {dims} nodes ({n} DoF, {its} "
                  f"iterations, {dt:.2f} s per solve): each job = {r['ranks']} processes on the process grid {'x'.join(map(str, part))} "
                  "of the GPU run, additive matrices, consistent/additive/unique vectors, interface exchange and all-reduce "
                  f"through shared memory, levels <= {r['gather']} gathered on one process (ugcore's MPI path emulated, "
                  "oracle/partitioned.py)")
        what = f"{n} DoF per job, {r['jobs']} jobs x {r['ranks']} processes"
    else:
        bmult = 2 * args.base_mult if strong else args.base_mult
        r = cpu_replicas(refs, nproc, args.steps, args.warmup, args.workload, bmult, order)
        n, its, val, dt = r["n"], r["its"], r["value"], r["dt_per_step"]
        nodes = bmult * 2 ** refs + 1
        sample = (f"{r['cores']} concurrent serial solves x {args.steps} steps of {spec['label']} {nodes}^3 nodes ({n} DoF, {its} "
                  f"iterations each, {dt:.2f} s per solve): one process per host core, no threads inside ugcore on this path, "
                  "no MPI on the box -> no interface exchange (upper bound of the MPI weak-scaling throughput)")
        what = f"{n} DoF per process, {r['cores']} processes"
    nproc = r["cores"]           # may have been reduced to fit the host memory
    out = {"impl": "reference", "metric": spec["metric"], "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"{spec['label']} unit cube hexahedra numRefs={refs} ({what}) "
                                  f"{spec['method']}, base LU on level 0, ugcore CPU kernels", "iterations": its, "cpu_arm": arm},
           "cpu_baseline": {"value": val, "unit": UNIT, "cores": nproc, "kind": r["kind"], "sample": sample},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def kernel_roofline(s, prob, top, peak_gbs, peak_src):
    """Dominant kernel = the top-level fused smoothing step `sc += st; sd -= A st; st' = Dinv sd` (4 of the 5 top-level
    matrix sweeps per CG iteration), timed alone with CUDA events on the launching stream.  The vectors rotate over
    three sets, so no launch finds its operands in L2 (matrix stream + 3 x 5 vectors >> 126 MB).

    `roofline` reports the bytes the kernel HAS TO MOVE in its stored format against the measured copy bandwidth
    (frac) and against the 8 TB/s figure north_star quotes (frac_vs_spec).  The format is a lossless 4 B-per-entry
    encoding of the CRS matrix, so the same time against the CRS bytes of SURVEY.md §8d (12 B per entry) is a larger
    number; it is kept as achieved_vs_crs_bytes / frac_vs_crs_bytes and is NOT an HBM fraction.
    `roofline_plain` is the same sweep on the plain 12 B-per-entry stream (UG4B200_MAT_NO_COMPRESS) — the kernel every
    matrix without a small value dictionary gets; there stored bytes = CRS bytes (+ 1.5 % SELL padding)."""
    import ctypes as C
    from ugcore_b200 import capi
    from ugcore_b200.solver import host_ctx, DeviceBuffer
    dev = capi.dev
    ctx = host_ctx()
    A = prob.matrix(top)
    n, nnz = A.nrows, A.nnz
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rng = np.random.default_rng(0)
    NSET = 3
    sets = [[DeviceBuffer.from_numpy(rng.standard_normal(n)) for _ in range(4)] for _ in range(NSET)]   # sd, st, st2, sc
    dinv = DeviceBuffer(n)
    e0, e1 = C.c_void_p(), C.c_void_p()
    dev.ug4b200_event_create(ctx, C.byref(e0)); dev.ug4b200_event_create(ctx, C.byref(e1))

    def timeit(fn, reps=21):
        for k in range(3):
            fn(k % NSET)
        dev.ug4b200_sync(ctx)
        dev.ug4b200_event_record(ctx, e0)
        for k in range(reps):
            fn(k % NSET)
        dev.ug4b200_event_record(ctx, e1)
        dev.ug4b200_event_sync(ctx, e1)
        ms = C.c_float()
        dev.ug4b200_event_elapsed_ms(ctx, e0, e1, C.byref(ms))
        return ms.value / reps

    gbs = lambda b, ms: b / (ms * 1e-3) / 1e9
    # CRS bytes of the reference sweep, SURVEY.md §8d / DESIGN.md §4
    b_apply = 12 * nnz + 4 * (n + 1) + 8 * n + 8 * n
    b_minus = b_apply + 8 * n
    b_fused = 12 * nnz + 4 * (n + 1) + 64 * n   # + st_in, sd r/w, dinv, st_out, sc r/w
    flags = capi.SMOOTH_ADD_IN | capi.SMOOTH_JACOBI
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tr = json.load(f)
    except Exception:
        tr = {}
    out = {}
    for name, mflags in (("roofline", 0), ("roofline_plain", capi.MAT_NO_COMPRESS)):
        m = C.c_void_p()
        capi.check(dev.ug4b200_matrix_upload_crs(ctx, 1, n, n, vp(A.rowptr), vp(A.cols), vp(A.vals), mflags, C.byref(m)), ctx)
        info = capi.MatrixInfo()
        dev.ug4b200_matrix_get_info(m, C.byref(info))
        capi.check(dev.ug4b200_jacobi_prepare(ctx, m, 0.66, 1, dinv.ptr), ctx)
        t_fused = timeit(lambda k: dev.ug4b200_jacobi_smooth_fused(ctx, m, dinv.ptr, sets[k][0].ptr, sets[k][1].ptr, sets[k][2].ptr, sets[k][3].ptr, flags))
        t_spmv = timeit(lambda k: dev.ug4b200_matrix_matmul_minus(ctx, m, sets[k][0].ptr, sets[k][1].ptr, 1))
        t_apply = timeit(lambda k: dev.ug4b200_matrix_apply(ctx, m, sets[k][0].ptr, sets[k][1].ptr, 1))
        dev.ug4b200_matrix_destroy(ctx, m)
        ns, pnnz = int(info.num_slices), int(info.padded_nnz)
        vec_bytes = 8 * n * int(info.fused_vector_streams) if hasattr(info, "fused_vector_streams") and info.fused_vector_streams else 64 * n
        if info.value_indexed:
            # one 32-bit word per (padded) entry + row lengths + slice offsets + column bases + the vector streams
            b_stored = 4 * pnnz + 4 * n + 8 * (ns + 1) + 4 * ns + vec_bytes
            kname, tkey = "tma::spmv1_vi_kernel<-1,INPLACE,FUSE_JACOBI> (ug4b200_jacobi_smooth_fused, value-indexed SELL-32 stream, 4 B/entry)", "spmv1_vi_kernel"
        elif info.x_staged:
            # one 32-bit word per (padded) entry + row lengths + per-slice header (16 B) and run slots (8 B each) + the
            # vector streams (x counted once, 8 B per column: its staged segments overlap and are served from L2)
            b_stored = 4 * pnnz + 4 * n + ns * (16 + 8 * int(info.x_staged_runs)) + vec_bytes
            kname, tkey = "tma::spmv1_xs_kernel<-1,INPLACE,FUSE_JACOBI> (ug4b200_jacobi_smooth_fused, x-staged value-indexed SELL-32 stream, 4 B/entry)", "spmv1_xs_kernel"
        else:
            b_stored = 12 * pnnz + 4 * n + 8 * (ns + 1) + vec_bytes
            kname, tkey = "tma::spmv1_tma_kernel<-1,INPLACE,FUSE_JACOBI> (ug4b200_jacobi_smooth_fused, plain SELL-32 stream, 12 B/entry)", "spmv1_tma_kernel"
        out[name] = {
            "bound": "hbm", "kernel": kname, "achieved": gbs(b_stored, t_fused), "peak": peak_gbs, "unit": "GB/s",
            "frac": gbs(b_stored, t_fused) / peak_gbs, "peak_source": peak_src,
            "traffic": tr.get(tkey), "traffic_source": (tr.get("_sources") or {}).get(tkey, tr.get("_source")),
            "bytes_per_launch": b_stored, "ms_per_launch": t_fused,
            "spec_peak": 8000.0, "frac_vs_spec": gbs(b_stored, t_fused) / 8000.0,
            "crs_bytes_per_launch": b_fused, "achieved_vs_crs_bytes": gbs(b_fused, t_fused), "frac_vs_crs_bytes": gbs(b_fused, t_fused) / peak_gbs,
            "note": "bytes_per_launch = what the kernel has to move in its stored format (entry stream incl. SELL padding, row / slice "
                    "metadata, 8 vector streams); achieved = that / CUDA-event time per launch, operands rotated so that nothing is "
                    "L2-resident; frac is against the measured copy bandwidth, frac_vs_spec against north_star's 8 TB/s. "
                    "*_vs_crs_bytes use the CRS bytes of SURVEY.md 8d (12 B per entry) and exceed 1 for the 4 B encoding: not an HBM fraction.",
            "other_sweeps": {
                "matmul_minus": {"ms_per_launch": t_spmv, "crs_bytes_per_launch": b_minus, "stored_bytes_per_launch": b_stored - vec_bytes + 24 * n,
                                 "achieved": gbs(b_stored - vec_bytes + 24 * n, t_spmv), "frac": gbs(b_stored - vec_bytes + 24 * n, t_spmv) / peak_gbs},
                "apply": {"ms_per_launch": t_apply, "crs_bytes_per_launch": b_apply, "stored_bytes_per_launch": b_stored - vec_bytes + 16 * n,
                          "achieved": gbs(b_stored - vec_bytes + 16 * n, t_apply), "frac": gbs(b_stored - vec_bytes + 16 * n, t_apply) / peak_gbs}}}
    dev.ug4b200_event_destroy(ctx, e0); dev.ug4b200_event_destroy(ctx, e1)
    return out["roofline"], out["roofline_plain"]


def workload_roofline(workload, prob, top, peak_gbs, peak_src):
    """Dominant kernel of the non-default workloads, timed alone like kernel_roofline:
    convdiff   — one forward multicolour Gauss-Seidel sweep over the colour-sorted top-level matrix
                 (ug4b200_gs_step; a sweep reads the lower triangle incl. the diagonal: 12 nnz_lower + 4(n+1) + 24 n bytes);
    elasticity — y -= A x with 3x3 blocks (ug4b200_matrix_matmul_minus; (72+4) nnzb + 4(nb+1) + 72 nb bytes)."""
    import ctypes as C
    from ugcore_b200 import capi
    from ugcore_b200.solver import host_ctx, DeviceBuffer
    dev = capi.dev
    ctx = host_ctx()
    A = prob.matrix(top)
    n, nnz, b = A.nrows, A.nnz, A.block
    rng = np.random.default_rng(0)
    e0, e1 = C.c_void_p(), C.c_void_p()
    dev.ug4b200_event_create(ctx, C.byref(e0)); dev.ug4b200_event_create(ctx, C.byref(e1))

    def timeit(fn, reps=20):
        for _ in range(3):
            fn()
        dev.ug4b200_sync(ctx)
        dev.ug4b200_event_record(ctx, e0)
        for _ in range(reps):
            fn()
        dev.ug4b200_event_record(ctx, e1)
        dev.ug4b200_event_sync(ctx, e1)
        ms = C.c_float()
        dev.ug4b200_event_elapsed_ms(ctx, e0, e1, C.byref(ms))
        return ms.value / reps

    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    m = C.c_void_p()
    try:
        if workload == "convdiff":
            color = np.zeros(n, np.int32)
            nc = C.c_int()
            dev.ug4b200_color_greedy(n, vp(A.rowptr), vp(A.cols), vp(color), C.byref(nc))
            order = np.argsort(color, kind="stable")
            perm = np.empty(n, np.int64); perm[order] = np.arange(n)
            cptr = np.concatenate([[0], np.cumsum(np.bincount(color, minlength=nc.value))]).astype(np.int64)
            rows = np.repeat(np.arange(n), np.diff(A.rowptr))
            pr_, pc_ = perm[rows], perm[A.cols]
            key = np.lexsort((pc_, pr_))
            rp = np.concatenate([[0], np.cumsum(np.bincount(pr_, minlength=n))]).astype(np.int64)
            ci = pc_[key].astype(np.int32); va = np.ascontiguousarray(A.vals[key])
            capi.check(dev.ug4b200_matrix_upload_crs(ctx, 1, n, n, vp(rp), vp(ci), vp(va), 0, C.byref(m)), ctx)
            d, c = DeviceBuffer.from_numpy(rng.standard_normal(n)), DeviceBuffer.from_numpy(np.zeros(n))
            t = timeit(lambda: dev.ug4b200_gs_step(ctx, m, cptr.size - 1, vp(cptr), 0, C.c_double(1.0), c.ptr, d.ptr))
            # a forward sweep reads, per row, the connections up to and including the diagonal in the colour-sorted
            # numbering (c_i = (d_i - sum_{j<i} a_ij c_j) / a_ii): about half of the stored entries — NOT the whole matrix
            # (the lines of profiles/r02a, r02f, r02h counted 12 * nnz and overstated this fraction by 1.8x)
            n_lower = int(np.count_nonzero(pc_ <= pr_))
            nbytes = 12 * n_lower + 4 * (n + 1) + 24 * n
            kname = (f"gs_color_kernel x {cptr.size - 1} colours (ug4b200_gs_step, forward sweep, plain SELL-32 stream; "
                     f"{n_lower} of {nnz} entries read)")
        else:
            capi.check(dev.ug4b200_matrix_upload_crs(ctx, b, n, n, vp(A.rowptr), vp(A.cols), vp(A.vals), 0, C.byref(m)), ctx)
            y, x = DeviceBuffer.from_numpy(rng.standard_normal(n * b)), DeviceBuffer.from_numpy(rng.standard_normal(n * b))
            t = timeit(lambda: dev.ug4b200_matrix_matmul_minus(ctx, m, y.ptr, x.ptr, b))
            nbytes = (8 * b * b + 4) * nnz + 4 * (n + 1) + 3 * 8 * b * n
            kname = f"spmvB_kernel<{b}> (ug4b200_matrix_matmul_minus, {b}x{b} blocks)"
        gbs = nbytes / (t * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": kname, "achieved": gbs, "peak": peak_gbs, "unit": "GB/s", "frac": gbs / peak_gbs,
                "peak_source": peak_src, "traffic": None, "bytes_per_launch": nbytes, "ms_per_launch": t}
    finally:
        dev.ug4b200_event_destroy(ctx, e0); dev.ug4b200_event_destroy(ctx, e1)
        if m:
            dev.ug4b200_matrix_destroy(ctx, m)


def cpu_baseline_sample(refs, workload="poisson", base_mult=1, order=0):
    """Bounded CPU sample for the default run: the SAME workload solved once by every host core
    concurrently with the reference's kernels (~10-30 s including set-up)."""
    r = cpu_replicas(refs, host_cores(), 1, 0, workload, base_mult, order)
    nproc = r["cores"]
    serial_equiv = r["n"] / r["dt_per_step"] / 1e6
    return {"value": r["value"], "unit": UNIT, "cores": nproc, "kind": r["kind"],
            "sample": f"{nproc} concurrent serial solves of the same workload ({r['n']} DoF, {r['its']} CG iterations, "
                      f"{r['dt_per_step']:.1f} s each), ugcore SparseMatrix/Vector kernels compiled from the reference; one process "
                      "per core (no threading on this path, no MPI on the box: no interface exchange)",
            "solve_s": r["dt_per_step"], "per_core_value": serial_equiv, "history_last": r["hist"][-1]}, np.array(r["hist"])


def run_ours(args):
    import ctypes as C
    import torch
    from ugcore_b200 import capi, problems as pr
    from ugcore_b200 import solver as S
    from ugcore_b200.capi import check_host, host

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one line (the JSON record): NCCL's version banner would precede it
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ.pop("NCCL_DEBUG")
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local_rank)
    # ONE explicit stream for torch and for the library: torch's default stream has handle 0, which the library would
    # take as "create your own (non-blocking) stream" — the CUDA events below would then sit on an idle stream and
    # x_dev.zero_() would race with the solver's first copy (seen at 257^3: a start defect of 5.3 instead of 0.0105)
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    S.host_init(local_rank, C.c_void_p(stream.cuda_stream))
    ctx = S.host_ctx()
    dev = capi.dev
    refs = args.refs
    spec = workload_spec(args.workload)
    desc = solver_desc(refs, workload=args.workload, base_solver=args.base_solver)
    bm = args.base_mult

    strong = args.scaling == "strong"
    order = {"lex": pr.ORDER_LEX, "hier": pr.ORDER_HIER}[args.order]
    reorder = None if args.reorder == "none" else args.reorder
    part = PART[world]
    # global base grid: weak scaling = bm cells per GPU and direction, strong scaling = 2 bm cells per direction whatever N
    gbase = tuple(2 * bm for _ in range(3)) if strong else tuple(bm * p for p in part)
    if world > 1:
        import torch.distributed as dist
        from ugcore_b200 import dist as ugdist
        if reorder:
            raise SystemExit("--reorder is a single-GPU option")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        gen_kw = dict(spec["kw"])
        gen_kw["base"] = gbase
        gen_kw["order"] = order
        prob, s = ugdist.build_partitioned_solver(desc, refs, part, rank, dist, problem=spec["problem"], flags=args.flags, **gen_kw)
        barrier = lambda: (dist.barrier(), torch.cuda.synchronize())
    else:
        prob = pr.Problem(dim=3, num_refs=refs, problem=spec["problem"], base=gbase, order=order, **spec["kw"])
        s = S.Solver.from_problem(desc, prob, flags=args.flags, order=reorder)
        barrier = lambda: torch.cuda.synchronize()
    t_setup = time.perf_counter()
    s.init()
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup      # solver:init — uploads, smoother preprocess, base factorisation (not part of a step)
    n_local = prob.num_dofs
    dims = [gbase[d] * 2 ** refs + 1 for d in range(3)]
    n_global = dims[0] * dims[1] * dims[2] * spec["block"]

    rhs = np.array(prob.rhs())
    if s.perm is not None:      # --reorder: the solver works in its own (Cuthill-McKee) numbering; hand it b in that numbering
        pb = np.empty_like(rhs); pb[s.perm] = rhs; rhs = pb
    b_host = torch.from_numpy(rhs).pin_memory()
    x_host = torch.zeros(n_local, dtype=torch.float64).pin_memory()
    b_dev = b_host.cuda()
    x_dev = torch.zeros(n_local, dtype=torch.float64, device="cuda")

    def solve_device():
        x_dev.zero_()
        ok = s.apply_device(x_dev.data_ptr(), b_dev.data_ptr())
        assert ok, "solver did not converge"

    def solve_e2e():
        # solver:apply(u, b) with u = 0 through the reference-facing call with HOST vectors: H2D of b, the start vector
        # is set on the device (ug4b200_solver_apply_zero_guess), D2H of the solution
        ok = check_host(host.ug4b200_solver_apply_zero_guess(s.h, C.c_void_p(x_host.data_ptr()), C.c_void_p(b_host.data_ptr()))) == 0
        assert ok, "solver did not converge"

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        l0 = s.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        wall = (time.perf_counter() - w0) * 1e3 / steps
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, wall, (s.launch_count() - l0) // steps

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, wall_ms, launches = timed(solve_device, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    its = s.steps
    hist = s.history()
    ms_e2e, wall_e2e, _ = timed(solve_e2e, args.steps, args.warmup)

    if rank != 0:
        return
    peak, peak_src = measured_peak_gbs()
    top = prob.matrix(refs)
    top_gb = (8 * spec["block"] ** 2 + 4) * top.nnz / 1e9
    nodes = f"{dims[0]}x{dims[1]}x{dims[2]}" + (f" nodes x {spec['block']}" if spec["block"] > 1 else "")
    out = {"metric": spec["metric"], "value": n_global / (ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"{spec['label']} unit-cell hexahedra, base grid {gbase[0]}x{gbase[1]}x{gbase[2]} cells, numRefs={refs}, "
                                  f"{nodes} = {n_global} DoF ({n_local} per GPU, boxes {part[0]}x{part[1]}x{part[2]}), {spec['method']}, "
                                  f"base {'LU' if args.base_solver == 'lu' else 'on-device CG (1e-14)'} on level 0, DoF order {args.order}" + (f" + {reorder} at upload" if reorder else ""),
                      "iterations": its, "solve_s": ms * 1e-3, "init_s": setup_s,
                      "l2_policy": f"inputs larger than L2 (top-level matrix {top_gb:.2f} GB per GPU)",
                      "wall_ms_per_step": wall_ms, "final_reduction": float(hist[-1] / hist[0]) if len(hist) else None,
                      "history": [float(v) for v in hist]},
           "e2e": {"value": n_global / (ms_e2e * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": 8 * n_local,
                   "d2h_bytes_per_step": 8 * n_local, "ms_per_step": ms_e2e, "wall_ms_per_step": wall_e2e},
           "gpu_launches": int(launches * args.steps), "gpu_launches_per_step": int(launches), "clocks": clocks}
    if world == 1:
        if args.workload == "poisson":
            out["roofline"], out["roofline_plain"] = kernel_roofline(s, prob, refs, peak, peak_src)
            # whole solve against the bytes of the UNFUSED reference sequence (~2.9 kB per fine DoF and
            # iteration) — a yardstick for fusion + encoding, explicitly not a roofline fraction
            out["solve_gbs_vs_unfused_reference_bytes"] = 2.9e3 * n_global * max(its, 1) / (ms * 1e-3) / 1e9
        else:
            out["roofline"] = workload_roofline(args.workload, prob, refs, peak, peak_src)
        if not args.no_cpu_baseline:
            cb, h_cpu = cpu_baseline_sample(refs, args.workload, gbase[0], order)
            out["cpu_baseline"] = cb
            if args.workload != "convdiff":   # the CPU arm sweeps in ugcore's lexicographic order: another smoother
                k = min(len(h_cpu), len(hist))
                out["config"]["history_rel_err_vs_cpu"] = float(np.max(np.abs(hist[:k] - h_cpu[:k]) / np.abs(h_cpu[:k])))
            out["config"]["iterations_cpu"] = len(h_cpu) - 1
            out["config"]["history_cpu"] = [float(v) for v in h_cpu]
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--refs", type=int, default=7, help="refinements per GPU sub-box (7 -> 129^3)")
    ap.add_argument("--cpu-refs", type=int, default=None, help="refinements of the CPU reference arm (default: --refs)")
    ap.add_argument("--workload", default="poisson", choices=["poisson", "convdiff", "elasticity"],
                    help="poisson: BASELINE configs[1]/[2] (default); convdiff: configs[3]; elasticity: configs[4]")
    ap.add_argument("--flags", type=int, default=0, help="UG4B200_FLAG_* bits for the solver (32: device-resident BiCGStab)")
    ap.add_argument("--base-mult", type=int, default=1, help="base-grid elements per GPU and direction "
                    "(--workload elasticity --base-mult 3 --refs 5: 97^3 nodes per GPU, 193^3 x 3 = 21.6 M DoF on 8 GPUs)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): one box of base-mult cells per GPU; strong: the same 2x2x2-cell base grid (257^3 at "
                         "--refs 7, BASELINE configs[2]) on every GPU count")
    ap.add_argument("--order", default="lex", choices=["lex", "hier"], help="DoF numbering of the generator: lexicographic, "
                    "or hierarchical like ugcore's global refinement (coarse vertices first)")
    ap.add_argument("--reorder", default="none", choices=["none", "cmk", "rcmk"], help="(reverse) Cuthill-McKee of every level before upload")
    ap.add_argument("--base-solver", default="lu", choices=["lu", "cg"], help="GMG base solver: dense LU (ugcore's default) or the small "
                    "on-device CG of north_star (single CTA, reduction 1e-14) — for base grids of many cells, where the dense triangular "
                    "solves of an LU are a serial latency chain on the device")
    ap.add_argument("--cpu-procs", type=int, default=0, help="processes of the CPU arm (0 = all host cores)")
    ap.add_argument("--cpu-arm", default="auto", choices=["auto", "replicas", "partitioned"],
                    help="CPU arm at --gpus N > 1: partitioned (auto for poisson / elasticity) = the N-GPU grid over N CPU processes "
                         "per job with ugcore's parallel protocol; replicas = exchange-free serial solves of one GPU's box")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

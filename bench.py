#!/usr/bin/env python
"""bench.py - BASELINE.json's headline metric on BASELINE.json's config, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference] [--config 3]

Workload (N = 1): config 3, the synthetic sparse LP 1M x 2M with 16 nnz/col (32 M nonzeros),
the configuration the metric's target is quoted on; it fits one B200 and its per-iteration
working set (~1 GB) is far larger than L2, so no explicit L2 flush is needed.
A "step" is ONE COMPLETE SOLVE of that LP from a zero start to 1e-6 relative KKT.

  metric / value : PDHG iterations per second with the model resident in HBM
                   (total iterations of the K timed solves / their wall time)
  time_to_tol_s  : seconds per solve to 1e-6 relative KKT (resident)
  e2e            : the same metric from HOST buffers, everything inside the timed region.  N = 1: through the
                   reference-facing plugin call - miplib_create_solver(Cuda) + miplib_cuda_load_csr +
                   Solver::solve() + values, a fresh solver every step - with the C-ABI figure beside it
                   (e2e.c_abi: mipcu_load_lp, H2D from pinned memory + transposes, + mipcu_solve_lp +
                   mipcu_get_x / get_y); N > 1: the C ABI on every rank
  roofline       : the slower of the two fused kernels (k_row_step / k_col_step), timed in situ by
                   CUDA events on the solver's stream during the timed solves; both are listed
  cpu_baseline   : the multi-core C/OpenMP port of the same iteration (oracle/pdhg_port.c) on the
                   same LP, a bounded number of iterations, on this box's host cores

  parity         : the (x, y) of the last timed solve, y gathered across the ranks, checked on rank 0
                   by the independent FP64 certificate verifier (oracle/kkt_verify.py) and against
                   the golden objective; a failed check makes the run exit non-zero
  nodes          : bounded secondary block (config 4): rounds of 128 branch-and-bound node LPs per
                   GPU, dealt to the ranks with no collective; root LP certificate-checked
  mip            : N = 1 only, informative (never changes the exit code): config 4 as a MIP at FULL size
                   for 12 s through the plugin C API; incumbent recomputed integral and row-feasible,
                   dual bound of the interrupted tree, gap
  cpu_baseline_highs : N = 1 only - SciPy-HiGHS (the stand-in for the reference's SCIP / lp_solve
                   backends) on config 1 to optimality and on the bench config under a time cap

N > 1 (torchrun, one process per GPU): the same LP row-sharded across the ranks (strong scaling):
every rank owns m/N rows of A and n/N columns of A' (block-owner exchange: the slices of the iterates are
pushed into peer memory over NVLink from the kernels' epilogues; --shard-exchange 0 = one NCCL all-reduce of
the K'y partial sums per iteration).  --impl reference: rank 0 times the CPU port; other ranks exit 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "pdhg_iters_per_s"
UNIT = "iter/s"


def workload_name(cfg: int) -> str:
    return {1: "C1: synthetic sparse LP 1k x 2k, 8 nnz/col",
            2: "C2: synthetic sparse LP 100k x 200k, 8 nnz/col",
            3: "C3: synthetic sparse LP 1M x 2M, 16 nnz/col (32M nnz), solved to 1e-6 rel. KKT"}[cfg]


def config_dict(cfg: int, m: int, n: int, nnz: int) -> dict:
    """Identical in the CUDA arm and the reference arm (the driver compares them)."""
    return {"workload": workload_name(cfg), "m": int(m), "n": int(n), "nnz": int(nnz),
            "l2": "inputs larger than L2 (no flush)" if nnz > 8e6 else "L2-resident working set"}


GOLDEN_KEY = {1: "C1_lp", 2: "C2_lp", 3: "C3_lp"}


def parity_block(L, cfg: int, x, y, reported_obj: float) -> dict:
    """Independent certificate of the returned point (oracle/kkt_verify.py shares no code with the
    CUDA path) + relative objective distance to the golden optimum of tests/golden/objectives.json
    (C1: HiGHS simplex; C2 / C3: HiGHS' own PDLP at 1e-6, so 2e-6 is the agreement one can ask)."""
    from oracle import kkt_verify
    cert = kkt_verify.certificate(L, x, y)
    golden = json.loads((ROOT / "tests" / "golden" / "objectives.json").read_text())[GOLDEN_KEY[cfg]]
    rel_obj = abs(cert["primal_obj"] - golden["objective"]) / (1.0 + abs(golden["objective"]))
    worst = max(cert["rel_primal_res"], cert["rel_dual_res"], cert["rel_gap"])
    ok = bool(worst <= 1.05e-6 and rel_obj <= 2e-6 and
              abs(cert["primal_obj"] - reported_obj) <= 1e-9 * (1.0 + abs(reported_obj)))
    return {"checker": "oracle.kkt_verify.certificate on the gathered (x, y) of the last timed (end-to-end) solve",
            "rel_primal_res": cert["rel_primal_res"], "rel_dual_res": cert["rel_dual_res"],
            "rel_gap": cert["rel_gap"], "primal_obj": cert["primal_obj"], "dual_obj": cert["dual_obj"],
            "max_bound_violation": cert["max_bound_violation"],
            "golden_objective": golden["objective"], "golden_kind": golden["kind"],
            "rel_objective_vs_golden": rel_obj, "tolerance": {"kkt": 1.05e-6, "objective": 2e-6}, "ok": ok}


def cpu_baseline_highs(cfg: int, cap_s: float = 60.0) -> dict:
    """SciPy-HiGHS on the host cores as the stand-in for the reference's SCIP / lp_solve
    backends (SCIPsolve reference miplib/scip/solver.cpp:370, ::solve miplib/lpsolve/solver.cpp:156 -
    not installable offline).  Config 1 to optimality; the bench config in a child process under a
    hard wall-clock cap (HiGHS' own time limit is not checked inside presolve / factorisation)."""
    import scipy
    from oracle import gen, highs_ref
    out = {"label": "HiGHS stand-in for the reference's SCIP/lp_solve backend (not installable offline)",
           "scipy": scipy.__version__, "host_cores": os.cpu_count()}
    L1 = gen.config(1)
    for method in ("highs-ds", "highs-ipm"):
        t0 = time.perf_counter()
        r = highs_ref.solve_lp(L1, method)
        out[f"C1_{method}"] = {"status": r["status"], "objective": r["obj"], "iterations": r["nit"],
                               "seconds": time.perf_counter() - t0}
    if cfg != 1:
        code = ("import sys, time; sys.path.insert(0, %r)\n"
                "from oracle import gen, highs_ref\n"
                "L = gen.config(%d); t0 = time.perf_counter()\n"
                "r = highs_ref.solve_lp(L, 'highs-ipm', time_limit=%f)\n"
                "print('HIGHS', r['status'], r['obj'], time.perf_counter() - t0)\n") % (str(ROOT), cfg, cap_s)
        t0 = time.perf_counter()
        try:
            pr = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                                timeout=cap_s + 45.0)   # generation + model copy ride on top of the cap
            tail = [ln for ln in pr.stdout.splitlines() if ln.startswith("HIGHS")]
            if tail and tail[-1].split()[1] == "0":
                f = tail[-1].split()
                out[f"C{cfg}_highs-ipm"] = {"status": 0, "objective": float(f[2]), "seconds": float(f[3])}
            else:
                out[f"C{cfg}_highs-ipm"] = {"status": "did not finish", "cap_seconds": cap_s,
                                           "detail": (tail[-1] if tail else pr.stderr[-200:])}
        except subprocess.TimeoutExpired:
            out[f"C{cfg}_highs-ipm"] = {"status": "did not finish", "cap_seconds": cap_s,
                                       "detail": f"killed after {time.perf_counter() - t0:.0f} s wall"}
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def pinned_like(a: np.ndarray) -> np.ndarray:
    """Copy of `a` in page-locked host memory (so the timed H2D runs at PCIe speed)."""
    import torch
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0].copy()).dtype, pin_memory=True)
    out = t.numpy()
    out[...] = a
    out_holder.append(t)
    return out


out_holder = []


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed ncu summary, or None."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(kernel)
        except Exception:
            return None
    return None


def cpu_baseline(L, seconds_budget: float = 15.0):
    """The multi-core C port on the same LP: bounded number of iterations, all host cores."""
    from oracle import pdhg_ref, port
    t0 = time.perf_counter()
    port.use_all_threads()
    S = pdhg_ref.prepare(L, step_size=0.998)     # scaling only (eta pinned: no power iteration)
    st = port.PortState(S)
    setup = time.perf_counter() - t0
    probe = st.iterate(2)
    iters = int(max(4, min(256, seconds_budget / max(probe / 2, 1e-6))))
    secs = st.iterate(iters)
    return {"value": iters / secs, "unit": UNIT, "cores": port.num_threads(), "kind": "port",
            "sample": f"{iters} PDHG iterations (both fused passes) of the same LP in {secs:.1f} s; "
                      f"scaling setup {setup:.1f} s not included",
            "iterations": iters, "seconds": secs}


def run_reference(args, rank):
    if rank != 0:
        return
    from oracle import gen
    L = gen.config(args.config)
    per_step = []
    from oracle import pdhg_ref, port
    port.use_all_threads()
    S = pdhg_ref.prepare(L, step_size=0.998)
    st = port.PortState(S)
    probe = st.iterate(2)
    budget = 120.0 / max(1, args.steps + args.warmup)          # whole run within a few minutes
    iters = int(max(2, min(128, budget / max(probe / 2, 1e-6))))
    for _ in range(args.warmup):
        st.iterate(iters)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        per_step.append(st.iterate(iters))
    total = time.perf_counter() - t0
    value = iters * args.steps / total
    cores = port.num_threads()
    sample = (f"each step = {iters} PDHG iterations (both fused passes) of the same LP by the C/OpenMP "
              f"port oracle/pdhg_port.c on {cores} host threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_dict(args.config, L.m, L.n, L.nnz),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference's own SCIP / lp_solve backends are not installable offline; this arm "
                "times the CPU port of the same PDHG iteration",
    }), flush=True)


def run_cuda(args, rank, world, local_rank):
    import torch
    from miplib_b200 import Engine
    from oracle import gen

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    from miplib_b200 import sharding
    L = gen.config(args.config)
    A = L.A.tocsr()
    A.sort_indices()
    m, n = A.shape
    # row block of this rank (the whole matrix when world == 1)
    r0, r1 = sharding.row_block(m, rank, world)
    ptr, idx, val, lo, hi = sharding.shard_csr(A, L.row_lo, L.row_hi, rank, world)
    host = dict(indptr=pinned_like(ptr), indices=pinned_like(idx), data=pinned_like(val),
                c=pinned_like(L.c), lb=pinned_like(L.lb), ub=pinned_like(L.ub),
                row_lo=pinned_like(lo), row_hi=pinned_like(hi))
    h2d = sum(v.nbytes for v in host.values())
    d2h = 8 * (n + (r1 - r0))

    if world > 1:
        uid = [Engine.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng = Engine(local_rank, rank, world, uid[0])
        eng.set_param("shard_exchange", args.shard_exchange)
    else:
        eng = Engine(local_rank)

    def load():
        eng.load(host["indptr"], host["indices"], host["data"], host["c"], host["lb"], host["ub"],
                 host["row_lo"], host["row_hi"], L.maximize, n=n)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident: model loaded once, K timed solves ------------------------------------------
    load()
    for _ in range(args.warmup):
        st = eng.solve()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    stats = [eng.solve() for _ in range(args.steps)]
    barrier()
    resident_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None
    iters = sum(s.iterations for s in stats)
    launches = sum(s.kernel_launches for s in stats)
    dev_iter_s = max_over_ranks(sum(s.iter_seconds for s in stats))

    # ---- end to end through the C ABI with host buffers ---------------------------------------
    def e2e_step():
        load()
        s = eng.solve()
        x = eng.x(); y = eng.y()
        return s, float(x[0] + y[0])
    e2e_step()                                   # one warm-up of the full path
    barrier()
    t0 = time.perf_counter()
    e2e_stats = [e2e_step()[0] for _ in range(args.steps)]
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_iters = sum(s.iterations for s in e2e_stats)

    # ---- parity of the last timed solve: gather y (row blocks) on rank 0, certificate there -------
    eng_x = eng.x()
    y_loc = eng.y()
    if dist is not None:
        sizes = [sharding.row_block(m, r, world) for r in range(world)]
        pad = max(b - a for a, b in sizes)
        mine = torch.zeros(pad, dtype=torch.float64, device="cuda")
        mine[:y_loc.size] = torch.from_numpy(y_loc).cuda()
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        y_all = np.concatenate([p[:b - a].cpu().numpy() for p, (a, b) in zip(parts, sizes)])
    else:
        y_all = y_loc
    exchange = int(eng.get_param("shard_exchange")) if world > 1 else None
    multicast = int(eng.get_param("multicast")) if world > 1 else None
    last = stats[-1]
    peak, peak_src = measured_peak_gbs()
    row_us = float(np.mean([s.row_kernel_us for s in stats]))
    col_us = float(np.mean([s.col_kernel_us for s in stats]))
    spmv = {}
    if rank == 0:
        if row_us <= 0:   # graph mode (small configs): time the kernels with the hook instead
            row_us, _ = eng.time_kernel(2, 50, True)
            col_us, _ = eng.time_kernel(3, 50, True)
        if world == 1:
            # the two plain products (B_Ax / B_ATy over >= 100 repetitions, CUDA events on
            # the library's stream through the mipcu_time_kernel hook); after the timed solves
            for which, name in ((0, "Ax"), (1, "ATy")):
                us, nbytes = eng.time_kernel(which, 100, False)
                spmv[name] = {"launch_us": us, "bytes": int(nbytes), "GBps": nbytes / us / 1e3,
                              "frac_of_peak": nbytes / us / 1e3 / peak}
    barrier()
    eng.close()

    # ---- N = 1: the same end-to-end step through the reference-facing plugin call -----------------
    plugin = None
    if world == 1:
        plugin = plugin_e2e(args, L, A, host)

    # ---- bounded secondary block: node LPs of config 4 dealt across the ranks ---------------------
    nodes = None
    if not args.no_nodes:
        nodes = nodes_block(4, args.nodes_per_gpu, steps=2, warmup=1, rank=rank, world=world,
                            local_rank=local_rank, dist=dist)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    parity = parity_block(L, args.config, eng_x, y_all, e2e_stats[-1].primal_obj)
    kernels = {"k_row_step": (row_us, last.row_kernel_bytes), "k_col_step": (col_us, last.col_kernel_bytes)}
    dom = max(kernels, key=lambda k: kernels[k][0])
    dom_us, dom_bytes = kernels[dom]
    achieved = dom_bytes / dom_us / 1e3
    e2e = {"value": e2e_iters / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "seconds_per_solve": e2e_s / args.steps,
           "through": "C ABI (include/mipcu.h): mipcu_load_lp + mipcu_solve_lp + mipcu_get_x / get_y, host buffers"}
    if plugin is not None:
        # headline e2e at N = 1 is the plugin call; the C-ABI figure stays beside it
        e2e = dict(plugin, c_abi=e2e)
    line = {
        "metric": METRIC, "value": iters / resident_s, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * resident_s / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_dict(args.config, m, n, int(A.nnz)),
        "layout": {"matrix_format": "SELL-32-256, one lane per row (CSR lanes-per-row kernels for long-row models); rows sorted by "
                                    "length, columns and rows binned by their first index at load (single-GPU contexts)",
                   "parallelism": "single GPU" if world == 1 else
                   f"row-sharded x{world}, " + {
                       2: "block-owner exchange: each rank streams its row block of K and its column block of K', "
                          "yhat / xbar slices pushed into peer memory over NVLink from the kernels' epilogues",
                       1: "fused P2P reduce-scatter + primal step + all-gather kernel over NVLink",
                       0: "NCCL all-reduce of K'y"}[exchange],
                   "multicast": multicast},
        "time_to_tol_s": resident_s / args.steps,
        "iterations_per_solve": iters / args.steps,
        "device_iter_seconds_per_solve": dev_iter_s / args.steps,
        "kkt": {"rel_primal_res": last.rel_primal_res, "rel_dual_res": last.rel_dual_res,
                "rel_gap": last.rel_gap, "status": last.status, "primal_obj": last.primal_obj},
        "parity": parity,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "frac_of_nominal_8TBs": achieved / 8000.0,
                     "peak_source": peak_src, "traffic": ncu_traffic(dom) if world == 1 else None,
                     "algorithmic_bytes_per_launch": int(dom_bytes), "launch_us": dom_us,
                     "all_kernels": {k: {"launch_us": v[0], "bytes": int(v[1]), "GBps": v[1] / v[0] / 1e3,
                                         "frac": v[1] / v[0] / 1e3 / peak}
                                     for k, v in kernels.items()}},
    }
    if world > 1:
        it_us = 1e6 * dev_iter_s / max(1, iters)
        line["roofline"]["iteration_split_us"] = {
            "iteration_mean": it_us, "row_step_kernel": row_us, "col_step_kernel": col_us,
            "note": "iteration_mean = device time of the iteration loop / iterations (controller, KKT pass and "
                    "restarts included); the two kernel times are CUDA events around one launch per 64-iteration "
                    "block and include the cross-GPU entry barrier and the pushes of the rank's slice"}
    if spmv:
        line["spmv"] = spmv
    if nodes is not None:
        line["nodes"] = nodes
    if world == 1 and not args.no_nodes:
        line["mip"] = mip_block(4, 12.0)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(L)
        line["cpu_baseline_highs"] = cpu_baseline_highs(args.config, args.highs_cap)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if not parity["ok"] or (nodes is not None and not nodes.get("ok", True)):
        print("bench.py: PARITY CHECK FAILED", json.dumps(parity), file=sys.stderr, flush=True)
        sys.exit(3)


def plugin_e2e(args, L, A, host) -> dict:
    """One end-to-end step = what a user of the reference's API runs: miplib_create_solver(Cuda) ->
    miplib_cuda_load_csr (bulk ingress, host buffers) -> Solver::solve() -> values of all columns,
    all inside the timed region (libmiplib.so, the C++ front-end + Backend::Cuda adapter)."""
    import ctypes as C
    from miplib_b200 import build
    lib = C.CDLL(str(build.FRONT_LIB))
    lib.miplib_get_last_error.restype = C.c_char_p
    lib.miplib_cuda_load_csr.argtypes = [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 9 + [C.c_int]
    n = A.shape[1]
    lo = pinned_like(np.maximum(host["row_lo"], -1e30))
    hi = pinned_like(np.minimum(host["row_hi"], 1e30))
    x = pinned_like(np.zeros(n))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)

    def step():
        solver = C.c_void_p()
        if lib.miplib_create_solver(C.byref(solver), 5):
            raise RuntimeError(lib.miplib_get_last_error().decode())
        try:
            lib.miplib_cuda_set_verbose(solver, 0)
            rc = lib.miplib_cuda_load_csr(solver, A.shape[0], n, vp(host["indptr"]), vp(host["indices"]),
                                          vp(host["data"]), vp(host["c"]), vp(host["lb"]), vp(host["ub"]),
                                          None, vp(lo), vp(hi), int(L.maximize))
            res, has = C.c_int(), C.c_int()
            rc = rc or lib.miplib_cuda_solve(solver, C.byref(res), C.byref(has))
            obj = C.c_double()
            rc = rc or lib.miplib_cuda_get_values(solver, vp(x), C.byref(obj))
            if rc:
                raise RuntimeError(lib.miplib_get_last_error().decode())
            it = C.c_int64()
            lib.miplib_cuda_get_info(solver, C.byref(it), None, None, None)
            return res.value, obj.value, it.value
        finally:
            lib.miplib_destroy_solver(solver)

    import torch
    step()                                       # warm-up of the whole path
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    outs = [step() for _ in range(args.steps)]
    secs = time.perf_counter() - t0
    iters = sum(o[2] for o in outs)
    h2d = sum(host[k].nbytes for k in ("indptr", "indices", "data", "c", "lb", "ub")) + lo.nbytes + hi.nbytes
    return {"value": iters / secs, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(8 * n),
            "seconds_per_solve": secs / args.steps, "result": outs[-1][0], "objective": outs[-1][1],
            "through": "plugin: miplib_create_solver(Cuda) + miplib_cuda_load_csr + Solver::solve() + values "
                       "(libmiplib.so; host buffers, a fresh solver every step)"}


def mip_block(cfg: int, seconds: float) -> dict:
    """Bounded secondary block (N = 1): BASELINE config 4 as a MIP at FULL size through the plugin C API
    (miplib_create_solver(Cuda) + miplib_cuda_load_csr with Binary columns + Solver::solve() under a time
    limit).  Nobody here can prove its optimum, so the outcome is checked size-independently: the incumbent
    is integral and feasible for every row (recomputed in FP64), its objective is c'x, the dual bound is on
    the right side of it.  Informative only: never changes the exit code."""
    import ctypes as C
    from miplib_b200 import build
    from oracle import gen
    try:
        L = gen.config(cfg)
        A = L.A.tocsr(); A.sort_indices()
        lib = C.CDLL(str(build.FRONT_LIB))
        lib.miplib_get_last_error.restype = C.c_char_p
        lib.miplib_cuda_load_csr.argtypes = [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 9 + [C.c_int]
        lib.miplib_cuda_set_time_limit.argtypes = [C.c_void_p, C.c_double]
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        arrs = [A.indptr.astype(np.int64), A.indices.astype(np.int32), A.data.astype(np.float64),
                np.ascontiguousarray(L.c, dtype=np.float64), np.ascontiguousarray(L.lb, dtype=np.float64),
                np.ascontiguousarray(L.ub, dtype=np.float64), np.ones(L.n, np.int32),
                np.maximum(L.row_lo, -1e30), np.minimum(L.row_hi, 1e30)]
        s = C.c_void_p()
        if lib.miplib_create_solver(C.byref(s), 5):
            raise RuntimeError(lib.miplib_get_last_error().decode())
        try:
            lib.miplib_cuda_set_verbose(s, 0)
            rc = lib.miplib_cuda_load_csr(s, L.m, L.n, *[vp(a) for a in arrs], int(L.maximize))
            rc = rc or lib.miplib_cuda_set_time_limit(s, float(seconds))
            res, has = C.c_int(), C.c_int()
            t0 = time.perf_counter()
            rc = rc or lib.miplib_cuda_solve(s, C.byref(res), C.byref(has))
            secs = time.perf_counter() - t0
            if rc:
                raise RuntimeError(lib.miplib_get_last_error().decode())
            it, nodes, bound, proven = C.c_int64(), C.c_int64(), C.c_double(), C.c_int()
            lib.miplib_cuda_get_info(s, C.byref(it), None, C.byref(nodes), None)
            lib.miplib_cuda_get_bound(s, C.byref(bound), C.byref(proven))
            out = {"workload": f"C{cfg} as a MIP at full size ({L.m} x {L.n} binaries), {seconds:g} s time limit, "
                               "branch and bound with batched node LPs on one GPU (plugin C API)",
                   "seconds": secs, "result": res.value, "has_incumbent": bool(has.value), "nodes": nodes.value,
                   "lp_iterations": it.value, "dual_bound": bound.value, "proven_optimal": bool(proven.value)}
            if has.value:
                x = np.empty(L.n); obj = C.c_double()
                lib.miplib_cuda_get_values(s, vp(x), C.byref(obj))
                ax = L.A @ x
                sgn = -1.0 if L.maximize else 1.0
                out.update(incumbent=obj.value, relative_gap=abs(obj.value - bound.value) / max(1e-9, abs(obj.value)),
                           max_integrality_violation=float(np.abs(x - np.round(x)).max()),
                           max_row_violation=float(max((L.row_lo - ax)[np.isfinite(L.row_lo)].max(initial=0.0),
                                                       (ax - L.row_hi)[np.isfinite(L.row_hi)].max(initial=0.0), 0.0)),
                           objective_recomputed=float(L.c @ x))
                out["ok"] = bool(out["max_integrality_violation"] == 0.0 and out["max_row_violation"] <= 1e-9 and
                                 abs(out["objective_recomputed"] - obj.value) <= 1e-9 * (1 + abs(obj.value)) and
                                 sgn * bound.value <= sgn * obj.value)
            else:
                out["ok"] = False
            return out
        finally:
            lib.miplib_destroy_solver(s)
    except Exception as e:       # informative block: report, never fail the bench
        return {"error": str(e)[:200], "ok": False}


def nodes_block(cfg, per_gpu, steps, warmup, rank, world, local_rank, dist) -> dict | None:
    """Rounds of branch-and-bound node LP relaxations that share the matrix (configs 4 / 5), `per_gpu`
    nodes per GPU per round, dealt to the ranks with no device-side exchange (weak scaling).  A step =
    one round on every rank.  Every node is the root box with `depth` binaries fixed and - as in a
    branch and bound, where a node is opened after its parent was solved - starts from the root
    relaxation's (x, y, primal weight) through mipcu_solve_lp_batch_warm; the root solve itself is
    outside the timed region and is certificate-checked (oracle/kkt_verify.py).  One extra cold round
    (all nodes from zero, the round-1 behaviour) is timed beside it.
    Returns the result dict on rank 0, None elsewhere."""
    import torch
    from miplib_b200 import Engine
    from oracle import gen, kkt_verify
    L = gen.config(cfg)
    depth = 12 if cfg == 4 else 30                       # binaries fixed per node
    rng = np.random.default_rng(1000 + rank)

    def boxes():
        lb = np.tile(L.lb, (per_gpu, 1)); ub = np.tile(L.ub, (per_gpu, 1))
        for b in range(per_gpu):
            cols = rng.choice(L.n, depth, replace=False)
            vals = (rng.random(depth) < (0.5 if cfg == 4 else 0.1)).astype(float)
            lb[b, cols] = vals; ub[b, cols] = vals
        return lb, ub

    eng = Engine(local_rank)
    eng.load_lp(L)
    eng.set_param("iter_limit", 60000)
    root = eng.solve()
    xr, yr = eng.x(), eng.y()
    x0 = np.ascontiguousarray(np.tile(xr, (per_gpu, 1)))
    y0 = np.ascontiguousarray(np.tile(yr, (per_gpu, 1)))
    w0 = np.full(per_gpu, root.primal_weight)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(rounds, warm, cutoff=None, branch_tol=None):
        barrier()
        t0 = time.perf_counter()
        res = [eng.solve_batch(lb, ub, cutoff=cutoff, want_x=False, x0=x0 if warm else None, y0=y0 if warm else None,
                               weight0=w0 if warm else None, branch_tol=branch_tol)[0] for lb, ub in rounds]
        barrier()
        secs = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([secs], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs = float(t.item())
        return res, secs

    for _ in range(warmup):
        eng.solve_batch(*boxes(), want_x=False, x0=x0, y0=y0, weight0=w0)
    rounds = [boxes() for _ in range(steps)]
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    results, secs = timed(rounds, warm=True)
    clocks = sampler.stop() if rank == 0 else None
    cold, cold_secs = timed(rounds[:1], warm=False)
    # the same rounds as the branch and bound runs them (miplib_b200/miplib/cuda/bnb.cpp): an incumbent 2 % off
    # the root bound as cutoff, and nodes that provably cannot be pruned by bound handed back for
    # branching at 1e-4 (MIPCU_BRANCH) instead of being polished to 1e-6
    inc = root.primal_obj * (0.98 if L.maximize else 1.02)
    tree, tree_secs = timed(rounds, warm=True, cutoff=np.full(per_gpu, inc), branch_tol=np.full(per_gpu, 1e-4))
    tree_status = {}
    for st in tree:
        for s_ in st:
            tree_status[s_.status] = tree_status.get(s_.status, 0) + 1
    decided = ("optimal", "cutoff", "infeasible")
    solved = sum(sum(s.status in decided for s in st) for st in results)
    # the same boxes from a cold start must reach the same objectives (the start point changes the path only)
    agree = max([abs(w.primal_obj - c.primal_obj) / (1 + abs(c.primal_obj))
                 for w, c in zip(results[0], cold[0]) if w.status == "optimal" and c.status == "optimal"] or [0.0])
    if dist is not None:
        t = torch.tensor([float(solved)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        solved = int(t.item())
        t = torch.tensor([agree], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        agree = float(t.item())
    eng.close()
    if rank != 0:
        return None
    iters = [s.iterations for st in results for s in st]
    cold_iters = [s.iterations for s in cold[0]]
    cert = kkt_verify.certificate(L, xr, yr)
    ok = bool(root.status == "optimal" and
              max(cert["rel_primal_res"], cert["rel_dual_res"], cert["rel_gap"]) <= 1.05e-6 and
              abs(cert["primal_obj"] - root.primal_obj) <= 1e-9 * (1 + abs(root.primal_obj)) and
              agree <= 5e-6 and solved >= int(0.95 * per_gpu * world * steps))
    s0 = results[0][0]
    peak, _ = measured_peak_gbs()
    out = {
        "metric": "node_lps_per_s", "value": per_gpu * world * steps / secs, "unit": "node LP/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * secs / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C{cfg}: node LP relaxations of the synthetic "
                               f"{'set-cover' if cfg == 4 else 'multi-dimensional knapsack'} MIP "
                               f"({L.m} x {L.n}), {per_gpu} nodes per GPU per round, {depth} binaries fixed per node, "
                               f"warm-started from the root relaxation (parent)",
                   "parallelism": "nodes dealt to GPUs, no collective"},
        "node_lps_decided": solved, "node_lps_total": per_gpu * world * steps,
        "iterations_mean": float(np.mean(iters)), "iterations_max": int(np.max(iters)),
        "cold_start": {"value": per_gpu * world / cold_secs, "unit": "node LP/s", "rounds": 1,
                       "iterations_mean": float(np.mean(cold_iters)), "iterations_max": int(np.max(cold_iters)),
                       "max_rel_objective_difference_warm_vs_cold": agree},
        "tree_mode": {"value": per_gpu * world * steps / tree_secs, "unit": "node LP decided/s",
                      "what": "same rounds as the branch and bound issues them: cutoff = incumbent 2 % off the root "
                              "bound, nodes that cannot be pruned by bound returned for branching at 1e-4 "
                              "(MIPCU_BRANCH, dual bound still valid); rank 0's outcomes below",
                      "outcomes_rank0": tree_status,
                      "iterations_mean": float(np.mean([s_.iterations for st in tree for s_ in st]))},
        "root_lp": {"status": root.status, "objective": root.primal_obj, "iterations": root.iterations,
                    "certificate": {k: cert[k] for k in ("rel_primal_res", "rel_dual_res", "rel_gap", "primal_obj")}},
        "ok": ok,
        "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak,
                     "note": "one launch of each batched kernel in the first block of a round (every node still in flight)",
                     "k_row_step_batch": {"launch_us": s0.row_kernel_us, "bytes": s0.row_kernel_bytes,
                                          "GBps": s0.row_kernel_bytes / max(s0.row_kernel_us, 1e-9) / 1e3,
                                          "frac": s0.row_kernel_bytes / max(s0.row_kernel_us, 1e-9) / 1e3 / peak},
                     "k_col_step_batch": {"launch_us": s0.col_kernel_us, "bytes": s0.col_kernel_bytes,
                                          "GBps": s0.col_kernel_bytes / max(s0.col_kernel_us, 1e-9) / 1e3,
                                          "frac": s0.col_kernel_bytes / max(s0.col_kernel_us, 1e-9) / 1e3 / peak}},
        "gpu_launches": int(sum(st[0].kernel_launches for st in results)), "clocks": clocks,
    }
    return out


def run_nodes(args, rank, world, local_rank):
    """`python bench.py --workload nodes --config 4|5`: the node-LP block as the whole run."""
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    line = nodes_block(args.config, args.nodes_per_gpu, args.steps, args.warmup, rank, world, local_rank, dist)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and not line["ok"]:
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3, 4, 5])
    ap.add_argument("--workload", default="lp", choices=["lp", "nodes"])
    ap.add_argument("--nodes-per-gpu", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-nodes", action="store_true", help="skip the bounded config-4 node-LP block")
    ap.add_argument("--highs-cap", type=float, default=60.0,
                    help="wall-clock cap (s) of the HiGHS stand-in on the bench config (N = 1)")
    ap.add_argument("--shard-exchange", type=int, default=2, choices=[0, 1, 2],
                    help="N > 1: 2 = block-owner exchange (row block of K + column block of K' per rank), "
                         "1 = fused P2P reduce-scatter / primal step / all-gather kernel, 0 = ncclAllReduce")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.workload == "nodes":
        if args.config not in (4, 5):
            args.config = 4
        run_nodes(args, rank, world, local_rank)
    elif args.impl == "reference":
        run_reference(args, rank)
    else:
        run_cuda(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

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
  e2e            : the same metric through the C ABI with HOST buffers: mipcu_load_lp (H2D of the
                   CSR + vectors from pinned memory, CSC build) + mipcu_solve_lp (scaling, step
                   size, iterations) + mipcu_get_x / get_y (D2H), all inside the timed region
  roofline       : the dominant kernel (fused K.xbar + dual step, k_row_step), timed in situ by
                   CUDA events on the solver's stream during the timed solves
  cpu_baseline   : the multi-core C/OpenMP port of the same iteration (oracle/pdhg_port.c) on the
                   same LP, a bounded number of iterations, on this box's host cores

N > 1 (torchrun, one process per GPU): the same LP row-sharded across the ranks (strong scaling):
every rank owns m/N rows of A; the K'y partial sums are combined by one NCCL all-reduce per
iteration.  --impl reference: rank 0 times the CPU port; other ranks exit 0.
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
        "config": {"workload": workload_name(args.config), "m": L.m, "n": L.n, "nnz": L.nnz},
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

    if rank != 0:
        return
    last = stats[-1]
    peak, peak_src = measured_peak_gbs()
    row_us = float(np.mean([s.row_kernel_us for s in stats]))
    col_us = float(np.mean([s.col_kernel_us for s in stats]))
    if row_us <= 0:   # graph mode (small configs): time the kernels with the hook instead
        row_us, _ = eng.time_kernel(2, 50, True)
        col_us, _ = eng.time_kernel(3, 50, True)
    kernels = {"k_row_step": (row_us, last.row_kernel_bytes), "k_col_step": (col_us, last.col_kernel_bytes)}
    dom = max(kernels, key=lambda k: kernels[k][0])
    dom_us, dom_bytes = kernels[dom]
    achieved = dom_bytes / dom_us / 1e3
    line = {
        "metric": METRIC, "value": iters / resident_s, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * resident_s / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args.config), "m": m, "n": n, "nnz": int(A.nnz),
                   "l2": "inputs larger than L2 (no flush)" if A.nnz > 8e6 else "L2-resident working set",
                   "parallelism": "single GPU" if world == 1 else
                   f"row-sharded x{world}, " + {
                       2: "block-owner exchange: each rank streams its row block of K and its column block of K', "
                          "yhat / xbar slices pushed into peer memory over NVLink from the kernels' epilogues",
                       1: "fused P2P reduce-scatter + primal step + all-gather kernel over NVLink",
                       0: "NCCL all-reduce of K'y"}[int(eng.get_param("shard_exchange"))]},
        "time_to_tol_s": resident_s / args.steps,
        "iterations_per_solve": iters / args.steps,
        "device_iter_seconds_per_solve": dev_iter_s / args.steps,
        "kkt": {"rel_primal_res": last.rel_primal_res, "rel_dual_res": last.rel_dual_res,
                "rel_gap": last.rel_gap, "status": last.status, "primal_obj": last.primal_obj},
        "e2e": {"value": e2e_iters / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "seconds_per_solve": e2e_s / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "frac_of_nominal_8TBs": achieved / 8000.0,
                     "peak_source": peak_src, "traffic": ncu_traffic(dom) if world == 1 else None,
                     "algorithmic_bytes_per_launch": int(dom_bytes), "launch_us": dom_us,
                     "all_kernels": {k: {"launch_us": v[0], "bytes": int(v[1]), "GBps": v[1] / v[0] / 1e3}
                                     for k, v in kernels.items()}},
    }
    line["config"]["matrix_format"] = "SELL-32-256, one lane per row (CSR lanes-per-row kernels for long-row models)"
    if world > 1:
        line["roofline"]["col_step_split_us"] = {
            "partial_spmv": float(np.mean([s.col_partial_us for s in stats])),
            "exchange_and_primal": float(np.mean([s.col_exchange_us for s in stats]))}
    if world == 1:
        # the two plain products (B_Ax / B_ATy over >= 100 repetitions, CUDA events on
        # the library's stream through the mipcu_time_kernel hook); after the timed solves
        spmv = {}
        for which, name in ((0, "Ax"), (1, "ATy")):
            us, nbytes = eng.time_kernel(which, 100, False)
            spmv[name] = {"launch_us": us, "bytes": int(nbytes), "GBps": nbytes / us / 1e3,
                          "frac_of_peak": nbytes / us / 1e3 / peak}
        line["spmv"] = spmv
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(L)
    print(json.dumps(line), flush=True)
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def run_nodes(args, rank, world, local_rank):
    """Secondary workload (configs 4 / 5): rounds of branch-and-bound node LP relaxations that share
    the matrix, 128 nodes per GPU per round, dealt to the ranks with no device-side exchange (weak
    scaling).  A step = one round on every rank.  `python bench.py --workload nodes --config 4`."""
    import torch
    from miplib_b200 import Engine
    from oracle import gen, kkt_verify
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    L = gen.config(args.config)
    per_gpu = args.nodes_per_gpu
    depth = 12 if args.config == 4 else 30                       # binaries fixed per node
    rng = np.random.default_rng(1000 + rank)

    def boxes():
        lb = np.tile(L.lb, (per_gpu, 1)); ub = np.tile(L.ub, (per_gpu, 1))
        for b in range(per_gpu):
            if rank == 0 and b == 0:
                continue                                           # node 0 of rank 0 is the root
            cols = rng.choice(L.n, depth, replace=False)
            vals = (rng.random(depth) < (0.5 if args.config == 4 else 0.1)).astype(float)
            lb[b, cols] = vals; ub[b, cols] = vals
        return lb, ub

    eng = Engine(local_rank)
    eng.load_lp(L)
    eng.set_param("iter_limit", 60000)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        eng.solve_batch(*boxes(), want_x=False)
    rounds = [boxes() for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    results = [eng.solve_batch(lb, ub, want_x=(rank == 0 and i == 0)) for i, (lb, ub) in enumerate(rounds)]
    barrier()
    secs = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([secs], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    solved = sum(sum(s.status in ("optimal", "cutoff", "infeasible") for s in st) for st, _ in results)
    if dist is not None:
        t = torch.tensor([float(solved)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        solved = int(t.item())
    if rank != 0:
        return
    st0, x0 = results[0]
    root = st0[0]
    iters = [s.iterations for st, _ in results for s in st]
    line = {
        "metric": "node_lps_per_s", "value": per_gpu * world * args.steps / secs, "unit": "node LP/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C{args.config}: node LP relaxations of the synthetic "
                               f"{'set-cover' if args.config == 4 else 'multi-dimensional knapsack'} MIP "
                               f"({L.m} x {L.n}), {per_gpu} nodes per GPU per round, {depth} binaries fixed per node",
                   "parallelism": "nodes dealt to GPUs, no collective"},
        "node_lps_decided": solved, "node_lps_total": per_gpu * world * args.steps,
        "iterations_mean": float(np.mean(iters)), "iterations_max": int(np.max(iters)),
        "root_lp": {"status": root.status, "objective": root.primal_obj, "rel_primal_res": root.rel_primal_res,
                    "rel_gap": root.rel_gap},
        "gpu_launches": int(sum(st[0].kernel_launches for st, _ in results)), "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


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

#!/usr/bin/env python
"""bench.py — MIRK Newton steps/s on BASELINE config C2 (MIRK6, n = 16 states, 20 000 mesh nodes, FP64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one Newton step of the collocation system on the fixed C2 mesh: residual Phi + boundary
rows, all (N-1) Jacobian blocks [L_i R_i] + boundary blocks, the almost-block-diagonal solve, the
update y -= delta, and |F|_inf (unit of work M1).  Every step starts from the same
stored guess so each does identical work.

  value    device-resident throughput: K steps timed with CUDA events on the solver's stream
  e2e      the same step through the C ABI with HOST buffers: mesh + guess copied host->device from
           pinned memory, one Newton step, solution and |F|_inf copied back, every step
  roofline the dominant kernel's algorithmic bytes / its CUDA-event duration vs measured HBM copy peak
  cpu_baseline  the CPU oracle (a restatement of the reference's algorithm; the Julia reference
           cannot run in this image) timed on this box's host cores, 1 thread like the reference

N > 1 (torchrun, one rank per GPU): the Newton path of ONE problem this size does not need more than
one GPU, so ranks run independent C2 problems (an ensemble of large BVPs: independent units, no
data-path collective), weak scaling; the timed region is bracketed by barriers and the slowest rank
counts.  `--impl reference` runs the CPU oracle on rank 0 only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "mirk_newton_steps_per_sec"
UNIT = "newton_steps/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_steps(cfg, steps):
    """K Newton steps of the CPU oracle on the full C2 problem (abstol = 0 so it never stops early)."""
    from oracle import oracle as O
    ws = O.Workspace(O.builtin(cfg.problem), cfg.order, cfg.p, cfg.mesh, cfg.y0)
    t = time.perf_counter()
    ret, it, nrm = ws.newton(abstol=0.0, maxiters=steps)
    dt = time.perf_counter() - t
    return it, dt, nrm


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.warmup > 0:
        cpu_steps(cfg, min(args.warmup, 2))
    it, dt, nrm = cpu_steps(cfg, args.steps)
    value = it / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / it, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _config(cfg, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"{it} full-size Newton steps (oracle/mirk_oracle.c orc_newton; the reference's "
                                   "hot path is single-threaded and Julia is absent from this image)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _config(cfg, n_gpus):
    return {"workload": f"C2: {cfg.desc}; Newton step = residual + ABD Jacobian + block-cyclic-reduction solve + update",
            "problem": cfg.problem, "order": cfg.order, "n_states": cfg.n, "mesh_nodes": cfg.N,
            "unknowns": cfg.N * cfg.n, "problems_per_gpu": 1, "parallelism": f"independent problems x{n_gpus}",
            "l2_policy": "per-step working set (Jacobian blocks + elimination factors, ~4 x 82 MB) exceeds the 126 MB L2"}


def run_ensemble(args):
    """BASELINE config C3: EnsembleProblem pendulum sweep, 262 144 BVPs, MIRK4, dt = 0.05, adaptive.
    A step = one complete batched solve of this rank's shard.  strong scaling: the 262 144 trajectories are
    block-partitioned over the ranks (no data-path collective); weak: 262 144 per rank."""
    import math

    import numpy as np
    import torch
    import torch.distributed as dist

    import mirk_b200 as M
    from boundaryvaluediffeq_jl_b200 import configs, ensemble as E

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    total = args.trajectories * (world if args.ensemble_scaling == "weak" else 1)
    params_all = configs.c3_ensemble_params(total)
    first, count = E.partition(total, world)[rank]
    prob = M.BVProblem("pendulum", [math.pi / 2, math.pi / 2], (0.0, math.pi / 2), p=[9.81])
    h = E.EnsembleHandle(prob, M.MIRK4(), count, 0.05, device=local)
    pin = torch.empty(count, 1, dtype=torch.float64).pin_memory()
    pin.numpy()[:] = params_all[first:first + count]
    h.set_inputs(pin.numpy(), prob.u0)
    for _ in range(args.warmup):
        h.run()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ms = sum(h.run() for _ in range(args.steps))
    barrier()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    value = total * args.steps / (float(t.item()) * 1e-3)
    res = h.results()
    conv = torch.tensor([int(np.sum(res["retcodes"] == 0))], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(conv)
    # end to end: parameters from pinned host memory, batched solve, outcomes back to the host, every step
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        h.set_inputs(pin.numpy(), prob.u0)
        h.run()
        res = h.results()
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        cpu = None
        if world == 1:
            from oracle import oracle as O
            nthreads = os.cpu_count() or 1
            sample = 8192
            t1 = time.perf_counter()
            O.ensemble_solve(O.builtin("pendulum"), 4, params_all[:sample], prob.u0, prob.tspan, 32, nthreads=nthreads)
            dt = time.perf_counter() - t1
            cpu = {"value": sample / dt, "unit": "bvp_solves/s", "cores": nthreads, "kind": "port",
                   "sample": f"first {sample} trajectories of the sweep, OpenMP over {nthreads} host threads (mirrors EnsembleThreads)"}
        its = float(np.mean(res["newton_iters"]))
        nm = float(np.mean(res["n_mesh"]))
        print(json.dumps({
            "metric": "ensemble_bvp_solves_per_sec", "value": value, "unit": "bvp_solves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(t.item()) / args.steps,
            "higher_is_better": True, "scaling": args.ensemble_scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "C3: EnsembleProblem pendulum parameter sweep, g/L ~ U(8,12), MIRK4, dt=0.05, adaptive, abstol=1e-6",
                       "trajectories_total": total, "trajectories_per_gpu": count, "parallelism": f"trajectory shards x{world}",
                       "l2_policy": "per-trajectory state slab (8.3 GB per 262144 trajectories) exceeds L2"},
            "converged_fraction": float(conv.item()) / total, "mean_newton_iters": its, "mean_final_nodes": nm,
            "e2e": {"value": total * args.steps / float(te.item()), "unit": "bvp_solves/s",
                    "h2d_bytes_per_step": 8 * count + 16, "d2h_bytes_per_step": count * (4 * 4 + 8 * 2 + 16)},
            "gpu_launches": 2 * args.steps, "cpu_baseline": cpu, "clocks": clocks}), flush=True)
    h.close()
    if world > 1:
        dist.destroy_process_group()


def run_partitioned(args):
    """Mesh-partitioned single problem: C2's chain (n = 16, MIRK6), `--nint` intervals PER RANK (weak scaling),
    one NCCL all-gather of the reduced interface relation + one 8-byte all-reduce per Newton step."""
    import torch
    import torch.distributed as dist

    import mirk_b200 as M
    from boundaryvaluediffeq_jl_b200 import configs, partition

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    maker = configs.c5_chain16 if args.workload == "c5part" else configs.c2_chain8
    c = maker(args.nint * world + (world - 1))
    prob = M.BVProblem(c.problem, c.y0, c.tspan, p=c.p, mesh=c.mesh)
    cache, (lo, hi) = partition.init_partitioned(prob, M.MIRK6(), device=local, chunk=args.chunk)
    cache.bench_newton_steps(max(args.warmup, 5))
    dist.barrier()
    torch.cuda.synchronize()
    st, ms, ph, launches = cache.bench_newton_steps(args.steps)
    dist.barrier()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        per_step = float(t.item()) / args.steps
        print(json.dumps({
            "metric": "mirk_newton_steps_per_sec_mesh_partitioned", "value": 1e3 / per_step, "unit": "newton_steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step, "higher_is_better": True,
            "scaling": "weak", "dtype": "f64", "data": "synthetic",
            "mesh_interval_updates_per_sec": (c.N - 1) * 1e3 / per_step,
            "config": {"workload": f"{c.desc}; mesh partitioned into {world} segments", "mesh_nodes_total": c.N,
                       "mesh_nodes_per_gpu": hi - lo + 1, "n_states": c.n,
                       "collectives_per_step": "1 all-reduce(max) of 8 B + 1 all-gather of (2n^2+n+2Ln+L) doubles per rank"},
            "phases_us_rank0": [round(1e3 * p / args.steps, 1) for p in ph[:7]], "gpu_launches": int(launches)}), flush=True)
    cache.close()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nint", type=int, default=19999, help="mesh intervals (default: C2's 19 999)")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c2part", "c5part"],
                    help="c2: the headline Newton-step metric (default); c3: ensemble sweep; c2part/c5part: mesh-partitioned")
    ap.add_argument("--trajectories", type=int, default=262144)
    ap.add_argument("--ensemble-scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--profile", action="store_true", help="device-timed steps only (for ncu runs; prints no bench line)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "b200" and args.workload == "c3":
        return run_ensemble(args)
    if args.impl == "b200" and args.workload in ("c2part", "c5part"):
        return run_partitioned(args)

    import numpy as np

    import mirk_b200 as M
    from boundaryvaluediffeq_jl_b200 import configs

    cfg = configs.c2_chain8(args.nint)
    if args.impl == "reference":
        return run_reference(args, cfg)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n, N = cfg.n, cfg.N
    mesh_pinned = torch.empty(N, dtype=torch.float64).pin_memory()
    y_pinned = torch.empty(N, n, dtype=torch.float64).pin_memory()
    out_t = torch.empty(N, dtype=torch.float64).pin_memory()
    out_y = torch.empty(N, n, dtype=torch.float64).pin_memory()
    mesh_pinned.numpy()[:] = cfg.mesh
    y_pinned.numpy()[:] = cfg.y0

    cache = M.init(M.BVProblem(cfg.problem, y_pinned.numpy(), cfg.tspan, p=cfg.p, mesh=mesh_pinned.numpy()),
                   M.MIRK6(), adaptive=False, device=local, chunk=args.chunk)
    # ---- device-resident timing ------------------------------------------------------------------
    st, _, _, _ = cache.bench_newton_steps(args.warmup)
    assert st == 0, "warm-up Newton step failed"
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    st, total_ms, phases, launches = cache.bench_newton_steps(args.steps)
    barrier()
    assert st == 0
    if args.profile:
        print(json.dumps({"profile_run": True, "ms_per_step": total_ms / args.steps,
                          "phases_us": [round(1e3 * p / args.steps, 1) for p in phases[:7]]}))
        cache.close()
        return
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * args.steps / (total_ms_max * 1e-3)

    # ---- end to end through the C ABI with host buffers ---------------------------------------------
    import ctypes as C
    from boundaryvaluediffeq_jl_b200 import _lib as B
    L = B.lib()
    dptr = lambda tns: C.cast(tns.data_ptr(), B.dp)  # noqa: E731
    nrm = C.c_double(0)

    def e2e_step():
        B.check(L.mirk_set_mesh_guess(cache._h, N, dptr(mesh_pinned), dptr(y_pinned)))
        B.check(L.mirk_newton_step(cache._h, C.byref(nrm)))
        B.check(L.mirk_get_solution(cache._h, dptr(out_t), dptr(out_y)))

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps / float(te.item())
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        # ---- roofline of the dominant kernel -------------------------------------------------------
        names = ["(unused)", "residual+jacobian_blocks", "abd_reduce_level0", "abd_reduce_upper", "abd_tail+closing_solve",
                 "abd_backsub", "update"]
        per_step_ms = [p / args.steps for p in phases[:7]]
        k = int(np.argmax(per_step_ms))
        nn, ni = n * n, N - 1
        s = 5 if cfg.order == 6 else 3
        alg_bytes = {
            "(unused)": 0,
            "residual+jacobian_blocks": 8 * (n * N + N + (s + 1) * n * ni + 2 * nn * ni),
            # reads L_i, R_i, Phi_i; writes the elimination factors of every eliminated node and the
            # collapsed relations: together again (2 n^2 + n) doubles per interval
            "abd_reduce_level0": 8 * 2 * (2 * nn + n) * ni,
            "abd_reduce_upper": 8 * 2 * (2 * nn + n) * ni // 7,
            "abd_tail+closing_solve": 8 * 2 * (2 * nn + n) * 64,
            "abd_backsub": 8 * ((2 * nn + n) * ni + n * N),
            "update": 8 * 3 * n * N,
        }
        hbm_peak, peak_src = _peaks()
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath) and cfg.N == 20000:   # ncu-measured DRAM bytes per launch of that kernel (static capture)
            with open(tpath) as fh:
                traffic = json.load(fh).get(names[k])
        achieved = alg_bytes[names[k]] / (per_step_ms[k] * 1e-3) * 1e-9
        step_bytes = 8 * (2 * n * N + N + 2 * 2 * nn * ni)  # B = 32 n^2 N + 16 n N
        # ---- CPU baseline (bounded sample: a few full-size steps, ~1 s each) ---------------------------
        cpu = None
        if world == 1:
            it, dt, _ = cpu_steps(cfg, 8)
            cpu = {"value": it / dt, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"{it} full-size C2 Newton steps of the CPU oracle (restatement of the reference "
                             "algorithm, single thread like the reference's hot path; Julia is absent here)"}
        fp64, hbm_live = C.c_double(0), C.c_double(0)
        B.check(L.mirk_measure_peaks(local, C.byref(fp64), C.byref(hbm_live)))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": _config(cfg, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * (N + N * n),
                    "d2h_bytes_per_step": 8 * (N + N * n) + 8},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": names[k], "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes[names[k]], "kernel_ms": per_step_ms[k],
                         "whole_step": {"algorithmic_bytes": step_bytes,
                                        "achieved_gbs": step_bytes / (total_ms_max / args.steps * 1e-3) * 1e-9},
                         "live_peaks": {"fp64_fma_tflops": fp64.value, "hbm_copy_gbs": hbm_live.value}},
            "phases_ms_per_step": dict(zip(names, per_step_ms)),
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    cache.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — MIRK Newton steps/s on BASELINE config C2 (MIRK6, n = 16 states, 20 000 mesh nodes, FP64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one Newton step of the collocation system on a fixed mesh of C2's shape: residual Phi + boundary
rows, all Jacobian blocks [L_i R_i] + boundary blocks, the almost-block-diagonal solve, the update
y -= delta, and |F|_inf (unit of work M1).  Every step starts from the same stored guess so
each does identical work.

  N = 1   one C2 problem on one GPU.
  N > 1   (torchrun, one rank per GPU) the MESH-PARTITIONED mode the north star names: ONE boundary value
          problem on 19 999 x N intervals, every rank holding a C2-sized segment (weak scaling).  Per Newton
          step every rank reduces its segment to one interface relation, the relations are exchanged (remote
          stores into peer memory over NVLink from inside the pack kernel — `--exchange nccl` selects
          ncclAllGather instead), every rank solves the interface system and back-substitutes its segment; one
          more 24-byte exchange max-reduces |F|_inf and the status word.  A step over N segments counts as N units
          (C2-sized Newton steps), so `value` = N x steps / time and ideal weak scaling is N x the 1-GPU value.
          Rank 0 also solves the whole problem on one GPU and the line carries the partitioned iterate's distance
          from it (`parity_vs_single_gpu`, asserted <= 1e-10 relative).

  value    device-resident throughput: K steps timed with CUDA events on the solver's stream (max over ranks)
  e2e      the same step through the C ABI with HOST buffers: (this rank's segment of) mesh + guess copied
           host->device from pinned memory, one Newton step, solution copied back, every step.  `--e2e-handles`
           (default 12) handles on as many host threads keep independent problem instances in flight, so copies
           overlap kernels and the latency-bound phases of one elimination share the GPU with those of another
           (`e2e.serial_value` is the one-handle, one-thread figure).
  roofline the dominant kernel's algorithmic bytes (its share of SURVEY §8d's B = 32 n^2 N + 16 n N) / its
           CUDA-event duration against the measured HBM copy peak, plus the FP64 fraction (C2 sits at
           machine balance) for that kernel and for the whole step
  cpu_baseline  the CPU oracle (a restatement of the reference's algorithm; the Julia reference cannot run in
           this image) timed on this box's host cores, 1 thread like the reference's hot path
  extra    `c3_strong` (BASELINE config C3: 262 144 pendulum BVPs sharded over the ranks, solves/s, converged
           fraction) and `c5part` (config C5: n = 32, 2 000 000 nodes, mesh-partitioned over the ranks, ms per
           Newton step) measured in the same run (`--no-extra` skips them)

`--impl reference` runs the CPU oracle on rank 0 only and never maps libmirkb200.so.
"""
import argparse
import hashlib
import json
import math
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
C2_NINT = 19999


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernel_source_sha16():
    """Hash of the CUDA sources: profiles/r02_traffic.json is only trusted for the build it was captured on."""
    d = os.path.join(ROOT, "boundaryvaluediffeq.jl_b200", "csrc")
    h = hashlib.sha256()
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".cpp")) or f == "Makefile":
            with open(os.path.join(d, f), "rb") as fh:
                h.update(f.encode())
                h.update(fh.read())
    return h.hexdigest()[:16]


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


# ---- the CPU arm ------------------------------------------------------------------------------------
def cpu_threads():
    """host threads of the CPU arm: every core for the interval-parallel loops (residual, Jacobian blocks); the
    almost-block-diagonal elimination stays sequential, as the reference's LU is"""
    return max(1, os.cpu_count() or 1)


def cpu_steps(cfg, steps, nthreads=1):
    """K Newton steps of the CPU oracle on the full C2 problem (abstol = 0 so it never stops early)."""
    from oracle import oracle as O
    O.set_interval_threads(nthreads)
    try:
        ws = O.Workspace(O.builtin(cfg.problem), cfg.order, cfg.p, cfg.mesh, cfg.y0)
        t = time.perf_counter()
        ret, it, nrm = ws.newton(abstol=0.0, maxiters=steps)
        dt = time.perf_counter() - t
    finally:
        O.set_interval_threads(1)
    return it, dt, nrm


CPU_NOTE = ("oracle/mirk_oracle.c orc_newton: a plain C restatement of the reference algorithm (gcc -O3 "
            "-march=x86-64-v3, FMA contraction off; register-blocked unit-stride dense products, the 6 n^3 products per "
            "interval MIRK6 needs, sequential row-pivoted ABD elimination), single thread like the reference's hot "
            "path; the Julia reference cannot run in this image, so this is context, not a tuned-CPU comparison")
CPU_NOTE_MT = (" Threads: the residual and the Jacobian blocks run over mesh intervals on all host threads (more than the "
               "reference's own single-threaded loop uses), the elimination is sequential; single_thread_value is the same "
               "run on one thread.")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import mirk_b200  # noqa: F401  (package alias only; the shared library is NOT loaded on this arm)
    from boundaryvaluediffeq_jl_b200 import configs
    from oracle import oracle as O
    configs.set_mesh_provider(O.mesh_uniform)
    cfg = configs.c2_chain8(args.nint)
    nt = cpu_threads()
    if args.warmup > 0:
        cpu_steps(cfg, min(args.warmup, 2), nt)
    it, dt, nrm = cpu_steps(cfg, args.steps, nt)
    value = it / dt
    it1, dt1, _ = cpu_steps(cfg, min(args.steps, 3), 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / it, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _config(cfg, args.gpus, None),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nt, "kind": "port", "single_thread_value": it1 / dt1,
                         "sample": f"{it} full-size C2 Newton steps; " + CPU_NOTE + CPU_NOTE_MT},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _config(cfg, n_gpus, exchange):
    if n_gpus <= 1 or exchange is None:
        par = "one problem on one GPU" if n_gpus <= 1 else f"(CPU arm) 1 thread; the GPU arm partitions one mesh over {n_gpus} GPUs"
        nodes_total = cfg.N
    else:
        par = (f"mesh-partitioned x{n_gpus}: one BVP on {cfg.nint * n_gpus + 1} nodes, one C2-sized segment per GPU; "
               f"interface relations exchanged by {'peer-memory pushes over NVLink (CUDA IPC)' if exchange == 'p2p' else 'ncclAllGather'}")
        nodes_total = cfg.nint * n_gpus + 1
    return {"workload": f"C2: {cfg.desc}; Newton step = residual + ABD Jacobian + block-cyclic-reduction solve + update",
            "problem": cfg.problem, "order": cfg.order, "n_states": cfg.n, "mesh_nodes": cfg.N,
            "mesh_nodes_total": nodes_total, "unknowns": nodes_total * cfg.n, "parallelism": par,
            "unit_of_value": "C2-sized Newton steps/s (a partitioned step over N segments counts N)",
            "l2_policy": "per-step working set (Jacobian blocks + elimination factors, ~4 x 82 MB per GPU) exceeds the 126 MB L2"}


# ---- work model -------------------------------------------------------------------------
def work_model(n, N, order):
    """Algorithmic bytes and flops of one Newton step, split over the kernels so that the parts SUM to §8(d)'s
    B = 8 [2 n N + N + 2 * 2 n^2 (N-1)] and F = F_f + F_J + F_S.  Factor write-back / re-reads of the elimination
    are overhead traffic, not algorithmic bytes."""
    ni, nn = N - 1, n * n
    s, g = (5, 6) if order == 6 else (3, 2)
    cf = 6 * n + (n // 2) * 20       # RHS flop count of the pendulum chain / Bratu lines (sin, exp ~ 20 flops)
    Ff = ni * (s * cf + 2 * n * (s * (s + 1) // 2) + (2 * s + 1) * n)
    FJ = ni * (s * 4 * n + 2 * g * n ** 3)
    FS = (14.0 / 3.0) * n ** 3 * ni
    bytes_ = {
        "residual+jacobian_blocks": 8 * (n * N + N + 2 * nn * ni),   # read y, mesh; write L_i, R_i
        "abd_reduce_level0": 8 * 2 * nn * ni,                          # read L_i, R_i
        "abd_reduce_upper": 0, "abd_tail+closing_solve": 0,
        "abd_backsub": 8 * n * N,                                      # write y_new (update fused)
        "update": 0, "(unused)": 0,
    }
    # F_S = 14/3 n^3 per eliminated node (sequential partial-pivoting count; the ~1.65x extra of cyclic reduction is
    # not credited), attributed to the level that eliminates the node: level 0 collapses groups of c0 intervals
    c0 = min(16, max(8, -(-ni // (148 * 12)))) if n <= 16 else 8
    f0 = (c0 - 1.0) / c0
    flops = {
        "residual+jacobian_blocks": Ff + FJ,
        "abd_reduce_level0": FS * f0, "abd_reduce_upper": FS * (1.0 - f0) * 0.99, "abd_tail+closing_solve": FS * (1.0 - f0) * 0.01,
        "abd_backsub": 4.0 * n * n * N, "update": 0, "(unused)": 0,
    }
    return bytes_, flops, sum(bytes_.values()), Ff + FJ + FS + 4.0 * n * n * N


PHASES = ["(unused)", "residual+jacobian_blocks", "abd_reduce_level0", "abd_reduce_upper", "abd_tail+closing_solve",
          "abd_backsub", "update"]


def _dist_setup():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    return rank, world, local, barrier, allmax


def _make_cache(M, partition, cfg_full, world, local, args, pinned=None):
    """A C2-shaped handle: the whole problem (world = 1) or this rank's segment of the partitioned one."""
    if world == 1:
        y, mesh = (pinned if pinned is not None else (cfg_full.y0, cfg_full.mesh))
        prob = M.BVProblem(cfg_full.problem, y, cfg_full.tspan, p=cfg_full.p, mesh=mesh)
        return M.init(prob, M.MIRK6() if cfg_full.order == 6 else M.MIRK4(), adaptive=False, device=local,
                      chunk=args.chunk), (0, cfg_full.N - 1)
    prob = M.BVProblem(cfg_full.problem, cfg_full.y0, cfg_full.tspan, p=cfg_full.p, mesh=cfg_full.mesh)
    return partition.init_partitioned(prob, M.MIRK6() if cfg_full.order == 6 else M.MIRK4(), device=local,
                                      chunk=args.chunk, exchange=args.exchange)


def run_newton(args):
    """The headline line: C2 Newton steps (one GPU) / mesh-partitioned C2-sized segments (N GPUs)."""
    import ctypes as C

    import numpy as np
    import torch

    import mirk_b200 as M
    from boundaryvaluediffeq_jl_b200 import _lib as B
    from boundaryvaluediffeq_jl_b200 import configs, partition

    rank, world, local, barrier, allmax = _dist_setup()
    cfg1 = configs.c2_chain8(args.nint)                       # one segment's shape (for the work model / CPU arm)
    cfg = cfg1 if world == 1 else configs.c2_chain8(args.nint * world)   # the whole problem
    n = cfg.n
    L = B.lib()
    cache, (lo, hi) = _make_cache(M, partition, cfg, world, local, args)
    Nloc = hi - lo + 1

    # ---- device-resident timing ------------------------------------------------------------------
    st, _, _, _ = cache.bench_newton_steps(max(args.warmup, 3))
    assert st == 0, "warm-up Newton step failed"
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    st, total_ms, phases, launches = cache.bench_newton_steps(args.steps)
    barrier()
    assert st == 0
    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": total_ms / args.steps,
                              "phases_us": [round(1e3 * p / args.steps, 1) for p in phases[:7]]}))
        cache.close()
        return
    total_ms_max = allmax(total_ms)
    value = world * args.steps / (total_ms_max * 1e-3)

    # ---- parity of the partitioned iterate against the single-GPU solve of the same problem ---------
    parity = None
    if world > 1:
        import torch.distributed as dist
        B.check(L.mirk_set_mesh_guess(cache._h, Nloc, cfg.mesh[lo:hi + 1].ctypes.data_as(B.dp),
                                      np.ascontiguousarray(cfg.y0[lo:hi + 1]).ctypes.data_as(B.dp)))
        norms = []
        for _ in range(3):
            st, nrm = cache.newton_step()
            norms.append(nrm)
        full = partition.gather_solution(cache, cfg.N)
        if rank == 0:
            ref = M.init(M.BVProblem(cfg.problem, cfg.y0, cfg.tspan, p=cfg.p, mesh=cfg.mesh), M.MIRK6(), adaptive=False,
                         device=local, chunk=args.chunk)
            ref_norms = [ref.newton_step()[1] for _ in range(3)]
            _, u = ref.solution()
            ref.close()
            rel = float(np.max(np.abs(full - u)) / np.max(np.abs(u)))
            parity = {"newton_steps": 3, "max_rel_diff": rel, "resid_norms_partitioned": norms,
                      "resid_norms_single_gpu": ref_norms, "tolerance": 1e-10}
            assert rel <= 1e-10, f"partitioned iterate differs from the single-GPU one: {rel:.3e}"
        dist.barrier()

    # ---- end to end through the C ABI with host buffers ---------------------------------------------
    dptr = lambda tns: C.cast(tns.data_ptr(), B.dp)  # noqa: E731

    def pinned_set():
        mp = torch.empty(Nloc, dtype=torch.float64).pin_memory()
        yp = torch.empty(Nloc, n, dtype=torch.float64).pin_memory()
        ot = torch.empty(Nloc, dtype=torch.float64).pin_memory()
        oy = torch.empty(Nloc, n, dtype=torch.float64).pin_memory()
        mp.numpy()[:] = cfg.mesh[lo:hi + 1]
        yp.numpy()[:] = cfg.y0[lo:hi + 1]
        return mp, yp, ot, oy

    def e2e_loop(handle, bufs, count):
        mp, yp, ot, oy = bufs
        nrm = C.c_double(0)
        for _ in range(count):
            B.check(L.mirk_set_mesh_guess(handle, Nloc, dptr(mp), dptr(yp)))
            B.check(L.mirk_newton_step(handle, C.byref(nrm)))
            B.check(L.mirk_get_solution(handle, dptr(ot), dptr(oy)))

    bufs0 = pinned_set()
    e2e_loop(cache._h, bufs0, args.warmup)
    barrier()
    t0 = time.perf_counter()
    e2e_loop(cache._h, bufs0, args.steps)
    barrier()
    e2e_serial = world * args.steps / allmax(time.perf_counter() - t0)
    # two independent problem instances in flight (handles are re-entrant, one stream each): copies of one
    # overlap the kernels of the other.  Partitioned handles are collective, so both host threads of every rank
    # would have to interleave identically — the pipelined figure is taken at N = 1 only.
    e2e_value, e2e_mode = e2e_serial, "one handle, one host thread"
    concurrent = None
    # (partitioned handles are collective: every rank drives handle k from its host thread k, and each handle has its
    #  own peer-memory exchange buffers, so the handles' collectives cannot interleave wrongly — not true of NCCL, whose
    #  communicators would need a global launch order, so that flavour keeps the serial figure)
    if world == 1 or getattr(cache, "exchange", None) == "p2p":
        nh = max(2, args.e2e_handles)
        if world > 1:
            # partitioned handles spin on their peers inside kernels: more concurrent streams than hardware queues
            # (CUDA_DEVICE_MAX_CONNECTIONS = 8) would put one handle's push behind another handle's wait
            nh = min(nh, 4)
        extra_caches = [_make_cache(M, partition, cfg, world, local, args)[0] for _ in range(nh - 1)]
        handles = [(cache._h, bufs0)] + [(c._h, pinned_set()) for c in extra_caches]
        for h, b in handles[1:]:
            e2e_loop(h, b, args.warmup)
        # at least 24 steps per handle: with few the start-up of the host threads is what gets timed
        per = max(24, (args.steps + nh - 1) // nh)
        # three passes, the median reported (a pass is ~50 ms of wall clock: one descheduled host thread shows)
        e2e_runs = []
        for _ in range(3):
            th = [threading.Thread(target=e2e_loop, args=(h, b, per)) for h, b in handles]
            barrier()
            t0 = time.perf_counter()
            for t in th:
                t.start()
            for t in th:
                t.join()
            torch.cuda.synchronize()
            e2e_runs.append(world * nh * per / allmax(time.perf_counter() - t0))
        e2e_value = sorted(e2e_runs)[1]
        e2e_mode = (f"{nh} handles on {nh} host threads, {per} steps each, median of 3 passes ({', '.join('%.0f' % v for v in e2e_runs)}): "
                    "independent problem instances in flight, so the copies of one overlap "
                    "the kernels of another and the narrow upper levels of one elimination share the GPU with the wide phases of another")
        # the same handles stepping device-resident problems concurrently: aggregate Newton steps/s when several
        # independent problems are in flight (the narrow upper levels of one elimination leave most SMs idle for the
        # wide phases of another) — reported beside `value`, which stays the single-problem figure
        res_ms = [0.0] * nh

        def dev_loop(k, c):
            res_ms[k] = c.bench_newton_steps(args.steps)[1]

        if world == 1:
            all_caches = [cache] + extra_caches
            for c in extra_caches:
                c.bench_newton_steps(3)
            th = [threading.Thread(target=dev_loop, args=(k, c)) for k, c in enumerate(all_caches)]
            barrier()
            t0 = time.perf_counter()
            for t in th:
                t.start()
            for t in th:
                t.join()
            torch.cuda.synchronize()
            # (bench_newton_steps repeats its steps once more for the per-phase events: 2 x steps per handle)
            concurrent = {"handles": nh, "value": 2 * nh * args.steps / (time.perf_counter() - t0), "unit": UNIT,
                          "note": "device-resident, independent C2 problems on separate streams; wall clock around the threads"}
        barrier()
        for c in extra_caches:
            c.close()
    clocks = sampler.stop() if rank == 0 else None

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel -------------------------------------------------------
        per_step_ms = [p / args.steps for p in phases[:7]]
        k = int(np.argmax(per_step_ms))
        kb, kf, step_bytes, step_flops = work_model(n, Nloc, cfg.order)
        hbm_peak, peak_src = _peaks()
        fp64, hbm_live = C.c_double(0), C.c_double(0)
        B.check(L.mirk_measure_peaks(local, C.byref(fp64), C.byref(hbm_live)))
        traffic, traffic_note = None, "no capture for this build"
        tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                tj = json.load(fh)
            if tj.get("src_sha16") == kernel_source_sha16() and Nloc == tj.get("mesh_nodes"):
                traffic = tj.get("dram_bytes_per_launch", {}).get(PHASES[k])
                traffic_note = f"ncu --set full capture of this build ({tj.get('capture')})"
            else:
                traffic_note = "profiles/r02_traffic.json was captured on another build or mesh size: not used"
        step_s = total_ms_max / args.steps * 1e-3
        kern_s = per_step_ms[k] * 1e-3
        achieved = kb[PHASES[k]] / kern_s * 1e-9
        cpu = None
        if world == 1:
            nt = cpu_threads()
            it, dt, _ = cpu_steps(cfg1, 8, nt)
            it1, dt1, _ = cpu_steps(cfg1, 4, 1)
            cpu = {"value": it / dt, "unit": UNIT, "cores": nt, "kind": "port", "single_thread_value": it1 / dt1,
                   "sample": f"{it} full-size C2 Newton steps; " + CPU_NOTE + CPU_NOTE_MT}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": _config(cfg1, world, getattr(cache, "exchange", None)),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * (Nloc + Nloc * n) * world,
                    "d2h_bytes_per_step": (8 * (Nloc + Nloc * n) + 8) * world, "mode": e2e_mode,
                    "serial_value": e2e_serial},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": PHASES[k], "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "traffic_note": traffic_note,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": kb[PHASES[k]], "kernel_ms": per_step_ms[k],
                         "fp64": {"achieved_tflops": kf[PHASES[k]] / kern_s * 1e-12, "peak_tflops": fp64.value,
                                  "frac": kf[PHASES[k]] / kern_s * 1e-12 / fp64.value,
                                  "algorithmic_flops_per_launch": kf[PHASES[k]]},
                         "whole_step": {"algorithmic_bytes": step_bytes, "achieved_gbs": step_bytes / step_s * 1e-9,
                                        "hbm_frac": step_bytes / step_s * 1e-9 / hbm_peak,
                                        "algorithmic_flops": step_flops, "achieved_tflops": step_flops / step_s * 1e-12,
                                        "fp64_frac": step_flops / step_s * 1e-12 / fp64.value,
                                        "note": "per GPU (per segment when partitioned); the per-kernel algorithmic bytes sum to this"},
                         "per_phase": {PHASES[i]: {"ms": per_step_ms[i],
                                                   "hbm_frac": kb[PHASES[i]] / (per_step_ms[i] * 1e-3) * 1e-9 / hbm_peak,
                                                   "fp64_frac": kf[PHASES[i]] / (per_step_ms[i] * 1e-3) * 1e-12 / fp64.value}
                                       for i in range(7) if per_step_ms[i] > 0 and (kb[PHASES[i]] or kf[PHASES[i]])},
                         "live_peaks": {"fp64_fma_tflops": fp64.value, "hbm_copy_gbs": hbm_live.value}},
            "phases_ms_per_step": dict(zip(PHASES, per_step_ms)),
            "concurrent_problems": concurrent,
            "parity_vs_single_gpu": parity,
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
    cache.close()
    return line


# ---- BASELINE config C3: the ensemble sweep -----------------------------------------------------------
def measure_ensemble(args, scaling="strong", with_cpu=True):
    """EnsembleProblem pendulum sweep, 262 144 BVPs, MIRK4, dt = 0.05, adaptive.  A step = one complete batched
    solve of this rank's shard.  strong: the trajectories are block-partitioned over the ranks (no data-path
    collective); weak: `--trajectories` per rank."""
    import numpy as np
    import torch

    import mirk_b200 as M
    from boundaryvaluediffeq_jl_b200 import configs, ensemble as E

    rank, world, local, barrier, allmax = _dist_setup()
    total = args.trajectories * (world if scaling == "weak" else 1)
    params_all = configs.c3_ensemble_params(total)
    first, count = E.partition(total, world)[rank]
    prob = M.BVProblem("pendulum", [math.pi / 2, math.pi / 2], (0.0, math.pi / 2), p=[9.81])
    h = E.EnsembleHandle(prob, M.MIRK4(), count, 0.05, device=local)
    pin = torch.empty(count, 1, dtype=torch.float64).pin_memory()
    pin.numpy()[:] = params_all[first:first + count]
    h.set_inputs(pin.numpy(), prob.u0)
    steps = max(1, min(args.steps, 5))
    for _ in range(2):
        h.run()
    barrier()
    ms = sum(h.run() for _ in range(steps))
    barrier()
    ms = allmax(ms)
    value = total * steps / (ms * 1e-3)
    res = h.results()
    conv = torch.tensor([int(np.sum(res["retcodes"] == 0))], dtype=torch.int64, device="cuda")
    its = torch.tensor([float(np.sum(res["newton_iters"])), float(np.sum(res["n_mesh"]))], dtype=torch.float64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(conv)
        dist.all_reduce(its)
    # end to end: parameters from pinned host memory, batched solve, outcomes back to the host, every step
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        h.set_inputs(pin.numpy(), prob.u0)
        h.run()
        res = h.results()
    barrier()
    te = allmax(time.perf_counter() - t0)
    h.close()
    out = {"metric": "ensemble_bvp_solves_per_sec", "value": value, "unit": "bvp_solves/s", "n_gpus": world,
           "steps": steps, "ms_per_step": ms / steps, "scaling": scaling, "trajectories_total": total,
           "trajectories_per_gpu": count, "converged_fraction": float(conv.item()) / total,
           "mean_newton_iters": float(its[0].item()) / total, "mean_final_nodes": float(its[1].item()) / total,
           "e2e": {"value": total * steps / te, "unit": "bvp_solves/s", "h2d_bytes_per_step": 8 * total + 16,
                   "d2h_bytes_per_step": total * (4 * 4 + 8 * 2 + 16)},
           "workload": "C3: EnsembleProblem pendulum parameter sweep, g/L ~ U(8,12), MIRK4, dt=0.05, adaptive, abstol=1e-6"}
    if rank == 0 and with_cpu and world == 1:
        from oracle import oracle as O
        nthreads = os.cpu_count() or 1
        sample = 8192
        t1 = time.perf_counter()
        O.ensemble_solve(O.builtin("pendulum"), 4, params_all[:sample], prob.u0, prob.tspan, 32, nthreads=nthreads)
        dt = time.perf_counter() - t1
        out["cpu_baseline"] = {"value": sample / dt, "unit": "bvp_solves/s", "cores": nthreads, "kind": "port",
                               "sample": f"first {sample} trajectories of the sweep, OpenMP over {nthreads} host threads (mirrors EnsembleThreads)"}
    return out


# ---- BASELINE config C5: n = 32, 2 000 000 nodes, mesh-partitioned ------------------------------------------
def measure_c5(args, which="c5"):
    """C5 (n = 32, 2 000 000 nodes, mesh-partitioned over the ranks) or, with which = "c4", BASELINE config C4
    (n = 128, N = 4000, MIRK4; one GPU)."""
    import numpy as np  # noqa: F401

    import mirk_b200 as M
    from boundaryvaluediffeq_jl_b200 import configs, partition

    rank, world, local, barrier, allmax = _dist_setup()
    c = configs.c5_chain16(args.c5_nint) if which == "c5" else configs.c4_bratu64(args.c4_nint)
    cache, (lo, hi) = _make_cache(M, partition, c, world, local, args)
    steps = max(1, min(args.steps, 5))
    st, _, _, _ = cache.bench_newton_steps(2)
    assert st == 0
    barrier()
    st, ms, ph, launches = cache.bench_newton_steps(steps)
    barrier()
    ms = allmax(ms)
    cache.close()
    per = ms / steps
    n, ni = c.n, c.N - 1
    _, _, B_, F_ = work_model(n, c.N, c.order)
    return {"metric": "mirk_newton_steps_per_sec_mesh_partitioned", "ms_per_step": per, "value": 1e3 / per,
            "unit": "newton_steps/s", "n_gpus": world, "steps": steps, "n_states": n, "mesh_nodes_total": c.N,
            "mesh_nodes_per_gpu": hi - lo + 1, "mesh_interval_updates_per_sec": ni * 1e3 / per,
            "algorithmic_gflop_per_step": F_ * 1e-9, "achieved_tflops_all_gpus": F_ / (per * 1e-3) * 1e-12,
            "algorithmic_gb_per_step": B_ * 1e-9, "achieved_gbs_all_gpus": B_ / (per * 1e-3) * 1e-9,
            "exchange": getattr(cache, "exchange", None),
            "phases_ms_rank0": dict(zip(PHASES, [p / steps for p in ph[:7]])), "gpu_launches": int(launches),
            "workload": f"{c.key}: {c.desc}; mesh partitioned into {world} segment(s)"}


# ---- BASELINE config C1: the reference's own benchmark suite (benchmark/simple_pendulum.jl:66-70) -----------------
def measure_c1(args):
    """`solve(prob_iip, alg(), dt = 0.05)` for alg in MIRK2 … MIRK6 on the simple pendulum — the five MIRK lines of the
    reference's benchmark suite, one complete adaptive solve each, through the public API (host wall clock: a solve of 33
    nodes is launch- and synchronisation-bound, a device-only time would flatter it).  Beside it the CPU restatement on
    one host core (the problem the reference is built for: n = 2, 33 -> ~50 nodes) and, as an independent solver,
    scipy.integrate.solve_bvp on the two-point relative of the problem (theta(0) = -pi/2 instead of the interior
    condition; same 4th-order formula as MIRK4, other mesh refinement)."""
    import statistics

    import numpy as np

    import mirk_b200 as M
    from oracle import oracle as O

    u0, tspan, p = [math.pi / 2, math.pi / 2], (0.0, math.pi / 2), [9.81]
    prob = M.BVProblem("pendulum", u0, tspan, p=p)
    reps = 15
    out = {"metric": "bvp_solve_wall_ms", "unit": "ms per solve (median of %d, host wall clock)" % reps, "lower_is_better": True,
           "workload": "C1: benchmark/simple_pendulum.jl, solve(prob, alg(), dt = 0.05), adaptive, abstol = 1e-6", "algs": {}}
    for name, order in (("MIRK2", 2), ("MIRK3", 3), ("MIRK4", 4), ("MIRK5", 5), ("MIRK6", 6)):
        alg = getattr(M, name)()
        sol = M.solve(prob, alg, dt=0.05)  # warm-up (graph capture, first launches)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            sol = M.solve(prob, alg, dt=0.05)
            ts.append(time.perf_counter() - t0)
        # the same on a cache that already exists (solve!(cache) after a new guess: no allocation, no stream / graph set-up)
        from boundaryvaluediffeq_jl_b200 import _lib as B
        import ctypes as C
        cache = M.init(prob, alg, dt=0.05)
        M.solve_b(cache)
        u0a = np.ascontiguousarray(np.array(u0, dtype=np.float64))
        tr = []
        for _ in range(reps):
            t0 = time.perf_counter()
            B.check(B.lib().mirk_set_uniform_guess(cache._h, C.c_double(tspan[0]), C.c_double(tspan[1]), C.c_double(0.05),
                                                   u0a.ctypes.data_as(C.POINTER(C.c_double))))
            solr = M.solve_b(cache)
            tr.append(time.perf_counter() - t0)
        reused_same = bool(np.array_equal(solr.u, sol.u))
        cache.close()
        ref = O.solve_dt(O.builtin("pendulum"), order, p, u0, tspan, 0.05)
        tc = []
        for _ in range(reps):
            t0 = time.perf_counter()
            ref = O.solve_dt(O.builtin("pendulum"), order, p, u0, tspan, 0.05)
            tc.append(time.perf_counter() - t0)
        same = sol.original["hist_n_mesh"] == ref.hist_N and sol.original["hist_newton"] == ref.hist_newton
        err = float(np.max(np.abs(sol.u - ref.u)) / np.max(np.abs(ref.u))) if len(sol.t) == len(ref.t) else None
        out["algs"][name] = {"gpu_ms": 1e3 * statistics.median(ts), "gpu_reused_cache_ms": 1e3 * statistics.median(tr),
                             "reused_cache_same_solution": reused_same,
                             "cpu_port_ms": 1e3 * statistics.median(tc), "retcode": int(sol.retcode),
                             "mesh_history": list(sol.original["hist_n_mesh"]), "newton_history": list(sol.original["hist_newton"]),
                             "same_histories_as_cpu_port": bool(same), "max_rel_diff_vs_cpu_port": err}
    # the same solve as ONE kernel launch: a one-trajectory ensemble (solve(EnsembleProblem(prob), alg; trajectories = 1):
    # the warp-per-trajectory whole-solve kernel, instantiated for MIRK4 / MIRK6) — parameters in, final mesh and solution out
    try:
        from boundaryvaluediffeq_jl_b200 import ensemble as E
        for name, order in (("MIRK4", 4), ("MIRK6", 6)):
            h = E.EnsembleHandle(prob, getattr(M, name)(), 1, 0.05)
            par = np.array([[9.81]])
            tk = []
            for i in range(reps + 2):
                t0 = time.perf_counter()
                h.set_inputs(par, prob.u0)
                h.run()
                res = h.results()
                mesh, y = h.trajectory(0)
                if i >= 2:
                    tk.append(time.perf_counter() - t0)
            ref = O.solve_dt(O.builtin("pendulum"), order, p, u0, tspan, 0.05)
            ok = int(res["retcodes"][0]) == 0 and len(mesh) == len(ref.t)
            out["algs"][name]["gpu_one_kernel_ms"] = 1e3 * statistics.median(tk)
            out["algs"][name]["one_kernel_max_rel_diff_vs_cpu_port"] = float(np.max(np.abs(y - ref.u)) / np.max(np.abs(ref.u))) if ok else None
            h.close()
    except Exception as e:  # noqa: BLE001
        out["one_kernel_error"] = str(e)[:200]
    try:
        from scipy.integrate import solve_bvp
        x = np.linspace(tspan[0], tspan[1], 33)
        y = np.full((2, 33), math.pi / 2)
        f = lambda t, u: np.vstack([u[1], -9.81 * np.sin(u[0])])  # noqa: E731
        bc = lambda ua, ub: np.array([ua[0] + math.pi / 2, ub[0] - math.pi / 2])  # noqa: E731
        tsp = []
        for _ in range(reps):
            t0 = time.perf_counter()
            r = solve_bvp(f, bc, x, y, tol=1e-6)
            tsp.append(time.perf_counter() - t0)
        out["scipy_solve_bvp"] = {"ms": 1e3 * statistics.median(tsp), "status": int(r.status), "nodes": int(r.x.size),
                                  "note": "two-point relative of C1 (theta(0) = -pi/2), independent solver, 1 core"}
    except Exception as e:  # noqa: BLE001
        out["scipy_solve_bvp"] = {"error": str(e)[:200]}
    out["cpu_baseline"] = {"kind": "port", "cores": 1, "sample": "the same 15 solves per method, oracle/mirk_oracle.c orc_solve"}
    return out


def _guarded(fn, *a, **kw):
    """extras must not take the headline line down with them"""
    try:
        return fn(*a, **kw)
    except BaseException as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nint", type=int, default=C2_NINT, help="mesh intervals per GPU (default: C2's 19 999)")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--exchange", default=None, choices=["p2p", "nccl"], help="interface exchange of the partitioned mode")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c5part", "c4"],
                    help="c2: the headline Newton-step line (default, with the extras); c3 / c5part: that workload alone")
    ap.add_argument("--no-extra", action="store_true", help="skip the C3 / C5 extras of the default run")
    ap.add_argument("--trajectories", type=int, default=262144)
    ap.add_argument("--ensemble-scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--c5-nint", type=int, default=1999999)
    ap.add_argument("--c4-nint", type=int, default=3999)
    ap.add_argument("--e2e-handles", type=int, default=12, help="independent handles / host threads of the end-to-end leg")
    ap.add_argument("--extra-timeout", type=float, default=300.0)
    ap.add_argument("--profile", action="store_true", help="device-timed steps only (for ncu runs; prints no bench line)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    if args.workload == "c3":
        out = measure_ensemble(args, args.ensemble_scaling)
        if rank == 0:
            print(json.dumps(out), flush=True)
    elif args.workload in ("c5part", "c4"):
        out = measure_c5(args, "c5" if args.workload == "c5part" else "c4")
        if rank == 0:
            print(json.dumps(out), flush=True)
    else:
        line = run_newton(args)
        if not args.profile:
            extra = None
            if not args.no_extra:
                # the extras are collective: if one rank fails the others would wait forever, so a watchdog prints the
                # headline line without them and ends the process
                def give_up():
                    if rank == 0:
                        line["extra"] = {"error": f"extras did not finish within {args.extra_timeout} s"}
                        print(json.dumps(line), flush=True)
                    os._exit(0)
                dog = threading.Timer(args.extra_timeout, give_up)
                dog.daemon = True
                dog.start()
                extra = {"c3_strong": _guarded(measure_ensemble, args, "strong", True), "c5part": _guarded(measure_c5, args)}
                if int(os.environ.get("WORLD_SIZE", "1")) == 1:
                    extra["c4"] = _guarded(measure_c5, args, "c4")
                    extra["c1"] = _guarded(measure_c1, args)
                dog.cancel()
            if rank == 0:
                line["extra"] = extra
                print(json.dumps(line), flush=True)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except ImportError:
        pass


if __name__ == "__main__":
    main()

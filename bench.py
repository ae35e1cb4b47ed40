#!/usr/bin/env python
"""bench.py -- SQP(LM) iterations per second on a batched OCP, on N B200s of one node, next to the reference's CPU path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU solver on the host cores

One "step" = one cold batched solve of the workload: trajectory initialisation + `iterations` Levenberg-Marquardt outer
iterations for every instance (LevenbergMarquardtSparse always runs all of them).  metric = instances x iterations / time.
  value      device-resident: x0 already in HBM, nothing crosses PCIe inside the timed region (CUDA events, max over ranks)
  e2e        the same through b200sqp_step with pinned HOST buffers: H2D of x0/xref, solve, D2H of trajectories+chi2+status
  roofline   the LM kernel alone against the measured HBM copy bandwidth with SURVEY.md section 8d's algorithmic bytes
  cpu_baseline  the reference (oracle/_ref, kind "reference") or the oracle port timed on this box's host cores (rank 0, N=1)
Instances shard over ranks with no data-path collective; one all-gather of the per-instance chi2 per step is the stop-test
exchange.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from control_box_rst_b200 import _abi as abi  # noqa: E402
from control_box_rst_b200 import problems  # noqa: E402

METRIC = "sqp_lm_iterations_per_second"
UNIT = "iters/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=1, help="BASELINE.json configs[] index of the OCP (1 = Van der Pol N=50)")
    ap.add_argument("--batch", type=int, default=4096, help="instances per GPU (north_star target batch; weak scaling)")
    ap.add_argument("--cpu-sample", type=int, default=4096, help="instances of the same workload timed on the host cores")
    ap.add_argument("--precision", default=None, choices=["f64", "f32"],
                    help="arithmetic of the solve: f64 (the reference's; default) or f32 (default for --config 4, which BASELINE.json quotes in fp32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the short runs of configs[2..4] appended to the default N=1 line")
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"],
                    help="stop-test exchange at N>1: chi2 stored into every rank's gather buffer by the LM kernel itself over NVLink peer "
                         "memory (p2p, default) or a separate NCCL all-gather after the solve (nccl)")
    return ap.parse_args()


def workload_name(cfg, ocp, batch, iterations, precision="f64"):
    names = {0: "van_der_pol_fd_n20", 1: "van_der_pol_fd_n50", 2: "unicycle_time_optimal_n50", 3: "cart_pole_shooting_n100", 4: "quadrotor_fd_n60"}
    return f"{names[cfg]}_batch{batch}_per_gpu_{'fp32' if precision == 'f32' else 'fp64'}_{iterations}_lm_iterations_cold_start"


def default_precision(args):
    return args.precision or ("f32" if args.config == 4 else "f64")


# ---------------------------------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md)
# ---------------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML in a thread every few ms (the timed region of this
    bench is tens of ms, too short for `nvidia-smi -lms`), with the profiling recipe's nvidia-smi query as fallback."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_s=0.002):
        self.gpu_index = gpu_index
        self.period_s = period_s
        self.samples = []   # (sm_mhz, reasons bitmask, power_w)
        self.smax = None
        self.thread = None
        self.stop_flag = threading.Event()
        self.nvml = None
        self.proc = None
        self.lines = []

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            # NVML indexes physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu_index
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu_index])
                except (ValueError, IndexError):
                    idx = self.gpu_index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                reasons = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                power = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                self.samples.append((mhz, reasons, power))
            except Exception:
                pass
            time.sleep(self.period_s)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=1)
            n = self.nvml
            names = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            reasons = sorted(k for k, bit in names.items() if any(r & bit for _, r, _ in self.samples))
            sm = [m for m, _, _ in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.smax, "reasons": reasons, "samples": len(sm),
                    "power_w_max": max((p for _, _, p in self.samples), default=None), "source": "nvml, sampled during the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi -lms 100"}


# ---------------------------------------------------------------------------------------------------------------------------
# the reference's own CPU implementation of the path, all host threads (oracle/_ref when it exists, else the oracle port)
# ---------------------------------------------------------------------------------------------------------------------------
def cpu_checker():
    from oracle import bindings

    if bindings.Reference.available():
        return bindings.Reference(), "reference"
    if not bindings.Oracle.available():
        bindings.build_oracle()
    return bindings.Oracle(), "port"


def run_cpu(ocp, opts, sample, seed):
    """-> dict: the reference's CPU path on `sample` instances over all host threads.  `value` follows SURVEY.md section 8d: iterations
    per second timed around the solver call (`_solver->solve`, summed per thread -> x threads / sum), fresh solver objects per instance;
    `wall_value` is the same batch by the wall clock including object construction and the grid update."""
    checker, kind = cpu_checker()
    cores = os.cpu_count() or 1
    x0, xref = problems.instance_data(ocp, sample, seed=seed)
    t0 = time.perf_counter()
    _, _, _, secs = checker.solve_batch(ocp, opts, x0, xref, threads=cores)
    wall = time.perf_counter() - t0
    threads = min(cores, sample)
    solve_wall = float(secs[1]) / threads if secs[1] > 0 else wall  # per-thread sums run side by side
    return {"value": sample * opts.iterations / solve_wall, "wall_value": sample * opts.iterations / wall, "wall_s": wall,
            "solver_s_per_thread": solve_wall, "kind": kind, "cores": cores, "sample": sample}


def b200_config_keys(args, ocp, n_params, world):
    return {"workload": workload_name(args.config, ocp, args.batch, problems.config(args.config)[1]["iterations"], default_precision(args)),
            "instances_total": args.batch * world, "n_grid": ocp.n_grid, "n_params": n_params}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from control_box_rst_b200 import solver

    ocp, kw, _ = problems.config(args.config)
    opts = abi.LmOptions.defaults(iterations=kw["iterations"], weights=kw["weights"])
    sample = args.batch if args.config == 1 else min(args.batch, args.cpu_sample)  # the headline workload: every instance the GPU arm solves per GPU
    for _ in range(max(0, min(args.warmup, 1))):
        run_cpu(ocp, opts, 64, seed=99)
    solver_s, wall_s, r = 0.0, 0.0, None
    for s in range(args.steps):
        r = run_cpu(ocp, opts, sample, seed=1234 + args.config)
        solver_s += r["solver_s_per_thread"]
        wall_s += r["wall_s"]
    value = sample * opts.iterations * args.steps / solver_s
    cfg = b200_config_keys(args, ocp, solver.dims_of(ocp).n_params, 1)
    cfg["note"] = ("reference CPU path (LevenbergMarquardtSparse through its hypergraph problem), fresh solver objects per instance; timed around "
                   "the solver call as SURVEY.md section 8d defines the metric; wall_value includes object construction and the grid update")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * solver_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cfg,
        "wall_value": sample * opts.iterations * args.steps / wall_s, "wall_ms_per_step": 1e3 * wall_s / args.steps,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                         "sample": f"{sample} instances of the workload per step, {args.steps} steps, std::thread static partition, solver-call time"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


CPU_SAMPLE_CAP = {0: 4096, 1: 4096, 2: 1024, 3: 256, 4: 128}  # instances of the workload the host cores solve for cpu_baseline (seconds of work)


def cpu_baseline_entry(r):
    return {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "wall_value": r["wall_value"],
            "sample": f"{r['sample']} instances of the same workload on {r['cores']} threads: {r['solver_s_per_thread']:.2f} s inside the solver call per "
                      f"thread, {r['wall_s']:.2f} s wall with object construction"}


def csrc_hash():
    """sha256 over the kernel sources and build flags: identifies the binary a profile was captured from (the .so itself is not in git)"""
    import hashlib

    h = hashlib.sha256()
    d = os.path.join(ROOT, "control_box_rst_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh", ".h", ".cpp")) or name == "Makefile":
            h.update(name.encode())
            h.update(open(os.path.join(d, name), "rb").read())
    return h.hexdigest()[:16]


def measured_traffic(key):
    """DRAM bytes of one launch of the dominant kernel from the `ncu --set full` capture in profiles/traffic.json -- only if the capture
    was taken from THIS source tree (the entry carries the hash of the kernel sources; a stale capture reads as null + the reason)."""
    try:
        t_all = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if key not in t_all:
            return None, "no ncu capture of this workload in profiles/traffic.json"
        e = t_all[key]
        if e.get("csrc_sha256") != csrc_hash():
            return None, f"stale: profiles/traffic.json was captured from kernel sources {e.get('csrc_sha256')}, this tree is {csrc_hash()}"
        return e["dram_bytes_read"] + e["dram_bytes_write"], e["source"]
    except (OSError, ValueError, KeyError) as exc:
        return None, f"profiles/traffic.json unreadable: {exc}"


def fp64_roofline(ocp, dims, B, iterations, kernel_ms, fp64_peak):
    """SURVEY.md section 8d "also report the fp64 FMA bound": algorithmic flops of the launch (control_box_rst_b200/roofline.py) over the
    kernel's time, against the fp64 FMA throughput measured live on this GPU (b200sqp_measure_fp64_peak)."""
    from control_box_rst_b200 import roofline

    flops = roofline.flops_per_iteration(ocp, dims) * B * iterations
    achieved = flops / (kernel_ms * 1e-3) / 1e12
    return {"bound": "fp64_fma", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak else None,
            "flops_per_iteration": roofline.flops_per_iteration(ocp, dims), "algorithmic_flops_per_launch": flops,
            "peak_source": "b200sqp_measure_fp64_peak: independent DFMA chains on every SM, best of 3 launches, measured in this run",
            "note": "binding roof of the fused kernel (J, H and L never reach DRAM); one count per +,-,*,/ and libm call, rejected steps not counted"}


def plugin_e2e(batch, iterations):
    """The reference's OWN API route: `batch` StructuredOptimalControlProblem objects of the workload (built and grid-updated by the
    unmodified reference), solved by one corbo::SolverB200Lm::solveBatch call -- the C++ plugin of control_box_rst_b200/adapter over the
    same C ABI.  Timed inside the drop-in binary (tests/adapter/dropin_test.cpp --bench) on the host clock around solveBatch: hypergraph
    walk, parameter gather from / scatter to the vertex objects, H2D, device solve, D2H.  The binary links the compiled reference, so it
    only exists where /root/reference was present at build time (it travels to the GPU box with the snapshot)."""
    exe = os.path.join(ROOT, "tests", "adapter", "_build", "dropin_test")
    if not os.path.exists(exe):
        return {"unavailable": "tests/adapter/_build/dropin_test not built (needs /root/reference at build time)"}
    try:
        out = subprocess.run([exe, "--bench", str(batch), "5"], capture_output=True, text=True, timeout=300)
        r = json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as exc:
        return {"unavailable": f"{type(exc).__name__}: {exc}"}
    if "error" in r:
        return {"unavailable": r["error"]}
    n = 147
    return {"value": r["value"], "unit": UNIT, "ms_per_step": r["solve_batch_ms_mean"], "ms_per_step_best": r["solve_batch_ms_best"],
            "objects": r["objects"], "kernel_ms": r["kernel_ms"], "h2d_bytes_per_step": batch * (2 * 2 + n) * 8, "d2h_bytes_per_step": batch * (n * 8 + 8 + 4),
            "build_objects_s": r["build_objects_s"],
            "note": "corbo::SolverB200Lm::solveBatch over reference OCP objects (host clock around the call; per-object vertex gather/scatter on one host thread)"}


def other_configs(device, stream, hbm_peak, fp64_peak, with_cpu):
    """BASELINE.json configs[2..4] at their full batch on this GPU, a few steps each (parity-test cases, not bench lines: recorded so
    that the driver's run carries their throughput, roofline fractions and CPU baseline too).  Same timing rules as the main line."""
    import torch

    from control_box_rst_b200 import solver

    out = {}
    for cfg, precision, key in ((2, "f64", "2"), (3, "f64", "3"), (4, "f32", "4"), (4, "f64", "4_f64")):
        try:
            ocp, kw, B = problems.config(cfg)
            iterations = kw["iterations"]
            lm = solver.BatchedLevenbergMarquardt(ocp, B, device=device)
            lm.setIterations(iterations)
            lm.setPenaltyWeights(*kw["weights"])
            lm.set_precision(precision)
            lm.set_stream(stream.cuda_stream)
            x0, xref = problems.instance_data(ocp, B, seed=1234 + cfg)
            lm.set_problem_data(x0, xref)
            flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=f"cuda:{device}")
            steps, total_ms, kernel_ms = (3 if cfg == 4 else 5), 0.0, 0.0
            for it in range(steps + 1):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                lm.initialize_trajectories()
                lm.solve(new_run=True, fetch=False)
                e1.record(stream)
                e1.synchronize()
                if it > 0:  # the first step is the warm-up
                    total_ms += e0.elapsed_time(e1)
                    kernel_ms += lm.last_solve_ms()
            extra = {}
            if cfg == 3:
                # measured, not the default: a 4-way split per instance (two 128-thread blocks per SM) beats the automatic 8-way split in
                # this multi-wave launch of a long horizon.  The automatic choice depends on the horizon only, so that an instance's result
                # never depends on the size of the batch it shares (tests/test_gpu_parity.py::test_full_size_properties_other_configs).
                lm.set_threads_per_instance(4)
                t4 = 0.0
                for it in range(steps + 1):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    lm.initialize_trajectories()
                    lm.solve(new_run=True, fetch=False)
                    e1.record(stream)
                    e1.synchronize()
                    if it > 0:
                        t4 += e0.elapsed_time(e1)
                extra["value_threads_per_instance_4"] = B * iterations * steps / (t4 * 1e-3)
            alg = lm.dims.algorithmic_bytes_per_iteration // (2 if precision == "f32" else 1) * B * iterations
            achieved = alg / (kernel_ms / steps * 1e-3) / 1e9
            entry = {"workload": workload_name(cfg, ocp, B, iterations, precision), "dtype": precision, "value": B * iterations * steps / (total_ms * 1e-3), "unit": UNIT,
                     "steps": steps, "ms_per_step": total_ms / steps, "kernel_ms": kernel_ms / steps,
                     "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak},
                     "roofline_fp64": fp64_roofline(ocp, lm.dims, B, iterations, kernel_ms / steps, fp64_peak)}
            entry.update(extra)
            lm.clear()
            del flush
            torch.cuda.empty_cache()
            if with_cpu and key != "4_f64":
                opts = abi.LmOptions.defaults(iterations=iterations, weights=kw["weights"])
                entry["cpu_baseline"] = cpu_baseline_entry(run_cpu(ocp, opts, CPU_SAMPLE_CAP[cfg], seed=1234 + cfg))
            out[key] = entry
        except Exception as exc:  # the main line must not depend on the extra configurations
            out[key] = {"error": f"{type(exc).__name__}: {exc}"}
    return out


def adaptive_config(device, with_cpu):
    """SURVEY 8f row 2 on BASELINE configs[2]'s controller: 4096 time-optimal unicycle controllers whose grids adapt independently
    (NonUniformFiniteDifferencesVariableGrid::adaptGridTimeBasedSingleStep, 3 OCP iterations per controller step), bucketed by grid size
    on the device.  Timed end to end through b200sqp_adaptive_step with host buffers (host clock: the buckets solve on their own streams)."""
    import torch

    from control_box_rst_b200 import solver

    try:
        ocp, kw, B = problems.config(2)
        iterations, m, steps = kw["iterations"], 3, 6
        rng = np.random.default_rng(77)
        r, ang = rng.uniform(0.5, 9.0, B), rng.uniform(-1.0, 1.0, B)
        x0 = np.zeros((B, 3))
        xf = np.stack([r * np.cos(ang), r * np.sin(ang), ang + rng.uniform(-0.5, 0.5, B)], axis=1)
        ad = solver.AdaptiveGridBatch(ocp, B, 5, 100, 0.1, warm_start=True, device=device)
        ad.setIterations(iterations)
        ad.setPenaltyWeights(*kw["weights"])
        ad.reserve()  # all buckets up front: a bucket allocates for the whole batch
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=f"cuda:{device}")
        total, n_hist = 0.0, []
        for it in range(steps + 5):  # five warm-up steps: the grids spread over their sizes first
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            u0, chi2, status, n = ad.step(x0 + 0.01 * it * (xf - x0), xf, num_ocp_iterations=m)
            dt = time.perf_counter() - t0
            if it >= 5:
                total += dt
            n_hist.append([int(n.min()), float(n.mean()), int(n.max())])
        stats = ad.statistics()
        ad.close()
        del flush
        torch.cuda.empty_cache()
        entry = {"workload": f"unicycle_time_optimal_adaptive_grid_n{ocp.n_grid}_b{B}_it{iterations}x{m}", "dtype": "f64",
                 "value": B * iterations * m * steps / total, "unit": UNIT, "steps": steps, "ms_per_step": total / steps * 1e3,
                 "ocp_iterations_per_step": m, "timing": "host clock around b200sqp_adaptive_step (H2D states, adaptation, solves, D2H controls)",
                 "grid_size_min_mean_max": n_hist[-1], "occupied_buckets": stats["occupied_buckets"], "splits": stats["splits"],
                 "merges": stats["merges"]}
        if with_cpu:
            from oracle import bindings  # the one other place bench.py may use oracle/: the reported CPU baseline

        if with_cpu and bindings.Reference.available():
            ref = bindings.Reference()
            opts = abi.LmOptions.defaults(iterations=iterations, weights=kw["weights"])
            sample = 8
            t0 = time.perf_counter()
            done = 0
            for i in range(sample):
                seq = np.stack([x0[i] + 0.01 * s * (xf[i] - x0[i]) for s in range(2)])
                # forked: the reference's adaptation has inputs it does not survive (oracle/bindings.py isolated)
                if bindings.isolated(lambda: ref.adaptive_steps(ocp, opts, seq, xf[i], 5, 100, 0.1, True, m)[0]) is not None:
                    done += 1
            dt = time.perf_counter() - t0
            entry["cpu_baseline"] = {"value": done * iterations * m * 2 / dt, "unit": UNIT, "cores": 1, "kind": "reference",
                                     "sample": f"{done} instances x 2 controller steps x {m} OCP iterations, one thread, wall clock incl. grid update"}
        return entry
    except Exception as exc:
        return {"error": f"{type(exc).__name__}: {exc}"}


def main_b200(args):
    import torch
    import torch.distributed as dist

    from control_box_rst_b200 import distributed, solver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    solver.load_library()  # before the first CUDA call: the library asks for one hardware queue per concurrent bucket stream (api.cpp)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path only exists as sm_100a kernels (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world

    ocp, kw, _ = problems.config(args.config)
    B = args.batch
    lm = solver.BatchedLevenbergMarquardt(ocp, B, device=local_rank)
    lm.setIterations(kw["iterations"])
    lm.setPenaltyWeights(*kw["weights"])
    iterations = kw["iterations"]
    precision = default_precision(args)
    lm.set_precision(precision)  # f32 raises for structures without a reduced-precision path
    alg_bytes_per_iteration = lm.dims.algorithmic_bytes_per_iteration // (2 if precision == "f32" else 1)  # s = 8 (fp64) / 4 (fp32)
    # a dedicated (non-default) torch stream: the library launches on it, torch's events, the L2 flush and NCCL are ordered on it
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    lm.set_stream(stream.cuda_stream)

    # every rank solves its own contiguous shard of the global instance list
    x0, xref = problems.instance_data(ocp, B, seed=1234 + args.config, offset=rank * B)
    lm.set_problem_data(x0, xref)
    ptrs = lm.device_pointers()
    chi2_local = distributed.device_view(ptrs["chi2"], B, f"cuda:{local_rank}")
    chi2_all = torch.empty(B * world, dtype=torch.float64, device=f"cuda:{local_rank}") if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=f"cuda:{local_rank}")  # > 126 MB L2
    rendezvous = torch.zeros(1, dtype=torch.float32, device=f"cuda:{local_rank}")

    # the single stop-test exchange: control_box_rst_b200.distributed.StopTestExchange -- fused into the LM kernel
    # over NVLink peer memory (p2p) or one NCCL all-gather after the solve (nccl)
    exch = distributed.StopTestExchange(lm, mode=args.collective, device=torch.device("cuda", local_rank))
    use_p2p = exch.mode == "p2p"
    exchange = exch.wait

    def device_step():
        lm.initialize_trajectories()
        lm.solve(new_run=True, fetch=False)
        exchange()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        total_ms, kernel_ms = 0.0, 0.0
        for _ in range(steps):
            if world > 1:
                # device-side rendezvous, outside the timed region: all ranks enter the step together, so a step's time is its own
                # work + exchange and not another rank's late start (the per-step analogue of the bracket barrier).  The L2 flush and a
                # short spin follow the rendezvous: equal device work on every rank, during which each rank's host thread enqueues the
                # launches of the step -- otherwise the step would start whenever that rank's Python got there (measured: 40-90 us of
                # skew between ranks at 8 GPUs, which the stop-test wait then turns into step time).
                dist.all_reduce(rendezvous)
            flush.zero_()  # evict L2 between timed iterations (outside the timed region)
            if world > 1:
                torch.cuda._sleep(2000000)  # ~1 ms: every rank's host thread has the step enqueued before the device gets to it
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            total_ms += e0.elapsed_time(e1)
            kernel_ms += lm.last_solve_ms()
        return total_ms, kernel_ms

    # ---- device-resident value ---------------------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()
    if use_p2p:
        # the fused gather must deliver exactly what an NCCL all-gather of the same solve delivers
        gathered = exch.gathered().clone()
        dist.all_gather_into_tensor(chi2_all, chi2_local)
        torch.cuda.synchronize()
        if not torch.equal(gathered, chi2_all):
            raise SystemExit("bench.py: peer-memory gather differs from the NCCL all-gather of the same solve")
        barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lm.launch_count()
    wall0 = time.perf_counter()
    total_ms, kernel_ms = timed(device_step, args.steps)
    barrier()
    wall = time.perf_counter() - wall0
    launches = lm.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms, kernel_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    per_rank = None
    if world > 1:
        # what every rank measured on its own device (the reported figures are the maxima): separates a slow GPU / rank skew from
        # a cost that every rank pays (the exchange)
        every = torch.empty(2 * world, dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_gather_into_tensor(every, t)
        every = (every.view(world, 2) / args.steps).cpu().tolist()
        per_rank = {"step_ms": [round(e[0], 5) for e in every], "kernel_ms": [round(e[1], 5) for e in every]}
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms = float(t[0]), float(t[1])
    value = B * world * iterations * args.steps / (total_ms * 1e-3)
    if lm.dims.block_dim < 8:
        # one more (untimed) step with the in-kernel phase clock: SM cycles of the whole LM kernel and of its exchange epilogue on each rank
        # (cycles x clock = time inside the kernel body; the rest of kernel_ms is launch and drain)
        lm.set_phase_profile(True)
        device_step()
        torch.cuda.synchronize()
        pc = lm.phase_cycles()
        lm.set_phase_profile(False)
        ex = torch.tensor([pc["exchange"], sum(pc.values())], dtype=torch.float64, device=f"cuda:{local_rank}")
        ex_all = ex.clone()
        if world > 1:
            ex_all = torch.empty(2 * world, dtype=torch.float64, device=f"cuda:{local_rank}")
            dist.all_gather_into_tensor(ex_all, ex)
        ex_all = ex_all.view(world, 2).cpu().tolist()
        per_rank = per_rank or {"step_ms": [round(total_ms / args.steps, 5)], "kernel_ms": [round(kernel_ms / args.steps, 5)]}
        per_rank["exchange_epilogue_kcycles"] = [round(e[0] / 1e3, 2) for e in ex_all]
        per_rank["kernel_kcycles"] = [round(e[1] / 1e3, 1) for e in ex_all]

    # ---- end to end through the C ABI with host buffers -----------------------------------------------------------------------
    n = lm.dims.n_params
    h_x0 = torch.from_numpy(x0).pin_memory()
    h_xref = torch.from_numpy(xref).pin_memory()
    h_params = torch.empty((B, n), dtype=torch.float64).pin_memory()
    h_chi2 = torch.empty(B, dtype=torch.float64).pin_memory()
    h_status = torch.empty(B, dtype=torch.int32).pin_memory()

    def e2e_step():
        lm.step_raw(h_x0.data_ptr(), h_xref.data_ptr(), h_params.data_ptr(), h_chi2.data_ptr(), h_status.data_ptr(), cold_start=True)
        exchange()

    for _ in range(2):
        e2e_step()
    barrier()
    e2e_ms, _ = timed(e2e_step, args.steps)
    barrier()
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t[0])
    e2e_value = B * world * iterations * args.steps / (e2e_ms * 1e-3)
    h2d = 2 * B * ocp.nx * 8
    d2h = B * n * 8 + B * 8 + B * 4

    # ---- the closed-loop front-end: measured states in, first controls out (b200sqp_mpc_step, cold start) -----------------------------
    h_u0 = torch.empty((B, ocp.nu), dtype=torch.float64).pin_memory()

    def mpc_step():
        lm.mpc_step_raw(0, h_x0.data_ptr(), h_xref.data_ptr(), h_u0.data_ptr(), h_chi2.data_ptr(), h_status.data_ptr())
        exchange()

    for _ in range(2):
        mpc_step()
    barrier()
    mpc_ms, _ = timed(mpc_step, args.steps)
    barrier()
    t = torch.tensor([mpc_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    mpc_ms = float(t[0])
    mpc_value = B * world * iterations * args.steps / (mpc_ms * 1e-3)
    stats = lm.statistics()

    # ---- the whole closed loop on the device (b200sqp_closed_loop): x0 in, the log out, nothing crosses PCIe between the MPC steps ------
    closed_loop = None
    if world == 1 and lm.ocp.grid != abi.GRID_FD_NONUNIFORM_VARDT:
        loop_steps = 20
        h_ul = torch.empty((loop_steps, B, ocp.nu), dtype=torch.float64).pin_memory()
        h_xl = torch.empty((loop_steps + 1, B, ocp.nx), dtype=torch.float64).pin_memory()

        def loop_call():
            lm.closed_loop_raw(2, 1, ocp.dt_ref, loop_steps, h_x0.data_ptr(), h_xref.data_ptr(), h_ul.data_ptr(), h_xl.data_ptr(), 0, 0)

        loop_call()
        reps = 3
        loop_ms, _ = timed(loop_call, reps)
        closed_loop = {"mpc_steps_per_s": B * loop_steps * reps / (loop_ms * 1e-3), "value": B * loop_steps * iterations * reps / (loop_ms * 1e-3),
                       "unit": UNIT, "ms_per_mpc_step_of_the_batch": loop_ms / reps / loop_steps, "closed_loop_steps": loop_steps,
                       "h2d_bytes_per_call": h2d, "d2h_bytes_per_call": int(h_ul.numel() + h_xl.numel()) * 8,
                       "note": "b200sqp_closed_loop: moving-horizon warm start + solve + RK4 plant step per MPC step, all on the device"}

    fp64_peak = solver.measure_fp64_peak(local_rank) if rank == 0 else None  # live microbenchmark (b200sqp_measure_fp64_peak), after the timed regions
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
        else:
            peak, peak_src = 6650.0, "B200_PROFILING.md fallback"
        alg_bytes = alg_bytes_per_iteration * B * iterations  # per launch of the LM kernel
        traffic, traffic_src = measured_traffic(workload_name(args.config, ocp, B, iterations).split("_per_gpu")[0])
        achieved = alg_bytes / (kernel_ms / args.steps * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": precision,
            "data": "synthetic",
            "config": {**b200_config_keys(args, ocp, n, world), "parallelism": (f"instance-sharded x{world}, stop-test gather of chi2 fused into the LM kernel over NVLink peer memory"
                                       if use_p2p else f"instance-sharded x{world}, one NCCL chi2 all-gather per step") if world > 1 else "single GPU",
                       "timing": "CUDA events per step on the launch stream, L2 flushed (256 MiB memset) between steps"
                                 + (", ranks rendezvous on the device before each timed step" if world > 1 else "") + ", max over ranks",
                       "wall_s_timed_region_incl_flush": wall},
            "clocks": clocks,
            "per_rank": per_rank,
            # the call a controller makes per MPC step (PredictiveController::step needs the first control, predictive_controller.cpp:46-79):
            # pinned host buffers, H2D of the measured states + references, cold start, solve, D2H of first controls + chi2 + status
            "e2e": {"value": mpc_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": B * ocp.nu * 8 + B * 8 + B * 4,
                    "ms_per_step": mpc_ms / args.steps,
                    "call": "b200sqp_mpc_step: measured states in, first controls + chi2 + status out; the optimised trajectories stay in HBM "
                            "(b200sqp_get_params fetches them on demand)"},
            # the same step with EVERY optimised trajectory copied back as well (b200sqp_step): 4.9 MB per 4096 instances
            "e2e_full_trajectories": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                      "ms_per_step": e2e_ms / args.steps, "call": "b200sqp_step"},
            "closed_loop": closed_loop,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": "lmSolve", "kernel_ms": kernel_ms / args.steps, "algorithmic_bytes_per_launch": alg_bytes,
                         "peak_source": peak_src},
            "roofline_fp64": fp64_roofline(ocp, lm.dims, B, iterations, kernel_ms / args.steps, fp64_peak),
            "lm": {"inner_passes_per_instance": float(stats["inner_passes"].mean()), "rejects_per_instance": float(stats["rejects"].mean()),
                   "relinearizations_per_instance": float(stats["relinearizations"].mean())},
        }
        if n_gpus == 1 and not args.no_cpu_baseline:
            opts = abi.LmOptions.defaults(iterations=iterations, weights=kw["weights"])
            line["cpu_baseline"] = cpu_baseline_entry(run_cpu(ocp, opts, min(args.cpu_sample, CPU_SAMPLE_CAP[args.config]), seed=1234 + args.config))
        if n_gpus == 1 and args.config == 1 and not args.no_other_configs:
            lm.clear()
            del flush
            torch.cuda.empty_cache()
            line["configs"] = other_configs(local_rank, stream, peak, fp64_peak, not args.no_cpu_baseline)
            line["configs"]["2_adaptive"] = adaptive_config(local_rank, not args.no_cpu_baseline)
            line["e2e_plugin"] = plugin_e2e(B, iterations)
        emit(line)
    if use_p2p:
        if exch.timed_out():
            raise SystemExit("bench.py: a peer never arrived in b200sqp_peer_wait (2 s bound)")
        exch.close()
    lm.clear()
    if world > 1:
        dist.destroy_process_group()
    return 0


_JSON_FD = None


def emit(line):
    """the ONE JSON line, on the process's original stdout"""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


if __name__ == "__main__":
    a = parse_args()
    # libraries print to file descriptor 1 behind Python's back (NCCL's "NCCL version ..." banner at communicator creation): keep the
    # original stdout for the JSON line only and send everything else that lands on fd 1 to stderr
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    sys.exit(main_reference(a) if a.impl == "reference" else main_b200(a))

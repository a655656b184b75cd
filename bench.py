#!/usr/bin/env python3
"""Benchmark of the hot path: log-likelihood evaluations per second (waveform + response + inner product).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 1|2|4|5]

Workload (default): BASELINE.json configs[1] -- IMRPhenomPv2 precessing BBH, 3 detectors (H1/L1/V1), a parallel-tempered
ensemble of 8 temperatures x 512 walkers = 4096 walkers per GPU on 16384 frequency bins; synthetic, seeded inputs
(gw_analysis_tools_b200/workloads.py).  One "step" = one pass of the hot path over the whole ensemble of one GPU.
Scaling is weak: every rank owns its own 4096-walker ensemble (walkers are independent; no data-path collective).

JSON line keys (see the task contract):
  value      whole-job evals/s with sampling vectors and outputs resident in HBM (device-pointer C-ABI entry point)
  e2e        the same through the host-buffer C-ABI call a GWAT user makes: pinned host params -> H2D -> kernels -> D2H logL
  roofline   dominant kernel (k_loglike): algorithmic flop-equivalents (670+90*D per active bin for
             IMRPhenomPv2, 270+90*D for IMRPhenomD ...) / device time of that kernel, against the FP64 FMA peak MEASURED
             on this GPU by a DFMA microbenchmark (MEASURED_PEAKS.json carries HBM and bf16 only); the HBM view is given
             beside it as `hbm`.
  cpu_baseline  the reference's own CPU code (oracle/_ref, the reference sources compiled unmodified) on all host threads
             over a bounded sample of the same walkers.
`--impl reference` times that CPU implementation alone, with the same config/metric/unit.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from gw_analysis_tools_b200 import workloads  # noqa: E402

METRIC = "log-likelihood evals/sec (waveform+response+inner product)"
UNIT = "evals/s"

# FP64-pipe flop-equivalents per active (walker, bin), counted on the reference's schedule
FLOP_EQ = {"IMRPhenomD": (270, 90), "IMRPhenomPv2": (670, 90), "IMRPhenomD_NRT": (530, 90), "dCS_IMRPhenomD": (320, 90)}


def ncu_traffic(config):
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_loglike launch of this workload, from the committed
    `ncu --set full` capture (profiles/traffic.json names the capture); None when no capture exists for the config."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t["cfg%d" % config]["k_loglike_dram_bytes"]
    except (OSError, ValueError, KeyError):
        return None


def ncu_pipe_pct(config):
    """FP64-pipe utilisation (sm__inst_executed_pipe_fp64, % of peak sustained active) of k_loglike from the same capture."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t["cfg%d" % config]["fp64_pipe_active_pct"]
    except (OSError, ValueError, KeyError):
        return None


def flop_eq_per_bin(method, D):
    a, b = FLOP_EQ[method]
    return a + b * D


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.

    In-process NVML (nvidia_ml_py) on a thread, one sample every 5 ms: the timed region of the short configurations is
    tens of milliseconds, less than the start-up time of an `nvidia-smi -lms` child.  Falls back to that child when NVML
    cannot be loaded."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []   # (sm_mhz, reasons bitmask)
        self.sm_max = None
        self._stop = threading.Event()

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.sm_max = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self._stop.is_set():
            try:
                self.samples.append((float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)),
                                     int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))))
            except Exception:
                pass
            self._stop.wait(0.005)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            n = self.nvml
            masks = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            reasons = sorted(nm for nm, m in masks.items() if any(r & m for _, r in self.samples))
            sm = [c for c, _ in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.sm_max, "samples": len(sm),
                    "reasons": reasons, "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                smax.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


def make_injection(ctx, wl):
    """Zero-noise injection computed by the product itself: data_d = response_d(theta_inj)."""
    ctx.set_network(wl.detectors, wl.f, wl.psd)
    src = ctx.repack_mcmc_batch(wl.method, wl.inj[None, :], wl.gmst, wl.mod)
    src[0].tc = wl.T_segment - src[0].tc
    wl.data = ctx.coherent_response_batch(wl.method, src)[0]
    ctx.set_network(wl.detectors, wl.f, wl.psd, wl.data)


def host_threads():
    """All host threads this process may use.  torchrun exports OMP_NUM_THREADS=1, so the OpenMP default is not it."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference_rate(wl, sample, nthreads=0, repeats=1):
    """evals/s of the reference's CPU path (oracle/_ref) on `sample` walkers of the workload, all host threads."""
    from oracle import gwat_ref
    nthreads = nthreads or host_threads()
    p = wl.params[:sample]
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        gwat_ref.loglike_mcmc_batch(wl.method, wl.mod, p, wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, wl.data,
                                    nthreads=nthreads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return sample / best, nthreads, best


def workload_config(wl, n_gpus):
    return {"workload": "%s: %s, %d detectors (%s), %d walkers/GPU x %d bins, MCMC sampling dim %d" %
                        (wl.name, wl.method, wl.D, "/".join(wl.detectors), wl.W, wl.L, wl.P),
            "method": wl.method, "walkers_per_gpu": wl.W, "bins": wl.L, "detectors": wl.D, "dimension": wl.P,
            "parallelism": "walkers sharded over %d GPU(s), no data-path collective" % n_gpus,
            "l2": "L2 flushed (256 MiB device memset) between timed steps; a fresh walker set every step"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import gwat_ref
    if not gwat_ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libgwat_ref.so not built (needs /root/reference at build time)"}))
        return 0
    wl = workloads.make(args.config, W=args.walkers, L=args.bins)
    # the data the GPU arm uses comes from the product; here the oracle makes the same injection itself
    _, src = gwat_ref.loglike_mcmc_batch(wl.method, wl.mod, wl.inj[None, :], wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd,
                                         None, return_sources=True)
    wl.data = gwat_ref.coherent_response(wl.method, src[0], wl.detectors, wl.f)
    nthreads = host_threads()
    sample = min(wl.W, args.cpu_sample)
    for _ in range(args.warmup):
        cpu_reference_rate(wl, min(sample, 4 * nthreads), nthreads)
    t0 = time.perf_counter()
    for k in range(args.steps):
        lo = (k * sample) % max(1, wl.W - sample + 1)
        sub = wl.params[lo:lo + sample]
        gwat_ref.loglike_mcmc_batch(wl.method, wl.mod, sub, wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd, wl.data,
                                    nthreads=nthreads)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    desc = "%d of the %d walkers per step, %d OpenMP threads over walkers (one chain per thread, as the reference's pool)" % (
        sample, wl.W, nthreads)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(wl, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "reference", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def run_b200(args):
    import torch
    import torch.distributed as dist

    from gw_analysis_tools_b200 import engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world

    # every rank owns its own ensemble (weak scaling): same grid/injection, different walker draws
    wl = workloads.make(args.config, W=args.walkers, L=args.bins, seed=workloads.SEED0 + args.config + 1000 * rank)
    ctx = engine.Context(local)
    make_injection(ctx, wl)
    W, P, D, L = wl.W, wl.P, wl.D, wl.L

    # a few distinct walker sets so that consecutive steps do not repeat the same parameter points
    nsets = 4
    rng = np.random.default_rng(7 + rank)
    host_sets = []
    for s in range(nsets):
        p = wl.params.copy()
        p[:, [0, 2, 4]] += 1e-3 * rng.standard_normal((W, 3))
        host_sets.append(torch.from_numpy(p).pin_memory())
    dev_sets = [h.cuda(non_blocking=True) for h in host_sets]
    d_out = torch.empty(W, dtype=torch.float64, device="cuda")
    h_out = torch.empty(W, dtype=torch.float64).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    # a non-default stream: the C ABI treats a NULL stream as "the context's own", and torch's default stream handle is 0
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()

    def step_resident(k):
        ctx.loglike_mcmc_batch_dev(wl.method, dev_sets[k % nsets].data_ptr(), W, P, wl.gmst, wl.T_segment, d_out.data_ptr(),
                                   wl.mod, stream.cuda_stream)

    def step_e2e(k):
        h = host_sets[k % nsets]
        rc = ctx._lib.gwat_b200_loglike_mcmc_batch(ctx._h, wl.method.encode(), _mod_ref(wl.mod), P, W,
                                                   _ptr(h.data_ptr()), _dbl(wl.gmst), _dbl(wl.T_segment), _ptr(h_out.data_ptr()))
        ctx._check(rc)

    for k in range(args.warmup):
        step_resident(k)
        step_e2e(k)
    torch.cuda.synchronize()
    fp64_peak = ctx.measure_fp64_peak()
    launches0 = ctx.launch_count

    # ---- resident-input timing: CUDA events on the launching stream, L2 flushed between steps ----------------------------
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms, active = [], []
    with torch.cuda.stream(stream):
        for k in range(args.steps):
            flush.zero_()
            ev[k][0].record(stream)
            step_resident(k)
            ev[k][1].record(stream)
    torch.cuda.synchronize()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    t_resident = sum(step_ms) * 1e-3
    launches_resident = ctx.launch_count - launches0

    # ---- end-to-end timing through the host-buffer entry point (H2D + kernels + D2H inside the timed region) ----------
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(args.steps):
        step_e2e(k)
        kernel_ms.append(ctx.last_kernel_ms)
        active.append(ctx.last_active_bins)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    clocks = sampler.stop()
    launches_total = ctx.launch_count - launches0
    logl_checksum = float(np.nansum(h_out.numpy()))

    if world > 1:
        tt = torch.tensor([t_resident, t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_resident, t_e2e = float(tt[0]), float(tt[1])
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    K = args.steps
    value = n_gpus * W * K / t_resident
    e2e_value = n_gpus * W * K / t_e2e
    k_ms = float(np.mean(kernel_ms))
    act = float(np.mean(active))
    feq = flop_eq_per_bin(wl.method, D)
    achieved_tf = act * feq / (k_ms * 1e-3) / 1e12
    alg_bytes = L * (8 + 24 * D) + W * 8 * (P + 1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    roofline = {"bound": "fp64", "kernel": "k_loglike", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved_tf / fp64_peak if fp64_peak > 0 else None,
                "peak_source": "DFMA-chain microbenchmark run in this process (gwat_b200_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
                "work": "%d flop-eq per active (walker,bin) [SURVEY 8(d)] x %.4g active bins per launch (%.1f%% of W*L)" % (
                    feq, act, 100.0 * act / (W * L)),
                "kernel_ms": k_ms, "traffic": ncu_traffic(args.config), "fp64_pipe_active_pct_ncu": ncu_pipe_pct(args.config),
                "hbm": {"achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback",
                        "algorithmic_bytes": alg_bytes}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": args.warmup,
            "ms_per_step": t_resident / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(wl, n_gpus),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": W * P * 8, "d2h_bytes_per_step": W * 8 + 8,
                    "ms_per_step": t_e2e / K * 1e3},
            "gpu_launches": int(launches_total), "gpu_launches_per_step": launches_resident / K,
            "clocks": clocks, "roofline": roofline, "logL_checksum": logl_checksum}

    # ---- CPU baseline: the reference's own code on the host cores of this box, bounded sample (N=1 only) ----------------
    if n_gpus == 1 and not args.no_cpu_baseline:
        try:
            from oracle import gwat_ref
            if gwat_ref.available():
                sample = min(W, args.cpu_sample)
                rate, cores, secs = cpu_reference_rate(wl, sample)
                rate1, _, _ = cpu_reference_rate(wl, max(8, sample // 16), nthreads=1)
                line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "reference",
                                        "single_thread_value": rate1,
                                        "sample": "%d of the %d walkers of this workload, one pass, %.2f s wall, OpenMP over walkers" % (
                                            sample, W, secs)}
            else:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                        "sample": "oracle/_ref not built on this box"}
        except Exception as exc:  # the baseline must never take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "failed: %r" % (exc,)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


# small ctypes helpers for the raw C-ABI call in step_e2e
import ctypes as _C  # noqa: E402


def _ptr(x):
    return _C.c_void_p(x)


def _dbl(x):
    return _C.c_double(x)


def _mod_ref(mod):
    return _C.byref(mod) if mod is not None else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 4, 5])
    ap.add_argument("--walkers", type=int, default=None, help="walkers per GPU (default: the config's)")
    ap.add_argument("--bins", type=int, default=None, help="frequency bins (default: the config's)")
    ap.add_argument("--cpu-sample", type=int, default=2048, help="walkers in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python3
"""Benchmark of the hot path: log-likelihood evaluations per second (waveform + response + inner product).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 1|2|3|4|5] [--masses heavy|light]
                    [--scaling weak|strong] [--workload likelihood|sampler]

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
  sustained  the same two timed loops repeated for >= 1 s each (the K-step region of the short configs is ~10 ms)
  extra      (N=1 default run) the other lines north_star names, measured in the same process, each with value / e2e /
             roofline / cpu_baseline: the (10,8) Msun mass set, configs 1, 4, 5 and config 3 (10^5 Fisher matrices)
`--impl reference` times that CPU implementation alone, with the same config/metric/unit.
`--config 3` makes the Fisher batch the main line (unit Fisher/s); `--scaling strong` splits the config's ensemble over
the GPUs instead of giving each its own; `--workload sampler` times device-resident PTMCMC steps with the PT-swap
exchange over NCCL (tools/bench_sampler.py).
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


def ncu_value(key, field):
    """A per-launch figure from the committed `ncu --set full` captures (profiles/traffic.json names each capture):
    dram__bytes_read.sum + dram__bytes_write.sum (`*_dram_bytes`) or sm__inst_executed_pipe_fp64 in % of peak sustained
    active (`fp64_pipe_active_pct`).  None when no capture exists for that workload."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[key][field]
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


def physical_core_groups(cpus):
    """The logical CPUs of `cpus` grouped by physical core (hyperthread siblings together), in core order."""
    groups, seen = [], set()
    for c in sorted(cpus):
        if c in seen:
            continue
        sib = {c}
        try:
            txt = open("/sys/devices/system/cpu/cpu%d/topology/thread_siblings_list" % c).read().strip()
            for part in txt.split(","):
                lo, _, hi = part.partition("-")
                sib.update(range(int(lo), int(hi or lo) + 1))
        except (OSError, ValueError):
            pass
        g = sorted(x for x in sib if x in cpus)
        seen.update(g)
        groups.append(g)
    return groups


def pin_rank_cores(local, world):
    """Give every rank of a multi-GPU run its own, disjoint share of the host's PHYSICAL cores, hyperthread siblings together
    (eight ranks on the same 32 logical CPUs made the host-buffer path's max-over-ranks time a scheduler
    lottery -- measured again in round 2: 54-58 M evals/s through host buffers on 8 GPUs unpinned, 60 M pinned; a split by logical
    CPU number alone had put the two ranks of a 2-GPU box on sibling threads of the same cores and cost 8 %).  Returns the CPUs kept."""
    try:
        cpus = set(os.sched_getaffinity(0))
        groups = physical_core_groups(cpus)
        if world <= 1 or len(groups) < world:
            return sorted(cpus)
        share = len(groups) // world
        mine = sorted(c for g in groups[local * share:(local + 1) * share] for c in g)
        os.sched_setaffinity(0, mine)
        return mine
    except (AttributeError, OSError):
        return []


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


def workload_config(wl, n_gpus, scaling="weak", masses="heavy"):
    par = ("walkers sharded over %d GPU(s), no data-path collective" % n_gpus if scaling == "weak" else
           "the config's ensemble split over %d GPU(s) (%d walkers each), no data-path collective" % (n_gpus, wl.W))
    return {"workload": "%s: %s, %d detectors (%s), %d walkers/GPU x %d bins, MCMC sampling dim %d, %s masses" %
                        (wl.name, wl.method, wl.D, "/".join(wl.detectors), wl.W, wl.L, wl.P, masses),
            "method": wl.method, "walkers_per_gpu": wl.W, "bins": wl.L, "detectors": wl.D, "dimension": wl.P, "masses": masses,
            "parallelism": par,
            "l2": "L2 flushed (256 MiB device memset) between timed steps; a fresh walker set every step"}


LIGHT = {1: (10.0, 8.0), 2: (10.0, 8.0), 4: (10.0, 8.0)}   # the low-mass set, every bin below 0.2/M active


def make_workload(config, masses, W=None, L=None, seed=None):
    m = LIGHT.get(config) if masses == "light" else None
    return workloads.make(config, W=W, L=L, masses=m, seed=seed)


def peaks_file():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        return {}


# ---------------------------------------------------------------------------------------------------------------------------
# the likelihood path (configs 1, 2, 4, 5)
# ---------------------------------------------------------------------------------------------------------------------------

class LikelihoodBench:
    """One workload resident on one GPU: pinned host walker sets, device copies, and the two timed call paths."""

    def __init__(self, ctx, wl, rank, stream, flush, nsets=4):
        import torch
        self.torch = torch
        self.ctx, self.wl, self.stream, self.flush = ctx, wl, stream, flush
        make_injection(ctx, wl)
        rng = np.random.default_rng(7 + rank)
        self.host_sets = []
        for _ in range(nsets):   # a few distinct walker sets so that consecutive steps do not repeat the same parameter points
            p = wl.params.copy()
            p[:, [0, 2, 4]] += 1e-3 * rng.standard_normal((wl.W, 3))
            self.host_sets.append(torch.from_numpy(p).pin_memory())
        self.dev_sets = [h.cuda(non_blocking=True) for h in self.host_sets]
        self.d_out = torch.empty(wl.W, dtype=torch.float64, device="cuda")
        self.h_out = torch.empty(wl.W, dtype=torch.float64).pin_memory()
        self.nsets = nsets
        torch.cuda.synchronize()

    def step_resident(self, k):
        wl = self.wl
        self.ctx.loglike_mcmc_batch_dev(wl.method, self.dev_sets[k % self.nsets].data_ptr(), wl.W, wl.P, wl.gmst, wl.T_segment,
                                        self.d_out.data_ptr(), wl.mod, self.stream.cuda_stream)

    def step_e2e(self, k):
        wl, ctx = self.wl, self.ctx
        h = self.host_sets[k % self.nsets]
        rc = ctx._lib.gwat_b200_loglike_mcmc_batch(ctx._h, wl.method.encode(), _mod_ref(wl.mod), wl.P, wl.W, _ptr(h.data_ptr()),
                                                   _dbl(wl.gmst), _dbl(wl.T_segment), _ptr(self.h_out.data_ptr()))
        ctx._check(rc)

    def warm(self, n):
        # the two call paths share the context's lane-0 scratch: never in flight together (include/gwat_b200.h)
        for k in range(n):
            self.step_resident(k)
        self.torch.cuda.synchronize()
        for k in range(n):
            self.step_e2e(k)
        self.torch.cuda.synchronize()

    def time_resident(self, steps):
        """CUDA events on the launching stream around every step, L2 flushed before each; returns seconds (sum over steps)."""
        torch = self.torch
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        with torch.cuda.stream(self.stream):
            for k in range(steps):
                self.flush.zero_()
                ev[k][0].record(self.stream)
                self.step_resident(k)
                ev[k][1].record(self.stream)
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in ev) * 1e-3

    def time_e2e(self, steps, kernel_pass=True):
        """Host-buffer C-ABI call (H2D + kernels + D2H + sync inside), wall clock, as a user makes it.  The per-launch k_loglike time
        comes from a second pass over the same steps with the library's CUDA events around that kernel switched on
        (gwat_b200_set_kernel_timing): the events cost ~6 us per call, so the library records them only on request."""
        active = []
        self.torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(steps):
            self.step_e2e(k)
            active.append(self.ctx.last_active_bins)
        self.torch.cuda.synchronize()
        t = time.perf_counter() - t0
        kernel_ms = []
        if kernel_pass:
            self.ctx.set_kernel_timing(True)
            for k in range(steps):
                self.step_e2e(k)
                kernel_ms.append(self.ctx.last_kernel_ms)
            self.ctx.set_kernel_timing(False)
        return t, float(np.mean(kernel_ms)) if kernel_ms else 0.0, float(np.mean(active))

    def roofline(self, k_ms, act, fp64_peak, config_key):
        wl = self.wl
        feq = flop_eq_per_bin(wl.method, wl.D)
        achieved_tf = act * feq / (k_ms * 1e-3) / 1e12
        alg_bytes = wl.L * (8 + 24 * wl.D) + wl.W * 8 * (wl.P + 1)
        peaks = peaks_file()
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        return {"bound": "fp64", "kernel": "k_loglike", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": achieved_tf / fp64_peak if fp64_peak > 0 else None,
                "peak_source": "DFMA-chain microbenchmark run in this process (gwat_b200_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
                "work": "%d flop-eq per active (walker,bin) on the reference's schedule [SURVEY 8(d)] x %.4g active bins per launch (%.1f%% of W*L)" % (
                    feq, act, 100.0 * act / (wl.W * wl.L)),
                "kernel_ms": k_ms, "traffic": ncu_value(config_key, "k_loglike_dram_bytes"),
                "fp64_pipe_active_pct_ncu": ncu_value(config_key, "fp64_pipe_active_pct"),
                "hbm": {"achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback",
                        "algorithmic_bytes": alg_bytes}}

    def cpu_baseline(self, sample):
        from oracle import gwat_ref
        wl = self.wl
        if not gwat_ref.available():
            return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref not built on this box"}
        sample = min(wl.W, sample)
        rate, cores, secs = cpu_reference_rate(wl, sample)
        rate1, _, _ = cpu_reference_rate(wl, max(4, sample // 16), nthreads=1)
        return {"value": rate, "unit": UNIT, "cores": cores, "kind": "reference", "single_thread_value": rate1,
                "sample": "%d of the %d walkers of this workload, one pass, %.2f s wall, OpenMP over walkers" % (sample, wl.W, secs)}

    def brief(self, steps, warmup, fp64_peak, config_key, cpu_sample):
        """value / e2e / roofline / cpu_baseline of this workload in one dict (used for the extra lines of the default run)."""
        self.warm(warmup)
        t_res = self.time_resident(steps)
        t_e2e, k_ms, act = self.time_e2e(steps)
        wl = self.wl
        out = {"config": workload_config(wl, 1, masses=config_key.split("_")[-1] if "_" in config_key else "heavy"),
               "value": wl.W * steps / t_res, "unit": UNIT, "steps": steps, "ms_per_step": t_res / steps * 1e3,
               "e2e": {"value": wl.W * steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": wl.W * wl.P * 8,
                       "d2h_bytes_per_step": wl.W * 8 + 8, "ms_per_step": t_e2e / steps * 1e3},
               "roofline": self.roofline(k_ms, act, fp64_peak, config_key)}
        if cpu_sample:
            try:
                out["cpu_baseline"] = self.cpu_baseline(cpu_sample)
            except Exception as exc:
                out["cpu_baseline"] = {"value": None, "sample": "failed: %r" % (exc,)}
        return out


# ---------------------------------------------------------------------------------------------------------------------------
# config 3: batched Fisher matrices
# ---------------------------------------------------------------------------------------------------------------------------

FISHER_METRIC = "order-4 numerical Fisher matrices/sec (IMRPhenomD, 11 parameters, 3 detectors summed, 4096 bins)"
FISHER_DETS = ["Hanford", "Livingston", "Virgo"]


def fisher_grid(L=4096):
    f = 20.0 + 0.25 * np.arange(L)
    return f, np.tile(workloads.aligo_analytic_psd(f), (3, 1))


def fisher_flop_eq(srcs, f, D=3, dim=11):
    """SURVEY 8(d), reference schedule: per source D*4*dim*L_active*(270+65) for the stencil evaluations, D*dim*L*10 for
    combining the points, D*dim(dim+1)/2*L*10 for the assembly.  L_active = bins at or below 0.2/M."""
    total = 0.0
    L = f.size
    for s in srcs:
        fcut = 0.2 / ((s.mass1 + s.mass2) * workloads.MSOL_SEC)
        la = int(np.searchsorted(f, fcut, side="right"))
        total += D * 4 * dim * la * 335.0 + D * dim * L * 10.0 + D * (dim * (dim + 1) // 2) * L * 10.0
    return total


def fisher_measure(ctx, S, steps, warmup, fp64_peak, cpu_sample, seed_offset=0):
    """Fisher/s through the host-buffer C ABI (the only entry point a GWAT user has for Fishers: sources in host memory,
    matrices back in host memory), device time from the library's events, roofline on the reference-schedule count."""
    f, psd = fisher_grid()
    ctx.set_network(FISHER_DETS, f, psd)
    sets = [workloads.fisher_sources(S, seed=workloads.SEED0 + 3 + 17 * k + seed_offset) for k in range(2)]
    arrs = [(workloads.abi.Source * S)(*s) for s in sets]
    for k in range(max(1, min(warmup, 3))):
        ctx.fisher_numerical_batch("IMRPhenomD", arrs[k % 2], 11, order=4)
    dev_ms, t0 = [], time.perf_counter()
    for k in range(steps):
        F = ctx.fisher_numerical_batch("IMRPhenomD", arrs[k % 2], 11, order=4)
        dev_ms.append(ctx.last_kernel_ms)
    t_e2e = time.perf_counter() - t0
    t_dev = sum(dev_ms) * 1e-3
    work = 0.5 * (fisher_flop_eq(sets[0], f) + fisher_flop_eq(sets[1], f)) if steps > 1 else fisher_flop_eq(sets[0], f)
    achieved = work * steps / t_dev / 1e12
    nonfinite = int(np.sum(~np.all(np.isfinite(F.reshape(S, -1)), axis=1)))
    out = {"metric": FISHER_METRIC, "value": S * steps / t_dev, "unit": "Fisher/s", "steps": steps, "sources_per_step": S,
           "ms_per_step": t_dev / steps * 1e3,
           "value_note": "device time of the whole pass (chunk loop incl. the H2D of the sources and the D2H of the matrices, CUDA events inside the library)",
           "e2e": {"value": S * steps / t_e2e, "unit": "Fisher/s", "h2d_bytes_per_step": S * C_sizeof_source(),
                   "d2h_bytes_per_step": S * 121 * 8, "ms_per_step": t_e2e / steps * 1e3},
           "equivalent_response_evals_per_s": S * steps / t_dev * 132,
           "roofline": {"bound": "fp64", "kernel": "k_fisher_fused (+ k_fisher_setup: whole pass)", "achieved": achieved,
                        "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak else None,
                        "work": "reference schedule [SURVEY 8(d)]: %.4g flop-eq per source on average (3 det x 44 stencil responses x L_active x 335 + combine + assembly); per source in HBM: the 11 x 4 coefficient blocks (61.6 kB) the fused kernel stages, see traffic" % (work / S),
                        "fp64_pipe_active_pct_ncu": ncu_value("cfg3", "fp64_pipe_active_pct"),
                        "traffic": ncu_value("cfg3", "k_fisher_dram_bytes_per_source"),
                        "peak_source": "DFMA-chain microbenchmark run in this process"},
           "nonfinite_sources": nonfinite}
    if cpu_sample:
        try:
            from oracle import gwat_ref
            if gwat_ref.available():
                n = min(cpu_sample, S)
                nt = host_threads()
                t0 = time.perf_counter()
                R = gwat_ref.fisher_numerical_batch("IMRPhenomD", sets[(steps - 1) % 2][:n], FISHER_DETS, f, psd, 11, order=4,
                                                    detector_index=-1, reference_index=0, nthreads=nt)
                dtc = time.perf_counter() - t0
                ok = np.all(np.isfinite(R.reshape(n, -1)), axis=1)
                dg = np.sqrt(np.abs(np.einsum("sii->si", R[ok])))
                nerr = np.abs(F[:n][ok] - R[ok]) / (dg[:, :, None] * dg[:, None, :])
                out["cpu_baseline"] = {"value": n / dtc, "unit": "Fisher/s", "cores": nt, "kind": "reference",
                                       "sample": "%d of the %d sources, fisher_numerical per detector and summed, %.2f s wall, OpenMP over sources" % (n, S, dtc)}
                out["parity_on_sample"] = {"normalised_error_median": float(np.median(nerr)), "normalised_error_max": float(nerr.max()),
                                           "measure": "|dF_ij|/sqrt(F_ii F_jj) vs the reference on the CPU sample"}
        except Exception as exc:
            out["cpu_baseline"] = {"value": None, "sample": "failed: %r" % (exc,)}
    return out


def C_sizeof_source():
    import ctypes
    return ctypes.sizeof(workloads.abi.Source)


# ---------------------------------------------------------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------------------------------------------------------

def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import gwat_ref
    if not gwat_ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libgwat_ref.so not built (needs /root/reference at build time)"}))
        return 0
    nthreads = host_threads()
    if args.config == 3:
        f, psd = fisher_grid()
        n = min(args.fisher_sources, max(nthreads * 4, 64))
        srcs = workloads.fisher_sources(n)
        for _ in range(min(args.warmup, 1)):
            gwat_ref.fisher_numerical_batch("IMRPhenomD", srcs[:nthreads], FISHER_DETS, f, psd, 11, order=4, nthreads=nthreads)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            gwat_ref.fisher_numerical_batch("IMRPhenomD", srcs, FISHER_DETS, f, psd, 11, order=4, nthreads=nthreads)
        dt = time.perf_counter() - t0
        value = n * args.steps / dt
        desc = "%d of the %d sources per step, %d OpenMP threads over sources" % (n, args.fisher_sources, nthreads)
        line = {"impl": "reference", "metric": FISHER_METRIC, "value": value, "unit": "Fisher/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": fisher_config(args.fisher_sources, args.gpus),
                "cpu_baseline": {"value": value, "unit": "Fisher/s", "cores": nthreads, "kind": "reference", "sample": desc},
                "e2e": {"value": value, "unit": "Fisher/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0
    W_local = args.walkers
    wl = make_workload(args.config, args.masses, W=W_local, L=args.bins)
    if args.scaling == "strong" and args.gpus > 1:
        wl.params = wl.params[:wl.W // args.gpus]
    # the data the GPU arm uses comes from the product; here the oracle makes the same injection itself
    _, src = gwat_ref.loglike_mcmc_batch(wl.method, wl.mod, wl.inj[None, :], wl.gmst, wl.T_segment, wl.detectors, wl.f, wl.psd,
                                         None, return_sources=True)
    wl.data = gwat_ref.coherent_response(wl.method, src[0], wl.detectors, wl.f)
    sample = min(wl.W, args.cpu_sample if args.config != 5 else min(args.cpu_sample, 4 * nthreads))
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
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(wl, args.gpus, args.scaling, args.masses),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "reference", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def fisher_config(S, n_gpus):
    return {"workload": "cfg3_IMRPhenomD_Fisher: %d sources/GPU (m in U(3,100) Msun, testing/fisher_comparison.cpp draw), 11 parameters, order-4 stencil, "
                        "3 detectors (Hanford/Livingston/Virgo) summed, 4096 bins from 20 Hz at df = 1/4 Hz" % S,
            "method": "IMRPhenomD", "sources_per_gpu": S, "bins": 4096, "detectors": 3, "dimension": 11, "order": 4,
            "parallelism": "sources sharded over %d GPU(s), no collective" % n_gpus,
            "l2": "inputs larger than L2: the derivative buffer of one pass is 2 GiB; two alternating source sets"}


# ---------------------------------------------------------------------------------------------------------------------------
# product arm
# ---------------------------------------------------------------------------------------------------------------------------

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
    # several ranks: each gets its own physical cores unless --no-pin-cores (the reference arm never pins: it runs on all host threads)
    pin = args.pin_cores if args.pin_cores is not None else world > 1
    cores = pin_rank_cores(local, world) if pin else sorted(os.sched_getaffinity(0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world
    ctx = engine.Context(local)
    stream = torch.cuda.Stream()   # non-default: the C ABI treats a NULL stream as "the context's own"
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def reduce_max(vals):
        if world == 1:
            return vals, [vals]
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        allv = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        per_rank = [[float(x) for x in a] for a in allv]
        return [max(col) for col in zip(*per_rank)], per_rank

    if args.workload == "sampler":
        from tools import bench_sampler
        return bench_sampler.run_under_bench(args, ctx, world, rank, local, dist if world > 1 else None, ClockSampler)

    # ---- config 3: Fisher batches ---------------------------------------------------------------------------------------
    if args.config == 3:
        S = args.fisher_sources if args.scaling == "weak" else args.fisher_sources // world
        fp64_peak = ctx.measure_fp64_peak()
        launches0 = ctx.launch_count
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local)
        sampler.start()
        res = fisher_measure(ctx, S, args.steps, args.warmup, fp64_peak, 0 if (world > 1 or args.no_cpu_baseline) else args.fisher_cpu_sample,
                             seed_offset=1000 * rank)
        clocks = sampler.stop()
        (t_dev, t_e2e), per_rank = reduce_max([res["ms_per_step"], res["e2e"]["ms_per_step"]])
        if rank != 0:
            if world > 1:
                dist.destroy_process_group()
            return 0
        line = {"metric": FISHER_METRIC, "value": n_gpus * S / (t_dev * 1e-3), "unit": "Fisher/s", "n_gpus": n_gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t_dev, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": fisher_config(S, n_gpus),
                "e2e": dict(res["e2e"], value=n_gpus * S / (t_e2e * 1e-3), ms_per_step=t_e2e),
                "gpu_launches": int(ctx.launch_count - launches0), "clocks": clocks, "roofline": res["roofline"],
                "equivalent_response_evals_per_s": n_gpus * S / (t_dev * 1e-3) * 132, "nonfinite_sources": res["nonfinite_sources"]}
        for k in ("cpu_baseline", "parity_on_sample"):
            if k in res:
                line[k] = res[k]
        print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- likelihood configs ---------------------------------------------------------------------------------------------
    # weak scaling: every rank owns its own ensemble of the config's size (same grid/injection, different walker draws);
    # strong scaling: the config's ONE ensemble is split over the ranks (whole temperature rungs per rank)
    wl = make_workload(args.config, args.masses, W=args.walkers, L=args.bins,
                       seed=workloads.SEED0 + args.config + (1000 * rank if args.scaling == "weak" else 0))
    W_total = wl.W * (n_gpus if args.scaling == "weak" else 1)
    if args.scaling == "strong" and world > 1:
        per = wl.W // world
        wl.params = wl.params[rank * per:(rank + 1) * per].copy()
    lb = LikelihoodBench(ctx, wl, rank, stream, flush)
    W, P, D, L = wl.W, wl.P, wl.D, wl.L
    lb.warm(args.warmup)
    fp64_peak = ctx.measure_fp64_peak()
    launches0 = ctx.launch_count

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    t_resident = lb.time_resident(args.steps)
    launches_resident = ctx.launch_count - launches0
    if world > 1:
        dist.barrier()
    t_e2e, k_ms, act = lb.time_e2e(args.steps)
    clocks = sampler.stop()
    launches_total = ctx.launch_count - launches0
    logl_checksum = float(np.nansum(lb.h_out.numpy()))

    # ---- sustained figure: the same two loops for >= 1 s each (weak #9: the K-step region is ~10 ms) ---------------------
    sustained = None
    if not args.no_extras:
        n_sus = int(min(20000, max(args.steps, np.ceil(args.sustain_seconds / max(t_resident / args.steps, 1e-6)))))
        if world > 1:
            dist.barrier()
        s2 = ClockSampler(local)
        s2.start()
        ts_res = lb.time_resident(n_sus)
        ts_e2e, _, _ = lb.time_e2e(n_sus, kernel_pass=False)
        c2 = s2.stop()
        sustained = [ts_res, ts_e2e, n_sus, c2]

    (t_resident, t_e2e, ts_res, ts_e2e), per_rank = reduce_max([t_resident, t_e2e, sustained[0] if sustained else 0.0,
                                                                 sustained[1] if sustained else 0.0])
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    K = args.steps
    key = "cfg%d" % args.config + ("_light" if args.masses == "light" else "")
    line = {"metric": METRIC, "value": W_total * K / t_resident, "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": args.warmup,
            "ms_per_step": t_resident / K * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(wl, n_gpus, args.scaling, args.masses),
            "e2e": {"value": W_total * K / t_e2e, "unit": UNIT, "h2d_bytes_per_step": W * P * 8, "d2h_bytes_per_step": W * 8 + 8,
                    "ms_per_step": t_e2e / K * 1e3, "per_rank_ms_per_step": [r[1] / K * 1e3 for r in per_rank],
                    "host_cores_per_rank": len(cores)},
            "gpu_launches": int(launches_total), "gpu_launches_per_step": launches_resident / K,
            "clocks": clocks, "roofline": lb.roofline(k_ms, act, fp64_peak, key), "logL_checksum": logl_checksum}
    if sustained:
        n_sus, c2 = sustained[2], sustained[3]
        line["sustained"] = {"steps": n_sus, "seconds_resident": ts_res, "seconds_e2e": ts_e2e, "value": W_total * n_sus / ts_res,
                             "e2e_value": W_total * n_sus / ts_e2e, "unit": UNIT, "clocks": c2,
                             "note": "the same two timed loops run for >= %.1f s each (L2 flushed between steps)" % args.sustain_seconds}

    # ---- CPU baseline: the reference's own code on the host cores of this box, bounded sample (N=1 only) ----------------
    if n_gpus == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = lb.cpu_baseline(args.cpu_sample)
        except Exception as exc:  # the baseline must never take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "failed: %r" % (exc,)}

    # ---- the other lines north_star names, measured in the same process (N=1 default run only) --------------------------
    if n_gpus == 1 and not args.no_extras and args.walkers is None and args.bins is None:
        extras = {}
        del lb
        torch.cuda.empty_cache()

        def extra_like(cfg, masses, steps, cpu_sample):
            w2 = make_workload(cfg, masses)
            b2 = LikelihoodBench(ctx, w2, 0, stream, flush, nsets=2)
            return b2.brief(steps, 3, fp64_peak, "cfg%d" % cfg + ("_light" if masses == "light" else ""),
                            0 if args.no_cpu_baseline else cpu_sample)
        try:
            other = "light" if args.masses == "heavy" else "heavy"
            if args.config in LIGHT:
                extras["%s_masses" % other] = extra_like(args.config, other, 10, 256)
            for cfg, steps, cs in ((1, 20, 512), (2, 10, 512), (4, 10, 512), (5, 3, 32)):
                if cfg != args.config:
                    extras["cfg%d" % cfg] = extra_like(cfg, "heavy", steps, cs)
            if args.config != 3:
                extras["cfg3"] = fisher_measure(ctx, args.fisher_sources, 2, 1, fp64_peak, 0 if args.no_cpu_baseline else args.fisher_cpu_sample)
                extras["cfg3"]["config"] = fisher_config(args.fisher_sources, 1)
        except Exception as exc:
            extras["error"] = repr(exc)
        line["extra"] = extras
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
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5])
    ap.add_argument("--masses", default="heavy", choices=["heavy", "light"],
                    help="heavy: (36,29) Msun, ~30%% of the bins below 0.2/M; light: (10,8) Msun, every bin active (configs 1, 2, 4)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the config's ensemble per GPU; strong: the config's ensemble split over the GPUs")
    ap.add_argument("--workload", default="likelihood", choices=["likelihood", "sampler"],
                    help="sampler: device-resident PTMCMC steps, chains sharded over the GPUs, PT swap over NCCL")
    ap.add_argument("--walkers", type=int, default=None, help="walkers per GPU (default: the config's)")
    ap.add_argument("--bins", type=int, default=None, help="frequency bins (default: the config's)")
    ap.add_argument("--cpu-sample", type=int, default=2048, help="walkers in the CPU-baseline sample")
    ap.add_argument("--fisher-sources", type=int, default=100000, help="config 3: sources per GPU and step")
    ap.add_argument("--fisher-cpu-sample", type=int, default=256)
    ap.add_argument("--sustain-seconds", type=float, default=1.0)
    ap.add_argument("--pin-cores", dest="pin_cores", action="store_true", default=None,
                    help="give every rank a disjoint share of the host's physical cores (default when there are several ranks)")
    ap.add_argument("--no-pin-cores", dest="pin_cores", action="store_false")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the sustained figure and the other configs' lines")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.config == 3 and args.steps > 20 and "--steps" not in sys.argv:
        args.steps = 5
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""bench.py -- benchmarks of the probayes hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c2|c3|c4|c5]

Workloads = the configurations of BASELINE.json:
  c2 (default, the headline) 4096 independent Metropolis-Hastings chains per GPU x 10^4
      steps on the 2-D correlated normal of examples/mcmc/mcmc_prob4a.py, fp64, native
      Philox RNG, every step recorded.  One bench "step" = --walks-per-step whole walks
      (so that a step is >= 50 ms of GPU work and the timed region >= 1 s).  Weak scaling:
      every rank runs its own 4096 chains (global Philox chain ids rank*4096 + c).
      metric: mh_chain_steps_per_sec.
  c3  Bayesian linear regression, N = 10^6 observations, 3 parameters, 16384 MH chains
      SHARDED over the ranks (strong scaling); one bench step = --mh-steps MH steps of
      all chains (one log-likelihood evaluation of N terms per chain per MH step).
      metric: loglik_evals_per_sec.
  c4  discrete grid exact inference, 4096 x 4096 (mu, sigma) grid, N = 10^5 observations,
      mu-row slabs SHARDED over the ranks; one bench step = log-joint of every cell +
      normaliser (all-reduce max / sum) + posterior + both marginals (all-reduce of the
      sigma marginal, all-gather of the mu slabs).  metric: loglik_evals_per_sec (one
      evaluation = one grid cell = N terms).
  c5  Gibbs sampling of a d = 64 multivariate normal through its conditional covariances,
      65536 chains SHARDED over the ranks; one bench step = --sweeps sweeps of 64 coordinate
      updates, the target density of every kept state on the FP64 tensor cores.
      metric: mh_chain_steps_per_sec (one chain-step = one coordinate update, the
      reference's step with tsteps=1).

Prints ONE JSON line (rank 0):
  value     whole-job throughput, inputs resident in HBM, CUDA events, max over ranks
  e2e       the same metric through the public API (pb.SP / pb.SD / dist.dgei_sharded) with
            HOST inputs and results: H2D and D2H copies inside the timed region
  roofline  the bound that applies to the dominant kernel (FP64 pipe for all four; the
            HBM-streaming likelihood regime is reported under roofline_stream)
  cpu_baseline   the oracle's C restatement on the box's host cores (kind "port") and, for
            c2, the REAL reference's NumPy sampler from oracle/_ref timed in the same run
  secondary_n    (default workload only) short c3 / c4 / c5 passes at the same N GPUs
  secondary      (N = 1 only) kernel-level timings of every kernel family
`--impl reference` times the CPU implementation of the same workload on the host cores.
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

COV = np.array([[2.0, 1.2], [1.2, 2.0]])
MEAN = np.array([0.0, 0.0])
INIT = np.array([0.0, 1.0])


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"])
    # c2
    ap.add_argument("--chains", type=int, default=0,
                    help="chains (c2: per GPU, default 4096; c3: total 16384; c5: total 65536)")
    ap.add_argument("--walk-steps", type=int, default=10000, help="c2: MH steps per walk")
    ap.add_argument("--walks-per-step", type=int, default=100, help="c2: walks per bench step")
    ap.add_argument("--thin", type=int, default=1)
    ap.add_argument("--accept", default="log", choices=["reference", "log"])
    ap.add_argument("--variant", type=int, default=0,
                    help="c2 K1 kernel: 0 auto, 1 per-thread, 2 warp-specialised")
    # c3 / c4 / c5
    ap.add_argument("--n-obs", type=int, default=0, help="c3: 10^6, c4: 10^5 by default")
    ap.add_argument("--mh-steps", type=int, default=20, help="c3: MH steps per bench step")
    ap.add_argument("--grid", type=int, default=4096, help="c4: grid points per axis")
    ap.add_argument("--sweeps", type=int, default=100, help="c5: sweeps per bench step")
    ap.add_argument("--cpu-sample-steps", type=int, default=0,
                    help="c2: MH steps of the bounded CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the secondary / secondary_n blocks")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--quick", action="store_true", help="smaller secondary workloads")
    return ap.parse_args()


class ClockSampler:
    """SM clock / throttle-reason samples DURING the timed region.  The timed region of
    the headline is a few tens of milliseconds, shorter than one `nvidia-smi` query, so
    the samples come from NVML in-process (pynvml, ~1 kHz from a thread that runs while
    the main thread waits in cudaStreamSynchronize); `nvidia-smi -lms` is the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.nvml = None
        self.samples = []
        self._stop = False

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(
                ("GPU-" + uuid if not uuid.startswith("GPU-") else uuid).encode())
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.max_mhz = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle,
                                                                     self.nvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self._stop:
            try:
                mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                try:
                    rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                try:
                    pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                except Exception:
                    pw = 0.0
                self.samples.append((float(mhz), int(rs), pw))
            except Exception:
                break
            time.sleep(0.0003)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop = True
            self.thread.join(timeout=2)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"]}
            reasons = set()
            for _, rs, _ in self.samples:
                for bit, name in self.REASONS.items():
                    if rs & bit:
                        reasons.add(name)
            return {"sm_mhz": float(np.median([m for m, _, _ in self.samples])),
                    "sm_max_mhz": self.max_mhz,
                    "power_w_max": float(max(p for _, _, p in self.samples)),
                    "samples": len(self.samples), "source": "nvml", "reasons": sorted(reasons)}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(power)), "samples": len(sm), "source": "nvidia-smi",
                "reasons": sorted(reasons)}


def bind_to_gpu_cpus(index):
    """Pins this rank to the CPUs NVML reports as local to its GPU (same NUMA node /
    PCIe root), before any pinned host buffer is allocated: with 8 ranks streaming
    samples device->host at once, remote-node pinned memory halves the aggregate rate.
    Returns a short description for the bench line."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(index).uuid)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(
                ("GPU-" + uuid if not uuid.startswith("GPU-") else uuid).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        ideal = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        bind_to_gpu_cpus.original = allowed
        use = ideal & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            return "bound to %d of %d allowed CPUs (GPU-local)" % (len(use), len(allowed))
        return "no narrower GPU-local CPU set (%d ideal, %d allowed)" % (len(ideal), len(allowed))
    except Exception as e:                                   # diagnostics only
        return "unbound (%s)" % type(e).__name__



def secondary(eng, peaks, fp64_peak, quick=False):
    """log-likelihood evals/s (C3 shape), HBM-streaming likelihood roofline, DGEI
    (C4 shape) and Gibbs (C5 shape) kernel timings; kernel-only, CUDA events."""
    import torch
    out = {}
    rng = np.random.default_rng(2024)

    def timeit(fn, reps=5, warm=3):
        for _ in range(warm):
            fn()
        ms = []
        for _ in range(reps):
            fn()
            ms.append(eng.last_kernel_ms())
        return float(np.median(ms))

    lims = np.array([[-6., 6.], [-6., 6.], [0.001, 10.]])
    ex = np.array([[0, 0], [0, 0], [1, 0]]); lg = np.zeros(3, int)
    # ---- C3: 16384 chains x N = 1e6 observations, tiles kernel (FP64 bound) --------
    N, C = 1_000_000, 16384
    x = eng.to_device(rng.normal(0, 1, N))
    y = eng.to_device(-1 + 1.5 * x.cpu().numpy() + rng.normal(0, .5, N))
    theta = eng.to_device(np.stack([rng.normal(-1, .001, C), rng.normal(1.5, .001, C),
                                    rng.uniform(.49, .51, C)]))
    ms = timeit(lambda: eng.normreg_logjoint(theta, y, x, lims, ex, lg, variant=1))
    flops = 5.0 * N * C
    out["c3_loglik"] = {"workload": "C3: normal log-likelihood, N=1e6 obs x 16384 chains "
                                    "(obs tiles via TMA, chains in registers)",
                        "ms": ms, "loglik_evals_per_s": C / (ms * 1e-3),
                        "terms_per_s": N * C / (ms * 1e-3),
                        "fp64_tflops": flops / (ms * 1e-3) / 1e12,
                        "fp64_frac_of_measured_peak": flops / (ms * 1e-3) / 1e12 / fp64_peak,
                        "note": "3 FP64 instr (fma, add, fma = 5 flop) per term: pipe-time "
                                "ceiling is 5/6 of the FMA-only flop peak"}
    sd = 0.5 / np.sqrt(N)
    st = eng.to_device(np.tile(np.array([[-1.], [1.5], [.5]]), (1, C)))
    T = 3 if quick else 10
    burn = eng.mh_normreg(st, y, x, 2, lims, ex, lg, [2.4 * sd] * 3, seed=1, record=False)
    ms_mh = timeit(lambda: eng.mh_normreg(st, y, x, T, lims, ex, lg, [2.4 * sd] * 3, seed=1,
                                          step0=2, state_lp=burn["state_lp"], record=False),
                   reps=3, warm=1)
    out["c3_mh"] = {"workload": "C3: MH, 16384 chains, N=1e6, %d steps per call" % T,
                    "ms_per_mh_step": ms_mh / T, "chain_steps_per_s": C * T / (ms_mh * 1e-3),
                    "loglik_evals_per_s": C * T / (ms_mh * 1e-3)}
    # opt-in algorithmic variant (NOT the measured path of the roofline figures): the same
    # walk from centred sufficient statistics, O(1) per likelihood evaluation
    st2 = eng.to_device(np.tile(np.array([[-1.], [1.5], [.5]]), (1, C)))
    Tss = 1000
    ms_ss = timeit(lambda: eng.mh_normreg(st2, y, x, Tss, lims, ex, lg, [2.4 * sd] * 3, seed=1,
                                          variant=3, record=False), reps=3, warm=1)
    out["c3_mh_suffstat"] = {"workload": "C3 with variant 3 (sufficient statistics, opt-in): "
                                         "16384 chains x %d steps, N=1e6, one launch" % Tss,
                             "ms_per_walk": ms_ss,
                             "chain_steps_per_s": C * Tss / (ms_ss * 1e-3),
                             "note": "same values as the term-by-term kernels to fp64 round-off; "
                                     "3 passes over the observations per call, then O(1) per "
                                     "evaluation"}
    del theta, st, st2
    # ---- small problems (the reference-feasible sizes): whole walk in one launch vs one
    # kernel launch per MH step
    Nsm, Csm, Tsm = 1000, 64, 2000
    xs_, ys_ = x[:Nsm].contiguous(), y[:Nsm].contiguous()
    sd_s = 0.5 / np.sqrt(Nsm)
    for tag, var in (("resident_obs_one_launch", 0), ("launch_per_step", 1)):
        st_s = eng.to_device(np.tile(np.array([[-1.], [1.5], [.5]]), (1, Csm)))
        ms_s = timeit(lambda: eng.mh_normreg(st_s, ys_, xs_, Tsm, lims, ex, lg, [2.4 * sd_s] * 3,
                                             seed=1, variant=var, record=True), reps=3, warm=1)
        out["c3_small_" + tag] = {"workload": "MH, %d chains, N=%d obs, %d steps (%s)"
                                              % (Csm, Nsm, Tsm, tag.replace("_", " ")),
                                  "ms_per_walk": ms_s, "us_per_mh_step": 1e3 * ms_s / Tsm,
                                  "chain_steps_per_s": Csm * Tsm / (ms_s * 1e-3)}
    # ---- HBM-streaming regime: <= 8 chains, N = 2^27 (2 GiB of observations > L2) ----
    Ns = (1 << 24) if quick else (1 << 27)
    xs = torch.randn(Ns, dtype=torch.float64, device=eng.device)
    ys = (-1 + 1.5 * xs + 0.5 * torch.randn(Ns, dtype=torch.float64, device=eng.device))
    for Cs in (1, 4, 8):
        th = eng.to_device(np.stack([np.full(Cs, -1.), np.full(Cs, 1.5), np.full(Cs, .5)]))
        ms = timeit(lambda: eng.normreg_logjoint(th, ys, xs, lims, ex, lg, variant=2))
        gbs = 16.0 * Ns / (ms * 1e-3) / 1e9
        out["stream_c%d" % Cs] = {"ms": ms, "hbm_gbs": gbs, "frac": gbs / peaks["hbm_gbs"],
                                  "loglik_evals_per_s": Cs / (ms * 1e-3),
                                  "terms_per_s": Cs * Ns / (ms * 1e-3)}
    # the streaming MH STEP (north_star's roofline target): 4 chains, every step streams all
    # observations once, proposal + priors + accept test + record fused into the same launch
    Tst = 6
    sd_st = 0.5 / np.sqrt(Ns)
    st4 = eng.to_device(np.tile(np.array([[-1.], [1.5], [.5]]), (1, 4)))
    ms_mh = timeit(lambda: eng.mh_normreg(st4, ys, xs, Tst, lims, ex, lg, [2.4 * sd_st] * 3,
                                          seed=2, variant=2, record=True), reps=3, warm=1) / Tst
    gbs_mh = 16.0 * Ns / (ms_mh * 1e-3) / 1e9
    out["stream_mh_c4"] = {"ms_per_mh_step": ms_mh, "hbm_gbs": gbs_mh,
                           "frac": gbs_mh / peaks["hbm_gbs"],
                           "chain_steps_per_s": 4 / (ms_mh * 1e-3),
                           "terms_per_s": 4 * Ns / (ms_mh * 1e-3)}
    out["roofline_stream"] = {"bound": "hbm", "kernel": "nr_stream_kernel<8,true> (one launch "
                                                        "per MH step, accept fused)",
                              "achieved": gbs_mh, "peak": peaks["hbm_gbs"],
                              "unit": "GB/s", "frac": gbs_mh / peaks["hbm_gbs"],
                              "algorithmic_bytes_per_launch": 16.0 * Ns,
                              "workload": "streaming MH likelihood step, 4 chains, N=%d (16 B "
                                          "per observation per step), %d steps per call; the "
                                          "evaluate-only figure is secondary.stream_c4" % (Ns, Tst)}
    del st4
    del xs, ys
    # ---- C4: DGEI 4096 x 4096 grid, N = 1e5 ---------------------------------------------
    Ng, M, S = (10_000 if quick else 100_000), 4096, 4096
    data = eng.to_device(rng.normal(50., 10., Ng))
    mu = eng.to_device(np.linspace(40, 60, M + 2)[1:-1])
    sg = eng.to_device(np.exp(np.linspace(np.log(5), np.log(20), S + 2)[1:-1]))
    lpm = eng.to_device(np.full(M, -np.log(20.)))
    lps = eng.to_device(np.full(S, -np.log(np.log(4.))))
    lj = eng.empty(M, S)
    ms = timeit(lambda: eng.grid_norm_logjoint(data, mu, sg, lpm, lps, out=lj), reps=3, warm=1)
    cellobs = float(Ng) * M * S
    out["c4_logjoint"] = {"workload": "C4: DGEI log-joint, %dx%d grid, N=%d" % (M, S, Ng),
                          "ms": ms, "terms_per_s": cellobs / (ms * 1e-3),
                          "loglik_evals_per_s": M * S / (ms * 1e-3),
                          "fp64_tflops": 4.0 * cellobs / (ms * 1e-3) / 1e12,
                          "fp64_frac_of_measured_peak": 4.0 * cellobs / (ms * 1e-3) / 1e12 / fp64_peak,
                          "note": "2 FP64 instr (mul, fma = 3 flop; SURVEY counts 4) per cell-obs"}
    ms_ss = timeit(lambda: eng.grid_norm_logjoint(data, mu, sg, lpm, lps, out=lj, suffstat=True),
                   reps=3, warm=1)
    out["c4_logjoint_suffstat"] = {"workload": "C4 log-joint from sufficient statistics (opt-in)",
                                   "ms": ms_ss, "algorithmic_gbs": 8.0 * M * S / (ms_ss * 1e-3) / 1e9,
                                   "frac": 8.0 * M * S / (ms_ss * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                   "note": "HBM-bound on writing the 134 MB grid"}
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        r = eng.grid_conditionalise(lj)
    torch.cuda.synchronize()
    t0.record()
    for _ in range(5):
        r = eng.grid_conditionalise(lj)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 5
    alg = 2.0 * 8 * M * S + 8.0 * (M + S)
    out["c4_normalise_marginals"] = {"ms": ms, "algorithmic_gbs": alg / (ms * 1e-3) / 1e9,
                                     "frac": alg / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                     "note": "two passes: online (max, sum-exp), then posterior + "
                                             "both marginals (2 reads + 1 write of the grid; "
                                             "algorithmic = 1 read + 1 write)"}
    del lj, r
    # ---- C5: Gibbs d = 64, 65536 chains, one sweep = 64 coordinate steps ----------------
    from probayes_b200.cond_cov import CondCov
    d, Cg = 64, (8192 if quick else 65536)
    A = rng.standard_normal((d, d))
    cov = A @ A.T / d + np.eye(d)
    mean = rng.standard_normal(d)
    cc = CondCov(mean, cov, np.tile([-10., 10.], (d, 1)))
    stg = eng.to_device(np.tile(mean[:, None], (1, Cg)))
    sweeps = 4
    ms = timeit(lambda: eng.gibbs_mvn(stg, cc, sweeps * d, thin=d, seed=5, want_prob=True),
                reps=3, warm=1)
    out["c5_gibbs"] = {"workload": "C5: Gibbs cond_cov d=64, %d chains, %d sweeps (+ DMMA "
                                   "density per sweep)" % (Cg, sweeps),
                       "ms_per_sweep": ms / sweeps,
                       "coordinate_updates_per_s": Cg * d * sweeps / (ms * 1e-3),
                       "chain_sweeps_per_s": Cg * sweeps / (ms * 1e-3),
                       # 2 d^2 flop of conditional means + 2 d^2 + 2 d of density
                       "fp64_tflops": (4.0 * d * d + 2 * d) * Cg * sweeps / (ms * 1e-3) / 1e12,
                       "fp64_frac_of_measured_peak": (4.0 * d * d + 2 * d) * Cg * sweeps
                       / (ms * 1e-3) / 1e12 / fp64_peak,
                       "note": "conditional means on the FP64 tensor cores (blocked Gauss-Seidel "
                               "in closed form, 18 DMMA per 8 coordinates), table ndtri, one "
                               "Philox block per two updates; DMMA/DFMA share the FP64 pipe"}
    xs64 = eng.to_device(rng.standard_normal((d, Cg)))
    ms = timeit(lambda: eng.mvn_logpdf(xs64, mean, cov))
    out["c5_mvn_logpdf_dmma"] = {"ms": ms, "points_per_s": Cg / (ms * 1e-3),
                                 "fp64_tflops": 2.0 * d * d * Cg / (ms * 1e-3) / 1e12}
    del xs64, stg
    # ---- K6: PD post-processing on a 4096^2 grid's worth of doubles / a sample set -------
    def timeit_ev(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(reps)]
        for a, b in ev:
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        return float(np.median([a.elapsed_time(b) for a, b in ev]))

    n6 = (1 << 22) if quick else (1 << 26)
    keys = torch.randn(n6, dtype=torch.float64, device=eng.device)
    ms = timeit_ev(lambda: eng.argsort(keys, want_keys=True))
    # 8 executed passes x (8 B histogram read + 12 B read + 12 B write) + 8 B and/or read
    sort_bytes = (8 * 32 + 8) * float(n6)
    out["k6_argsort"] = {"workload": "PD.sorted: stable argsort of %d fp64 keys (+ sorted keys)" % n6,
                         "ms": ms, "keys_per_s": n6 / (ms * 1e-3),
                         "algorithmic_gbs": sort_bytes / (ms * 1e-3) / 1e9,
                         "frac": sort_bytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                         "note": "LSD radix, 8 passes of an 8-bit digit: 264 B of traffic per key"}
    lp = torch.log(torch.rand(n6, dtype=torch.float64, device=eng.device)) - 100.0
    ms = timeit_ev(lambda: eng.cumprob(lp, True))
    out["k6_cumprob"] = {"workload": "PD.quantile: normalised cumulative probability of %d "
                                     "log-pscale cells (exp + reduce-then-scan)" % n6,
                         "ms": ms, "algorithmic_gbs": 24.0 * n6 / (ms * 1e-3) / 1e9,
                         "frac": 24.0 * n6 / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
    vals = torch.rand(2, n6, dtype=torch.float64, device=eng.device)
    ms = timeit_ev(lambda: eng.expectation_sums(lp, True, None, vals))
    out["k6_expectation"] = {"workload": "PD.expectation: sum p, sum p*v for 2 value arrays "
                                         "over %d log-pscale samples" % n6,
                             "ms": ms, "algorithmic_gbs": 24.0 * n6 / (ms * 1e-3) / 1e9,
                             "frac": 24.0 * n6 / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
    del lp, vals
    # ---- C1: the reference's own example (mcmc_prob4a: ONE chain, 12288 steps) through the
    # public API, wall clock incl. process(samples) -- what a user of the reference sees
    try:
        import scipy.stats
        import probayes_b200 as pb
        xr = pb.RV('x', vtype=float, vset=(-np.inf, np.inf))
        yr = pb.RV('y', vtype=float, vset=(-np.inf, np.inf))
        proc = pb.SP(xr & yr)
        proc.set_prob(scipy.stats.multivariate_normal, list(MEAN), COV.tolist())
        proc.set_tran(lambda **kw: 1.)
        proc.set_delta(scipy.stats.norm(0., 1.))
        proc.set_scores('hastings')
        proc.set_update('metropolis')
        T1 = 12288
        walls = []
        for rep in range(4):
            t0 = time.perf_counter()
            summ = proc(proc.walk(proc.sampler({'x': INIT[0], 'y': INIT[1]}, stop=T1, seed=rep)))
            n_acc = summ.u.count(True)
            walls.append(time.perf_counter() - t0)
        w = float(np.median(walls[1:]))
        out["c1_api_single_chain"] = {"workload": "C1 (BASELINE.json configs[0]): mcmc_prob4a, 1 chain "
                                                  "x %d steps through SP.sampler -> walk -> "
                                                  "process(samples), wall clock" % T1,
                                      "seconds": w, "chain_steps_per_s": T1 / w,
                                      "accepted": int(n_acc)}
    except Exception as e:                                   # diagnostics only
        out["c1_api_single_chain"] = {"error": repr(e)}
    return out



# ---------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


def k1_counters():
    """ncu counters of ONE K1 launch on the default c2 workload, written by
    scripts/ncu_counters.py from the --set full capture of the same kernel build
    (profiles/k1_counters.json names the capture)."""
    # (a copy sits next to the package: profiles/ is listed in .gpurunignore and may not travel)
    for path in (os.path.join(ROOT, "profiles", "k1_counters.json"),
                 os.path.join(ROOT, "probayes_b200", "k1_counters.json")):
        if os.path.exists(path):
            with open(path) as f:
                return json.load(f)
    return None


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def liboracle_all_cores():
    from oracle.c import liboracle
    liboracle.load()
    return liboracle, liboracle.use_all_cores()


def timed_repeats(fn, budget_s=10.0, max_reps=40):
    """Median rate over repeated calls of fn() -> (units, seconds) for ~budget_s."""
    rates, tot, reps = [], 0.0, 0
    while tot < budget_s and reps < max_reps:
        units, dt = fn(reps)
        rates.append(units / dt)
        tot += dt
        reps += 1
    return float(np.median(rates)), reps, tot


class Dist:
    """rank / world / barrier / max-over-ranks, one process per GPU."""

    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.cpu_binding = bind_to_gpu_cpus(self.local)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, v):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, v):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


# ---------------------------------------------------------------------------
# workloads.  Each has:  setup(eng, d) ; step(k, ev=None) -- one bench step on the device,
# inputs resident; units (whole job, per step); roofline(kernel_ms); e2e_step(k) through the
# public API with host buffers; cpu_sample() on the host cores; finish() -> quality dict
# ---------------------------------------------------------------------------
class C2:
    name, metric, unit, scaling = "c2", "mh_chain_steps_per_sec", "chain-steps/s", "weak"

    def __init__(self, args, world):
        self.a = args
        self.C = args.chains or 4096
        self.T, self.thin, self.W = args.walk_steps, args.thin, args.walks_per_step
        self.world = world
        self.units = world * self.C * self.T * self.W
        self.D = 2
        self.R = self.T // self.thin
        self.out_bytes = self.R * (self.D + 1) * self.C * 8

    def config(self):
        return {"workload": "C2 (BASELINE.json configs[1]): %d MH chains/GPU x %d steps, 2-D "
                            "correlated normal (mcmc_prob4a), thin=%d, Philox4x32-10; one bench "
                            "step = %d walks" % (self.C, self.T, self.thin, self.W),
                "chains_per_gpu": self.C, "walk_steps": self.T, "thin": self.thin,
                "walks_per_step": self.W, "accept": self.a.accept,
                "parallelism": "chains x%d (weak: every rank runs its own chains)" % self.world,
                "l2": "outputs per walk (%.0f MB) exceed the 126 MB L2; no explicit flush"
                      % (self.out_bytes / 1e6)}

    def setup(self, eng, d):
        self.eng, self.d = eng, d
        init_host = np.tile(INIT[:, None], (1, self.C))
        self.state = eng.to_device(init_host)
        self.init_dev = eng.to_device(init_host)
        self.bufs = {"x": eng.empty(self.R, self.D, self.C), "prob": eng.empty(self.R, self.C)}
        self.chain0 = d.rank * self.C
        self.last = None

    def walk(self, seed, ev=None):
        self.state.copy_(self.init_dev)
        self.last = self.eng.mh_mvn(self.state, MEAN, COV, self.T, thin=self.thin, seed=seed,
                                    chain0=self.chain0, accept=self.a.accept,
                                    variant=self.a.variant, out=self.bufs, events=ev)
        return self.last

    def step(self, k, ev=None):
        for w in range(self.W):              # CUDA events bracket the first walk's kernel
            self.walk(1000 + k * self.W + w, ev if w == 0 else None)

    kernel = "mh_mvn_wd_kernel<2>"

    def roofline(self, kernel_ms, peaks, which, fp64_peak, sm_mhz):
        eng = self.eng
        sms = int(eng.info.sm_count)
        clk = (sm_mhz or 1965.0) * 1e6
        cs = self.C * self.T
        out = {"bound": "fp64_pipe", "kernel": self.kernel if self.a.variant != 1 else "mh_mvn_kernel<2>",
               "kernel_ms": kernel_ms, "unit": "G FP64 warp-instr/s",
               "peak": sms * 4 * 0.5 * clk / 1e9,
               "peak_source": "%d SMs x 4 sub-partitions x 1 FP64 warp-instruction per 2 cycles x "
                              "%.0f MHz (the SM clock sampled during the timed region)"
                              % (sms, clk / 1e6),
               "output_gbs": self.out_bytes / (kernel_ms * 1e-3) / 1e9,
               "output_frac_of_hbm": self.out_bytes / (kernel_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
               "hbm_peak_source": which,
               "algorithmic_bytes_per_launch": self.out_bytes,
               "note": "K1 is bound by the FP64 pipe / issue slots of the draw (Philox + "
                       "Box-Muller + log) and record code, not by HBM: it only WRITES (D+1)*8 B "
                       "per recorded chain-step"}
        kc = k1_counters()
        if kc and kc.get("chains") == self.C and kc.get("steps") == self.T and \
                kc.get("thin") == self.thin and self.a.variant != 1:
            fp64_per_cs, inst_per_cs = kc["fp64_warp_inst"] / cs, kc["warp_inst"] / cs
            out["achieved"] = fp64_per_cs * cs / (kernel_ms * 1e-3) / 1e9
            out["frac"] = out["achieved"] / out["peak"]
            out["issue"] = {"warp_inst_per_launch": kc["warp_inst"],
                            "achieved_ginst_s": inst_per_cs * cs / (kernel_ms * 1e-3) / 1e9,
                            "peak_ginst_s": sms * 4 * clk / 1e9,
                            "frac": inst_per_cs * cs / (kernel_ms * 1e-3) / (sms * 4 * clk)}
            out["traffic"] = kc.get("dram_bytes")
            out["counters_source"] = kc.get("source")
        else:
            out["achieved"] = out["frac"] = out["traffic"] = None
            out["counters_source"] = "no ncu counters for this kernel build / workload shape"
        return out

    # ---- public API, host buffers ---------------------------------------------------------
    def e2e_setup(self):
        import scipy.stats
        import probayes_b200 as pb
        xr = pb.RV('x', vtype=float, vset=(-np.inf, np.inf))
        yr = pb.RV('y', vtype=float, vset=(-np.inf, np.inf))
        self.process = pb.SP(xr & yr)
        self.process.set_prob(scipy.stats.multivariate_normal, list(MEAN), COV.tolist())
        self.process.set_tran(lambda **kw: 1.)
        self.process.set_delta(scipy.stats.norm(0., 1.))
        self.process.set_scores('hastings')
        self.process.set_update('metropolis')
        self.hostbuf = {}
        self.e2e_units = self.world * self.C * self.T
        self.h2d = self.D * self.C * 8
        self.d2h = self.out_bytes + (2 * self.D + 2) * self.C * 8 + self.D * self.C * 8
        self.e2e_note = ("pb.SP(...).sampler(init, chains=%d, host_stream=True) -> walk -> "
                         "process(samples): one walk per e2e step, every recorded sample and "
                         "density streamed D2H into pinned host buffers" % self.C)

    def e2e_step(self, k):
        smp = self.process.sampler({'x': INIT[0], 'y': INIT[1]}, stop=self.T, chains=self.C,
                                   thin=self.thin, seed=4000 + k, accept=self.a.accept,
                                   host_stream=True, host_buffers=self.hostbuf)
        summary = self.process(self.process.walk(smp))
        return summary.v['x'][0, -1] + summary.u.count(True)     # touch the result

    def e2e_ceiling(self, d):
        """What bounds the end-to-end figure: the raw pinned device->host copy rate of the
        SAME bytes (every rank at once, as in the e2e loop), measured here.  The walk's
        samples must cross PCIe into ONE host's memory; the kernel is ~30x faster."""
        torch = d.torch
        hx, hp = self.hostbuf.get('x'), self.hostbuf.get('prob')
        if hx is None:
            return None
        reps = 5
        for _ in range(2):
            hx.copy_(self.bufs['x'], non_blocking=True)
            hp.copy_(self.bufs['prob'], non_blocking=True)
        d.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            hx.copy_(self.bufs['x'], non_blocking=True)
            hp.copy_(self.bufs['prob'], non_blocking=True)
        torch.cuda.synchronize()
        dt = d.max(time.perf_counter() - t0)
        gbs = self.world * self.out_bytes * reps / dt / 1e9
        return {"d2h_gbs_all_ranks": gbs, "d2h_gbs_per_rank": gbs / self.world,
                "chain_steps_per_s_at_ceiling": self.world * self.C * self.T * reps / dt,
                "note": "pinned cudaMemcpyAsync D2H of one walk's samples + densities (%.0f MB "
                        "per rank), all ranks concurrently" % (self.out_bytes / 1e6)}

    # ---- CPU ---------------------------------------------------------------------------------
    def cpu_sample(self, budget_s=10.0):
        lo, cores = liboracle_all_cores()
        C = self.C
        init = np.tile(INIT, (C, 1))

        def one(steps, seed):
            t0 = time.perf_counter()
            lo.mh_mvn_walk(init, MEAN, COV, steps, seed, accept=self.a.accept, record=True)
            return C * steps, time.perf_counter() - t0
        steps = self.a.cpu_sample_steps
        if not steps:
            u, dt = one(20, 1)
            steps = int(max(50, min(10000, u / dt * 1.2 / C)))
        rate, reps, tot = timed_repeats(lambda r: one(steps, 1234 + r), budget_s)
        return rate, cores, "%d x (%d chains x %d steps), %.1f s in total, median (C restatement " \
                            "oracle/c, OpenMP)" % (reps, C, steps, tot)

    def finish(self):
        out, eng, d = self.last, self.eng, self.d
        st = eng.chain_stats(out["stat_sum"], out["stat_sumsq"], self.T)
        if d.world > 1:                      # the path's only collective: [D, 4] summaries
            d.dist.all_reduce(st)
        from probayes_b200.dist import rhat_from_stats
        rhat = rhat_from_stats(st, self.T)
        acc = float(out["accept_count"].sum().item()) / (self.C * self.T)
        return {"accept_rate": acc, "rhat": [float(v) for v in rhat]}


LR_LIMS = np.array([[-6., 6.], [-6., 6.], [0.001, 10.]])
LR_EX = np.array([[0, 0], [0, 0], [1, 0]])
LR_LG = np.zeros(3, int)


def linreg_data(N, seed=2024):
    rng = np.random.default_rng(seed)
    x = rng.normal(0, 1, N)
    y = -1 + 1.5 * x + rng.normal(0, .5, N)
    return x, y


class C3:
    name, metric, unit, scaling = "c3", "loglik_evals_per_sec", "evals/s", "strong"

    def __init__(self, args, world, chains=None, n_obs=None, mh_steps=None):
        self.a = args
        self.Ctot = chains or args.chains or 16384
        self.N = n_obs or args.n_obs or 1_000_000
        self.T = mh_steps or args.mh_steps
        self.world = world
        self.units = self.Ctot * self.T

    def config(self):
        return {"workload": "C3 (BASELINE.json configs[2]): Bayesian linear regression posterior, "
                            "normal likelihood over N=%d synthetic obs, 3 params, %d MH chains "
                            "sharded over %d GPU(s); one bench step = %d MH steps of every chain"
                            % (self.N, self.Ctot, self.world, self.T),
                "chains_total": self.Ctot, "n_obs": self.N, "mh_steps_per_step": self.T,
                "parallelism": "chains sharded x%d (strong), observations replicated" % self.world,
                "l2": "the 16 MB of observations are L2-resident by design (shared by all chains "
                      "of a CTA through TMA-staged shared-memory tiles); the kernel is FP64-bound",
                "terms_per_eval": self.N}

    def setup(self, eng, d):
        from probayes_b200.dist import shard_range
        self.eng, self.d = eng, d
        self.chain0, self.C = shard_range(self.Ctot, d.rank, d.world)
        self.xh, self.yh = linreg_data(self.N)
        self.x, self.y = eng.to_device(self.xh), eng.to_device(self.yh)
        self.sd = 0.5 / np.sqrt(self.N)
        self.state = eng.to_device(np.tile(np.array([[-1.], [1.5], [.5]]), (1, self.C)))
        burn = eng.mh_normreg(self.state, self.y, self.x, 2, LR_LIMS, LR_EX, LR_LG,
                              [2.4 * self.sd] * 3, seed=1, chain0=self.chain0, record=False)
        self.lp = burn["state_lp"]
        self.step0 = 2
        self.acc = 0
        self.nsteps = 0

    def step(self, k, ev=None):
        if ev is not None:
            ev[0].record(self.eng.stream)
        out = self.eng.mh_normreg(self.state, self.y, self.x, self.T, LR_LIMS, LR_EX, LR_LG,
                                  [2.4 * self.sd] * 3, seed=1, step0=self.step0,
                                  chain0=self.chain0, state_lp=self.lp, record=True, stats=False)
        if ev is not None:
            ev[1].record(self.eng.stream)
        self.step0 += self.T
        self.last = out

    kernel = "nr_tiles_kernel<4>"

    def roofline(self, kernel_ms, peaks, which, fp64_peak, sm_mhz):
        # kernel_ms brackets T launches (one per MH step) of this rank's C chains
        flops = 5.0 * self.N * self.C * self.T
        ach = flops / (kernel_ms * 1e-3) / 1e12
        return {"bound": "fp64", "kernel": self.kernel, "kernel_ms": kernel_ms / self.T,
                "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                "peak_source": "pbx_fp64_peak (dependent-free DFMA stream, measured in this run)",
                "traffic": None,
                "algorithmic_flops_per_launch": 5.0 * self.N * self.C,
                "terms_per_s": self.N * self.C * self.T / (kernel_ms * 1e-3),
                "note": "3 FP64 instr (fma, add, fma = 5 flop) per (chain, observation): the pipe-"
                        "time ceiling is 5/6 of the FMA-only flop peak; observation tiles are read "
                        "once per CTA (TMA) and shared by 128 x KC chains"}

    def e2e_setup(self):
        import scipy.stats
        import probayes_b200 as pb
        x = pb.RV('x', vtype=float, vset=[-np.inf, np.inf])
        y = pb.RV('y', vtype=float, vset=[-np.inf, np.inf])
        beta_0 = pb.RV('beta_0', vtype=float, vset=[-6., 6.], pscale='log')
        beta_1 = pb.RV('beta_1', vtype=float, vset=[-6., 6.], pscale='log')
        y_sigma = pb.RV('y_sigma', vtype=float, vset=[(0.001), 10.], pscale='log')

        def norm_reg(x, y, beta_0, beta_1, y_sigma):
            return scipy.stats.norm.logpdf(y, loc=beta_0 + beta_1 * x, scale=y_sigma)
        paras = beta_0 & beta_1 & y_sigma
        self.process = pb.SP(x & y, paras)
        self.process.set_prob(norm_reg, pscale='log')
        paras.set_tran(lambda **k: 0.)
        paras.set_delta([2.4 * self.sd])
        self.process.set_tran(paras)
        self.process.set_delta(paras)
        self.process.set_scores('metropolis')
        self.e2e_units = self.Ctot * self.T
        self.h2d = 16 * self.N + 3 * self.C * 8
        self.d2h = self.T * 4 * self.C * 8 + self.C * 8
        self.e2e_note = ("pb.SP(stats, paras).sampler(init, {'x,y': [x_obs, y_obs]}, chains=%d, "
                         "iid=True, joint=True) -> walk -> process(samples): host observations "
                         "uploaded and every recorded sample read back each step" % self.C)

    def e2e_step(self, k):
        smp = self.process.sampler({'beta_0': -1., 'beta_1': 1.5, 'y_sigma': 0.5},
                                   {'x,y': [self.xh, self.yh]}, stop=self.T, iid=True, joint=True,
                                   chains=self.C, seed=50 + k, chain0=self.chain0)
        summary = self.process(self.process.walk(smp))
        return summary.v['beta_1'][0, -1] + summary.u.count(True)

    def cpu_sample(self, budget_s=10.0):
        lo, cores = liboracle_all_cores()
        rng = np.random.default_rng(3)
        xh, yh = linreg_data(self.N)

        def one(Cs, seed):
            th = np.stack([rng.normal(-1, .001, Cs), rng.normal(1.5, .001, Cs),
                           rng.uniform(.49, .51, Cs)], axis=1)
            t0 = time.perf_counter()
            lo.normreg_logjoint(th, xh, yh, LR_LIMS, LR_EX, LR_LG)
            return Cs, time.perf_counter() - t0
        u, dt = one(max(cores, 8), 0)
        Cs = int(max(cores, min(4096, u / dt * 1.5)))
        rate, reps, tot = timed_repeats(lambda r: one(Cs, r), budget_s)
        return rate, cores, "%d x (%d chains x N=%d log-joint evaluations), %.1f s in total, " \
                            "median (C restatement oracle/c, OpenMP)" % (reps, Cs, self.N, tot)

    def finish(self):
        acc = self.d.sum(float(self.last["accept_count"].sum().item()))
        return {"accept_rate_last_step": acc / (self.Ctot * self.T)}


class C4:
    name, metric, unit, scaling = "c4", "loglik_evals_per_sec", "evals/s", "strong"

    def __init__(self, args, world, grid=None, n_obs=None):
        self.a = args
        self.M = self.S = grid or args.grid
        self.N = n_obs or args.n_obs or 100_000
        self.world = world
        self.units = self.M * self.S

    def config(self):
        return {"workload": "C4 (BASELINE.json configs[3]): discrete grid exact inference of a "
                            "normal mean/std posterior, %dx%d grid over N=%d synthetic obs, mu-row "
                            "slabs sharded over %d GPU(s); one bench step = log-joint + normalise "
                            "+ posterior + both marginals" % (self.M, self.S, self.N, self.world),
                "grid": [self.M, self.S], "n_obs": self.N,
                "parallelism": "mu-row slabs x%d (strong); all-reduce(max), all-reduce(sum), "
                               "all-reduce of the sigma marginal, all-gather of the mu marginal"
                               % self.world,
                "l2": "the 134 MB log-joint exceeds the 126 MB L2 at N=1; no explicit flush",
                "terms_per_eval": self.N}

    def setup(self, eng, d):
        from probayes_b200.dist import shard_range
        self.eng, self.d = eng, d
        rng = np.random.default_rng(7)
        self.data_h = rng.normal(50., 10., self.N)
        self.mu_h = np.linspace(40, 60, self.M + 2)[1:-1]
        self.sg_h = np.exp(np.linspace(np.log(5), np.log(20), self.S + 2)[1:-1])
        self.lpm_h = np.full(self.M, -np.log(20.))
        self.lps_h = np.full(self.S, -np.log(np.log(4.)))
        self.r0, self.rows = shard_range(self.M, d.rank, d.world)
        sl = slice(self.r0, self.r0 + self.rows)
        self.data = eng.to_device(self.data_h)
        self.mu, self.sg = eng.to_device(self.mu_h[sl]), eng.to_device(self.sg_h)
        self.lpm, self.lps = eng.to_device(self.lpm_h[sl]), eng.to_device(self.lps_h)
        self.lj = eng.empty(self.rows, self.S)
        self.counts = [shard_range(self.M, r, d.world)[1] for r in range(d.world)]
        self.group = None
        self.ms_a = self.ms_b = None

    def step(self, k, ev=None):
        from probayes_b200 import dist as pdist
        eng = self.eng
        if ev is not None:
            ev[0].record(eng.stream)
        eng.grid_norm_logjoint(self.data, self.mu, self.sg, self.lpm, self.lps, out=self.lj)
        if ev is not None:
            ev[1].record(eng.stream)
        if k == -1:
            return
        r = eng.grid_conditionalise(self.lj, want_post=True, inplace=True,
                                    group=self.d.dist.group.WORLD if self.d.world > 1 else None)
        self.marg_mu = pdist.gather_slabs(r["marg_mu"], self.counts)
        self.last = r

    kernel = "grid_logjoint_kernel"

    def roofline(self, kernel_ms, peaks, which, fp64_peak, sm_mhz):
        cellobs = float(self.N) * self.rows * self.S
        ach = 4.0 * cellobs / (kernel_ms * 1e-3) / 1e12
        return {"bound": "fp64", "kernel": self.kernel, "kernel_ms": kernel_ms,
                "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                "peak_source": "pbx_fp64_peak (measured in this run)", "traffic": None,
                "algorithmic_flops_per_launch": 4.0 * cellobs,
                "terms_per_s": cellobs / (kernel_ms * 1e-3),
                "note": "2 FP64 instr (mul, fma) per (cell, observation), counted as 2 FMA slots = "
                        "4 flop (SURVEY 8d); the normalise/marginal passes are HBM-bound and "
                        "reported under normalise"}

    def e2e_setup(self):
        self.e2e_units = self.M * self.S
        self.h2d = 8 * (self.N + self.rows + self.S) + 8 * (self.rows + self.S)
        self.d2h = 8 * (self.M + self.S) + 16
        self.e2e_note = ("probayes_b200.dist.dgei_sharded(engine, data, mu, sigma, priors): host "
                         "observations / grids uploaded every step; both marginals and the "
                         "normaliser read back; the posterior slab [M/G, S] stays device-backed "
                         "(PD.prob copies it on first access)")

    def e2e_step(self, k):
        from probayes_b200 import dist as pdist
        r = pdist.dgei_sharded(self.eng, self.data_h, self.mu_h, self.sg_h, self.lpm_h, self.lps_h)
        mm = r["marg_mu"].cpu().numpy()
        ms = r["marg_sigma"].cpu().numpy()
        return float(mm.max() + ms.max() + r["gsum"].item())

    def cpu_sample(self, budget_s=10.0):
        lo, cores = liboracle_all_cores()
        rng = np.random.default_rng(7)
        data = rng.normal(50., 10., self.N)

        def one(m, seed):
            mu = np.linspace(40, 60, m + 2)[1:-1]
            sg = np.exp(np.linspace(np.log(5), np.log(20), m + 2)[1:-1])
            t0 = time.perf_counter()
            lj = lo.grid_norm_logjoint(data, mu, sg, np.full(m, -np.log(20.)),
                                       np.full(m, -np.log(np.log(4.))))
            lo.grid_posterior(lj)
            return m * m, time.perf_counter() - t0
        u, dt = one(16, 0)
        m = int(max(16, min(512, np.sqrt(u / dt * 1.5))))
        rate, reps, tot = timed_repeats(lambda r: one(m, r), budget_s)
        return rate, cores, "%d x (%dx%d grid cells x N=%d, log-joint + posterior), %.1f s in " \
                            "total, median (C restatement oracle/c, OpenMP)" % (reps, m, m, self.N, tot)

    def finish(self):
        torch = self.d.torch
        lp = self.last["post"]
        tot = self.d.sum(float(torch.exp(lp).sum().item()))
        return {"posterior_sum": tot,
                "marg_mu_sum": float(torch.exp(self.marg_mu).sum().item())}


class C5:
    name, metric, unit, scaling = "c5", "mh_chain_steps_per_sec", "chain-steps/s", "strong"

    def __init__(self, args, world, chains=None, sweeps=None):
        self.a = args
        self.d_ = 64
        self.Ctot = chains or args.chains or 65536
        self.sweeps = sweeps or args.sweeps
        self.world = world
        self.units = self.Ctot * self.d_ * self.sweeps

    def config(self):
        return {"workload": "C5 (BASELINE.json configs[4]): multivariate normal-covariance Gibbs "
                            "(cond_cov) d=64, %d chains sharded over %d GPU(s); one bench step = %d "
                            "sweeps of 64 coordinate updates, every sweep's state recorded with its "
                            "target density (FP64 DMMA)" % (self.Ctot, self.world, self.sweeps),
                "chains_total": self.Ctot, "dims": self.d_, "sweeps_per_step": self.sweeps,
                "chain_step": "one coordinate update (the reference's step with tsteps=1)",
                "parallelism": "chains sharded x%d (strong)" % self.world,
                "l2": "recorded states per step (%.0f MB per rank) exceed the L2; no explicit flush"
                      % (self.sweeps * self.d_ * (self.Ctot / self.world) * 8 / 1e6)}

    def model(self):
        rng = np.random.default_rng(0)
        d = self.d_
        A = rng.standard_normal((d, d))
        cov = A @ A.T / d + np.eye(d)
        mean = rng.standard_normal(d)
        return mean, cov

    def setup(self, eng, d):
        from probayes_b200.dist import shard_range
        from probayes_b200.cond_cov import CondCov
        self.eng, self.d = eng, d
        self.chain0, self.C = shard_range(self.Ctot, d.rank, d.world)
        self.mean, self.cov = self.model()
        self.cc = CondCov(self.mean, self.cov, np.tile([-10., 10.], (self.d_, 1)))
        self.state = eng.to_device(np.tile(self.mean[:, None], (1, self.C)))
        self.step0 = 0

    def step(self, k, ev=None):
        if ev is not None:
            ev[0].record(self.eng.stream)
        self.last = self.eng.gibbs_mvn(self.state, self.cc, self.sweeps * self.d_, thin=self.d_,
                                       seed=5, step0=self.step0, chain0=self.chain0,
                                       want_prob=True, stats=False)
        if ev is not None:
            ev[1].record(self.eng.stream)
        self.step0 += self.sweeps * self.d_

    kernel = "gibbs_mvn_mma_kernel (conditional means and the recorded states' density, both DMMA)"

    def roofline(self, kernel_ms, peaks, which, fp64_peak, sm_mhz):
        d = self.d_
        flops = (4.0 * d * d + 2 * d) * self.C * self.sweeps
        ach = flops / (kernel_ms * 1e-3) / 1e12
        return {"bound": "fp64", "kernel": self.kernel, "kernel_ms": kernel_ms,
                "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                "peak_source": "pbx_fp64_peak (measured in this run)", "traffic": None,
                "algorithmic_flops_per_launch": flops,
                "ms_per_sweep": kernel_ms / self.sweeps,
                "note": "SURVEY 8d: 2 d^2 flop of conditional means + 2 d^2 + 2 d of density per "
                        "chain-sweep, both on the FP64 tensor cores (DMMA m8n8k4, which shares the "
                        "FP64 pipe with DFMA); each coordinate update also needs half a Philox "
                        "block and one table-driven inverse normal cdf"}

    def e2e_setup(self):
        self.e2e_units = self.Ctot * self.d_ * self.e2e_sweeps()
        self.h2d = self.d_ * self.C * 8
        self.d2h = self.e2e_sweeps() * (self.d_ + 1) * self.C * 8
        self.e2e_note = ("pb.SP(x0 & ... & x63) with set_tran(scipy.stats.multivariate_normal, "
                         "mean, cov, tsteps=1), set_scores('gibbs'): sampler(init, chains=%d, "
                         "thin=64) -> walk -> process(samples); host initial state in, every kept "
                         "state and density out" % self.C)
        import scipy.stats
        import probayes_b200 as pb
        import functools
        rvs = [pb.RV('x%d' % i, vtype=float, vset=(-10., 10.)) for i in range(self.d_)]
        self.process = pb.SP(functools.reduce(lambda a, b: a & b, rvs))
        self.process.set_prob(scipy.stats.multivariate_normal, self.mean, self.cov)
        self.process.set_tran(scipy.stats.multivariate_normal, self.mean, self.cov, tsteps=1)
        self.process.set_scores('gibbs')
        self.init = {'x%d' % i: float(self.mean[i]) for i in range(self.d_)}
        self.hostbuf = {}

    def e2e_sweeps(self):
        return min(self.sweeps, 20)

    def e2e_step(self, k):
        smp = self.process.sampler(self.init, stop=self.e2e_sweeps() * self.d_, chains=self.C,
                                   thin=self.d_, seed=70 + k, chain0=self.chain0,
                                   host_buffers=self.hostbuf)
        summary = self.process(self.process.walk(smp))
        return summary.v['x0'][0, -1]

    def cpu_sample(self, budget_s=10.0):
        lo, cores = liboracle_all_cores()
        from probayes_b200.cond_cov import CondCov
        mean, cov = self.model()
        cc = CondCov(mean, cov, np.tile([-10., 10.], (self.d_, 1)))
        coef = cc.coef_matrix()

        def one(Cs, seed):
            x = np.tile(mean, (Cs, 1))
            t0 = time.perf_counter()
            lo.gibbs_mvn_walk(x, mean, coef, cc.stdv, cc.cdfs, 4 * self.d_, seed=seed)
            return Cs * 4 * self.d_, time.perf_counter() - t0
        u, dt = one(256, 0)
        Cs = int(max(256, min(65536, u / dt * 1.5 / (4 * self.d_))))
        rate, reps, tot = timed_repeats(lambda r: one(Cs, r), budget_s)
        return rate, cores, "%d x (%d chains x 4 sweeps of 64 coordinate updates), %.1f s in " \
                            "total, median (C restatement oracle/c, OpenMP; no density " \
                            "evaluation)" % (reps, Cs, tot)

    def finish(self):
        x = self.last["x"][-1]                       # [d, C] last kept sweep
        m = self.d.sum(float(x[0].sum().item())) / self.Ctot
        return {"mean_x0": m, "target_mean_x0": float(self.mean[0])}


WORKLOADS = {"c2": C2, "c3": C3, "c4": C4, "c5": C5}


# ---------------------------------------------------------------------------
# timing of one workload on the device (used for the headline and for secondary_n)
# ---------------------------------------------------------------------------
def time_workload(wl, d, steps, warmup, sampler=None):
    torch = d.torch
    eng = wl.eng
    for w in range(warmup):
        wl.step(-1000 + w)
    d.barrier()
    if sampler is not None:
        sampler.start()
    l0 = eng.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(steps)]
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for k in range(steps):
        wl.step(k, ev[k])                    # no host sync inside the timed region
    t_end.record()
    d.barrier()
    launches = eng.launches - l0
    clocks = sampler.stop() if sampler is not None else None
    total_ms = d.max(t_start.elapsed_time(t_end))
    kms = float(np.mean([a_.elapsed_time(b_) for a_, b_ in ev]))
    ms_per_step = total_ms / steps
    return dict(value=wl.units / (ms_per_step * 1e-3), ms_per_step=ms_per_step,
                kernel_ms=kms, launches=int(launches), clocks=clocks)


def time_e2e(wl, d, steps, warmup=2):
    wl.e2e_setup()
    for w in range(warmup):
        wl.e2e_step(-100 + w)
    d.barrier()
    t0 = time.perf_counter()
    for k in range(steps):
        wl.e2e_step(k)
    d.torch.cuda.synchronize()
    dt = d.max(time.perf_counter() - t0)
    out = {"value": wl.e2e_units * steps / dt, "unit": wl.unit,
           "h2d_bytes_per_step": int(wl.h2d), "d2h_bytes_per_step": int(wl.d2h),
           "steps": steps, "ms_per_step": 1e3 * dt / steps, "api": wl.e2e_note}
    if hasattr(wl, "e2e_ceiling"):
        ceil = wl.e2e_ceiling(d)
        if ceil:
            out["ceiling"] = ceil
            out["frac_of_ceiling"] = out["value"] / ceil["chain_steps_per_s_at_ceiling"]
    return out


def reference_numpy_c1():
    """The REAL reference's NumPy sampler (oracle/_ref, config C1) on this box's cores."""
    try:
        from oracle import ref_run
        if not ref_run.available():
            return {"unavailable": "oracle/_ref/probayes not shipped (run oracle/ref_run.py in "
                                   "the development container)"}
        r1, dt1, nacc = ref_run.c1_rate(4096, seed=0)
        rn, procs, wall = ref_run.c1_rate_all_cores(2048)
        return {"value": r1, "unit": "chain-steps/s", "cores": 1, "kind": "oracle/_ref",
                "sample": "examples/mcmc/mcmc_prob4a.py model, 1 chain x 4096 steps incl. "
                          "process(samples), np.random.seed(0): %.2f s, %d accepted" % (dt1, nacc),
                "all_cores": {"value": rn, "cores": procs,
                              "sample": "%d independent single-chain samplers x 2048 steps in %d "
                                        "processes, %.2f s wall" % (procs, procs, wall)}}
    except Exception as e:                                   # diagnostics only
        return {"unavailable": repr(e)}


def reference_numpy(workload):
    """The REAL reference (oracle/_ref) on this box's cores for the workload's config, at the
    largest size it evaluates sensibly."""
    if workload == "c2":
        return reference_numpy_c1()
    try:
        from oracle import ref_run
        if not ref_run.available():
            return {"unavailable": "oracle/_ref/probayes not shipped (run oracle/ref_run.py in "
                                   "the development container)"}
        if workload == "c3":
            r, dt = ref_run.c3_rate(100000, 40)
            return {"value": r, "unit": "evals/s", "cores": 1, "kind": "oracle/_ref",
                    "terms_per_s": r * 1e5,
                    "sample": "gibbs_linreg model as an MH sampler through SP.sampler, 1 chain x "
                              "40 steps, N = 100000 obs per log-likelihood evaluation (the full "
                              "config has N = 1000000): %.2f s" % dt}
        if workload == "c4":
            r, tps, dt = ref_run.c4_rate(1000, 256, 256)
            return {"value": r, "unit": "evals/s", "cores": 1, "kind": "oracle/_ref",
                    "terms_per_s": tps,
                    "sample": "dgei_norm1d_improved model: joint + conditionalise + both "
                              "marginals on a 256 x 256 grid over N = 1000 obs (one eval = one "
                              "cell = N terms; the full config has 4096 x 4096 x 100000, whose "
                              "[N, M, S] temporary the reference cannot hold): %.2f s" % dt}
        if workload == "c5":
            r, dt = ref_run.c5_rate(64, 4096)
            return {"value": r, "unit": "chain-steps/s", "cores": 1, "kind": "oracle/_ref",
                    "sample": "CondCov.interp at d = 64, 1 chain x 4096 coordinate updates "
                              "(cond_cov.py:42-65 as RF.eval_tfun calls it; through SP the "
                              "reference needs one numpy axis per RV and stops at 32): %.2f s" % dt}
    except Exception as e:                                   # diagnostics only
        return {"unavailable": repr(e)}
    return {"unavailable": "no reference leg for " + workload}


def run_reference(args):
    """--impl reference: the CPU implementation of the SAME workload on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload](args, max(1, args.gpus))
    for _ in range(min(args.warmup, 1)):
        wl.cpu_sample(budget_s=1.0)
    rates, t_all = [], time.perf_counter()
    for _ in range(args.steps):
        r, cores, sample = wl.cpu_sample(budget_s=max(2.0, 60.0 / max(1, args.steps)))
        rates.append(r)
        if time.perf_counter() - t_all > 150:
            break
    value = float(np.mean(rates))
    line = {"impl": "reference", "metric": wl.metric, "value": value, "unit": wl.unit,
            "n_gpus": args.gpus, "steps": len(rates), "warmup": args.warmup,
            "ms_per_step": 1e3 * wl.units / value, "higher_is_better": True,
            "scaling": wl.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": wl.config(),
            "cpu_baseline": {"value": value, "unit": wl.unit, "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": wl.unit, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    line["cpu_baseline"]["reference_numpy"] = reference_numpy(args.workload)
    print(json.dumps(line), flush=True)


def run_ours(args):
    d = Dist()
    torch = d.torch
    from probayes_b200.engine import get_engine
    eng = get_engine(d.local)
    wl = WORKLOADS[args.workload](args, d.world)
    wl.setup(eng, d)
    sampler = ClockSampler(d.local) if d.rank == 0 else None
    res = time_workload(wl, d, args.steps, args.warmup, sampler)
    quality = wl.finish()
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(3, min(args.steps, 10))
        try:
            e2e = time_e2e(wl, d, e2e_steps)
        except Exception as e:                               # never lose the headline line
            e2e = {"value": None, "error": repr(e)}
    sec_n = None
    if args.workload == "c2" and not args.no_secondary:
        # the other half of the metric at the SAME N: short sharded passes of c3 / c4 / c5
        sec_n = {}
        small = args.quick
        for name, mk in (("c3", lambda: C3(args, d.world, chains=16384,
                                          n_obs=100_000 if small else 1_000_000, mh_steps=10)),
                         ("c4", lambda: C4(args, d.world, grid=1024 if small else 4096,
                                          n_obs=10_000 if small else 100_000)),
                         ("c5", lambda: C5(args, d.world, chains=65536, sweeps=20))):
            try:
                w2 = mk()
                w2.setup(eng, d)
                r2 = time_workload(w2, d, 3, 2)
                sec_n[name] = {"metric": w2.metric, "value": r2["value"], "unit": w2.unit,
                               "n_gpus": d.world, "scaling": w2.scaling,
                               "ms_per_step": r2["ms_per_step"],
                               "workload": w2.config()["workload"]}
                del w2
            except Exception as e:
                sec_n[name] = {"error": repr(e)}
            torch.cuda.empty_cache()
    if d.rank == 0:
        peaks, which = measured_peaks()
        fp64_peak = eng.fp64_peak_tflops()
        clocks = res["clocks"]
        line = {
            "metric": wl.metric, "value": res["value"], "unit": wl.unit, "n_gpus": d.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": wl.config(),
            "e2e": e2e, "gpu_launches": res["launches"],
            "roofline": wl.roofline(res["kernel_ms"], peaks, which, fp64_peak,
                                    (clocks or {}).get("sm_mhz")),
            "fp64": {"peak_tflops_measured": fp64_peak},
            "clocks": clocks, "cpu_binding": d.cpu_binding, "quality": quality,
        }
        if sec_n is not None:
            line["secondary_n"] = sec_n
        if args.workload == "c2" and not args.no_secondary and d.world == 1:
            try:
                line["secondary"] = secondary(eng, peaks, fp64_peak, quick=args.quick)
                line["roofline_stream"] = line["secondary"].pop("roofline_stream")
            except Exception as e:
                line["secondary"] = {"error": repr(e)}
        if not args.no_cpu_baseline and d.world == 1:       # rank 0 at N = 1 only
            if getattr(bind_to_gpu_cpus, "original", None):  # the CPU arm gets every core
                os.sched_setaffinity(0, bind_to_gpu_cpus.original)
            try:
                r, cores, sample = wl.cpu_sample()
                line["cpu_baseline"] = {"value": r, "unit": wl.unit, "cores": cores,
                                        "kind": "port", "sample": sample}
                line["cpu_baseline"]["reference_numpy"] = reference_numpy(args.workload)
            except Exception as e:
                line["cpu_baseline"] = {"error": repr(e)}
        print(json.dumps(line), flush=True)
    d.close()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

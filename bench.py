#!/usr/bin/env python
"""bench.py -- headline benchmark of the probayes hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c2] [--chains C] [--walk-steps T]

Workload (BASELINE.json configs[1], "C2"): 4096 independent Metropolis-Hastings
chains x 10^4 steps on the 2-D correlated normal of examples/mcmc/mcmc_prob4a.py,
fp64, native Philox RNG, every step recorded (thin=1).  One bench "step" = one
whole walk (C x T chain-steps).  Under torchrun each rank runs its own 4096
chains (chains are independent: weak scaling, no data-path collective; the
per-chain summaries are all-reduced once after the timed region).

Prints ONE JSON line (rank 0):
  value   chain-steps/s, device-resident (outputs written to HBM)
  e2e     same metric through the host-buffer C-ABI call: H2D of the initial
          state + D2H of every recorded sample and density inside the timed region
  roofline / cpu_baseline / clocks / gpu_launches as the bench contract asks.
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
METRIC = "mh_chain_steps_per_sec"
UNIT = "chain-steps/s"
# DRAM traffic of one K1 launch on the default workload (4096 chains x 10^4 steps, D=2, thin=1):
# dram__bytes_read.sum (290 KB) + dram__bytes_write.sum (925.7 MB) from the ncu --set full capture in
# profiles/r1i_ncu_full_k1.csv.  Algorithmic bytes are 983.0 MB; the difference is the tail of
# the output still resident in the 126 MB L2 when the kernel ends.  No re-reads.
NCU_K1_DRAM_BYTES = 289536 + 925711104


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--chains", type=int, default=4096, help="chains per GPU")
    ap.add_argument("--walk-steps", type=int, default=10000, help="MH steps per walk")
    ap.add_argument("--thin", type=int, default=1)
    ap.add_argument("--accept", default="log", choices=["reference", "log"])
    ap.add_argument("--variant", type=int, default=0, help="K1 kernel: 0 auto, 1 per-thread, 2 warp-specialised")
    ap.add_argument("--cpu-sample-steps", type=int, default=0,
                    help="MH steps of the bounded CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the C3/C4/C5 secondary kernel timings")
    ap.add_argument("--quick", action="store_true", help="smaller secondary workloads")
    return ap.parse_args()


def config(args, n_gpus):
    return {"workload": "C2: %d MH chains/GPU x %d steps, 2-D correlated normal "
                        "(mcmc_prob4a), thin=%d, Philox4x32-10" % (args.chains, args.walk_steps,
                                                                  args.thin),
            "chains_per_gpu": args.chains, "walk_steps": args.walk_steps, "thin": args.thin,
            "accept": args.accept, "parallelism": "chains x%d" % n_gpus,
            "l2": "outputs per walk (%.0f MB) exceed the 126 MB L2; no explicit flush"
                  % (args.chains * (args.walk_steps // args.thin) * 24 / 1e6)}


# ---------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ---------------------------------------------------------------------------
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


# ---------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (cpu_baseline / --impl reference)
# ---------------------------------------------------------------------------
def cpu_walk_rate(chains, steps, accept, seed=1234):
    """Times the oracle port of the C2 walk on a bounded sample.  Prefers the C
    restatement (OpenMP, all host cores); falls back to the numpy restatement
    (1 core).  Returns (chain_steps_per_s, cores, kind_detail, seconds)."""
    try:
        from oracle.c import liboracle
        lib = liboracle.load()
    except Exception:
        lib = None
    if lib is not None:
        cores = liboracle.use_all_cores()
        t0 = time.perf_counter()
        liboracle.mh_mvn_walk(np.tile(INIT, (chains, 1)), MEAN, COV, steps, seed,
                              accept=accept, record=True)
        dt = time.perf_counter() - t0
        return chains * steps / dt, cores, "C restatement (oracle/c, OpenMP)", dt
    from oracle import np_oracle as o
    from oracle import philox
    t0 = time.perf_counter()
    Z = philox.normals(seed, steps, chains, 2)
    U = philox.thresholds(seed, steps, chains)
    o.mh_mvn_walk(np.tile(INIT, (chains, 1)), Z, U, MEAN, COV, accept=accept)
    dt = time.perf_counter() - t0
    return chains * steps / dt, 1, "numpy restatement (oracle/np_oracle.py)", dt


def auto_cpu_steps(chains, accept):
    """Sizes the CPU sample for roughly 10-20 s of work from a short probe."""
    rate, _, _, _ = cpu_walk_rate(chains, 20, accept)
    return int(max(50, min(10000, rate * 12.0 / chains)))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = args.cpu_sample_steps or auto_cpu_steps(args.chains, args.accept)
    rates = []
    for _ in range(args.warmup):
        cpu_walk_rate(args.chains, max(10, steps // 10), args.accept)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        r, cores, detail, dt = cpu_walk_rate(args.chains, steps, args.accept)
        rates.append(r)
        if time.perf_counter() - t_all > 150:
            break
    value = float(np.mean(rates))
    sample = "%d chains x %d steps per bench step (%s)" % (args.chains, steps, detail)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": len(rates), "warmup": args.warmup,
            "ms_per_step": 1e3 * args.chains * steps / value, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"



# ---------------------------------------------------------------------------
# secondary measurements: the other configs of BASELINE.json, one short pass each
# ---------------------------------------------------------------------------
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
    out["roofline_stream"] = {"bound": "hbm", "kernel": "nr_stream_kernel<8,true>",
                              "achieved": out["stream_c4"]["hbm_gbs"], "peak": peaks["hbm_gbs"],
                              "unit": "GB/s", "frac": out["stream_c4"]["frac"],
                              "algorithmic_bytes_per_launch": 16.0 * Ns,
                              "workload": "streaming normal log-likelihood, 4 chains, N=%d "
                                          "(16 B per observation per step)" % Ns}
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
                                     "note": "max + sumexp + posterior/marginal passes "
                                             "(3 reads + 1 write of the grid)"}
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
                       "note": "bounded by the serial coordinate dependency: per update one "
                               "Philox block + ndtri (~100 instr) next to 2 d flop"}
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
    return out


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


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    cpu_binding = bind_to_gpu_cpus(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from probayes_b200.engine import get_engine
    eng = get_engine(local)
    C, T, thin, D = args.chains, args.walk_steps, args.thin, 2
    R = T // thin
    chain0 = rank * C
    init_host = np.tile(INIT[:, None], (1, C))
    state = eng.to_device(init_host)
    init_dev = eng.to_device(init_host)
    out_bytes = R * (D + 1) * C * 8

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    bufs = {"x": eng.empty(R, D, C), "prob": eng.empty(R, C)}

    def walk(seed, events=None):
        state.copy_(init_dev)
        return eng.mh_mvn(state, MEAN, COV, T, thin=thin, seed=seed, chain0=chain0,
                          accept=args.accept, variant=args.variant, out=bufs, events=events)

    # ---- device-resident timing ------------------------------------------------
    for w in range(args.warmup):
        walk(1000 + w)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for k in range(args.steps):
        out = walk(2000 + k, ev[k])          # CUDA events around the kernel on its stream;
    t_end.record()                           # no host sync inside the timed region
    barrier()
    launches = eng.launches - l0
    total_ms = t_start.elapsed_time(t_end)
    kms = [a_.elapsed_time(b_) for a_, b_ in ev]
    clocks = sampler.stop() if rank == 0 else None
    tm = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms = float(tm.item())
    ms_per_step = total_ms / args.steps
    value = world * C * T / (ms_per_step * 1e-3)
    kernel_ms = float(np.mean(kms))

    # per-chain summaries -> R-hat inputs, all-reduced once (not in the timed region)
    st = eng.chain_stats(out["stat_sum"], out["stat_sumsq"], T)
    if world > 1:
        dist.all_reduce(st)
    st = st.cpu().numpy()
    Cn = st[:, 3]
    W = st[:, 2] / Cn
    B = T * (st[:, 1] - st[:, 0] ** 2 / Cn) / (Cn - 1)
    rhat = np.sqrt(((T - 1) / T * W + B / T) / W)
    acc_rate = float(out["accept_count"].sum().item()) / (C * T)

    # ---- end-to-end through the PUBLIC API (the call a user makes) ----------------
    # pb.SP(...).sampler(init, chains=, host_stream=True) -> walk -> process(samples):
    # host init state -> H2D, chunked kernel launches, every recorded sample and
    # density streamed D2H into pinned buffers, summary PDs built from them.
    import scipy.stats
    import probayes_b200 as pb
    xr = pb.RV('x', vtype=float, vset=(-np.inf, np.inf))
    yr = pb.RV('y', vtype=float, vset=(-np.inf, np.inf))
    process = pb.SP(xr & yr)
    process.set_prob(scipy.stats.multivariate_normal, list(MEAN), COV.tolist())
    process.set_tran(lambda **kw: 1.)
    process.set_delta(scipy.stats.norm(0., 1.))
    process.set_scores('hastings')
    process.set_update('metropolis')
    hostbuf = {}
    e2e_steps = max(3, min(args.steps, 10))

    def api_walk(seed):
        smp = process.sampler({'x': INIT[0], 'y': INIT[1]}, stop=T, chains=C, thin=thin,
                              seed=seed, accept=args.accept, host_stream=True,
                              host_buffers=hostbuf)
        summary = process(process.walk(smp))
        return summary.v['x'][0, -1] + summary.u.count(True)     # touch the result

    for w in range(2):
        api_walk(3000 + w)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        api_walk(4000 + k)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * C * T * e2e_steps / float(te.item())
    h2d = D * C * 8
    d2h = out_bytes + (2 * D + 2) * C * 8 + D * C * 8

    if rank == 0:
        peaks, which = measured_peaks()
        achieved = out_bytes / (kernel_ms * 1e-3) / 1e9
        fp64_peak = eng.fp64_peak_tflops()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                         "traffic": NCU_K1_DRAM_BYTES if (C, T, thin, args.variant) == (4096, 10000, 1, 0) else None,
                         "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, "
                                           "profiles/r1i_ncu_full_k1.csv",
                         "peak_source": which, "kernel": "mh_mvn_kernel<2>" if args.variant == 1 else "mh_mvn_ws_kernel<2>",
                         "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": out_bytes,
                         "note": "K1 writes (D+1)*8 B per recorded chain-step; it is "
                                 "FP64-pipe/latency bound, not HBM bound (see fp64)"},
            "fp64": {"peak_tflops_measured": fp64_peak},
            "clocks": clocks,
            "cpu_binding": cpu_binding,
            "quality": {"accept_rate": acc_rate, "rhat": [float(v) for v in rhat]},
        }
        if not args.no_secondary and world == 1:        # the other configs: N = 1 only
            try:
                line["secondary"] = secondary(eng, peaks, fp64_peak, quick=args.quick)
                line["roofline_stream"] = line["secondary"].pop("roofline_stream")
            except Exception as e:                           # never lose the headline line
                line["secondary"] = {"error": repr(e)}
        if not args.no_cpu_baseline and world == 1:     # reported on rank 0 at N = 1 only
            if getattr(bind_to_gpu_cpus, "original", None):     # the CPU arm gets every core
                os.sched_setaffinity(0, bind_to_gpu_cpus.original)
            steps = args.cpu_sample_steps or auto_cpu_steps(C, args.accept)
            rs, tot, reps = [], 0.0, 0
            while tot < 10.0 and reps < 40:            # ~10 s of CPU work, median over repeats
                r, cores, detail, dt = cpu_walk_rate(C, steps, args.accept, seed=1234 + reps)
                rs.append(r)
                tot += dt
                reps += 1
            r = float(np.median(rs))
            line["cpu_baseline"] = {"value": r, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d x (%d chains x %d steps), %.1f s in total, "
                                              "median (%s)" % (reps, C, steps, tot, detail),
                                    "reference_python_survey": "1332 chain-steps/s, 1 core "
                                                               "(BASELINE.md, config C1)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

"""K6 timings on one B200: argsort / cumprob / expectation / take_axis on a 4096^2
grid's worth of doubles (n = 2^24) and on the OMC sample set."""
import json
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probayes_b200.engine import get_engine


def timeit(fn, reps=10, warm=3):
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


def main():
    eng = get_engine(0)
    out = {}
    for n in (1 << 20, 1 << 24, 1 << 26):
        rng = np.random.default_rng(1)
        for kind in ("normal", "unit"):
            k = rng.standard_normal(n) if kind == "normal" else 1.0 + rng.random(n)
            kd = eng.to_device(k)
            ms = timeit(lambda: eng.argsort(kd, want_keys=True))
            out["argsort_%s_n%d" % (kind, n)] = dict(ms=ms, mkeys_per_s=n / ms / 1e3)
        p = eng.to_device(rng.random(n))
        lp = eng.to_device(np.log(rng.random(n)) - 100.0)
        ms = timeit(lambda: eng.cumprob(p, False))
        out["cumprob_lin_n%d" % n] = dict(ms=ms, gbs=24.0 * n / ms / 1e6)
        ms = timeit(lambda: eng.cumprob(lp, True))
        out["cumprob_log_n%d" % n] = dict(ms=ms, gbs=24.0 * n / ms / 1e6)
        ms = timeit(lambda: eng.expectation_sums(p, False, None, p.reshape(1, -1)))
        out["expect_1d_n%d" % n] = dict(ms=ms, gbs=16.0 * n / ms / 1e6)
    M = 4096
    g = eng.to_device(np.random.default_rng(2).random((M, M)))
    rv = eng.to_device(np.random.default_rng(3).random((1, M)))
    ms = timeit(lambda: eng.expectation_sums(g, False, rv, rv))
    out["expect_grid_4096"] = dict(ms=ms, gbs=8.0 * M * M / ms / 1e6)
    perm = torch.randperm(M, device=eng.device).to(torch.int32)
    for ax in (0, 1):
        ms = timeit(lambda: eng.take_axis(g, perm, ax))
        out["take_axis%d_4096" % ax] = dict(ms=ms, gbs=16.0 * M * M / ms / 1e6)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

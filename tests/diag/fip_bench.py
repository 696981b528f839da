# FIP-periodogram accumulation: device (rvl_fip_accumulate) vs the CPU oracle on one (run, k) block
# of the reference's size (nfreq = 50000, evidence/fip_criterion.py:229).
#   python tests/diag/fip_bench.py [n_samples] [k]
import sys, time, numpy as np
sys.path.insert(0, '.')
from evidence_b200 import fip
from oracle import fip_oracle as fo
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng = np.random.default_rng(0)
centres = np.array([3.1, 42.0, 290.0, 11.0, 600.0])[:k]
per = np.exp(rng.normal(np.log(centres), 0.01, (n, k)))
mask = rng.random((n, k)) < 0.2
per[mask] = np.exp(rng.uniform(0, np.log(1000.0), mask.sum()))
w = rng.random(n) + 1e-3
Pmin, Pmax, nfreq, Tobs = 1.0, 1000.0, 50000, 400.0
nu, nua, nub = fip.frequency_grid(Pmin, Pmax, nfreq, Tobs)
for alias in (False, True):
    row = np.ones(nfreq); fip.accumulate_block(row, nua, nub, per, w, 0.7, Pmin, Pmax, alias)  # warm-up
    kms, walls = [], []
    for _ in range(5):
        row = np.ones(nfreq)
        t0 = time.perf_counter(); ms = fip.accumulate_block(row, nua, nub, per, w, 0.7, Pmin, Pmax, alias)
        walls.append(time.perf_counter() - t0); kms.append(ms)
    t0 = time.perf_counter(); want = fo.accumulate(np.ones(nfreq), nua, nub, per, w, 0.7, Pmin, Pmax, alias); t_vec = time.perf_counter() - t0
    m = min(n, 5000)
    t0 = time.perf_counter(); fo.accumulate_literal(np.ones(nfreq), nua, nub, per[:m], w[:m], 0.7, Pmin, Pmax, alias); t_lit = (time.perf_counter() - t0) * n / m
    print(f"n={n} k={k} alias={alias}: kernels {min(kms):.3f} ms ({n/min(kms)/1e3:.1f} M samples/s), call incl. copies {min(walls)*1e3:.2f} ms; "
          f"oracle vectorised numpy {t_vec*1e3:.1f} ms; reference-style Python loop {t_lit:.2f} s (extrapolated from {m} samples); "
          f"max |d fapnu| = {np.max(np.abs(row - want)):.2e}")

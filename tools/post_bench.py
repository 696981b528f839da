"""Post-processing entry points after the round-2 polish: planet ordering (row-per-warp kernel) and
FIP accumulation, kernel time and whole-call time of repeated calls (scratch kept between calls).
usage: python tools/post_bench.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from evidence_b200 import fip
from test_fip import _posterior

names, s = _posterior(1, 400000, 4)
nbytes = s.size * 8 * 2
for rep in range(4):
    t0 = time.perf_counter()
    out = fip.order_planets(s, names, 4)
    dt = time.perf_counter() - t0
    ms = fip.order_planets.last_kernel_ms
    print(f"order_planets call {rep}: kernel {ms:.4f} ms = {nbytes / ms / 1e6:.0f} GB/s algorithmic "
          f"({s.shape[0]} x {s.shape[1]}), whole call {dt * 1e3:.1f} ms")
rng = np.random.default_rng(0)
per = np.exp(rng.uniform(0, np.log(1000), (200000, 3)))
for rep in range(4):
    t0 = time.perf_counter()
    nu, fap = fip.fip_periodogram([[None, None, None, (per, np.ones(len(per)))]], [-30.0, -20.0, -10.0, -1.0],
                                  Pmin=1.0, Pmax=1000.0, nfreq=50000, Tobs=3000.0, device=0)
    print(f"fip_periodogram call {rep}: whole call {(time.perf_counter() - t0) * 1e3:.1f} ms")

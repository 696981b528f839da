"""Planet ordering after the round-2 polish (tile-per-block kernel, scratch kept between calls): kernel
time and whole-call time of repeated calls.  usage: python tools/post_bench.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from evidence_b200 import fip
from test_fip import _posterior

for n in (400000, 4000000):
    names, s = _posterior(1, n, 4)
    nbytes = s.size * 8 * 2
    for rep in range(3):
        t0 = time.perf_counter()
        out = fip.order_planets(s, names, 4)
        dt = time.perf_counter() - t0
        ms = fip.order_planets.last_kernel_ms
        print(f"order_planets call {rep}: kernel {ms:.4f} ms = {nbytes / ms / 1e6:.0f} GB/s algorithmic "
              f"({s.shape[0]} x {s.shape[1]}), whole call {dt * 1e3:.1f} ms", flush=True)

"""rvl_create_multi: ONE process, ONE host call per batch, rows sharded over every GPU of the box
inside the library.  usage: python tools/multi_handle_rate.py [config] [rows_per_gpu]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from evidence_b200 import synth
from evidence_b200.rvmodel import RVModel

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
per = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
case = synth.make_case(cfg)
ngpu = torch.cuda.device_count()
for n_dev in sorted({1, 2, 4, ngpu} & set(range(1, ngpu + 1))):
    m = RVModel(case.fixedpardict, case.datadict(), case.parnames, devices=list(range(n_dev)))
    B = per * n_dev
    th = torch.from_numpy(case.draw_theta(B, seed=5)).pin_memory()
    out = torch.empty(B, dtype=torch.float64).pin_memory()
    for _ in range(3):
        m.log_likelihood_batch(th.numpy(), out=out.numpy())
    t0 = time.perf_counter()
    reps = 8
    for _ in range(reps):
        m.log_likelihood_batch(th.numpy(), out=out.numpy())
    dt = (time.perf_counter() - t0) / reps
    print(f"rvl_create_multi over {n_dev} GPU(s): {B} theta per call (config {cfg}), {dt * 1e3:.2f} ms per call, "
          f"{B / dt / 1e6:.2f} M lnL/s end to end (pinned host buffers, one host thread)", flush=True)
    m.close()

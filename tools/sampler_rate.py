"""Throughput of the device-resident sampler: native bookkeeping kernels (csrc/rvslice.cu) vs the
torch formulation vs the numpy-bookkeeping sampler.  usage: python tools/sampler_rate.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from evidence_b200 import synth
from evidence_b200.rvmodel import RVModel
from evidence_b200.sampler import nested_sample
from evidence_b200.sampler_dev import nested_sample_device

case = synth.make_case(2, seed=11, n_epochs=300)
model = RVModel(case.fixedpardict, case.datadict(), case.parnames)
model.set_priors(case.priordict)


def run(tag, fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{tag:44s} lnZ {r.logz:9.3f} +- {r.logzerr:5.2f}  ncall {r.ncall:10d}  {dt:6.2f} s  "
          f"{r.ncall / dt / 1e6:6.2f} M lnL/s", flush=True)


# population 2048 walkers per round
big = dict(nlive=4096, batch_fraction=0.5, nsteps=8, dlogz=5.0, frac_remain=0.5)
run("native, population 2048", lambda: nested_sample_device(model.transform_loglike_device, case.ndim, seed=1, **big))
run("torch,  population 2048", lambda: nested_sample_device(model.transform_loglike_device, case.ndim, seed=1, native=False, **big))
small = dict(nlive=200, nsteps=12)
run("native, nlive 200", lambda: nested_sample_device(model.transform_loglike_device, case.ndim, seed=3, **small))
run("torch,  nlive 200", lambda: nested_sample_device(model.transform_loglike_device, case.ndim, seed=3, native=False, **small))
run("numpy bookkeeping, nlive 200", lambda: nested_sample(model.log_likelihood_batch, model.prior_transform_batch, case.ndim,
                                                          fused=model.transform_loglike_batch, seed=3, **small))

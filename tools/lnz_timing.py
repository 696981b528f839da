# wall time of one seeded ln Z run (in-repo vectorised sampler) with separate transform + loglike
# calls vs the fused u -> theta -> lnL call
import sys, time, numpy as np
sys.path.insert(0, '.')
from evidence_b200 import synth
from evidence_b200.rvmodel import RVModel
from evidence_b200.sampler import nested_sample
case = synth.make_case(2, seed=4, n_epochs=300)
m = RVModel(case.fixedpardict, case.datadict(), case.parnames)
m.set_priors(case.priordict)
kw = dict(nlive=400, seed=3, nsteps=20)
nested_sample(m.log_likelihood_batch, m.prior_transform_batch, case.ndim, nlive=50, seed=1, nsteps=4)  # warm-up
for name, fused, spec in (("separate", None, 1), ("fused", m.transform_loglike_batch, 1),
                          ("fused, look-ahead 4", m.transform_loglike_batch, 4),
                          ("fused, look-ahead auto", m.transform_loglike_batch, None)):
    t0 = time.perf_counter()
    r = nested_sample(m.log_likelihood_batch, m.prior_transform_batch, case.ndim, fused=fused,
                      speculate=spec, **kw)
    dt = time.perf_counter() - t0
    print(f"{name:24s}: ln Z = {r.logz:.3f} +- {r.logzerr:.3f}, {r.ncall} likelihood calls, {dt:.2f} s wall, {r.ncall/dt/1e3:.1f} k lnL/s")
# the same run with a large population: nlive = 4096, half of the live points replaced per round
# (2048 walkers per device call) -- the regime in which a run is device-bound, not host-bound
kw = dict(nlive=4096, seed=3, nsteps=20, batch_fraction=0.5)
t0 = time.perf_counter()
r = nested_sample(m.log_likelihood_batch, m.prior_transform_batch, case.ndim, fused=m.transform_loglike_batch, **kw)
dt = time.perf_counter() - t0
print(f"population 2048: ln Z = {r.logz:.3f} +- {r.logzerr:.3f}, {r.ncall} likelihood calls, {dt:.2f} s wall, {r.ncall/dt/1e6:.2f} M lnL/s")
# the whole run on the device (sampler_dev): same model, same settings
import torch
from evidence_b200.sampler_dev import nested_sample_device
fd = lambda U: m.transform_loglike_device(U)
nested_sample_device(fd, case.ndim, nlive=50, seed=1, nsteps=4)  # warm-up
for name, kw in (("device-resident", dict(nlive=400, seed=3, nsteps=20)),
                 ("device-resident, population 2048", dict(nlive=4096, seed=3, nsteps=20, batch_fraction=0.5))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = nested_sample_device(fd, case.ndim, **kw)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name:34s}: ln Z = {r.logz:.3f} +- {r.logzerr:.3f}, {r.ncall} likelihood calls, {dt:.2f} s wall, {r.ncall/dt/1e6:.2f} M lnL/s")

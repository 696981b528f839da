"""Small launches of every kernel of the library, for compute-sanitizer (memcheck / racecheck /
synccheck):  compute-sanitizer --tool memcheck python tools/sanitize_probe.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from evidence_b200 import fip, synth
from evidence_b200.rvmodel import RVModel

# graded work list with setup items and split points (Sm = 1), both ILP builds, conservative build
case = synth.make_case(2, n_epochs=300)
m = RVModel(case.fixedpardict, case.datadict(), case.parnames, device=0)
th = case.draw_theta(600, seed=1)
for var, ilp in ((0, 2), (0, 4), (0, 1), (1, 1)):
    m.set_option("variant", var); m.set_option("ilp", ilp)
    a = m.log_likelihood_batch(th)
m.set_option("variant", 0); m.set_option("ilp", 0)
m.set_priors(case.priordict)
U = case.draw_unit(300, seed=2)
t2, l2 = m.transform_loglike_batch(U)
th_dev = torch.from_numpy(th).cuda()
out = m.log_likelihood_device(th_dev)
print("cfg2-like ok", float(a[0]), float(l2[0]), float(out[0]))
# fused gather on one GPU (two "ranks" = two buffers), completion flags + bounded wait
mine = torch.zeros(2 * 600 + 2, dtype=torch.float64, device="cuda")
other = torch.zeros(2 * 600 + 2, dtype=torch.float64, device="cuda")
other.view(torch.int64)[2 * 600 + 1] = 1  # the "peer" has signalled exchange 1 into OUR buffer? (slot of rank 1)
mine.view(torch.int64)[2 * 600 + 1] = 1
res = np.empty(1200)
m.log_likelihood_gather_host(th, res, [mine.data_ptr(), other.data_ptr()], 0, 1200, 1)
print("gather ok", res[0])
print("true anomaly", m.true_anomaly(np.linspace(0, 50, 100), 0.3)[:2])
m.close()
# two resident epoch ranges (Sm = 2)
case5 = synth.make_case(5)
m5 = RVModel(case5.fixedpardict, case5.datadict(), case5.parnames, device=0)
print("cfg5 ok", m5.log_likelihood_batch(case5.draw_theta(96, seed=3))[:2])
m5.close()
# post-processing kernels
rng = np.random.default_rng(0)
per = np.exp(rng.uniform(0, np.log(1000), (2000, 2)))
nu, fap = fip.fip_periodogram([[None, None, (per, np.ones(2000))]], [-10.0, -5.0, -1.0], Pmin=1.0, Pmax=1000.0,
                              nfreq=2000, Tobs=1000.0, device=0)
print("fip ok", float(fap.min()))
# slice-sampler bookkeeping kernels
from evidence_b200.sampler_dev import nested_sample_device
r = nested_sample_device(lambda X: (-10 + 20 * X, -0.5 * ((-10 + 20 * X) ** 2).sum(1)), 3, nlive=60, nsteps=4,
                         seed=1, device="cuda", dlogz=2.0)
print("slice ok", r.logz)

# run-to-run scatter of ln Z for the two in-repo samplers on the small RV case of tests/test_gpu_runner.py
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from evidence_b200 import synth
from evidence_b200.rvmodel import RVModel
from evidence_b200.sampler import nested_sample
from evidence_b200.sampler_dev import nested_sample_device
case = synth.make_case(1, seed=4, n_epochs=96)
m = RVModel(case.fixedpardict, case.datadict(), case.parnames)
m.set_priors(case.priordict)
nl, ns = int(sys.argv[1]) if len(sys.argv) > 1 else 200, int(sys.argv[2]) if len(sys.argv) > 2 else 12
for name in ("host", "device"):
    vals, errs, ts = [], [], []
    for seed in range(4):
        t0 = time.perf_counter()
        if name == "host":
            r = nested_sample(m.log_likelihood_batch, m.prior_transform_batch, case.ndim, fused=m.transform_loglike_batch, nlive=nl, nsteps=ns, seed=seed)
        else:
            r = nested_sample_device(lambda U: m.transform_loglike_device(U), case.ndim, nlive=nl, nsteps=ns, seed=seed)
        ts.append(time.perf_counter() - t0); vals.append(r.logz); errs.append(r.logzerr)
    print(f"{name:6s} nlive={nl} nsteps={ns}: ln Z {np.round(vals, 2)} mean {np.mean(vals):.2f} std {np.std(vals):.2f} reported {np.mean(errs):.2f}; {np.mean(ts):.1f} s per run")

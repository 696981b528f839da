# latency of small likelihood batches (what a step sampler feeds): device call + sync, and the
# host-buffer calls, for a few work-list settings
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from evidence_b200 import synth
from evidence_b200.rvmodel import RVModel
case = synth.make_case(2)
m = RVModel(case.fixedpardict, case.datadict(), case.parnames)
m.set_priors(case.priordict)
def timeit(fn, n=400):
    for _ in range(30): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e6
for B in (16, 64, 256, 1024):
    th = case.draw_theta(B, seed=1); U = case.draw_unit(B, seed=2)
    thd = torch.from_numpy(th).cuda(); out = torch.empty(B, dtype=torch.float64, device='cuda')
    for opts in ({}, {"max_split": 16}, {"max_split": 16, "phase_items": 400}):
        for k, v in {"max_split": 8, "phase_items": 200, **opts}.items(): m.set_option(k, v)
        m.set_option("timing", 1); m.log_likelihood_device(thd, out=out); torch.cuda.synchronize(); kms = m.last_kernel_ms(); m.set_option("timing", 0)
        def dev(): m.log_likelihood_device(thd, out=out); torch.cuda.synchronize()
        print(f"B={B:5d} {str(opts):45s} kernel {kms*1e3:6.1f} us | device call+sync {timeit(dev):6.1f} | rvl_loglike numpy {timeit(lambda: m.log_likelihood_batch(th)):6.1f} | fused transform_loglike {timeit(lambda: m.transform_loglike_batch(U)):6.1f} us")

# one config-3 sweep launch for ncu: python prof_sweep.py [batch]
import sys, numpy as np, torch
sys.path.insert(0, '.')
from evidence_b200 import synth
from evidence_b200.rvmodel import RVModel
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
case = synth.make_case(cfg)
m = RVModel(case.fixedpardict, case.datadict(), case.parnames)
ilp = int(sys.argv[3]) if len(sys.argv) > 3 else 1
m.set_option("ilp", ilp)
if len(sys.argv) > 4:
    m.set_option("warps", int(sys.argv[4]))
th = torch.from_numpy(case.draw_theta(B, seed=77)).cuda()
out = torch.empty(B, dtype=torch.float64, device='cuda')
for _ in range(3):
    m.log_likelihood_device(th, out=out)
    print(m.last_kernel_ms())
torch.cuda.synchronize()
c = m.counters()
print("cfg", cfg, "B", B, "ilp", ilp, "iters/solve", c["n_newton_iters"]/c["n_solves"], "lnL/s", B/(m.last_kernel_ms()*1e-3))

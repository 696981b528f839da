# one config-3 sweep launch for ncu: python prof_sweep.py [batch]
import sys, numpy as np, torch
sys.path.insert(0, '.')
from evidence_b200 import synth
from evidence_b200.rvmodel import RVModel
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
nep = [int(a.split('=')[1]) for a in sys.argv[4:] if a.startswith('epochs=')]
case = synth.make_case(cfg, n_epochs=nep[0]) if nep else synth.make_case(cfg)
m = RVModel(case.fixedpardict, case.datadict(), case.parnames)
m.set_option("timing", 1)
ilp = int(sys.argv[3]) if len(sys.argv) > 3 else 2
m.set_option("ilp", ilp)
if len(sys.argv) > 4 and "=" not in sys.argv[4]:
    m.set_option("warps", int(sys.argv[4]))
if len(sys.argv) > 5 and "=" not in sys.argv[5]:
    m.set_option("slices", int(sys.argv[5]))
for kv in sys.argv[4:]:
    if "=" in kv and not kv.startswith("epochs="):
        k, v = kv.split("=")
        m.set_option(k, int(v))
th = torch.from_numpy(case.draw_theta(B, seed=77)).cuda()
out = torch.empty(B, dtype=torch.float64, device='cuda')
ms = []
for _ in range(8):
    m.log_likelihood_device(th, out=out)
    ms.append(m.last_kernel_ms())
print(ms)
torch.cuda.synchronize()
c = m.counters()
print("cfg", cfg, "B", B, "ilp", ilp, "iters/solve", c["n_newton_iters"]/c["n_solves"], "best_ms", min(ms), "lnL/s", B/(min(ms)*1e-3), "args", sys.argv[1:])

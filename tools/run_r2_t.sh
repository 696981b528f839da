# round 2, GPU call T (1 GPU): experiment -- U = 6 (384 threads) and U = 8 (256 threads) builds of the likelihood kernel
# against U = 4; those instantiations were NOT kept (100.4 and 183 ms against 92.5), so "ilp" 6 / 8 is refused today
set -x
for ilp in 4 6 8; do
python tools/prof_sweep.py 3 524288 $ilp | tail -1
done
python tools/prof_sweep.py 3 524288 6 warps=10 | tail -1
python tools/prof_sweep.py 3 524288 8 warps=6 | tail -1
python tools/prof_sweep.py 5 131072 4 | tail -1
python tools/prof_sweep.py 5 131072 6 | tail -1
python tools/prof_sweep.py 2 4096 2 | tail -1
python tools/prof_sweep.py 2 4096 4 | tail -1
python - <<'PY'
# parity of the wider builds against the U=4 build
import sys, numpy as np, torch
sys.path.insert(0, '.')
from evidence_b200 import synth
from evidence_b200.rvmodel import RVModel
case = synth.make_case(3)
th = torch.from_numpy(case.draw_theta(20000, seed=5)).cuda()
res = {}
for ilp in (4, 6, 8):
    m = RVModel(case.fixedpardict, case.datadict(), case.parnames)
    m.set_option("ilp", ilp)
    res[ilp] = m.log_likelihood_device(th).cpu().numpy()
for ilp in (6, 8):
    print("ilp", ilp, "max |d lnL| vs ilp 4:", np.max(np.abs(res[ilp] - res[4])))
PY

import sys; sys.path.insert(0,'.')
import numpy as np
from evidence_b200 import synth
from evidence_b200.rvmodel import RVModel
case = synth.make_case(1, seed=4, n_epochs=96)
m = RVModel(case.fixedpardict, case.datadict(), case.parnames)
m.set_priors(case.priordict)
for B in [120, 24, 23, 7, 3, 2, 1, 4096, 300]:
    th = case.draw_theta(B, seed=B)
    try:
        out = m.log_likelihood_batch(th)
        print(B, "ok", out[:2])
    except Exception as e:
        print(B, "FAIL", e)

"""Diagnostic: the bench's high-eccentricity parity set on the device (every kernel build) against
the C port and the staged reference.  usage: python tests/diag/diag_highecc.py [config]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from evidence_b200 import synth
from evidence_b200.layout import compile_model
from evidence_b200.rvmodel import RVModel
from oracle import ref_runner, rv_oracle

case = synth.make_case(int(sys.argv[1]) if len(sys.argv) > 1 else 3)
name, variant, th, bar = bench.parity_sets(case)[3]
t, v, s, ids = case.arrays()
desc, _ = compile_model(case.parnames, case.fixedpardict, case.insts, t[0])
port, it, caps = rv_oracle.c_loglike_batch(bytes(desc), t, v, s, ids, case.n_inst, th)
print("port: iters/solve", it / (len(th) * len(t) * case.n_planets), "cap hits", caps)
ref = None
if ref_runner.available():
    m = ref_runner.make_model(case.fixedpardict, case.datadict(), case.parnames)
    ref = ref_runner.loglike_rows(m, th[:64])
    print("reference vs port (64 rows): max", np.abs(ref - port[:64]).max())
cols = [case.parnames.index(f"planet{k}_ecc") for k in range(1, case.n_planets + 1)]
model = RVModel(case.fixedpardict, case.datadict(), case.parnames, device=0)
for var, ilp in ((0, 2), (0, 1), (0, 3), (0, 4), (1, 1)):
    model.set_option("variant", var)
    model.set_option("ilp", ilp)
    model.reset_counters()
    got = model.log_likelihood_batch(th)
    one = np.array([model.log_likelihood(x) for x in th[:32]])
    c = model.counters()
    err = np.abs(got - port)
    o = np.argsort(-err)[:4]
    print(f"variant {var} ilp {ilp}: max {err.max():.3e}  n>1e-5: {(err > 1e-5).sum()}  cap hits {c['n_cap_hits']}  "
          f"batch-vs-scalar max {np.abs(one - got[:32]).max():.3e}  worst rows "
          + "; ".join(f"{i}: e={np.round(th[i, cols], 4).tolist()} err={err[i]:.2e}" for i in o))

"""
TEST INFRASTRUCTURE: the CPU reference ln Z of the evidence ladder (BASELINE.json configs[3]).

Runs the SAME seeded sampler that examples/evidence_ladder.py drives with the device likelihood
(evidence_b200.sampler.nested_sample, seed 100 + k, same data set, same model, same priors) on the
CPU checker -- the C restatement of the reference path, oracle/rvlnl_oracle.c, spread over the host
cores -- and prints one JSON line per k.  The sampler is deterministic for a seed and the two
likelihoods agree to ~1e-11, so the chains are identical unless an accept/reject decision falls
inside that margin (profiles/r2_evidence_ladder.txt: identical ln Z and call counts for k = 0, 1, 2).

    python tests/ladder_cpu.py --kmax 2 --epochs 300 --nlive 200
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))


def _cpu_block(args):
    """One block of rows through the C restatement of the reference path (worker process)."""
    from oracle import rv_oracle
    desc, t, v, s, ids, n_inst, block = args
    return rv_oracle.c_loglike_batch(desc, t, v, s, ids, n_inst, block)[0]


def cpu_ladder(kmax, epochs=300, nlive=200, true_planets=2, nlive_per_dim=0, cores=None, ks=None):
    """[{k, ndim, cpu_logz, cpu_logzerr, cpu_ncall, cpu_seconds, cpu_cores}] for k = 0..kmax."""
    import multiprocessing as mp
    from evidence_ladder import ladder_model  # the example's own model builder
    from evidence_b200 import priors, synth
    from evidence_b200.layout import compile_model
    from evidence_b200.sampler import nested_sample
    from oracle import rv_oracle
    rv_oracle.build()
    cores = cores or os.cpu_count() or 1
    data = synth.make_case(2, seed=11, n_epochs=epochs, n_planets=true_planets)
    t, v, s, ids = data.arrays()
    out = []
    with mp.get_context("fork").Pool(cores) as pool:
        for k in (ks if ks is not None else range(kmax + 1)):
            spec, fixed = ladder_model(data, k, true_planets)
            parnames = sorted(spec)
            pri = {p: priors.make_prior(*spec[p]) for p in parnames}
            desc = bytes(compile_model(parnames, fixed, data.insts, t[0])[0])

            def loglike(theta):
                theta = np.ascontiguousarray(theta)
                if len(theta) < 4 * cores:
                    return _cpu_block((desc, t, v, s, ids, data.n_inst, theta))
                return np.concatenate(pool.map(_cpu_block, [(desc, t, v, s, ids, data.n_inst, b)
                                                            for b in np.array_split(theta, cores)]))

            def transform(u):
                return np.column_stack([pri[p].ppf(u[:, i]) for i, p in enumerate(parnames)])
            n = nlive_per_dim * len(parnames) if nlive_per_dim else nlive
            t0 = time.perf_counter()
            res = nested_sample(loglike, transform, len(parnames), nlive=n, seed=100 + k)
            out.append({"k": k, "ndim": len(parnames), "nlive": n, "cpu_logz": res.logz,
                        "cpu_logzerr": res.logzerr, "cpu_ncall": res.ncall,
                        "cpu_seconds": time.perf_counter() - t0, "cpu_cores": cores})
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--kmax", type=int, default=2)
    ap.add_argument("--epochs", type=int, default=300)
    ap.add_argument("--nlive", type=int, default=200)
    ap.add_argument("--nlive-per-dim", type=int, default=0)
    ap.add_argument("--true-planets", type=int, default=2)
    a = ap.parse_args()
    for k in range(a.kmax + 1):
        for rec in cpu_ladder(k, a.epochs, a.nlive, a.true_planets, a.nlive_per_dim, ks=[k]):
            print(json.dumps(rec), flush=True)

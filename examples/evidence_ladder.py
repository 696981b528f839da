"""
Evidence ladder (BASELINE.json config 4): ln Z for k = 0..kmax planets on one data set, one
independent run per GPU -- replicas, no collective.

    python examples/evidence_ladder.py --kmax 3                      # one GPU, the k's in sequence
    torchrun --nproc-per-node 6 examples/evidence_ladder.py --kmax 5 # rank r takes k = r, r+6, ...


Each line of output is one JSON record {k, ndim, nlive, logz, logzerr, ncall, seconds, device}.  The
CPU reference ln Z that BASELINE.json's configs[3] asks for beside it comes from
tests/ladder_cpu.py (test infrastructure: the SAME seeded sampler on the CPU checker; the chains
are identical bit for bit, profiles/r2_evidence_ladder.txt).  On one GPU the
ladder ends with what the reference's fip_criterion.py does with such runs: p(k|y) from the
evidences and the FIP periodogram of the posterior periods, accumulated on the device
(evidence_b200.fip), and prints the periods where the false inclusion probability is lowest.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import numpy as np  # noqa: E402

from evidence_b200 import fip, priors, synth  # noqa: E402
from evidence_b200.rvmodel import RVModel  # noqa: E402
from evidence_b200.sampler import nested_sample  # noqa: E402

def ladder_model(data, k, true_planets):
    """(prior spec, fixed parameters) of the k-planet model on the ladder's data set."""
    spec = {p: v for p, v in data.prior_spec.items()
            if not p.startswith("planet") or int(p[6:p.index("_")]) <= k}
    for j in range(true_planets + 1, k + 1):  # more planets than the data were made with
        for nm in ("k1", "period", "ecc", "omega", "ma0"):
            spec[f"planet{j}_{nm}"] = data.prior_spec[f"planet1_{nm}"]
    fixed = {f"planet{j}_epoch": synth.EPOCH for j in range(1, k + 1)}
    fixed["drift_tref"] = synth.EPOCH
    return spec, fixed


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kmax", type=int, default=3)
    ap.add_argument("--epochs", type=int, default=300)
    ap.add_argument("--nlive", type=int, default=200)
    ap.add_argument("--true-planets", type=int, default=2)
    ap.add_argument("--nlive-per-dim", type=int, default=0,
                    help="nlive = this x ndim (the reference's default is 25, evidence/ultranest/__init__.py:333)")
    ap.add_argument("--sampler", default="host", choices=["host", "device"],
                    help="host: numpy bookkeeping (evidence_b200.sampler); device: the device-resident "
                         "sampler with the native bookkeeping kernels (evidence_b200.sampler_dev)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    data = synth.make_case(2, seed=11, n_epochs=args.epochs, n_planets=args.true_planets)
    runs, logzs = [None] * (args.kmax + 1), [None] * (args.kmax + 1)
    for k in range(rank, args.kmax + 1, world):
        spec, fixed = ladder_model(data, k, args.true_planets)
        pri = {p: priors.make_prior(*v) for p, v in spec.items()}
        model = RVModel(fixed, data.datadict(), list(spec), device=dev)
        model.set_priors(pri)
        nlive = args.nlive_per_dim * model.ndim if args.nlive_per_dim else args.nlive
        t0 = time.perf_counter()
        if args.sampler == "device":
            import torch
            from evidence_b200.sampler_dev import nested_sample_device
            torch.cuda.set_device(dev)
            res = nested_sample_device(model.transform_loglike_device, model.ndim, nlive=nlive,
                                       seed=100 + k, device=f"cuda:{dev}")
        else:
            res = nested_sample(model.log_likelihood_batch, model.prior_transform_batch, model.ndim,
                                nlive=nlive, seed=100 + k, fused=model.transform_loglike_batch)
        cols = [model.parnames.index(f"planet{j}_period") for j in range(1, k + 1)]
        runs[k] = (res.weighted_samples[:, cols], res.weights) if k else None
        logzs[k] = res.logz
        rec = {"k": k, "ndim": model.ndim, "nlive": nlive, "sampler": res.method, "logz": res.logz,
               "logzerr": res.logzerr, "ncall": res.ncall, "seconds": time.perf_counter() - t0, "device": dev}
        print(json.dumps(rec), flush=True)
        model.close()
    if world == 1 and args.kmax >= 1:
        t, _, _, _ = data.arrays()
        nu, fapnu = fip.fip_periodogram([runs], logzs, Pmin=1.0, Pmax=1000.0, nfreq=50000,
                                        Tobs=float(t.max() - t.min()), device=dev)
        pky = fip.posterior_of_k(logzs)
        best = np.argsort(fapnu[0])[:2000]
        peaks = []
        for j in best:  # lowest FIP first, one entry per well-separated period, FIP < 1/2 only
            if fapnu[0, j] >= 0.5:
                break
            if all(abs(np.log(nu[j] / nu[q])) > 0.05 for q in peaks):
                peaks.append(j)
            if len(peaks) == args.true_planets + 1:
                break
        print(json.dumps({"p(k|y)": [float(x) for x in pky],
                          "fip_minima": [{"period": float(2 * np.pi / nu[j]), "fip": float(max(fapnu[0, j], 1e-15))}
                                         for j in peaks],
                          "true_periods": [data.truth[f"planet{j}_period"]
                                           for j in range(1, args.true_planets + 1)]}), flush=True)


if __name__ == "__main__":
    main()

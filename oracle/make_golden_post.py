"""
make_golden_post.py — TEST INFRASTRUCTURE ONLY.

Golden vectors for the two post-processing loops produced by the REFERENCE'S
OWN STATEMENTS.  Neither script can be imported (fip_criterion.py executes at import and reads run
directories; post_processing.py needs matplotlib / corner), so the loop nests are taken from the
source files where they lie under /root/reference -- by line range, at generation time, nothing is
copied into the repository -- dedented and executed on seeded inputs with the names they expect:

  fip_ref.npz     evidence/fip_criterion.py:305-338   fapnu[run, nfreq]  (with_alias = False; the
                  alias branch of the reference raises NameError at :321 -- x_freqs is never
                  allocated -- so it has no executable reference behaviour to pin)
  order_ref.npz   evidence/post_processing.py:93-128  the ordered posterior samples

Run in the build container:   python oracle/make_golden_post.py
"""
import json
import os
import sys
import textwrap
import types

import numpy as np
import pandas as pd

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def ref_lines(path, first, last):
    """Source lines [first, last] (1-based, inclusive) of a reference file, dedented."""
    with open(os.path.join(REF, path)) as f:
        lines = f.readlines()[first - 1:last]
    return textwrap.dedent("".join(lines))


def save(name, meta, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, meta=np.array(json.dumps(meta)), **arrays)
    print(f"{name:18s} {os.path.getsize(path) / 1024:8.1f} KiB")


def fip_vectors():
    from oracle import fip_oracle as fo
    rng = np.random.default_rng(17)
    Pmin, Pmax, nfreq, Tobs = 1.0, 1000.0, 4000, 400.0
    nu, nua, nub = fo.frequency_grid(Pmin, Pmax, nfreq, Tobs)
    logZs = np.array([-100.0, -90.0, -88.0, -89.5])
    pky = fo.posterior_of_k(logZs)
    nmod, min_iters = len(logZs), 2
    centres = np.array([3.1, 42.0, 290.0])
    posteriors, flat = [], {}
    for r in range(min_iters):
        models = [{}]
        for k in range(1, nmod):
            n = 300
            per = np.exp(rng.normal(np.log(centres[:k]), 0.01, (n, k)))
            mask = rng.random((n, k)) < 0.15
            per[mask] = rng.uniform(0.3, 2500.0, mask.sum())
            w = rng.random(n) + 1e-3
            models.append({"samples": per, "weights": w.copy()})
            flat[f"samples_r{r}_k{k}"] = per
            flat[f"weights_r{r}_k{k}"] = w
        posteriors.append(models)
    src = ref_lines("evidence/fip_criterion.py", 305, 338)
    ns = {"np": np, "min_iters": min_iters, "nmod": nmod, "nfreq": nfreq, "posteriors": posteriors,
          "pky": pky, "nua": nua, "nub": nub, "Pmin": Pmin, "Pmax": Pmax,
          "args": types.SimpleNamespace(with_alias=False, recalculate_fip=True)}
    exec(compile(src, "fip_criterion.py:305-338", "exec"), ns)
    save("fip_ref", {"Pmin": Pmin, "Pmax": Pmax, "nfreq": nfreq, "Tobs": Tobs, "logZs": logZs.tolist(),
                     "n_runs": min_iters, "nmod": nmod, "source": "evidence/fip_criterion.py:305-338"},
         fapnu=ns["fapnu"], **flat)


def order_vectors():
    rng = np.random.default_rng(23)
    cases = {}
    meta = {"source": "evidence/post_processing.py:93-128", "cases": []}
    src = ref_lines("evidence/post_processing.py", 93, 128)
    for K in (1, 2, 3, 4):
        names = ["inst_jitter", "inst_offset", "drift_lin"]
        for p in range(1, K + 1):
            names += [f"planet{p}_{q}" for q in ("ecc", "k1", "ma0", "omega", "period")]
        names = sorted(names)
        n = 400
        s = rng.normal(size=(n, len(names)))
        for p in range(1, K + 1):
            s[:, names.index(f"planet{p}_period")] = np.exp(rng.uniform(0, 6, n))
        s[3, names.index("planet1_period")] = np.nan
        frame = pd.DataFrame(s.copy(), columns=names)
        ns = {"np": np, "order": True, "nplanets": K, "parnames": names, "samples": frame}
        exec(compile(src, "post_processing.py:93-128", "exec"), ns)
        cases[f"in_K{K}"] = s
        cases[f"out_K{K}"] = ns["samples"].values
        meta["cases"].append({"K": K, "parnames": names})
    save("order_ref", meta, **cases)


if __name__ == "__main__":
    fip_vectors()
    order_vectors()

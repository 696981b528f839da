"""
TEST DOUBLE of the third-party ``pypolychord`` package (absent from this image and from the
reference checkout: SURVEY.md 8c).  NOT PolyChord's code: a small scalar nested sampler behind
PolyChord's Python interface as the reference uses it (evidence/polychord/__init__.py:14-16, 190,
201-203, 421): ``run_polychord(loglikelihood, nDims, nDerived, settings, prior)`` calls
``prior(cube[nDims]) -> theta[nDims]`` and ``loglikelihood(theta) -> (lnL, derived_list)`` ONE point
at a time (PolyChord drives them from Fortran), and returns an output object with ``logZ``,
``logZerr``, ``base_dir``, ``file_root``, ``samples`` (DataFrame: weight, loglike, p0..) and
``make_paramnames_files``.  Every call is counted in ``output.nlike``.
"""
import os

import numpy as np

from . import priors, settings  # noqa: F401

__version__ = "0.0-test-double"


class PolyChordOutput:
    def __init__(self, base_dir, file_root):
        self.base_dir, self.file_root = base_dir, file_root

    def make_paramnames_files(self, paramnames):
        os.makedirs(self.base_dir, exist_ok=True)
        with open(os.path.join(self.base_dir, self.file_root + ".paramnames"), "w") as f:
            for name, latex in paramnames:
                f.write(f"{name}   {latex}\n")


def run_polychord(loglikelihood, nDims, nDerived, settings, prior=None, dumper=None):
    rng = np.random.default_rng(getattr(settings, "seed", -1) if getattr(settings, "seed", -1) >= 0 else 0)
    n, reps = int(settings.nlive), int(settings.num_repeats)
    ncall = [0]

    def evaluate(u):
        theta = np.asarray(prior(u), dtype=np.float64) if prior is not None else u
        if theta.shape != (nDims,):
            raise ValueError(f"prior returned shape {theta.shape}, expected {(nDims,)}")
        out = loglikelihood(theta)
        if not (isinstance(out, tuple) and len(out) == 2):
            raise ValueError("loglikelihood must return (logL, derived_list)")
        lnl, derived = out
        if len(derived) != nDerived:
            raise ValueError("wrong number of derived parameters")
        ncall[0] += 1
        return theta, float(lnl)

    us = rng.random((n, nDims))
    pts = [evaluate(u) for u in us]
    ps, ls = np.array([p for p, _ in pts]), np.array([l for _, l in pts])
    logz, it = -np.inf, 0
    dead_p, dead_l, dead_w = [], [], []
    while True:
        worst = int(np.argmin(ls))
        lmin = ls[worst]
        logw = -it / n + np.log1p(-np.exp(-1.0 / n))
        dead_p.append(ps[worst].copy()); dead_l.append(lmin); dead_w.append(logw)
        logz = np.logaddexp(logz, lmin + logw)
        it += 1
        if np.max(ls) - it / n < logz + np.log(settings.precision_criterion):
            break
        # slice sampling from a random other live point, num_repeats moves along whitened directions
        ctr = us.mean(axis=0)
        cov = np.cov((us - ctr).T).reshape(nDims, nDims) + 1e-14 * np.eye(nDims)
        chol = np.linalg.cholesky(cov)
        start = int(rng.choice([i for i in range(n) if i != worst and ls[i] > lmin] or [int(np.argmax(ls))]))
        u, p, l = us[start].copy(), ps[start].copy(), ls[start]
        for _ in range(reps):
            d = chol @ rng.standard_normal(nDims)
            d /= np.linalg.norm(np.linalg.solve(chol, d)) + 1e-300
            lo = -rng.random() * 2.0
            hi = lo + 2.0
            for _ in range(50):  # stepping out
                c = u + lo * d
                if np.any((c <= 0) | (c >= 1)) or evaluate(c)[1] <= lmin:
                    break
                lo *= 2.0
            for _ in range(50):
                c = u + hi * d
                if np.any((c <= 0) | (c >= 1)) or evaluate(c)[1] <= lmin:
                    break
                hi *= 2.0
            for _ in range(100):  # shrinkage
                t = lo + (hi - lo) * rng.random()
                c = u + t * d
                if np.all((c > 0) & (c < 1)):
                    pc, lc = evaluate(c)
                    if lc > lmin:
                        u, p, l = c, pc, lc
                        break
                if t < 0:
                    lo = t
                else:
                    hi = t
        us[worst], ps[worst], ls[worst] = u, p, l
    logw_live = -it / n - np.log(n)
    for k in np.argsort(ls):
        dead_p.append(ps[k].copy()); dead_l.append(ls[k]); dead_w.append(logw_live)
        logz = np.logaddexp(logz, ls[k] + logw_live)
    dead_l, dead_w = np.array(dead_l), np.array(dead_w)
    w = np.exp(dead_l + dead_w - logz)
    info = float(np.sum(w / w.sum() * (dead_l - logz)))
    out = PolyChordOutput(settings.base_dir, settings.file_root)
    out.logZ, out.logZerr = float(logz), float(np.sqrt(max(info, 0.0) / n))
    out.nlike, out.ndead, out.nlive = ncall[0], it, n
    import pandas as pd
    frame = {"weight": w / w.sum(), "loglike": dead_l}
    frame.update({f"p{i}": np.array(dead_p)[:, i] for i in range(nDims)})
    out.samples = pd.DataFrame(frame)
    os.makedirs(settings.base_dir, exist_ok=True)
    np.savetxt(os.path.join(settings.base_dir, settings.file_root + ".txt"),
               np.column_stack([w / w.sum(), -2 * dead_l, np.array(dead_p)]))
    with open(os.path.join(settings.base_dir, settings.file_root + ".stats"), "w") as f:
        f.write(f"log(Z) = {out.logZ} +/- {out.logZerr}\n")
    return out

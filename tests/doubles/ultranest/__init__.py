"""
TEST DOUBLE of the third-party ``ultranest`` package (absent from this image and from the
reference checkout: SURVEY.md 8c).  NOT UltraNest's code: a small nested sampler that reproduces
UltraNest's *interface and calling convention* as the reference uses it
(evidence/ultranest/__init__.py:165-185, 197, 217-221, 235), so that the ``which == "ultranest"``
branch of ``evidence_b200.ultranest.run`` executes against something that behaves like the real
package at the boundary:

* ``ReactiveNestedSampler(param_names, loglike, transform, log_dir=, num_test_samples=,
  wrapped_params=, num_bootstraps=, vectorized=, ndraw_min=, ndraw_max=, resume=, ...)``;
  with ``vectorized=True`` the callbacks receive 2-D arrays ``u[n, ndim]`` / ``theta[n, ndim]``
  and must return ``[n, ndim]`` / ``[n]``; the constructor tries them on ``num_test_samples``
  random points and raises on a wrong shape or a non-finite value, like the real one;
* ``sampler.stepsampler`` (``ultranest.stepsampler.RegionSliceSampler`` -> one point per
  likelihood call, ``ultranest.popstepsampler.PopulationSliceSampler`` -> ``popsize`` points per
  call); without a step sampler: region (bounding-ellipsoid) rejection sampling, ``ndraw_min`` ..
  ``ndraw_max`` candidates per call, doubling while the acceptance rate is low;
* ``sampler.run(min_num_live_points=, cluster_num_live_points=, dlogz=, frac_remain=, ...)``,
  ``sampler.results`` with ``logz, logzerr, ncall, niter, samples, weighted_samples, paramnames``,
  ``sampler.print_results()``, ``sampler.plot()``;
* ``log_dir/run1/chains/weighted_post.txt`` (``weight logl <params>``, space separated: what the
  reference's post-processing reads, evidence/post_processing.py:85-88) and
  ``equal_weighted_post.txt``, ``info/results.json``.

Every batch size the callbacks saw is recorded in ``sampler.call_sizes`` for the tests.
"""
import json
import os

import numpy as np

__version__ = "0.0-test-double"


def _logaddexp(a, b):
    return np.logaddexp(a, b)


class ReactiveNestedSampler:
    def __init__(self, param_names, loglike, transform=None, derived_param_names=[],
                 wrapped_params=None, resume="subfolder", run_num=None, log_dir=None,
                 num_test_samples=2, draw_multiple=True, num_bootstraps=30, vectorized=False,
                 ndraw_min=128, ndraw_max=65536, storage_backend="hdf5", warmstart_max_tau=-1):
        self.paramnames = list(param_names)
        self.x_dim = len(self.paramnames)
        self.vectorized = bool(vectorized)
        self.ndraw_min, self.ndraw_max = int(ndraw_min), int(ndraw_max)
        self.num_bootstraps = int(num_bootstraps)
        self.wrapped_axes = (np.nonzero(np.asarray(wrapped_params))[0]
                             if wrapped_params is not None else np.array([], dtype=int))
        self.stepsampler = None
        self.call_sizes = []
        self.ncall = 0
        self.results = None
        if self.vectorized:
            self._loglike, self._transform = loglike, transform
        else:  # scalar callbacks, the reference's own mode (:125-146)
            self._loglike = lambda th: np.array([loglike(row) for row in th])
            self._transform = (lambda u: np.array([transform(row) for row in u])) if transform else None
        self.logs = None
        if log_dir is not None:
            run_dir = os.path.join(log_dir, "run1" if run_num is None else f"run{run_num}")
            self.logs = {k: os.path.join(run_dir, k) for k in ("chains", "info", "results", "plots", "extra")}
            for d in self.logs.values():
                os.makedirs(d, exist_ok=True)
            self.logs["run_dir"] = run_dir
        # the real constructor tries the functions on a few random points
        rng = np.random.default_rng(1)
        u = rng.random((max(1, int(num_test_samples)), self.x_dim))
        p, logl = self._eval(u, count=False)
        if not np.all(np.isfinite(p)):
            raise ValueError("transform returned non-finite values for a test point")
        if not np.all(np.isfinite(logl)) and not np.all(logl > -1e300):
            raise ValueError("loglike returned non-finite values for a test point")

    # ---- the boundary: every likelihood evaluation goes through here -------------------------
    def _eval(self, u, count=True):
        u = np.ascontiguousarray(u, dtype=np.float64)
        n = len(u)
        p = np.asarray(self._transform(u)) if self._transform is not None else u
        if p.shape != (n, self.x_dim):
            raise ValueError(f"Error in transform function: returned shape {p.shape}, expected {(n, self.x_dim)}")
        logl = np.asarray(self._loglike(p))
        if logl.shape != (n,):
            raise ValueError(f"Error in loglikelihood function: returned shape {logl.shape}, expected {(n,)}")
        if count:
            self.call_sizes.append(n)
            self.ncall += n
        return p, logl

    # ---- replacement schemes --------------------------------------------------------------------
    def _whiten(self, us):
        ctr = us.mean(axis=0)
        cov = np.cov((us - ctr).T).reshape(self.x_dim, self.x_dim) + 1e-14 * np.eye(self.x_dim)
        return ctr, np.linalg.cholesky(cov)

    def _region_batch(self, rng, us, ndraw):
        ctr, L = self._whiten(us)
        y = np.linalg.solve(L, (us - ctr).T)
        r = np.sqrt(np.max(np.sum(y * y, axis=0))) * 1.3
        z = rng.standard_normal((ndraw, self.x_dim))
        z *= (rng.random(ndraw) ** (1.0 / self.x_dim) / np.linalg.norm(z, axis=1))[:, None]
        u = ctr + (z * r) @ L.T
        if len(self.wrapped_axes):
            u[:, self.wrapped_axes] %= 1.0
        return u[np.all((u > 0.0) & (u < 1.0), axis=1)]

    def _slice_population(self, rng, us, Ls, Lmin, popsize, nsteps):
        """`popsize` walkers from live points above Lmin take nsteps slice moves in lock-step."""
        ctr, Lc = self._whiten(us)
        start = rng.choice(np.nonzero(Ls > Lmin)[0], popsize)
        u, p, logl = us[start].copy(), None, Ls[start].copy()
        for _ in range(nsteps):
            d = rng.standard_normal((popsize, self.x_dim)) @ Lc.T
            d /= np.linalg.norm(np.linalg.solve(Lc, d.T), axis=0)[:, None] + 1e-300
            lo = -rng.random(popsize)
            hi = lo + 1.0
            for side in (-1, +1):  # stepping out
                active = np.ones(popsize, dtype=bool)
                for _ in range(20):
                    edge = lo if side < 0 else hi
                    cand = u + edge[:, None] * d
                    inside = np.all((cand > 0) & (cand < 1), axis=1) & active
                    if not inside.any():
                        break
                    _, l_c = self._eval(cand[inside])
                    grow = np.zeros(popsize, dtype=bool)
                    grow[np.nonzero(inside)[0][l_c > Lmin]] = True
                    active &= grow
                    if side < 0:
                        lo[grow] *= 2.0
                    else:
                        hi[grow] *= 2.0
            todo = np.ones(popsize, dtype=bool)
            for _ in range(60):  # shrinkage
                if not todo.any():
                    break
                t = lo + (hi - lo) * rng.random(popsize)
                cand = u + t[:, None] * d
                inside = np.all((cand > 0) & (cand < 1), axis=1)
                ev = todo & inside
                l_c = np.full(popsize, -np.inf)
                if ev.any():
                    _, l_c[ev] = self._eval(cand[ev])
                ok = ev & (l_c > Lmin)
                u[ok], logl[ok] = cand[ok], l_c[ok]
                todo &= ~ok
                shrink = todo
                lo[shrink & (t < 0)] = t[shrink & (t < 0)]
                hi[shrink & (t >= 0)] = t[shrink & (t >= 0)]
        p, logl = self._eval(u)  # final positions (physical parameters for the live set)
        return u, p, logl

    # ---- the run ------------------------------------------------------------------------------
    def run(self, update_interval_volume_fraction=0.8, update_interval_ncall=None, log_interval=None,
            show_status=True, viz_callback="auto", dlogz=0.5, dKL=0.5, frac_remain=0.01, Lepsilon=0.001,
            min_ess=400, max_iters=None, max_ncalls=None, max_num_improvement_loops=-1,
            min_num_live_points=400, cluster_num_live_points=40, insertion_test_window=10,
            insertion_test_zscore_threshold=4, region_class=None, widen_before_initial_plateau_num_warn=10000,
            widen_before_initial_plateau_num_max=50000, seed=0):
        rng = np.random.default_rng(seed)
        n = int(min_num_live_points)
        us = rng.random((max(n, 1), self.x_dim))
        ps, Ls = self._eval(us)
        logz, h_num = -np.inf, 0.0
        dead_p, dead_l, dead_logw = [], [], []
        pending_u, pending_p, pending_l = np.zeros((0, self.x_dim)), np.zeros((0, self.x_dim)), np.zeros(0)
        ndraw, it = self.ndraw_min, 0
        while True:
            worst = int(np.argmin(Ls))
            Lmin = Ls[worst]
            logw = -it / n + np.log1p(-np.exp(-1.0 / n))
            dead_p.append(ps[worst].copy()); dead_l.append(Lmin); dead_logw.append(logw)
            logz = _logaddexp(logz, Lmin + logw)
            it += 1
            remain = np.max(Ls) - it / n
            if (remain < logz + np.log(frac_remain) or _logaddexp(logz, remain) - logz < dlogz * 0.01
                    or (max_iters and it >= max_iters) or (max_ncalls and self.ncall >= max_ncalls)):
                break
            # replacement
            keep = pending_l > Lmin
            pending_u, pending_p, pending_l = pending_u[keep], pending_p[keep], pending_l[keep]
            while len(pending_l) == 0:
                if self.stepsampler is not None:
                    pop = getattr(self.stepsampler, "popsize", 1)
                    u_new, p_new, l_new = self._slice_population(rng, us, Ls, Lmin, pop, self.stepsampler.nsteps)
                else:
                    u_new = self._region_batch(rng, us, ndraw)
                    p_new, l_new = self._eval(u_new) if len(u_new) else (u_new, np.zeros(0))
                    if len(u_new) and np.mean(l_new > Lmin) < 0.05:
                        ndraw = min(self.ndraw_max, ndraw * 2)
                ok = l_new > Lmin
                pending_u, pending_p, pending_l = u_new[ok], p_new[ok], l_new[ok]
            us[worst], ps[worst], Ls[worst] = pending_u[0], pending_p[0], pending_l[0]
            pending_u, pending_p, pending_l = pending_u[1:], pending_p[1:], pending_l[1:]
        # remaining live points
        logw_live = -it / n - np.log(n)
        for k in np.argsort(Ls):
            dead_p.append(ps[k].copy()); dead_l.append(Ls[k]); dead_logw.append(logw_live)
            logz = _logaddexp(logz, Ls[k] + logw_live)
        dead_l, dead_logw, pts = np.array(dead_l), np.array(dead_logw), np.array(dead_p)
        w = np.exp(dead_l + dead_logw - logz)
        w /= w.sum()
        info = float(np.sum(w * (dead_l - logz)))
        # bootstrap over the dead points (UltraNest's num_bootstraps): scatter of ln Z
        boots = []
        for _ in range(max(1, self.num_bootstraps)):
            g = rng.gamma(1.0, size=len(w))
            boots.append(np.log(np.sum(np.exp(dead_l + dead_logw - logz) * g / g.mean())))
        logzerr = float(np.hypot(np.std(boots), np.sqrt(max(info, 0.0) / n)))
        nsamp = max(1, int(1.0 / np.sum(w ** 2)))
        idx = np.minimum(np.searchsorted(np.cumsum(w), (rng.random() + np.arange(nsamp)) / nsamp), len(w) - 1)
        self.results = dict(niter=it, logz=float(logz), logzerr=logzerr, logz_bs=float(logz),
                            logzerr_bs=float(np.std(boots)), logzerr_tail=0.0, ess=float(nsamp),
                            H=info, Herr=0.0, ncall=int(self.ncall), paramnames=list(self.paramnames),
                            samples=pts[idx],
                            weighted_samples=dict(points=pts, weights=w, logw=dead_logw, logl=dead_l,
                                                  upoints=None, bootstrapped_weights=None),
                            posterior=dict(mean=list(np.average(pts, weights=w, axis=0)),
                                           stdev=list(np.sqrt(np.average((pts - np.average(pts, weights=w, axis=0)) ** 2,
                                                                         weights=w, axis=0)))),
                            insertion_order_MWW_test=dict(independent_iterations=float("inf"), converged=True))
        if self.logs is not None:
            hdr = " ".join(["weight", "logl"] + self.paramnames)
            np.savetxt(os.path.join(self.logs["chains"], "weighted_post.txt"),
                       np.column_stack([w, dead_l, pts]), header=hdr, comments="")
            np.savetxt(os.path.join(self.logs["chains"], "equal_weighted_post.txt"), pts[idx],
                       header=" ".join(self.paramnames), comments="")
            with open(os.path.join(self.logs["info"], "results.json"), "w") as f:
                json.dump({k: v for k, v in self.results.items()
                           if k in ("niter", "logz", "logzerr", "ncall", "paramnames", "ess", "H")}, f)
        return self.results

    def print_results(self, use_unicode=True):
        r = self.results
        print(f"\nlogZ = {r['logz']:.3f} +- {r['logzerr']:.3f}   ({r['ncall']} likelihood calls, test double)")
        for name, m, s in zip(r["paramnames"], r["posterior"]["mean"], r["posterior"]["stdev"]):
            print(f"    {name:24s}: {m:.6g} +- {s:.3g}")

    def plot(self):
        pass  # (matplotlib / corner are absent from the image)

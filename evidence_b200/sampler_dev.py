"""
The slice-sampling nested sampler of ``evidence_b200.sampler`` with its bookkeeping on the device.

``sampler.nested_sample`` keeps the live points, brackets and masks in numpy and crosses the
host/device boundary for every likelihood batch; a run is then bound by that bookkeeping
(0.3 - 1.4 M lnL/s against 30 M lnL/s of the kernel).  Here every array of
the run -- live points, walkers, brackets, candidates, evidence terms -- is a torch tensor on the
likelihood's device and the likelihood is called through the device entry point
(``RVModel.transform_loglike_device``: u -> theta -> lnL without leaving the GPU).  The host only
sequences the launches and reads back two scalars per round (loop exits).

Same algorithm (the scheme of PolyChord / UltraNest's step samplers that the
reference configures at evidence/ultranest/__init__.py:175): per round the k worst live points are
retired with the shrinkage 1/n for n = nlive, nlive-1, ..; k walkers started from surviving live
points take ``nsteps`` slice moves along whitened random directions under L > L*_k, with the next
m stepping-out positions and shrinkage candidates of every walker evaluated speculatively in one
launch (see ``sampler._slice_moves``).  Differences from the numpy sampler: the random numbers come
from torch's generator (so the two are statistically, not bit-wise, equivalent), and candidates
outside the unit cube are clamped before they are evaluated and rejected afterwards (no compaction
of the batch, hence no synchronisation to learn its size).

There is no CPU fallback in the product: ``fused`` must be a device callable.  (The function itself
is device-agnostic torch code, which is how the CPU tests drive it with an analytic likelihood.)
"""
import math

import numpy as np

from .sampler import NestedResult, lineage_bootstrap, retire_groups


def _whitening(torch, u):
    d = u.shape[1]
    cov = torch.cov(u.T).reshape(d, d) + torch.eye(d, dtype=u.dtype, device=u.device) * 1e-18
    L, info = torch.linalg.cholesky_ex(cov)
    diag = torch.diag(torch.sqrt(torch.clamp(torch.diagonal(cov), min=1e-30)))
    return torch.where(info.reshape(1, 1) == 0, L, diag)


def _slice_moves(torch, gen, fused, u, theta, lmin, chol, nsteps, m, max_expand, max_shrink,
                 m_out=None):
    """k walkers, nsteps slice moves each, everything on u.device.  Returns (u, theta, lnL, ncall);
    lnL is NaN for a walker that never moved."""
    k, d = u.shape
    dev, f64 = u.device, u.dtype
    lcur = torch.full((k,), float("nan"), dtype=f64, device=dev)
    rows = torch.arange(k, device=dev)
    # stepping out has no sequential dependence at all (the positions are lo - j, hi + j): with
    # room in the launch every one of the max_expand positions of both sides is evaluated at once
    # and the phase needs ONE launch and no read-back
    mo = m if m_out is None else max(1, min(int(m_out), max_expand))
    steps = torch.arange(mo, device=dev, dtype=f64)
    isteps = torch.arange(mo, device=dev)
    hi_clamp = 1.0 - 2.0 ** -53
    ncall = 0

    def evaluate(points):
        """lnL (and theta) of points[k, n, d]; -inf outside the unit cube."""
        nonlocal ncall
        n = points.shape[1]
        inside = ((points >= 0.0) & (points < 1.0)).all(dim=-1)
        th, ll = fused(points.clamp(0.0, hi_clamp).reshape(k * n, d).contiguous())
        ncall += k * n
        ll = torch.where(inside, ll.reshape(k, n), torch.full_like(inside, -math.inf, dtype=f64))
        return ll, th.reshape(k, n, d)

    for _ in range(nsteps):
        z = torch.randn((k, d), generator=gen, dtype=f64, device=dev)
        z = z / torch.linalg.norm(z, dim=1, keepdim=True)
        dirn = z @ chol.T
        r = torch.rand((k,), generator=gen, dtype=f64, device=dev)
        lo, hi = -r, 1.0 - r
        # ---- stepping out: the next m unit steps of both sides in one launch
        grow_lo = torch.ones(k, dtype=torch.bool, device=dev)
        grow_hi = torch.ones(k, dtype=torch.bool, device=dev)
        done_lo = torch.zeros(k, dtype=torch.long, device=dev)
        done_hi = torch.zeros(k, dtype=torch.long, device=dev)
        n_out = (max_expand + mo - 1) // mo
        for _e in range(n_out):
            edges = torch.cat([lo[:, None] - steps[None, :], hi[:, None] + steps[None, :]], dim=1)
            mask = torch.cat([grow_lo[:, None] & (done_lo[:, None] + isteps[None, :] < max_expand),
                              grow_hi[:, None] & (done_hi[:, None] + isteps[None, :] < max_expand)], dim=1)
            ll, _ = evaluate(u[:, None, :] + edges[:, :, None] * dirn[:, None, :])
            above = (ll > lmin) & mask
            run_lo = torch.cumprod(above[:, :mo].to(torch.long), dim=1).sum(dim=1)
            run_hi = torch.cumprod(above[:, mo:].to(torch.long), dim=1).sum(dim=1)
            lo = lo - run_lo.to(f64)
            hi = hi + run_hi.to(f64)
            done_lo = done_lo + run_lo
            done_hi = done_hi + run_hi
            if n_out == 1:
                break
            grow_lo = grow_lo & (run_lo == mo) & (done_lo < max_expand)
            grow_hi = grow_hi & (run_hi == mo) & (done_hi < max_expand)
            if not bool((grow_lo | grow_hi).any()):  # one scalar back per launch
                break
        # ---- shrinkage: the next m candidates, each drawn as if the ones before it were rejected
        pending = torch.ones(k, dtype=torch.bool, device=dev)
        draws = torch.rand((max_shrink, k), generator=gen, dtype=f64, device=dev)
        it = 0
        while it < max_shrink:
            mm = min(m, max_shrink - it)
            lo_s, hi_s = lo, hi
            ts = []
            for j in range(mm):
                t = lo_s + (hi_s - lo_s) * draws[it + j]
                ts.append(t)
                lo_s = torch.where(t < 0, t, lo_s)
                hi_s = torch.where(t >= 0, t, hi_s)
            T = torch.stack(ts, dim=1)
            cand = u[:, None, :] + T[:, :, None] * dirn[:, None, :]
            ll, th = evaluate(cand)
            acc = (ll > lmin) & pending[:, None]
            hit = acc.any(dim=1)
            first = acc.to(torch.int8).argmax(dim=1)
            u = torch.where(hit[:, None], cand[rows, first], u)
            theta = torch.where(hit[:, None], th[rows, first], theta)
            lcur = torch.where(hit, ll[rows, first], lcur)
            pending = pending & ~hit
            lo = torch.where(pending, lo_s, lo)
            hi = torch.where(pending, hi_s, hi)
            it += mm
            if not bool(pending.any()):
                break
    return u, theta, lcur, ncall


class _NativeSlice:
    """
    The slice moves of a round through the hand-written kernels of csrc/rvslice.cu
    (``rvl_slice_phase``): per move 3-4 bookkeeping launches around 2-3 likelihood launches, nothing
    read back.  Stepping out evaluates all ``n_out`` positions of both sides at once; shrinkage runs
    ``rounds`` launches of ``m`` speculated candidates per walker (a walker whose bracket is not
    resolved after rounds * m candidates keeps its position -- counted in ``unresolved``; with the
    defaults, 24 candidates, that is a few 1e-3 of the moves on the RV posteriors).
    """

    def __init__(self, torch, k, d, dev, seed, n_out=8, m=8, rounds=3):
        import ctypes
        from . import _abi
        self.torch, self.k, self.d, self.n_out, self.m, self.rounds = torch, k, d, n_out, m, rounds
        self.lib = _abi.load()
        f64 = torch.float64
        z = lambda *shape, dt=f64: torch.zeros(shape, dtype=dt, device=dev)  # noqa: E731
        self.dirn, self.lo, self.hi, self.lo2, self.hi2 = z(k, d), z(k), z(k), z(k), z(k)
        self.tval, self.pending = z(k, m), z(k, dt=torch.int32)
        self.cand_o, self.inside_o = z(k * 2 * n_out, d), z(k * 2 * n_out, dt=torch.uint8)
        self.th_o, self.ll_o = z(k * 2 * n_out, d), z(k * 2 * n_out)
        self.cand, self.inside = z(k * m, d), z(k * m, dt=torch.uint8)
        self.th, self.ll = z(k * m, d), z(k * m)
        self.stats = z(2, dt=torch.int64)
        self.args = _abi.rvl_slice_args()
        self.args.k, self.args.d, self.args.n_out, self.args.m = k, d, n_out, m
        self.args.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        for name, t in (("dirn", self.dirn), ("lo", self.lo), ("hi", self.hi), ("lo2", self.lo2),
                        ("hi2", self.hi2), ("tval", self.tval), ("pending", self.pending),
                        ("cand_out", self.cand_o), ("inside_out", self.inside_o), ("ll_out", self.ll_o),
                        ("cand", self.cand), ("inside", self.inside), ("cand_ll", self.ll),
                        ("cand_th", self.th), ("stats", self.stats)):
            setattr(self.args, name, t.data_ptr())
        self._byref = ctypes.byref(self.args)
        self.move = 0

    def _phase(self, ph):
        stream = self.torch.cuda.current_stream().cuda_stream
        rc = self.lib.rvl_slice_phase(ph, self._byref, stream)
        if rc != 0:
            raise RuntimeError("rvl_slice_phase: " + self.lib.rvl_slice_last_error().decode())

    @staticmethod
    def _writes_in_place(fused):
        """Does ``fused`` take ``theta=`` / ``lnl=`` output tensors (RVModel.transform_loglike_device)?"""
        import inspect
        try:
            params = inspect.signature(fused).parameters
        except (TypeError, ValueError):
            return False
        return "theta" in params and "lnl" in params

    def _eval(self, fused, pts, th_out, ll_out):
        if self._inplace:
            fused(pts, theta=th_out, lnl=ll_out)
        else:
            th, ll = fused(pts)
            th_out.copy_(th)
            ll_out.copy_(ll)

    def moves(self, fused, u, theta, lcur, lmin, chol, nsteps):
        """nsteps slice moves of the k walkers u (updated in place with theta and lcur)."""
        a = self.args
        self._inplace = self._writes_in_place(fused)
        chol = chol.contiguous()
        a.lmin, a.chol = lmin.data_ptr(), chol.data_ptr()
        a.u, a.theta, a.lcur = u.data_ptr(), theta.data_ptr(), lcur.data_ptr()
        ncall = 0
        for _ in range(nsteps):
            a.move, a.shrink_round = self.move & 0xFFFFFFFF, 0
            self.move += 1
            self._phase(0)
            self._eval(fused, self.cand_o, self.th_o, self.ll_o)
            self._phase(1)
            for r in range(self.rounds):
                self._eval(fused, self.cand, self.th, self.ll)
                a.shrink_round = r + 1
                self._phase(2 if r + 1 < self.rounds else 3)
            ncall += self.k * (2 * self.n_out + self.rounds * self.m)
        return ncall


def nested_sample_device(fused, ndim, nlive=400, dlogz=0.5, frac_remain=0.01, seed=0, nsteps=None,
                         batch_fraction=0.2, speculate=None, device="cuda", max_expand=16,
                         max_shrink=64, max_calls=2_000_000_000, verbose=False, num_bootstraps=30,
                         native=None, native_m=8, native_rounds=3, native_out=8):
    """
    Nested sampling with every array on ``device``.  ``fused(U[n, ndim]) -> (theta[n, ndim],
    lnL[n])`` maps unit-cube points to parameters and log-likelihoods on that device
    (``RVModel.transform_loglike_device``).  Returns the same ``NestedResult`` as
    ``sampler.nested_sample`` (arrays as numpy).

    ``native`` (default: on a CUDA device): the slice moves run through the hand-written bookkeeping
    kernels of csrc/rvslice.cu (``_NativeSlice``) instead of ~150 small torch launches per move;
    ``native_out`` stepping-out positions per side, ``native_rounds`` x ``native_m`` shrinkage
    candidates per walker and move, all speculated.  The torch formulation below it is the same
    sampler for CPU tensors (tests) and the statistical cross-check of the kernels.
    """
    import torch
    dev = torch.device(device)
    f64 = torch.float64
    gen = torch.Generator(device=dev).manual_seed(int(seed))
    if native is None:
        native = dev.type == "cuda"
    nsteps = nsteps or max(4, 3 * ndim)  # the reference's default (evidence/ultranest/__init__.py:335)
    k = max(1, min(int(batch_fraction * nlive), nlive - 2))
    m = max(1, min(6, 512 // k)) if speculate is None else max(1, int(speculate))
    m_out = max(m, min(max_expand, 4096 // (2 * k)))  # a likelihood launch costs the same up to ~4096 points
    u_live = torch.rand((nlive, ndim), generator=gen, dtype=f64, device=dev)
    th_live, l_live = fused(u_live)
    th_live, l_live = th_live.clone(), l_live.clone()
    ncall = nlive
    # the shrinkage of one round is the same every round: n = nlive, nlive-1, .., nlive-k+1
    shrink = 1.0 / (nlive - torch.arange(k, dtype=f64, device=dev))
    before = -(torch.cumsum(shrink, 0) - shrink)          # ln X offset before each retirement
    logw_rel = before + torch.log1p(-torch.exp(-shrink))   # ln w_j - ln X(start of the round)
    round_shrink = float(shrink.sum())
    logx = 0.0
    logz = torch.tensor(-math.inf, dtype=f64, device=dev)
    dead_theta, dead_logl, dead_logw = [], [], []
    # threads (see sampler.lineage_bootstrap): slot and birth constraint of every point
    root_live = torch.arange(nlive, device=dev)
    birth_live = torch.full((nlive,), -math.inf, dtype=f64, device=dev)
    dead_root, dead_birth = [], []
    niter = 0
    natives, move_no = {}, 0
    l_sorted, order = torch.sort(l_live, stable=True)
    # points tied with the k-th worst are retired with it (a plateau goes as a whole: retire_groups);
    # the count rides on the one read-back per round
    kk = int((l_sorted <= l_sorted[k - 1]).sum())
    while True:
        kk = min(kk, nlive - 2)
        worst, keep = order[:kk], order[kk:]
        if kk == k:
            logw, dx = logx + logw_rel, round_shrink
        else:  # a plateau round (rare: invalid-Keplerian sentinels at the start of a run)
            dlx, lw = retire_groups(l_sorted[:kk].cpu().numpy(), nlive)
            logw, dx = logx + torch.from_numpy(lw).to(dev), -float(dlx.sum())
        logz = torch.logaddexp(logz, torch.logsumexp(l_sorted[:kk] + logw, 0))
        dead_theta.append(th_live[worst])
        dead_logl.append(l_sorted[:kk])
        dead_logw.append(logw)
        dead_root.append(root_live[worst])
        dead_birth.append(birth_live[worst])
        logx -= dx
        niter += kk
        lmin = l_sorted[kk - 1]
        chol = _whitening(torch, u_live[keep])
        starts = keep[torch.randint(0, len(keep), (kk,), generator=gen, device=dev)]
        if native:
            if kk not in natives:
                natives[kk] = _NativeSlice(torch, kk, ndim, dev, seed, n_out=native_out, m=native_m,
                                           rounds=native_rounds)
            natives[kk].move = move_no
            u_new, th_new = u_live[starts].clone(), th_live[starts].clone()
            l_new = torch.full((kk,), float("nan"), dtype=f64, device=dev)
            nc = natives[kk].moves(fused, u_new, th_new, l_new, lmin.reshape(1).clone(), chol, nsteps)
            move_no += nsteps
        else:
            u_new, th_new, l_new, nc = _slice_moves(torch, gen, fused, u_live[starts].clone(),
                                                    th_live[starts].clone(), lmin, chol, nsteps, m,
                                                    max_expand, max_shrink, m_out=m_out)
        ncall += nc
        stuck = ~torch.isfinite(l_new)  # a walker that never moved is a copy of its start point
        l_new = torch.where(stuck, l_live[starts], l_new)
        u_live[worst], th_live[worst], l_live[worst] = u_new, th_new, l_new
        birth_live[worst] = lmin  # (root = slot: the thread continues with the point written into it)
        if ncall > max_calls:
            raise RuntimeError("nested_sample_device: max_calls exceeded")
        # UltraNest's two criteria (evidence/ultranest/__init__.py:181-185); one read-back per round
        l_sorted, order = torch.sort(l_live, stable=True)
        log_remain = l_sorted[-1] + logx
        total = torch.logaddexp(logz, log_remain)
        tied = (l_sorted <= l_sorted[k - 1]).sum().to(f64)
        lr, tt, lz, kk = (float(x) for x in torch.stack([log_remain, total, logz, tied]).cpu())
        kk = int(kk)
        if verbose and (niter // k) % 20 == 0:
            print(f"it={niter} lnZ={lz:.3f} ln(remain/Z)={lr - lz:.2f} ncall={ncall}")
        if lr - tt < math.log(frac_remain) and tt - lz < dlogz:
            break
    logw_live = torch.full((nlive,), logx - math.log(nlive), dtype=f64, device=dev)
    logz = torch.logaddexp(logz, torch.logsumexp(l_sorted + logw_live, 0))
    dead_theta.append(th_live[order])
    dead_logl.append(l_sorted)
    dead_logw.append(logw_live)
    dead_root.append(root_live[order])
    dead_birth.append(birth_live[order])
    theta = torch.cat(dead_theta)
    logl = torch.cat(dead_logl)
    logwt = logl + torch.cat(dead_logw) - logz
    p = torch.exp(logwt)
    h_info = float((p * logl).sum() - logz)                 # H = sum p_i ln L_i - ln Z
    weights = torch.exp(logwt - logwt.max())
    weights = weights / weights.sum()
    nsamp = max(1, int(1.0 / float((weights ** 2).sum())))
    pos = (float(torch.rand((), generator=gen, dtype=f64, device=dev)) +
           torch.arange(nsamp, dtype=f64, device=dev)) / nsamp
    idx = torch.clamp(torch.searchsorted(torch.cumsum(weights, 0), pos), max=len(weights) - 1)
    skilling = math.sqrt(max(h_info, 0.0) / nlive)
    bs_std, _ = lineage_bootstrap(torch.cat(dead_birth).cpu().numpy(), logl.cpu().numpy(),
                                  torch.cat(dead_root).cpu().numpy(), nlive,
                                  np.random.default_rng([int(seed), 0xB007]), num_bootstraps)
    return NestedResult(logz=float(logz), logzerr=float(max(skilling, bs_std)),
                        logzerr_skilling=skilling, logzerr_bootstrap=bs_std,
                        ncall=int(ncall), niter=int(niter), information=h_info,
                        samples=theta[idx].cpu().numpy(), weighted_samples=theta.cpu().numpy(),
                        weights=weights.cpu().numpy(), logl=logl.cpu().numpy(), nlive=nlive,
                        seed=seed, method="slice-device" + ("-native" if native else ""),
                        unresolved_moves=int(sum(int(n.stats[0]) for n in natives.values())),
                        accepted_moves=int(sum(int(n.stats[1]) for n in natives.values())))

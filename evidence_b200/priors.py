"""
Priors: host-side description + the descriptors of the on-device unit-cube transform.

Keeps the reference's configuration vocabulary -- ``[PriorName, *shape]`` entries and
``prior_constructor(input_dict)`` returning ``{objkey_parkey: prior}`` with a ``.ppf(q)``
(evidence/priors.py:429-460, 472-505) -- so config files carry over unchanged.  Every prior also
knows how to describe itself to the device (``rvl_prior_desc``): closed forms are evaluated
in-kernel; distributions whose inverse CDF the reference obtains by tabulating the CDF on 1e4
points and interpolating (evidence/priors.py:118-124, 138-144, 195-202, 223-228, 282-287,
321-326, 349-354) ship that same table to the device ONCE (the reference rebuilds it on every
call); scipy special-function inverses (Beta, Gamma, Alpha) ship a dense inverse-CDF table.
"""
import numpy as np

from . import _abi

N_TABLE = 10000  # evidence/priors.py:9   N = 1e4
STEP = 1.0 / N_TABLE  # :10


class PriorError(Exception):  # evidence/priors.py:15-16
    pass


def _interp_inverse(cdf, x, q):
    """scipy interp1d(cdf, x)(q), kind='linear': hi = clip(searchsorted(cdf, q), 1, n-1)."""
    q = np.asarray(q, dtype=np.float64)
    hi = np.clip(np.searchsorted(cdf, q, side="left"), 1, len(cdf) - 1)
    lo = hi - 1
    slope = (x[hi] - x[lo]) / (cdf[hi] - cdf[lo])
    return slope * (q - cdf[lo]) + x[lo]


class Prior:
    """A frozen 1-D prior: ``ppf(q)`` on the host, ``descriptor()`` for the device."""
    name = "Prior"
    kind = None

    def __init__(self, *pars):
        self.pars = tuple(float(p) for p in pars)
        self._check()

    def _check(self):
        pass

    def ppf(self, q):
        raise NotImplementedError

    def table(self):
        """(cdf_knots, x_knots) for RVL_PRIOR_TABLE kinds, else None."""
        return None

    def descriptor(self, table_offset=0):
        d = _abi.rvl_prior_desc()
        d.kind = self.kind
        for i, p in enumerate(self._device_pars()):
            d.p[i] = p
        tab = self.table()
        if tab is not None:
            d.table_len = len(tab[0])
            d.table_offset = table_offset
        return d

    def _device_pars(self):
        return self.pars[:4]

    def __repr__(self):
        return f"{self.name}{self.pars}"


# ---- closed forms (evaluated in-kernel) ---------------------------------------------------
class Uniform(Prior):  # evidence/priors.py:22-42
    name, kind = "Uniform", _abi.RVL_PRIOR_UNIFORM

    def _check(self):
        if not self.pars[0] < self.pars[1]:
            raise PriorError("Uniform needs xmin < xmax")

    def ppf(self, q):
        xmin, xmax = self.pars
        return xmin + (xmax - xmin) * np.asarray(q, dtype=np.float64)

    def pdf(self, x):
        xmin, xmax = self.pars
        x = np.asarray(x, dtype=np.float64)
        return np.where((x >= xmin) * (x <= xmax), 1.0 / (xmax - xmin), 0.0)


class Jeffreys(Prior):  # :45-63
    name, kind = "Jeffreys", _abi.RVL_PRIOR_JEFFREYS

    def _check(self):
        if not (self.pars[0] > 0.0 and self.pars[1] > self.pars[0]):
            raise PriorError("Jeffreys needs 0 < xmin < xmax")

    def ppf(self, q):
        xmin, xmax = self.pars
        return xmin * (xmax / xmin) ** np.asarray(q, dtype=np.float64)

    def pdf(self, x):
        xmin, xmax = self.pars
        x = np.asarray(x, dtype=np.float64)
        bad = np.logical_or(x < xmin, x > xmax)
        return np.where(bad, 0.0, 1.0 / (x * np.log(xmax / xmin)))


class ModJeffreys(Prior):  # :66-83
    name, kind = "ModJeffreys", _abi.RVL_PRIOR_MODJEFFREYS

    def _check(self):
        if not (self.pars[1] > self.pars[0] > 0):
            raise PriorError("ModJeffreys needs 0 < x0 < xmax")

    def ppf(self, q):
        x0, xmax = self.pars
        return x0 * ((1 + float(xmax) / x0) ** np.asarray(q, dtype=np.float64)) - x0


class UniformFrequency(Prior):  # :85-101
    name, kind = "UniformFrequency", _abi.RVL_PRIOR_UNIFORMFREQ

    def _check(self):
        if not (self.pars[1] > self.pars[0] > 0):
            raise PriorError("UniformFrequency needs 0 < xmin < xmax")

    def ppf(self, q):
        xmin, xmax = self.pars
        return xmin / (1 - np.asarray(q, dtype=np.float64) * (xmax - xmin) / xmax)


class TruncatedRayleigh(Prior):  # :231-252
    name, kind = "TruncatedRayleigh", _abi.RVL_PRIOR_TRUNCRAYLEIGH

    def _check(self):
        if not self.pars[0] > 0:
            raise PriorError("TruncatedRayleigh needs sigma > 0")

    def ppf(self, q):
        sigma, xmax = self.pars
        A = 1 - np.exp(-xmax ** 2 / (2 * sigma ** 2))
        return np.sqrt(-2 * sigma ** 2 * np.log(1 - (np.asarray(q, dtype=np.float64) * A)))


class Normal(Prior):  # :436  stats.norm(loc=0, scale=1)
    name, kind = "Normal", _abi.RVL_PRIOR_NORMAL

    def __init__(self, loc=0.0, scale=1.0):
        super().__init__(loc, scale)

    def ppf(self, q):
        from scipy import special
        return self.pars[0] + self.pars[1] * special.ndtri(np.asarray(q, dtype=np.float64))


class LogNormal(Prior):  # :437  stats.lognorm(s, loc=0, scale=1)
    name, kind = "LogNormal", _abi.RVL_PRIOR_LOGNORMAL

    def __init__(self, s, loc=0.0, scale=1.0):
        super().__init__(s, loc, scale)

    def ppf(self, q):
        from scipy import special
        s, loc, scale = self.pars
        return loc + scale * np.exp(s * special.ndtri(np.asarray(q, dtype=np.float64)))


# ---- tabulated inverse CDFs -----------------------------------------------------------------
class _TablePrior(Prior):
    """ppf = interp1d(cdf(x_grid), x_grid)(q); the table is built once and cached."""
    kind = _abi.RVL_PRIOR_TABLE

    def _grid(self):
        raise NotImplementedError

    def _cdf(self, x):
        raise NotImplementedError

    def table(self):
        if not hasattr(self, "_tab"):
            x = np.ascontiguousarray(self._grid(), dtype=np.float64)
            cdf = np.ascontiguousarray(self._cdf(x), dtype=np.float64)
            self._tab = (cdf, x)
        return self._tab

    def _post(self, x):
        return x

    def ppf(self, q):
        cdf, x = self.table()
        return self._post(_interp_inverse(cdf, x, q))

    def _device_pars(self):
        return (0.0, 0.0, 0.0, 0.0)


def _arange_grid(xmin, xmax):
    dx = (xmax - xmin) * STEP
    return np.arange(xmin, xmax + dx, dx)


def _ncdf(x, mu, sigma):
    from scipy import stats
    return stats.norm.cdf(x, mu, sigma)


class Binormal(_TablePrior):  # :103-124
    name = "Binormal"

    def _check(self):
        mu1, s1, mu2, s2, A = self.pars
        if not (s1 > 0 and s2 > 0 and mu1 <= mu2 and -1.0 <= A <= 1.0):
            raise PriorError("bad Binormal parameters")

    def _grid(self):
        mu1, s1, mu2, s2, _ = self.pars
        return _arange_grid(mu1 - 9. * s1, mu2 + 9. * s2)

    def _cdf(self, x):
        mu1, s1, mu2, s2, A = self.pars
        return 0.5 * (_ncdf(x, mu1, s1) * (1. - A) + _ncdf(x, mu2, s2) * (1. + A))


class Log10Normal(_TablePrior):  # :127-144 (the reference's float `num` is taken as int(N))
    name = "Log10Normal"

    def _grid(self):
        mu, sigma = self.pars
        return np.linspace(mu - 9. * sigma, mu + 9. * sigma, N_TABLE)

    def _cdf(self, x):
        mu, sigma = self.pars
        return _ncdf(np.log10(10 ** x), mu, sigma)

    def _post(self, x):
        return 10 ** x

    def _device_pars(self):
        return (1.0, 0.0, 0.0, 0.0)  # p0 = 1: the device raises 10 to the interpolated value


class AsymmetricNormal(_TablePrior):  # :178-202
    name = "AsymmetricNormal"

    def _check(self):
        if not (self.pars[1] > 0 and self.pars[2] > 0):
            raise PriorError("bad AsymmetricNormal parameters")

    def _grid(self):
        mu, s1, s2 = self.pars
        return _arange_grid(mu - 9 * s1, mu + 9 * s2)

    def _cdf(self, x):
        mu, s1, s2 = self.pars
        k1 = 2.0 * s1 / (s1 + s2)
        k2 = 2.0 * s2 / (s1 + s2)
        c1 = _ncdf(x, mu, s1) * k1
        c2 = (_ncdf(x, mu, s2) - 0.5) * k2
        return np.where(x <= mu, c1, k1 * 0.5 + c2)


class TruncatedUNormal(_TablePrior):  # :205-228
    name = "TruncatedUNormal"

    def _check(self):
        if not self.pars[1] > 0:
            raise PriorError("TruncatedUNormal needs sigma > 0")

    def _grid(self):
        return _arange_grid(self.pars[2], self.pars[3])

    def _cdf(self, x):
        mu, sigma, xmin, xmax = self.pars
        A1 = _ncdf(xmax, mu, sigma) - _ncdf(xmin, mu, sigma)
        cdf = (_ncdf(x, mu, sigma) - _ncdf(xmin, mu, sigma)) / A1
        cdf = np.where(x >= xmin, cdf, 0.0)
        return np.where(x < xmax, cdf, 1.0)


class PowerLaw(_TablePrior):  # :267-287
    name = "PowerLaw"

    def _check(self):
        alpha, xmin, xmax = self.pars
        if not (xmax > xmin and xmin >= 0 and xmax > 0 and alpha != -1):
            raise PriorError("bad PowerLaw parameters")

    def _grid(self):
        return _arange_grid(self.pars[1], self.pars[2])

    def _cdf(self, x):
        alpha, xmin, xmax = self.pars
        Ap = 1.0 / (xmax ** (1.0 + alpha) - xmin ** (1.0 + alpha))
        cdf = Ap * (x ** (1.0 + alpha) - xmin ** (1.0 + alpha))
        cdf = np.where(x > xmin, cdf, 0.0)
        return np.where(x >= xmax, 1.0, cdf)


class DoublePowerLaw(_TablePrior):  # :290-326
    name = "DoublePowerLaw"

    def _check(self):
        alpha, beta, x0, xmin, xmax = self.pars
        if not (xmax > xmin and xmin >= 0 and xmax > 0 and alpha != -1):
            raise PriorError("bad DoublePowerLaw parameters")

    def _grid(self):
        return _arange_grid(self.pars[3], self.pars[4])

    def _cdf(self, x):
        alpha, beta, x0, xmin, xmax = self.pars
        a1 = (x0 ** (1.0 + alpha) - xmin ** (1.0 + alpha)) / (alpha + 1.0)
        a2 = (xmax ** (1.0 + beta) - x0 ** (1.0 + beta)) / (beta + 1.0)
        ratio = (x0 * 1.0) ** alpha / (x0 * 1.0) ** beta
        A = 1.0 / (a1 + ratio * a2)
        ca = A * (x ** (1.0 + alpha) - xmin ** (1.0 + alpha)) / (1.0 + alpha)
        cb = A * (x0 ** (1.0 + alpha) - xmin ** (1.0 + alpha)) / (1.0 + alpha) + \
            ratio * A * (x ** (1.0 + beta) - x0 ** (1.0 + beta)) / (1.0 + beta)
        cdf = np.where(x < x0, ca, cb)
        cdf = np.where(x > xmin, cdf, 0.0)
        return np.where(x >= xmax, 1.0, cdf)


class Sine(_TablePrior):  # :329-354, support clipped to [0, 180] degrees (:454-455)
    name = "Sine"

    def _check(self):
        if not self.pars[1] > self.pars[0]:
            raise PriorError("Sine needs xmin < xmax")

    def _grid(self):
        return _arange_grid(self.pars[0], self.pars[1])

    def _cdf(self, x):
        xmin = max(self.pars[0], 0.0)
        xmax = min(self.pars[1], 180.0)
        A = np.cos(xmin * np.pi / 180.0) - np.cos(xmax * np.pi / 180.0)
        cdf = (np.cos(xmin * np.pi / 180.0) - np.cos(x * np.pi / 180.0)) / A
        cdf = np.where(x >= xmin, cdf, 0.0)
        return np.where(x <= xmax, cdf, 1.0)


class _ScipyPrior(_TablePrior):
    """
    Host ppf straight from scipy (as the reference, evidence/priors.py:375-376, 397-398,
    424-425).  The device gets an inverse-CDF table x_k = ppf(q_k) on a q grid that is refined
    geometrically towards both ends, PLUS the exact slopes dx/dq = 1/pdf(x_k), and interpolates
    with a cubic Hermite spline (error O(h^4); linear where a slope is not finite).  The
    device-vs-scipy tolerance is a property of the table (measured <= 4e-11 relative in
    [1e-3, 1-1e-3] for Beta/Gamma/Alpha, tests state 1e-9); lnL parity is always defined on
    identical theta.
    """
    N_DEV = 1 << 14

    def _dist(self):
        raise NotImplementedError

    def ppf(self, q):
        return self._dist().ppf(np.asarray(q, dtype=np.float64))

    def table(self):
        if not hasattr(self, "_tab"):
            # knots: uniform in q, plus logistic in q (geometric refinement towards 0 and 1)
            n = self.N_DEV
            mid = np.linspace(0.0, 1.0, n + 1)[1:-1]
            logit = 1.0 / (1.0 + np.exp(-np.linspace(-40.0, 40.0, n)))
            tail = np.geomspace(1e-300, 1e-17, 256)
            q = np.unique(np.concatenate([[0.0], tail, logit, mid, 1.0 - tail[::-1], [1.0]]))
            dist = self._dist()
            with np.errstate(all="ignore"):
                x = dist.ppf(q)
            finite = np.isfinite(x)
            q, x = q[finite], x[finite]
            keep = np.concatenate([[True], np.diff(x) > 0])  # strictly increasing knots only
            q, x = q[keep], x[keep]
            with np.errstate(all="ignore"):
                slope = 1.0 / dist.pdf(x)
            slope = np.where(np.isfinite(slope) & (slope > 0), slope, -1.0)  # -1: use linear
            self._tab = (np.ascontiguousarray(q), np.ascontiguousarray(x))
            self._slope = np.ascontiguousarray(slope)
        return self._tab

    def slopes(self):
        self.table()
        return self._slope

    def _device_pars(self):
        return (0.0, 1.0, 0.0, 0.0)  # p1 = 1: a slope array follows the knots (Hermite)


class Alpha(_ScipyPrior):  # :356-376
    name = "Alpha"

    def _dist(self):
        from scipy import stats
        return stats.alpha(self.pars[0])


class Beta(_ScipyPrior):  # :378-398
    name = "Beta"

    def _check(self):
        if not (self.pars[0] > 0 and self.pars[1] > 0):
            raise PriorError("Beta needs a, b > 0")

    def _dist(self):
        from scipy import stats
        return stats.beta(self.pars[0], self.pars[1])


class Gamma(_ScipyPrior):  # :400-425
    name = "Gamma"

    def _check(self):
        if not (self.pars[0] > 0 and self.pars[1] > 0):
            raise PriorError("Gamma needs alpha, beta > 0")

    def _dist(self):
        from scipy import stats
        return stats.gamma(self.pars[0], scale=1.0 / self.pars[1])


# ---- PolyChord's sorted ("forced identifiability") priors ---------------------------------------
# The reference re-exports pypolychord.priors.SortedUniformPrior / LogSortedUniformPrior
# (evidence/priors.py:462-467) and its PolyChord runner applies ONE such prior object to the whole
# group of parameters that share it (evidence/polychord/__init__.py:145-160).  pypolychord is not
# part of the reference checkout and is absent from this image, so these follow its published
# definition (parity unpinned): t_N = x_N^(1/N), t_n = x_n^(1/n) t_{n+1}, then the plain
# (log-)uniform map.  Host only: PolyChord calls the prior one point at a time.
def forced_identifiability_transform(x):
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    t = np.zeros(n)
    t[n - 1] = x[n - 1] ** (1.0 / n)
    for k in range(n - 2, -1, -1):
        t[k] = x[k] ** (1.0 / (k + 1)) * t[k + 1]
    return t


class SortedUniformPrior:
    def __init__(self, a, b):
        self.a, self.b = float(a), float(b)

    def __call__(self, x):
        return self.a + (self.b - self.a) * forced_identifiability_transform(x)


class LogSortedUniformPrior(SortedUniformPrior):
    def __call__(self, x):
        return self.a * (self.b / self.a) ** forced_identifiability_transform(x)


SortedUniform, SortedLogUniform = SortedUniformPrior, LogSortedUniformPrior

distdict = {c.name: c for c in (Uniform, Jeffreys, ModJeffreys, UniformFrequency, Normal,
                                LogNormal, Log10Normal, Binormal, AsymmetricNormal,
                                TruncatedUNormal, TruncatedRayleigh, PowerLaw, DoublePowerLaw,
                                Sine, Alpha, Beta, Gamma)}


distdict.update({"SortedUniform": SortedUniformPrior, "SortedLogUniform": LogSortedUniformPrior})


def make_prior(priortype, *pars):
    try:
        cls = distdict[priortype]
    except KeyError:
        raise PriorError(f"Unknown type of prior: {priortype}")
    return cls(*pars)


def prior_constructor(input_dict, customprior_dict=None):
    """
    Same walk as evidence/priors.py:472-505: every ``[init, flag, [PriorName, *shape]]`` entry
    with ``flag != 0`` becomes ``priordict[objkey_parkey]``.
    """
    priordict = {}
    for objkey in input_dict.keys():
        for parkey in input_dict[objkey]:
            parlist = input_dict[objkey][parkey]
            if not isinstance(parlist, list):
                continue
            if parlist[1] == 0:
                continue
            priortype = parlist[2][0]
            pars = parlist[2][1:]
            if priortype not in distdict:  # the reference's message (evidence/priors.py:500-503)
                raise PriorError(f"Parameter {objkey}_{parkey}: Unknown type of prior.")
            try:
                priordict[objkey + "_" + parkey] = make_prior(priortype, *pars)
            except (PriorError, TypeError) as exc:  # bad shape parameters: keep the real cause
                raise PriorError(f"Parameter {objkey}_{parkey} ({priortype}): {exc}") from exc
    return priordict


def device_descriptors(priors):
    """``rvl_prior_desc`` list + the concatenated knot tables for a list of priors."""
    descs, chunks, off = [], [], 0
    for pr in priors:
        if not isinstance(pr, Prior):
            raise PriorError(f"{pr!r} cannot be staged on the device "
                             "(build priors with evidence_b200.priors.prior_constructor)")
        descs.append(pr.descriptor(table_offset=off))
        tab = pr.table()
        if tab is not None:
            chunks += [tab[0], tab[1]]
            off += 2 * len(tab[0])
            if hasattr(pr, "slopes"):  # third array: dx/dq at the knots
                chunks.append(pr.slopes())
                off += len(tab[0])
    tables = np.ascontiguousarray(np.concatenate(chunks) if chunks else np.zeros(0),
                                  dtype=np.float64)
    return descs, tables

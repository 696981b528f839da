"""
CPU restatement of the FIP-periodogram accumulation of the reference's fip_criterion.py.

TEST INFRASTRUCTURE ONLY: nothing under evidence_b200/ imports this module; it is the checker of
the device path (`evidence_b200.fip` -> `rvl_fip_accumulate`) in tests/ and the CPU arm of
tools/fip_bench.py.

Reference followed (paths relative to the reference checkout):
  evidence/fip_criterion.py:230-236   frequency grid  nu = linspace(2pi/Pmax, 2pi/Pmin, nfreq),
                                      window nu_window = coef * 2pi / Tobs, nua/nub = nu -/+ window/2
  evidence/fip_criterion.py:303-338   per run, per k-planet model, per posterior sample: mean motions
                                      2pi/P of the sample's planets (optionally their 1-day / 30-day
                                      aliases, clipped to the grid's range), the grid bins whose window
                                      contains one of them, and  fapnu[run, bins] -= p(k|y) * weight
  evidence/fip_criterion.py:264-266   p(k|y) = exp(logZ_k - logsumexp(logZ))

Parity status: the reference script executes at import, reads run directories of pickles and needs
matplotlib, so it cannot be imported, and none of its tests covers this path.  PINNED instead
against the reference's OWN STATEMENTS: oracle/make_golden_post.py takes lines 305-338 from the file
where it lies, executes them on seeded inputs and commits the result (tests/golden/fip_ref.npz);
`accumulate_literal` -- the same loop nest transcribed statement by statement (same numpy calls,
same fancy-index update, whose duplicate indices subtract ONCE) -- reproduces it bit for bit, and
`accumulate`, a vectorised form of the same arithmetic, to 1e-15 (tests/test_fip.py).  The alias
branch of the reference cannot run (NameError at :321) and stays unpinned.
"""
import numpy as np

TWO_PI = 2 * np.pi
SIDEREAL_DAY = 0.99727   # fip_criterion.py:322-323
MONTH = 30.0             # fip_criterion.py:324-325


def frequency_grid(Pmin, Pmax, nfreq, Tobs, coef_window=1.0):
    """(nu, nua, nub) of fip_criterion.py:233-236."""
    nu = np.linspace(TWO_PI / Pmax, TWO_PI / Pmin, nfreq)
    nu_window = coef_window * TWO_PI / Tobs
    return nu, nu - nu_window / 2, nu + nu_window / 2


def posterior_of_k(logZs):
    """p(k|y) from the per-model evidences, fip_criterion.py:264-266."""
    logZs = np.asarray(logZs, dtype=np.float64)
    m = logZs.max()
    return np.exp(logZs - (m + np.log(np.sum(np.exp(logZs - m)))))


def sample_frequencies(x, Pmin, Pmax, with_alias):
    """Mean motions of one sample (fip_criterion.py:318-331).  with_alias: the reference's own
    branch raises NameError (x_freqs is never allocated before :321), so it has no executable
    behaviour; this is its evident intent -- a fresh [5, k] array per sample."""
    x = np.asarray(x, dtype=np.float64)
    if not with_alias:
        return TWO_PI / x
    f = np.empty((5, len(x)))
    f[0, :] = TWO_PI / x
    f[1, :] = np.abs(TWO_PI / x + TWO_PI / SIDEREAL_DAY)
    f[2, :] = np.abs(TWO_PI / x - TWO_PI / SIDEREAL_DAY)
    f[3, :] = np.abs(TWO_PI / x + TWO_PI / MONTH)
    f[4, :] = np.abs(TWO_PI / x - TWO_PI / MONTH)
    f = f.flatten()
    f = f[f <= TWO_PI / Pmin]
    f = f[f >= TWO_PI / Pmax]
    return f


def accumulate_literal(fap_row, nua, nub, samples, weights, pk, Pmin, Pmax, with_alias=False):
    """One (run, k) block of fip_criterion.py:308-338, statement by statement.  In place."""
    weights = np.asarray(weights, dtype=np.float64)
    weights = weights / np.sum(weights)               # :313  normalise weights
    for i, x in enumerate(samples):                   # :317
        x_freqs = sample_frequencies(x, Pmin, Pmax, with_alias)
        beg = np.searchsorted(nub, x_freqs, 'right')  # :333
        end = np.searchsorted(nua, x_freqs, 'left')   # :334
        listind = []
        for bi, ei in zip(beg, end):                  # :336-337
            listind += range(bi, ei)
        fap_row[listind] -= pk * weights[i]           # :338  (duplicate bins subtract once)
    return fap_row


def accumulate(fap_row, nua, nub, samples, weights, pk, Pmin, Pmax, with_alias=False):
    """Vectorised form of `accumulate_literal` (difference array over the union of each sample's
    bin ranges); differs from it only by floating-point summation order."""
    samples = np.atleast_2d(np.asarray(samples, dtype=np.float64))
    weights = np.asarray(weights, dtype=np.float64)
    weights = weights / np.sum(weights)
    n, k = samples.shape
    f = TWO_PI / samples
    if with_alias:
        f = np.stack([f, np.abs(f + TWO_PI / SIDEREAL_DAY), np.abs(f - TWO_PI / SIDEREAL_DAY),
                      np.abs(f + TWO_PI / MONTH), np.abs(f - TWO_PI / MONTH)], axis=1).reshape(n, 5 * k)
        keep = (f <= TWO_PI / Pmin) & (f >= TWO_PI / Pmax)
    else:
        keep = np.ones_like(f, dtype=bool)
    beg = np.searchsorted(nub, f, 'right')
    end = np.searchsorted(nua, f, 'left')
    end = np.where(keep & (end > beg), end, beg)      # empty ranges
    # union of the ranges of one sample: sort by beg, clip each range to start after the running
    # maximum of the previous ends
    order = np.argsort(beg, axis=1, kind="stable")
    beg = np.take_along_axis(beg, order, 1)
    end = np.take_along_axis(end, order, 1)
    run_end = np.maximum.accumulate(end, axis=1)
    prev_end = np.concatenate([np.zeros((n, 1), dtype=run_end.dtype), run_end[:, :-1]], axis=1)
    b = np.maximum(beg, prev_end)
    e = np.maximum(end, b)
    w = (pk * weights)[:, None] * np.ones_like(b, dtype=np.float64)
    nz = e > b
    diff = np.zeros(len(fap_row) + 1)
    np.add.at(diff, b[nz], w[nz])
    np.add.at(diff, e[nz], -w[nz])
    fap_row -= np.cumsum(diff)[:-1]
    return fap_row


def fip_periodogram(runs, pky, Pmin, Pmax, nfreq, Tobs, coef_window=1.0, with_alias=False,
                    literal=False):
    """
    fapnu[run, nfreq] of fip_criterion.py:303-338.  ``runs[r][k]`` = (samples[n, k] periods,
    weights[n]) for k = 1..nmod-1 (``runs[r][0]`` is ignored: the 0-planet model has no periods).
    """
    nu, nua, nub = frequency_grid(Pmin, Pmax, nfreq, Tobs, coef_window)
    fapnu = np.ones([len(runs), nfreq])
    fn = accumulate_literal if literal else accumulate
    for r, models in enumerate(runs):
        for kmod in range(1, len(models)):
            samples, weights = models[kmod]
            fn(fapnu[r], nua, nub, samples, weights, pky[kmod], Pmin, Pmax, with_alias)
    return nu, fapnu

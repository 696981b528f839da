"""
FIP periodogram (false inclusion probability) of a set of nested-sampling runs, accumulated on the
device.  Host-side mirror of the block of the reference's ``fip_criterion.py`` that builds
``fapnu`` (evidence/fip_criterion.py:230-236, 264-266, 303-338); everything around it in that script
(reading run directories, tables, plots) is not rebuilt.

    nu, fapnu = fip_periodogram(runs, logZs, Pmin, Pmax, nfreq, Tobs)

``runs[r][k]`` is ``(samples[n, k], weights[n])`` -- the period columns of the posterior samples of
run ``r`` with ``k`` planets and their weights, as the script collects them (:186-198); entry 0
(the 0-planet model) is ignored.  The frequency grid and p(k|y) are formed on the host exactly as
the script does; the per-sample loop (:315-338) runs in ``rvl_fip_accumulate`` (one call per
(run, k) block).  There is no CPU fallback.
"""
import ctypes

import numpy as np

from . import _abi

TWO_PI = 2 * np.pi
_dp = ctypes.POINTER(ctypes.c_double)


class FIPError(RuntimeError):
    pass


def frequency_grid(Pmin, Pmax, nfreq, Tobs, coef_window=1.0):
    """(nu, nua, nub): evidence/fip_criterion.py:233-236."""
    nu = np.linspace(TWO_PI / Pmax, TWO_PI / Pmin, nfreq)
    nu_window = coef_window * TWO_PI / Tobs
    return nu, nu - nu_window / 2, nu + nu_window / 2


def posterior_of_k(logZs):
    """p(k|y) = exp(logZ - logsumexp(logZ)): evidence/fip_criterion.py:264-266."""
    logZs = np.asarray(logZs, dtype=np.float64)
    m = logZs.max()
    return np.exp(logZs - (m + np.log(np.sum(np.exp(logZs - m)))))


def accumulate_block(fap_row, nua, nub, samples, weights, pk, Pmin, Pmax, with_alias=False,
                     device=-1):
    """
    One (run, k) block of the loop nest (:308-338) on the device; ``fap_row`` (float64[nfreq],
    C-contiguous) is updated in place.  Returns the CUDA-event time of the kernels in ms.
    """
    lib = _abi.load()
    samples = np.ascontiguousarray(np.atleast_2d(samples), dtype=np.float64)
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    if samples.shape[0] != weights.shape[0]:
        raise ValueError("samples and weights disagree on the number of samples")
    if not (fap_row.flags.c_contiguous and fap_row.dtype == np.float64):
        raise ValueError("fap_row must be a C-contiguous float64 array")
    nua = np.ascontiguousarray(nua, dtype=np.float64)
    nub = np.ascontiguousarray(nub, dtype=np.float64)
    ms = ctypes.c_double(0.0)
    rc = lib.rvl_fip_accumulate(int(device), nua.ctypes.data_as(_dp), nub.ctypes.data_as(_dp),
                                len(nua), samples.ctypes.data_as(_dp), samples.shape[1],
                                weights.ctypes.data_as(_dp), samples.shape[0], float(pk),
                                1 if with_alias else 0, float(Pmin), float(Pmax),
                                fap_row.ctypes.data_as(_dp), ctypes.byref(ms))
    if rc != 0:
        raise FIPError(f"librvlnl: {_abi.RVL_ERRORS.get(rc, rc)}: "
                       f"{lib.rvl_fip_last_error().decode()}")
    return ms.value


def fip_periodogram(runs, logZs, Pmin, Pmax, nfreq=50000, Tobs=1.0, coef_window=1.0,
                    with_alias=False, device=-1):
    """
    ``(nu, fapnu[len(runs), nfreq])`` of evidence/fip_criterion.py:303-338 (``fapnu`` starts at 1
    and loses p(k|y) * weight wherever a sample's planet falls within the window of a frequency).
    ``logZs[k]``: the (median) evidence of the k-planet model (:256), k = 0..nmod-1.
    """
    nu, nua, nub = frequency_grid(Pmin, Pmax, nfreq, Tobs, coef_window)
    pky = posterior_of_k(logZs)
    fapnu = np.ones([len(runs), nfreq])
    for r, models in enumerate(runs):
        for kmod in range(1, len(models)):
            samples, weights = models[kmod]
            accumulate_block(fapnu[r], nua, nub, samples, weights, pky[kmod], Pmin, Pmax,
                             with_alias, device)
    return nu, fapnu


# ------------------------------------------------------------------------------------------
# posterior planet ordering (evidence/post_processing.py:93-128)
def planet_tables(parnames, nplanets):
    """The column lists the reference builds from the names (:94-102): planets[n-1] = columns
    whose name contains 'planet{n}', planet_idxs = their 'period' columns."""
    planets, planet_idxs = [], []
    for n in range(1, nplanets + 1):
        planets.append([i for i, par in enumerate(parnames) if f'planet{n}' in par])
        planet_idxs += [i for i in planets[-1] if 'period' in parnames[i]]
    return planets, planet_idxs


def order_planets(samples, parnames, nplanets, device=-1):
    """
    The reference's ``order`` option of ``postprocess`` (evidence/post_processing.py:104-128) for a
    whole posterior at once: rows whose planet periods are not non-decreasing get their planet
    columns permuted exactly as the reference's index list does.  Returns a new array.
    """
    lib = _abi.load()
    samples = np.ascontiguousarray(samples, dtype=np.float64)
    if samples.ndim != 2 or samples.shape[1] != len(parnames):
        raise ValueError("samples must have one column per parameter name")
    planets, planet_idxs = planet_tables(parnames, nplanets)
    if nplanets < 1 or len(planet_idxs) != nplanets or len({len(p) for p in planets}) != 1:
        raise ValueError("every planet needs one period column and the same number of parameters")
    K, Q = nplanets, len(planets[0])
    pc = (ctypes.c_int32 * K)(*planet_idxs)
    cols = (ctypes.c_int32 * (K * Q))(*[c for p in planets for c in p])
    out = np.empty_like(samples)
    ms = ctypes.c_double(0.0)
    rc = lib.rvl_order_planets(int(device), samples.ctypes.data_as(_dp), samples.shape[0],
                               samples.shape[1], pc, cols, K, Q, out.ctypes.data_as(_dp),
                               ctypes.byref(ms))
    if rc != 0:
        raise FIPError(f"librvlnl: {_abi.RVL_ERRORS.get(rc, rc)}: "
                       f"{lib.rvl_order_last_error().decode()}")
    order_planets.last_kernel_ms = ms.value
    return out

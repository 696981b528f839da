"""
CPU restatement of the posterior planet-ordering loop of the reference's post-processing.

TEST INFRASTRUCTURE ONLY (checker of `evidence_b200.fip.order_planets` -> `rvl_order_planets`).

Reference followed: evidence/post_processing.py:93-128 -- for every posterior sample whose planet
periods are not in non-decreasing order, the parameter columns of the planets are permuted with an
index list built from `np.argsort` of the periods.  Transcribed statement by statement (a numpy
array of rows stands in for the pandas frame), INCLUDING its gather-with-the-inverse-permutation
behaviour: column i of planet p receives the value of the planet slot whose index is p's RANK, so
for three or more planets in cyclic disorder the result is not sorted.  Identical results on
identical inputs is the contract, so the device path reproduces exactly this.

Parity status: post_processing.py needs matplotlib/corner at import and no reference test covers
this loop.  PINNED against the reference's OWN STATEMENTS: oracle/make_golden_post.py executes lines
93-128 of the file where it lies on seeded posteriors (1-4 planets, a NaN period) and commits input
and output (tests/golden/order_ref.npz); this transcription reproduces them bit for bit.
"""
import numpy as np


def planet_tables(parnames, nplanets):
    """planets[n-1] = columns whose name contains 'planet{n}'; planet_idxs = their period columns
    (post_processing.py:94-102)."""
    planets, planet_idxs = [], []
    for n in range(1, nplanets + 1):
        planets.append([])
        for i, par in enumerate(parnames):
            if f'planet{n}' in par:
                planets[n - 1].append(i)
                if 'period' in par:
                    planet_idxs.append(i)
    return planets, planet_idxs


def order_samples_literal(samples, parnames, nplanets, kind=None):
    """post_processing.py:104-128 on a float array samples[n, ndim]; returns a new array.
    ``kind``: passed to np.argsort; None is the reference's call.  numpy's default sort is NOT
    stable (its SIMD kernels order exactly equal periods platform-dependently), so rows with tied
    periods have no defined reference answer; the device breaks ties by planet index, which is
    ``kind='stable'``."""
    samples = np.array(samples, dtype=np.float64, copy=True)
    planets, planet_idxs = planet_tables(parnames, nplanets)
    for idx in range(len(samples)):
        sample_arr = samples[idx].copy()
        periods_tmp = sample_arr[planet_idxs]
        idxs = np.arange(samples.shape[1], dtype=int)
        if not np.all(periods_tmp[:-1] <= periods_tmp[1:]):
            sorted_periods_args = (np.argsort(periods_tmp) if kind is None
                                   else np.argsort(periods_tmp, kind=kind))
            for i, par in enumerate(parnames):
                if 'planet' not in par:
                    idxs[i] = i
                else:
                    planet = int(par[6])
                    new_pos = list(sorted_periods_args).index(planet - 1)
                    internal_pos = planets[planet - 1].index(i)
                    target = planets[new_pos][internal_pos]
                    idxs[i] = target
        samples[idx] = sample_arr[idxs]
    return samples

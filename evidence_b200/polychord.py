"""
PolyChord adapter (boundary only).

PolyChord calls the likelihood one point at a time from Fortran, so it cannot use large batches;
it gets the device model through the scalar protocol (a batch of 1 per call).  The entry point,
settings defaults and type checks follow evidence/polychord/__init__.py:32-258, 299-424.  The
sampler itself (pypolychord, Fortran/C++) is third-party and absent from this image: importing
this module works, ``run()`` raises ImportError without it.
"""
import datetime
import os

import numpy as np


def make_callbacks(model, priordict):
    """(prior, loglike) in PolyChord's convention: prior(cube)->theta, loglike(theta)->(lnL, [])."""
    parnames = model.parnames
    from .priors import LogSortedUniformPrior, SortedUniformPrior
    # Sorted (forced-identifiability) priors, evidence/polychord/__init__.py:137-160: the reference
    # collects EVERY parameter whose prior is a SortedUniformPrior into one group (and every
    # LogSortedUniformPrior into another), whatever object each parameter holds --
    # prior_constructor builds one object per parameter -- and applies the LAST such object seen to
    # the whole group, in parnames order.  (isinstance order as in the reference: the log variant
    # is tested second there, but as a subclass here it must be tested first.)
    groups = {"log": [None, []], "lin": [None, []]}
    plain = []
    for i, p in enumerate(parnames):
        pr = priordict[p]
        if isinstance(pr, LogSortedUniformPrior):
            groups["log"][0] = pr
            groups["log"][1].append(i)
        elif isinstance(pr, SortedUniformPrior):
            groups["lin"][0] = pr
            groups["lin"][1].append(i)
        else:
            plain.append(i)

    def prior(hypercube):
        hypercube = np.asarray(hypercube, dtype=np.float64)
        theta = np.ones_like(hypercube)
        for i in plain:
            theta[i] = priordict[parnames[i]].ppf(hypercube[i])
        for pr, idx in groups.values():
            if idx:
                theta[idx] = pr(hypercube[idx])
        return theta

    def loglike(x):
        return (model.log_likelihood(x), [])  # :166-171
    return prior, loglike


DEFAULTS = {"do_clustering": True, "write_resume": False, "read_resume": False, "feedback": 1,
            "precision_criterion": 0.001, "boost_posterior": 0.0}
_TYPES = {"nlive": int, "num_repeats": int, "do_clustering": bool, "read_resume": bool,
          "precision_criterion": float}


def default_settings(ndim, polysettings=None):
    """Defaults nlive = 25 ndim, num_repeats = 5 ndim, ... and the reference's type checks (:330-375)."""
    settings = dict(DEFAULTS, nlive=25 * ndim, num_repeats=5 * ndim)
    if polysettings is not None:
        if type(polysettings) is not dict:
            raise TypeError("polysettings has to be a dictionary")
        for key, typ in _TYPES.items():
            if key in polysettings and type(polysettings[key]) is not typ:
                raise TypeError(f"{key} has to be {typ.__name__} (got type {type(polysettings[key])})")
        settings.update(polysettings)
    return settings


def set_polysettings(rundict, polysettings, ndim, nderived, isodate, parnames, size=1):
    from pypolychord.settings import PolyChordSettings
    settings = default_settings(ndim, polysettings)
    rundict["target"] = rundict["target"].replace(" ", "")
    rundict["runid"] = rundict["runid"].replace(" ", "")
    file_root = rundict["target"] + "_" + rundict["runid"]
    if rundict.get("comment", "") != "":
        file_root += "-" + rundict["comment"]
    if rundict.get("nplanets") is not None:
        file_root += f'_k{rundict["nplanets"]}'
    drift_order = sum(1 for p in parnames
                      if "drift" in p and p[6:] in ("lin", "quad", "cub", "quar"))
    if drift_order > 0:
        file_root += f"_d{drift_order}"
    file_root += f'_nlive{settings["nlive"]}_ncores{size}_polychord_{isodate}'
    base_dir = os.path.join(rundict.get("save_dir", ""), file_root, "polychains")
    settings.update({"file_root": file_root, "base_dir": base_dir})
    return PolyChordSettings(ndim, nderived, **settings)


def run(model, rundict, priordict, polysettings=None):
    try:
        from pypolychord import run_polychord
    except ImportError:
        raise ImportError("Install PolyChord to use this module.")
    parnames = model.parnames
    ndim = len(parnames)
    isodate = datetime.datetime.today().isoformat()
    settings = set_polysettings(rundict, polysettings, ndim, 0, isodate, parnames)
    prior, loglike = make_callbacks(model, priordict)
    return run_polychord(loglike, ndim, 0, settings, prior)

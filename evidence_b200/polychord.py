"""
PolyChord adapter (boundary only).

PolyChord calls the likelihood one point at a time from Fortran, so it cannot use large batches;
it gets the device model through the scalar protocol (a batch of 1 per call).  The entry point,
settings defaults and type checks, output attributes and pickle follow
evidence/polychord/__init__.py:32-258, 261-297, 299-424.  The sampler itself (pypolychord,
Fortran/C++) is third-party and absent from this image: importing this module works, ``run()``
raises ImportError without it (the tests drive it through a test double, tests/doubles/pypolychord).
"""
import datetime
import os
import pickle
import time
from pathlib import Path

import numpy as np

try:  # MPI is optional, as in the reference (:21-29)
    from mpi4py import MPI
    comm = MPI.COMM_WORLD
    rank, size = comm.Get_rank(), comm.Get_size()
except ImportError:
    comm, rank, size = None, 0, 1


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
    """
    ``evidence.polychord.run`` on the device model (evidence/polychord/__init__.py:32-258): same
    signature, returns PolyChord's output object with the reference's extra attributes (:213-234) and
    writes the same pickle (:242, 261-297).  The reference's post-processing (matplotlib) only runs
    when ``rundict['postprocess']`` is true and the reference package is importable.
    """
    try:
        from pypolychord import run_polychord
    except ImportError:
        raise ImportError("Install PolyChord to use this module.")
    parnames = model.parnames
    ndim, nderived = len(parnames), 0
    isodate = datetime.datetime.today().isoformat()
    if size > 1:
        isodate = comm.bcast(isodate, root=0)
    settings = set_polysettings(rundict, polysettings, ndim, nderived, isodate, parnames, size=size)
    print(f'Saving results to {os.path.join(rundict.get("save_dir", ""), settings.file_root)}\n')
    prior, loglike = make_callbacks(model, priordict)
    ti = time.process_time()
    output = run_polychord(loglike, ndim, nderived, settings, prior)
    tf = time.process_time()
    if size > 1:
        ti = comm.reduce(ti, op=MPI.MIN, root=0)
        tf = comm.reduce(tf, op=MPI.MAX, root=0)
    if rank == 0:
        output.make_paramnames_files([(x, x) for x in parnames])  # :203
        output.runtime = datetime.timedelta(seconds=tf - ti)
        output.rundict = rundict.copy()
        output.datadict = dict(getattr(model, "datadict", {}))
        output.fixedpardict = dict(getattr(model, "fixedpardict", {}))
        model_path = getattr(model, "model_path", None)
        output.model_name = str(Path(model_path).stem) if model_path else type(model).__name__
        output.nlive = settings.nlive
        output.nrepeats = settings.num_repeats
        output.isodate = isodate
        output.ncores = size
        output.parnames = parnames
        output.ndim = ndim
        output.sampler = "PolyChord"
        if hasattr(model, "counters"):  # observability added by the device path
            output.device_counters = model.counters()
        if "prior_names" in rundict:
            output.priors = rundict["prior_names"]
        if "star_params" in rundict:
            output.starparams = rundict["star_params"]
        print(f"\nTotal run time was: {output.runtime}")
        dump2pickle_poly(output, output.file_root + ".pkl")
        if rundict.get("postprocess", False):
            from evidence.post_processing import postprocess  # the reference's own (unchanged)
            postprocess(str(Path(output.base_dir).parent.absolute()))
    return output


def dump2pickle_poly(output, filename, savedir=None):
    """Pickle PolyChord's output object next to the chains, evidence/polychord/__init__.py:261-297."""
    pickledir = Path(output.base_dir).parent if savedir is None else savedir
    os.makedirs(pickledir, exist_ok=True)
    with open(os.path.join(pickledir, filename), "wb") as f:
        pickle.dump(output, f)

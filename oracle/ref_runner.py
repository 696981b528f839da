"""
ref_runner.py — TEST INFRASTRUCTURE ONLY.  Not part of the product.

Runs the UNMODIFIED reference implementation of the path — ``evidence.rvmodel.RVModel`` with its
own C Kepler solver — from the copy that ``oracle/Makefile`` stages into the git-ignored
``oracle/_ref/evidence/`` (from /root/reference, in the build container; the staged files travel
to the GPU box with the working tree, the reference checkout itself does not).  Used by the
``cpu_baseline`` / ``--impl reference`` legs of bench.py (``kind: "reference"``) and by tests
that check the restatements in this directory against it.

The only shim is ``numpy.int = int``: evidence/rvmodel/__init__.py:53 uses the alias numpy
removed in 1.24.
"""
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")


def available():
    return (os.path.exists(os.path.join(_REF, "evidence", "rvmodel", "__init__.py"))
            and os.path.exists(os.path.join(_REF, "evidence", "rvmodel", "trueanomaly.so")))


def _module():
    if not available():
        from . import rv_oracle
        rv_oracle.build()  # stages _ref/ when the reference checkout is present
    if not available():
        raise ImportError("oracle/_ref/evidence is not staged (no reference checkout at build time)")
    if not hasattr(np, "int"):
        np.int = int  # evidence/rvmodel/__init__.py:53
    if _REF not in sys.path:
        sys.path.insert(0, _REF)
    import evidence.rvmodel as ref  # the reference's own module, from the staged copy
    assert os.path.realpath(ref.__file__).startswith(os.path.realpath(_REF)), ref.__file__
    return ref


def make_model(fixedpardict, tables, parnames):
    """
    The reference ``RVModel`` (evidence/rvmodel/__init__.py:84-154) on
    ``tables = {inst: {'data': {col: array}}}``; frames are built fresh because
    ``BaseModel.__init__`` adds an ``inst_id`` column to what it is given (:52).
    """
    import pandas as pd
    ref = _module()
    datadict = {inst: {"data": pd.DataFrame({k: np.asarray(v) for k, v in tab["data"].items()
                                             if k != "inst_id"})}
                for inst, tab in tables.items()}
    return ref.RVModel(dict(fixedpardict), datadict, list(parnames))


def loglike_rows(model, X):
    """``RVModel.log_likelihood`` (evidence/rvmodel/__init__.py:157-219), one row at a time: the
    loop a sampler's ``loglike`` closure drives (evidence/ultranest/__init__.py:141-146)."""
    return np.array([model.log_likelihood(x) for x in np.asarray(X, dtype=np.float64)])

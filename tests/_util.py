"""Shared helpers for the test-suite: golden fixtures, oracle models, tolerances."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# The parity bar of BASELINE.json's north_star: |lnL_gpu - lnL_ref| <= 1e-9 absolute on identical
# theta; above |lnL| = 1e4 one ulp of lnL itself approaches that, so the bound becomes relative
# (SURVEY.md H7).
ABS_TOL = 1e-9
REL_TOL = 1e-13


def lnl_close(got, want, abs_tol=ABS_TOL, rel_tol=REL_TOL):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    sentinel = want == -1e30
    ok_sentinel = np.array_equal(got[sentinel], want[sentinel])
    both_nan = np.isnan(got) & np.isnan(want)
    err = np.abs(got - want)
    bound = np.maximum(abs_tol, rel_tol * np.abs(want))
    ok = (err <= bound) | sentinel | both_nan
    worst = float(np.nanmax(np.where(sentinel | both_nan, 0.0, err))) if len(err) else 0.0
    return bool(ok_sentinel and ok.all()), worst


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return meta, z


def tables_of(meta, z):
    """{inst: {'data': {col: array}}} in the fixture's instrument order (a fresh copy)."""
    out = {}
    for inst in meta["insts"]:
        cols = {}
        for key in z.files:
            if key.startswith(f"data__{inst}__"):
                cols[key.split("__")[2]] = np.array(z[key])
        out[inst] = {"data": cols}
    return out


def oracle_model(meta, z, parnames=None, fixed=None):
    from oracle.rv_oracle import OracleRVModel
    return OracleRVModel(meta["fixed"] if fixed is None else fixed, tables_of(meta, z),
                         meta["parnames"] if parnames is None else parnames)


def device_model(meta, z, parnames=None, fixed=None, **kw):
    from evidence_b200.rvmodel import RVModel
    return RVModel(dict(meta["fixed"] if fixed is None else fixed), tables_of(meta, z),
                   list(meta["parnames"] if parnames is None else parnames), **kw)

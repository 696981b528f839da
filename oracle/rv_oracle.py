"""
rv_oracle.py — TEST INFRASTRUCTURE ONLY.  Not part of the product.

CPU restatement (numpy + the C Kepler solver) of the `evidence` RV log-likelihood path, used
as the checker by tests/, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of bench.py.  Nothing under ``evidence_b200/`` may import it.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the live reference
(/root/reference/evidence, in the build container) on the same inputs and the committed
fixtures under tests/golden/ hold the reference's outputs; tests/test_oracle.py checks this
restatement against them.  The reference's own tests hold no golden lnL.

The Kepler solve goes through the reference's own C routine when ``oracle/_ref/trueanomaly.so``
(compiled from /root/reference/evidence/rvmodel/trueanomaly.c by oracle/Makefile) is present,
otherwise through the restatement ``orc_trueanomaly`` in oracle/_build/librvoracle.so.

Citations are relative to the reference checkout.
"""
import ctypes
import os
import subprocess
from ctypes import POINTER, c_double, c_int, c_longlong

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_SO = os.path.join(_HERE, "_ref", "trueanomaly.so")
_ORC_SO = os.path.join(_HERE, "_build", "librvoracle.so")

_dp = POINTER(c_double)


def build(force=False):
    """Compile the C checker (and _ref/ when the reference checkout is present)."""
    if force or not os.path.exists(_ORC_SO) or (
        os.path.exists("/root/reference") and not os.path.exists(_REF_SO)
    ):
        subprocess.run(["make", "-s", "-C", _HERE, "all"], check=True)


_libs = {}


def _orc_lib():
    if "orc" not in _libs:
        build()
        lib = ctypes.CDLL(_ORC_SO)
        lib.orc_trueanomaly.argtypes = [_dp, c_int, c_double, _dp, c_int, c_double,
                                        POINTER(c_longlong)]
        lib.orc_trueanomaly.restype = c_int
        lib.orc_ppf_closed.argtypes = [c_int, _dp, c_double]
        lib.orc_ppf_closed.restype = c_double
        lib.orc_ppf_table.argtypes = [_dp, _dp, c_int, c_double]
        lib.orc_ppf_table.restype = c_double
        _libs["orc"] = lib
    return _libs["orc"]


def _ref_lib():
    """The reference's own solver (evidence/rvmodel/trueanomaly.h:4), or None."""
    if "ref" not in _libs:
        build()
        lib = None
        if os.path.exists(_REF_SO):
            lib = ctypes.CDLL(_REF_SO)
            # same argtypes the reference sets, evidence/rvmodel/__init__.py:151-152
            lib.trueanomaly.argtypes = [_dp, c_int, c_double, _dp, c_int, c_double]
            lib.trueanomaly.restype = c_int
        _libs["ref"] = lib
    return _libs["ref"]


def solver_kind():
    return "reference" if _ref_lib() is not None else "port"


def true_anomaly(ma, ecc, tol=1.0e-4, itmax=10000):
    """evidence/rvmodel/__init__.py:466-494 → trueanomaly.c:8-41.  Return code ignored (:490)."""
    ma = np.ascontiguousarray(ma, dtype=np.float64)
    nu = np.zeros_like(ma)
    ref = _ref_lib()
    if ref is not None:
        ref.trueanomaly(ma.ctypes.data_as(_dp), len(ma), float(ecc), nu.ctypes.data_as(_dp),
                        int(itmax), float(tol))
    else:
        _orc_lib().orc_trueanomaly(ma.ctypes.data_as(_dp), len(ma), float(ecc),
                                   nu.ctypes.data_as(_dp), int(itmax), float(tol), None)
    return nu


class OracleRVModel:
    """
    Restatement of BaseModel/RVModel (evidence/rvmodel/__init__.py:23-57, 94-154).

    ``datadict`` is ``{instrument: {'data': table}}`` with ``table['rjd'|'jdb']``,
    ``table['vrad']``, ``table['svrad']`` array-likes (a pandas DataFrame or a dict), in the
    reference's instrument order.
    """

    def __init__(self, fixedpardict, datadict, parnames, linpar_dict=None):
        self.fixedpardict = dict(fixedpardict)
        self.parnames = sorted(parnames)  # :43
        self.insts = list(datadict.keys())  # :46
        t, v, s, ids = [], [], [], []
        for i, inst in enumerate(self.insts):  # :50-55 instrument-major concatenation
            tab = datadict[inst]["data"]
            try:
                tcol = tab["rjd"]  # :141-144
            except KeyError:
                tcol = tab["jdb"]
            tcol = np.asarray(tcol, dtype=np.float64)
            t.append(tcol)
            v.append(np.asarray(tab["vrad"], dtype=np.float64))
            s.append(np.asarray(tab["svrad"], dtype=np.float64))
            ids.append(np.zeros(len(tcol), dtype=np.int32) + i)
        self.time = np.concatenate(t) if t else np.zeros(0)
        self.vrad = np.concatenate(v) if v else np.zeros(0)
        self.svrad = np.concatenate(s) if s else np.zeros(0)
        self.inst_id = np.concatenate(ids) if ids else np.zeros(0, dtype=np.int32)
        self._masks = [np.where(self.inst_id == i) for i in range(len(self.insts))]
        # structure from the FREE names only, :118-139
        self.nplanets = sum("k1" in p for p in self.parnames)
        self.drift_in_model = any("drift" in p for p in self.parnames)
        self.linpar_in_model = any("linpar" in p for p in self.parnames)
        self.jitter_in_model = any("jitter" in p for p in self.parnames)
        self.linpar_dict = dict(linpar_dict or {})

    # :59-80
    @staticmethod
    def logL(residuals, var):
        n = len(residuals)
        cte = -0.5 * n * np.log(2 * np.pi)
        return cte - np.sum(np.log(np.sqrt(var))) - np.sum(residuals ** 2 / (2 * var))

    # :388-463
    def modelk(self, pardict, time, planet):
        pre = f"planet{planet}_"
        if pre + "k1" in pardict:
            amp = pardict[pre + "k1"]
        else:
            amp = np.exp(pardict[pre + "logk1"])
        if pre + "period" in pardict:
            per = pardict[pre + "period"]
        else:
            per = np.exp(pardict[pre + "logperiod"])
        if pre + "secos" in pardict:  # :425-431
            c, s = pardict[pre + "secos"], pardict[pre + "sesin"]
            ecc = c ** 2 + s ** 2
            omega = np.arctan2(s, c)
            if ecc > 1:
                return None
        elif pre + "ecos" in pardict:  # :433-439
            c, s = pardict[pre + "ecos"], pardict[pre + "esin"]
            ecc = np.sqrt(c ** 2 + s ** 2)
            omega = np.arctan2(s, c)
            if ecc > 1:
                return None
        else:  # :441-447
            ecc = pardict[pre + "ecc"]
            omega = pardict[pre + "omega"]
        if pre + "ml0" in pardict:  # :449-454
            ma0 = pardict[pre + "ml0"] - omega
        else:
            ma0 = pardict[pre + "ma0"]
        epoch = pardict[pre + "epoch"]
        ma = 2 * np.pi / per * (time - epoch) + ma0  # :459
        nu = true_anomaly(ma, ecc)  # :461
        return amp * (np.cos(nu + omega) + ecc * np.cos(omega))  # :463

    # :343-385
    def kep_rv(self, pardict, time):
        rows = np.zeros((self.nplanets, len(time)))
        for k in range(1, self.nplanets + 1):
            r = self.modelk(pardict, time, k)
            if r is None:
                return None
            rows[k - 1] = r
        return rows.sum(axis=0)

    # :222-273
    @staticmethod
    def drift(pardict, time):
        lin = pardict.get("drift_lin", 0.0)
        quad = pardict.get("drift_quad", 0.0)
        cub = pardict.get("drift_cub", 0.0)
        quar = pardict.get("drift_quar", 0.0)
        tref = pardict["drift_tref"] if "drift_tref" in pardict else time[0]
        tt = (time - tref) / 365.25
        return lin * tt + quad * tt ** 2 + cub * tt ** 3 + quar * tt ** 4

    # :157-219
    def log_likelihood(self, x):
        pardict = {p: x[i] for i, p in enumerate(self.parnames)}
        pardict.update(self.fixedpardict)
        noise = np.zeros_like(self.svrad)
        rvm = np.zeros_like(self.vrad)
        for i, inst in enumerate(self.insts):
            idx = self._masks[i]
            rvm[idx] += pardict[f"{inst}_offset"]
            if self.jitter_in_model:
                noise[idx] = self.svrad[idx] ** 2 + pardict[f"{inst}_jitter"] ** 2
            else:
                noise[idx] = self.svrad[idx] ** 2
        if self.nplanets > 0:
            pred = self.kep_rv(pardict, self.time)
            if pred is None:
                return -1e30  # :203
            rvm += pred
        if self.drift_in_model:
            rvm += self.drift(pardict, self.time)
        if self.linpar_in_model:
            for name in self.linpar_dict:
                rvm += pardict[f"linpar_{name}"] * self.linpar_dict[name]
        return self.logL(self.vrad - rvm, noise)

    def log_likelihood_batch(self, X):
        X = np.asarray(X, dtype=np.float64)
        return np.array([self.log_likelihood(X[b]) for b in range(X.shape[0])])


# ------------------------------------------------------------------------------------------
# prior transform (evidence/ultranest/__init__.py:125-137; evidence/priors.py)
# ------------------------------------------------------------------------------------------
_CLOSED = {"Uniform": 0, "Jeffreys": 1, "ModJeffreys": 2, "UniformFrequency": 3,
           "TruncatedRayleigh": 4}


def ppf_closed(name, pars, q):
    """Closed-form inverse CDFs, evidence/priors.py:41-42, 62-63, 82-83, 100-101, 249-252."""
    p = (c_double * 4)(*([float(v) for v in pars] + [0.0] * (4 - len(pars))))
    lib = _orc_lib()
    q = np.atleast_1d(np.asarray(q, dtype=np.float64))
    return np.array([lib.orc_ppf_closed(_CLOSED[name], p, float(v)) for v in q])


def ppf_table(cdf, x, q):
    """interp1d(cdf, x)(q), evidence/priors.py:124 (and :202, :228, :287, :326, :354)."""
    cdf = np.ascontiguousarray(cdf, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    lib = _orc_lib()
    q = np.atleast_1d(np.asarray(q, dtype=np.float64))
    return np.array([lib.orc_ppf_table(cdf.ctypes.data_as(_dp), x.ctypes.data_as(_dp),
                                       len(x), float(v)) for v in q])


# ------------------------------------------------------------------------------------------
# batched plain-C evaluation of the whole path (fast checker for large parity cases)
# ------------------------------------------------------------------------------------------
def c_loglike_batch(desc_bytes, t, rv, err, inst, n_inst, theta, linpar_cols=()):
    """
    Whole-path C restatement (oracle/rvlnl_oracle.c: orc_loglike_batch) on a flattened model
    description (the bytes of an ``rvl_model_desc``).  Returns (lnL[B], newton_iters, cap_hits).
    """
    lib = _orc_lib()
    t = np.ascontiguousarray(t, dtype=np.float64)
    rv = np.ascontiguousarray(rv, dtype=np.float64)
    err = np.ascontiguousarray(err, dtype=np.float64)
    inst = np.ascontiguousarray(inst, dtype=np.int32)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    B = theta.shape[0]
    out = np.empty(B, dtype=np.float64)
    iters = c_longlong(0)
    caps = c_longlong(0)
    cols = [np.ascontiguousarray(c, dtype=np.float64) for c in linpar_cols]
    colptr = (_dp * max(1, len(cols)))(*[c.ctypes.data_as(_dp) for c in cols])
    buf = ctypes.create_string_buffer(bytes(desc_bytes), len(desc_bytes))
    fn = lib.orc_loglike_batch
    fn.restype = c_int
    fn.argtypes = [ctypes.c_void_p, _dp, _dp, _dp, POINTER(ctypes.c_int32), c_int, c_int,
                   POINTER(_dp), _dp, c_longlong, _dp, POINTER(c_longlong), POINTER(c_longlong)]
    rc = fn(ctypes.cast(buf, ctypes.c_void_p), t.ctypes.data_as(_dp), rv.ctypes.data_as(_dp),
            err.ctypes.data_as(_dp), inst.ctypes.data_as(POINTER(ctypes.c_int32)), len(t),
            int(n_inst), colptr, theta.ctypes.data_as(_dp), B, out.ctypes.data_as(_dp),
            ctypes.byref(iters), ctypes.byref(caps))
    if rc != 0:
        raise RuntimeError("orc_loglike_batch failed")
    return out, iters.value, caps.value

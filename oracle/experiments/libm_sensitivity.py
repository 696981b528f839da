"""
libm_sensitivity.py — TEST INFRASTRUCTURE / evidence for DESIGN.md section 3.  Not the product.

Question: the GPU parity tests allow 1e-5 instead of 1e-9 for e > 0.97
because "Newton from E = M is chaotic there and follows the last ulp of libm's sin/cos".  Is that
the reference's OWN platform sensitivity?  Experiment: the C restatement of the reference path
(oracle/rvlnl_oracle.c, bit-identical to the live reference on every BASELINE shape) is built
twice -- with the C library's double sin/cos (what the reference's trueanomaly.c calls), and with
sinl/cosl evaluated in 80-bit extended precision and rounded to double (a different, slightly more
accurate libm: the two agree except for a last-bit difference on a small fraction of arguments).
Both run the same theta; everything else is identical code.  Reference-vs-reference spread:

    python oracle/experiments/libm_sensitivity.py  >  profiles/r2_libm_sensitivity.txt
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from evidence_b200 import synth  # noqa: E402  (host-side data generator only)
from evidence_b200.layout import compile_model  # noqa: E402

dp = ctypes.POINTER(ctypes.c_double)


def build(tag, defs):
    so = os.path.join(ROOT, "oracle", "_build", f"librvoracle_{tag}.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", *defs, "-o", so,
                    os.path.join(ROOT, "oracle", "rvlnl_oracle.c"), "-lm"], check=True)
    lib = ctypes.CDLL(so)
    lib.orc_loglike_batch.restype = ctypes.c_int
    return lib


def run(lib, desc, t, v, s, ids, n_inst, theta):
    theta = np.ascontiguousarray(theta)
    out = np.empty(len(theta))
    iters, caps = ctypes.c_longlong(0), ctypes.c_longlong(0)
    buf = ctypes.create_string_buffer(bytes(desc), ctypes.sizeof(desc))
    cols = (dp * 1)()
    rc = lib.orc_loglike_batch(ctypes.cast(buf, ctypes.c_void_p), t.ctypes.data_as(dp),
                               v.ctypes.data_as(dp), s.ctypes.data_as(dp),
                               ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.c_int(len(t)),
                               ctypes.c_int(n_inst), cols, theta.ctypes.data_as(dp),
                               ctypes.c_longlong(len(theta)), out.ctypes.data_as(dp),
                               ctypes.byref(iters), ctypes.byref(caps))
    assert rc == 0
    return out, iters.value


def main():
    libm = build("libm", [])
    ext = build("ext", ["-DORC_SIN(x)=((double)sinl((long double)(x)))",
                        "-DORC_COS(x)=((double)cosl((long double)(x)))"])
    ulp16 = build("ulp16", ["-DORC_PERTURB", "-DORC_SIN(x)=orc_perturb(sin(x))", "-DORC_COS(x)=orc_perturb(cos(x))"])
    # how different are the two trig implementations themselves?
    x = np.random.default_rng(0).uniform(-2.0e4, 2.0e4, 400000)
    sl = np.sin(x.astype(np.longdouble)).astype(np.float64)
    print("# reference path with two libms: C double sin/cos  vs  80-bit sinl/cosl rounded to double")
    print(f"sin(x), x ~ U(-2e4, 2e4): the two differ (by one ulp) on {np.mean(np.sin(x) != sl) * 100:.3f} % "
          f"of 400000 arguments")
    case = synth.make_case(2)  # N = 1000, K = 2, 2 instruments, linear drift
    t, v, s, ids = case.arrays()
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    desc, _ = compile_model(case.parnames, case.fixedpardict, case.insts, t[0])
    cols = [case.parnames.index(f"planet{k}_ecc") for k in (1, 2)]
    rng = np.random.default_rng(7)
    print("third build: the C library's sin/cos with the last bit flipped on 1/16 of the results "
          "(a libm that is <= 1 ulp like CUDA's or this kernel's own)")
    print(f"{'e of planet 1':>22s} {'rows':>5s} {'iters/solve':>11s} {'rows differing':>14s} "
          f"{'median |dlnL|':>13s} {'max |dlnL|':>11s} | {'1/16-ulp libm: rows':>19s} {'max |dlnL|':>11s}")
    for lo, hi in ((0.0, 0.9), (0.90, 0.95), (0.95, 0.97), (0.97, 0.98), (0.98, 0.99), (0.99, 1.0)):
        theta = case.draw_theta(256, seed=int(lo * 1000) + 11)
        theta[:, cols[0]] = rng.uniform(lo, hi, len(theta))
        a, it = run(libm, desc, t, v, s, ids, case.n_inst, theta)
        b, _ = run(ext, desc, t, v, s, ids, case.n_inst, theta)
        c, _ = run(ulp16, desc, t, v, s, ids, case.n_inst, theta)
        d, d2 = np.abs(a - b), np.abs(a - c)
        print(f"{f'[{lo:.2f}, {hi:.2f})':>22s} {len(theta):5d} {it / (len(theta) * len(t) * 2):11.2f} "
              f"{int((d > 0).sum()):14d} {np.median(d):13.3e} {d.max():11.3e} | "
              f"{int((d2 > 0).sum()):19d} {d2.max():11.3e}")
    print("(same code, same theta, same data: only the last bit of some sin/cos values differs)")


if __name__ == "__main__":
    main()

"""
make_golden.py — TEST INFRASTRUCTURE ONLY.

Generates the committed fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference/evidence, imported in place; it cannot travel to the GPU box) on seeded inputs:

  kat_51peg.npz      Appendix-B known-answer cases on the reference's real 51 Peg fixture
                     (tests/test_examples/51Peg/51Peg.rv) -- every modelk parametrisation branch,
                     e > 0.99 clamp, e > 1 -> -1e30, fixed-jitter-is-ignored, drift tref default
  cfg{1,2,3,5}.npz   BASELINE.json shapes: data + theta batch + reference lnL
  edge_*.npz         high-eccentricity / zero-jitter / mixed parametrisation batches
  priors.npz         reference ppf values of every named distribution of evidence/priors.py
  trueanomaly.npz    outputs of the reference's shipped trueanomaly.so

Run in the build container:   python oracle/make_golden.py
The only compatibility shim is ``numpy.int = int`` (evidence/rvmodel/__init__.py:53 uses the
removed alias).
"""
import json
import os
import sys
import warnings

import numpy as np

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
np.int = int  # noqa: shim, see module docstring

import pandas as pd  # noqa: E402
from evidence import priors as ref_priors  # noqa: E402
from evidence.rvmodel import RVModel as RefRVModel  # noqa: E402

from evidence_b200 import synth  # noqa: E402

warnings.filterwarnings("ignore")


def ref_model(fixed, tables, parnames):
    """tables: {inst: {'rjd'|'jdb':..., 'vrad':..., 'svrad':...}} -> reference RVModel."""
    datadict = {k: {"data": pd.DataFrame({c: np.array(v) for c, v in t.items()})}
                for k, t in tables.items()}
    return RefRVModel(dict(fixed), datadict, list(parnames))


def ref_lnl(model, theta):
    return np.array([model.log_likelihood(np.array(row)) for row in np.atleast_2d(theta)])


def save(name, meta, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, meta=np.array(json.dumps(meta)), **arrays)
    print(f"{name:18s} {os.path.getsize(path) / 1024:8.1f} KiB")


def pack_tables(tables):
    arrays, insts = {}, []
    for inst, t in tables.items():
        insts.append(inst)
        for c, v in t.items():
            arrays[f"data__{inst}__{c}"] = np.asarray(v, dtype=np.float64)
    return insts, arrays


# ------------------------------------------------------------------------------------------
def kat_51peg():
    raw = pd.read_csv(os.path.join(REF, "tests/test_examples/51Peg/51Peg.rv"), sep="\t",
                      skiprows=(1,))
    tables = {"hamilton": {"rjd": raw["rjd"].values, "vrad": raw["vrad"].values,
                           "svrad": raw["svrad"].values}}
    A = ["hamilton_jitter", "hamilton_offset", "planet1_ecc", "planet1_k1", "planet1_ma0",
         "planet1_omega", "planet1_period"]
    fA = {"planet1_epoch": 51050}
    fE = {"planet1_epoch": 51050.0, "hamilton_offset": -2.0}
    cases = [
        ("A1", A, fA, [5.0, -2.0, 0.05, 56.0, 1.0, 0.5, 4.2308]),
        ("A2_e0", A, fA, [5.0, -2.0, 0.0, 56.0, 1.0, 0.5, 4.2308]),
        ("A3_e0.9", A, fA, [5.0, -2.0, 0.9, 56.0, 1.0, 0.5, 4.2308]),
        ("A4_e0.995_clamp", A, fA, [5.0, -2.0, 0.995, 56.0, 1.0, 0.5, 4.2308]),
        ("A5_jit0_P1", A, fA, [0.0, -2.0, 0.3, 56.0, 1.0, 0.5, 1.0]),
        ("A6_e_negative", A, fA, [5.0, -2.0, -0.2, 56.0, 1.0, 0.5, 4.2308]),
        ("B1_0planets", ["hamilton_jitter", "hamilton_offset"], {}, [5.0, -2.0]),
        ("C1_lindrift", ["drift_lin"] + A, {"drift_tref": 51050, "planet1_epoch": 51050},
         [-3.0, 5.0, -2.0, 0.05, 56.0, 1.0, 0.5, 4.2308]),
        ("C2_driftcfg", ["drift_lin"] + A, {"drift_tref": 51050, "planet1_epoch": 51050},
         [-26.0, 18.5, -2.5999999999999996, 0.11465599401597196, 74.0, 2.324778563656447,
          2.324778563656447, 1.5863699097355521]),
        ("D1_2planets", A + ["planet2_ecc", "planet2_k1", "planet2_ma0", "planet2_omega",
                             "planet2_period"], {"planet1_epoch": 51050, "planet2_epoch": 51050},
         [5.0, -2.0, 0.05, 56.0, 1.0, 0.5, 4.2308, 0.4, 10.0, 2.0, 3.0, 37.5]),
        ("E1_secos_ml0_logs", ["hamilton_jitter", "planet1_logk1", "planet1_logperiod",
                               "planet1_ml0", "planet1_secos", "planet1_sesin"], fE,
         [5.0, 4.02535169073515, 1.442391100471761, 2.0, 0.3, -0.4]),
        ("E2_secos_invalid", ["hamilton_jitter", "planet1_k1", "planet1_ml0", "planet1_period",
                              "planet1_secos", "planet1_sesin"], fE,
         [5.0, 56.0, 2.0, 4.2308, 0.8, -0.7]),
        ("E3_ecos_ma0", ["hamilton_jitter", "planet1_ecos", "planet1_esin", "planet1_k1",
                         "planet1_ma0", "planet1_period"], fE, [5.0, 0.3, -0.4, 56.0, 2.0, 4.2308]),
        ("E3b_ecos_invalid", ["hamilton_jitter", "planet1_ecos", "planet1_esin", "planet1_k1",
                              "planet1_ma0", "planet1_period"], fE,
         [5.0, 0.9, -0.6, 56.0, 2.0, 4.2308]),
        ("E4_jitter_fixed", ["planet1_ecc", "planet1_k1", "planet1_ma0", "planet1_omega",
                             "planet1_period"], dict(fE, hamilton_jitter=5.0),
         [0.1, 56.0, 2.0, 0.5, 4.2308]),
        ("E5_jitter_absent", ["planet1_ecc", "planet1_k1", "planet1_ma0", "planet1_omega",
                              "planet1_period"], fE, [0.1, 56.0, 2.0, 0.5, 4.2308]),
        ("E6_drift4_tref_default", ["drift_cub", "drift_lin", "drift_quad", "drift_quar",
                                    "hamilton_jitter"], {"hamilton_offset": -2.0},
         [-0.1, -3.0, 0.5, 0.01, 5.0]),
        ("E7_fixed_wins", A, dict(fA, planet1_k1=30.0), [5.0, -2.0, 0.05, 56.0, 1.0, 0.5, 4.2308]),
    ]
    meta_cases, lnls = [], []
    for name, parnames, fixed, theta in cases:
        m = ref_model(fixed, tables, parnames)
        assert list(m.parnames) == sorted(parnames)
        order = [parnames.index(p) for p in m.parnames]
        th = [theta[i] for i in order]
        val = float(m.log_likelihood(np.array(th)))
        meta_cases.append({"name": name, "parnames": list(m.parnames),
                           "fixed": {k: float(v) for k, v in fixed.items()}, "theta": th})
        lnls.append(val)
        print(f"   {name:24s} {val!r}")
    insts, arrays = pack_tables(tables)
    save("kat_51peg", {"insts": insts, "cases": meta_cases}, lnl=np.array(lnls), **arrays)


def config_case(cfg, B, seed=11, **over):
    case = synth.make_case(cfg, **over)
    tables = {k: v["data"] for k, v in case.datadict().items()}
    m = ref_model(case.fixedpardict, tables, case.parnames)
    assert list(m.parnames) == case.parnames
    theta = case.draw_theta(B, seed=seed)
    lnl = ref_lnl(m, theta)
    insts, arrays = pack_tables(tables)
    save(f"cfg{cfg}", {"insts": insts, "parnames": case.parnames,
                       "fixed": case.fixedpardict, "config": cfg,
                       "prior_spec": {k: list(v) for k, v in case.prior_spec.items()}},
         theta=theta, lnl=lnl, **arrays)


def edge_cases():
    # (a) high eccentricity 0.9..1.05 (direct parametrisation, so > 0.99 is clamped not invalid),
    #     zero jitter for a third of the rows: large |lnL|, long Newton tails
    case = synth.make_case(2, seed=21, n_epochs=500)
    tables = {k: v["data"] for k, v in case.datadict().items()}
    m = ref_model(case.fixedpardict, tables, case.parnames)
    rng = np.random.default_rng(5)
    theta = case.draw_theta(96, seed=3)
    names = case.parnames
    for k in (1, 2):
        theta[:, names.index(f"planet{k}_ecc")] = rng.uniform(0.9, 1.05, len(theta))
    for inst in case.insts:
        theta[::3, names.index(f"{inst}_jitter")] = 0.0
    insts, arrays = pack_tables(tables)
    save("edge_highecc", {"insts": insts, "parnames": names, "fixed": case.fixedpardict},
         theta=theta, lnl=ref_lnl(m, theta), **arrays)

    # (b) mixed parametrisations: planet1 secos/sesin+ml0+logk1+logperiod, planet2 ecos/esin+ma0,
    #     planet3 direct; fixed epoch differs per planet; quadratic drift with default tref;
    #     two instruments, one offset fixed; includes e > 1 rows (-1e30)
    rng = np.random.default_rng(8)
    n = 300
    t = np.sort(rng.uniform(53000, 54500, n))
    cut = 170
    tables = {"harps": {"jdb": t[:cut], "vrad": rng.normal(0, 8, cut), "svrad": rng.uniform(0.4, 1.5, cut)},
              "coralie": {"jdb": t[cut:], "vrad": rng.normal(3, 8, n - cut),
                          "svrad": rng.uniform(2, 6, n - cut)}}
    parnames = ["planet1_logk1", "planet1_logperiod", "planet1_secos", "planet1_sesin",
                "planet1_ml0", "planet2_k1", "planet2_period", "planet2_ecos", "planet2_esin",
                "planet2_ma0", "planet3_k1", "planet3_period", "planet3_ecc", "planet3_omega",
                "planet3_ma0", "harps_jitter", "coralie_jitter", "coralie_offset", "drift_lin",
                "drift_quad"]
    fixed = {"planet1_epoch": 53500.0, "planet2_epoch": 53750.5, "planet3_epoch": 54000.0,
             "harps_offset": 1.25}
    m = ref_model(fixed, tables, parnames)
    names = list(m.parnames)
    B = 128
    th = np.empty((B, len(names)))
    draw = {"planet1_logk1": lambda: rng.uniform(-1, 3, B), "planet1_logperiod": lambda: rng.uniform(0.5, 6, B),
            "planet1_secos": lambda: rng.uniform(-0.8, 0.8, B), "planet1_sesin": lambda: rng.uniform(-0.8, 0.8, B),
            "planet1_ml0": lambda: rng.uniform(0, 2 * np.pi, B),
            "planet2_k1": lambda: rng.uniform(0, 15, B), "planet2_period": lambda: rng.uniform(2, 500, B),
            "planet2_ecos": lambda: rng.uniform(-0.75, 0.75, B), "planet2_esin": lambda: rng.uniform(-0.75, 0.75, B),
            "planet2_ma0": lambda: rng.uniform(0, 2 * np.pi, B),
            "planet3_k1": lambda: rng.uniform(0, 15, B), "planet3_period": lambda: rng.uniform(1.1, 50, B),
            "planet3_ecc": lambda: rng.uniform(0, 0.9, B), "planet3_omega": lambda: rng.uniform(0, 2 * np.pi, B),
            "planet3_ma0": lambda: rng.uniform(0, 2 * np.pi, B),
            "harps_jitter": lambda: rng.uniform(0, 5, B), "coralie_jitter": lambda: rng.uniform(0, 5, B),
            "coralie_offset": lambda: rng.uniform(-5, 5, B),
            "drift_lin": lambda: rng.uniform(-2, 2, B), "drift_quad": lambda: rng.uniform(-0.5, 0.5, B)}
    for i, p in enumerate(names):
        th[:, i] = draw[p]()
    lnl = ref_lnl(m, th)
    print(f"   edge_mixed: {np.sum(lnl == -1e30)} invalid rows of {B}")
    insts, arrays = pack_tables(tables)
    save("edge_mixed", {"insts": insts, "parnames": names, "fixed": fixed},
         theta=th, lnl=lnl, **arrays)


def prior_vectors():
    q = np.concatenate([[0.0, 1e-12, 1e-6, 1e-3], np.linspace(0.01, 0.99, 50), [0.999, 1 - 1e-9]])
    specs = [("Uniform", (4.0, 6.0)), ("Uniform", (-10.0, 10.0)), ("Jeffreys", (10.0, 100.0)),
             ("Jeffreys", (0.1, 100.0)), ("ModJeffreys", (1.0, 100.0)),
             ("UniformFrequency", (1.0, 100.0)), ("UniformFrequency", (1.0, 1000.0)),
             ("TruncatedRayleigh", (0.2, 1.0)), ("Normal", (3.0, 0.5)), ("LogNormal", (0.5, 0.0, 2.0)),
             ("Binormal", (0.0, 1.0, 4.0, 0.5, 0.3)), ("AsymmetricNormal", (1.0, 0.5, 2.0)),
             ("TruncatedUNormal", (0.0, 1.0, -1.0, 2.0)), ("PowerLaw", (-0.5, 1.0, 10.0)),
             ("DoublePowerLaw", (-0.5, -2.0, 3.0, 1.0, 10.0)), ("Sine", (0.0, 180.0)),
             ("Alpha", (2.5,)), ("Beta", (0.867, 3.03)), ("Gamma", (2.0, 0.5))]
    meta, vals = [], []
    for name, pars in specs:
        dist = getattr(ref_priors, name)(*pars)
        out = np.full(len(q), np.nan)
        for i, qi in enumerate(q):
            try:
                out[i] = float(dist.ppf(qi))
            except Exception:  # e.g. interp1d bounds (TruncatedUNormal at the ends)
                pass
        meta.append({"name": name, "pars": list(pars)})
        vals.append(out)
    # the two values the reference's own tests pin (tests/test_priors.py:11-15, 28)
    pins = {"Uniform(4,6).ppf(0.5)": float(ref_priors.Uniform(4, 6).ppf(0.5)),
            "Uniform(4,6).pdf(5)": float(ref_priors.Uniform(4, 6).pdf(5)),
            "Jeffreys(10,100).pdf(10)": float(ref_priors.Jeffreys(10, 100).pdf(10))}
    save("priors", {"specs": meta, "pins": pins}, q=q, ppf=np.array(vals))


def trueanomaly_vectors():
    from ctypes import POINTER, c_double, c_int, cdll
    lib = cdll.LoadLibrary(os.path.join(REF, "evidence/rvmodel/trueanomaly.so"))
    dp = POINTER(c_double)
    lib.trueanomaly.argtypes = [dp, c_int, c_double, dp, c_int, c_double]
    rng = np.random.default_rng(17)
    eccs = [0.0, 0.05, 0.3, 0.6, 0.9, 0.95, 0.985, 0.99, 0.995, -0.2]
    Ms, nus = [], []
    for e in eccs:
        M = np.concatenate([rng.uniform(-50, 50, 96), rng.uniform(0, 2 * np.pi, 96),
                            rng.uniform(5000, 16000, 64)])
        nu = np.zeros_like(M)
        rc = lib.trueanomaly(M.ctypes.data_as(dp), len(M), e, nu.ctypes.data_as(dp), 10000, 1e-4)
        assert rc == 0
        Ms.append(M)
        nus.append(nu)
    save("trueanomaly", {"eccs": eccs, "tol": 1e-4, "itmax": 10000}, M=np.array(Ms), nu=np.array(nus))


def linpar_smoother_vectors():
    """RVModel.linear_parameter of the reference (evidence/rvmodel/__init__.py:276-340)."""
    rng = np.random.default_rng(3)
    t = np.sort(rng.uniform(0, 900, 120))
    ind = rng.normal(0, 1, 120)
    m = ref_model({"a_offset": 0.0}, {"a": {"rjd": t, "vrad": ind, "svrad": np.ones(120)}}, ["a_jitter"])
    out = {}
    for k in (None, "gaussian", "epanechnikov"):  # 'box' raises under numpy >= 2 (bool /= float)
        for ft in ("lp", "hp"):
            out[f"{k}_{ft}"] = m.linear_parameter(t, ind, kernel=k, timescale=0.3, filter_type=ft)
    np.savez_compressed(os.path.join(OUT, "linpar_smoother.npz"), t=t, ind=ind, **out)
    print("linpar_smoother")


if __name__ == "__main__":
    kat_51peg()
    config_case(1, 256)
    config_case(2, 256)
    config_case(3, 96)
    config_case(5, 24)
    edge_cases()
    prior_vectors()
    trueanomaly_vectors()
    linpar_smoother_vectors()

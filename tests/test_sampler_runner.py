"""
CPU tests: the seeded vectorised nested sampler and the runner plumbing, on the analytic
problems the reference's own integration tests use (tests/test_polychord.py:75-151 of the
reference: unit Gaussian exp(-x^2/2) with U(-10,10) priors, ln Z = d ln(sqrt(2 pi)/20)
= -2.0768 (1-D), -4.1536 (2-D), tolerance 0.5).
"""
import os
import pickle

import numpy as np
import pytest

from evidence_b200 import priors
from evidence_b200 import ultranest as runner
from evidence_b200.sampler import nested_sample


class GaussianModel:
    """The reference's toy model protocol (tests/test_examples/gaussian/model_gaussian_example.py)."""

    def __init__(self, ndim):
        self.parnames = sorted(f"par_x{i + 1}" for i in range(ndim))
        self.datadict, self.fixedpardict = {}, {}

    def log_likelihood(self, x):
        return -0.5 * np.sum((np.asarray(x) - 0.0) ** 2)


@pytest.mark.parametrize("method", ["slice", "ellipsoid"])
@pytest.mark.parametrize("ndim,want", [(1, -2.0768), (2, -4.1536)])
def test_analytic_gaussian_evidence(ndim, want, method):
    res = nested_sample(lambda th: -0.5 * np.sum(th ** 2, axis=1), lambda u: -10 + 20 * u, ndim,
                        nlive=400, ndraw=2048, seed=3, method=method)
    assert abs(res.logz - want) < 0.5  # the reference's own bar
    assert abs(res.logz - want) < 4 * res.logzerr + 0.05
    assert res.samples.shape[1] == ndim and abs(np.mean(res.samples)) < 0.3


def test_sampler_is_deterministic_for_a_seed():
    f = lambda th: -0.5 * np.sum(th ** 2, axis=1)  # noqa: E731
    g = lambda u: -10 + 20 * u  # noqa: E731
    a = nested_sample(f, g, 2, nlive=100, ndraw=512, seed=7)
    b = nested_sample(f, g, 2, nlive=100, ndraw=512, seed=7)
    assert a.logz == b.logz and a.ncall == b.ncall
    c = nested_sample(f, g, 2, nlive=100, ndraw=512, seed=8)
    assert c.logz != a.logz


def test_runner_output_contract(tmp_path):
    model = GaussianModel(2)
    priordict = {p: priors.Uniform(-10, 10) for p in model.parnames}
    rundict = {"target": "gauss ian", "runid": "2 d", "save_dir": str(tmp_path), "nplanets": 0}
    out = runner.run(model, rundict, priordict, {"nlive": 100, "sampler": "builtin", "seed": 5,
                                                 "ndraw_min": 512})
    assert abs(out.logZ + 4.1536) < 0.5
    for attr in ("runtime", "rundict", "fixedpardict", "model_name", "nlive", "nrepeats",
                 "isodate", "ncores", "parnames", "ndim", "sampler", "base_dir", "file_root",
                 "logZ", "logZerr", "nlike", "samples"):
        assert hasattr(out, attr), attr  # evidence/ultranest/__init__.py:200-229
    assert out.file_root.startswith("gaussian_2d_k0_nlive100_ncores1_ultranest_")
    assert list(out.samples.columns) == model.parnames
    pkl = os.path.join(os.path.dirname(out.base_dir), out.file_root + ".pkl")
    back = pickle.load(open(pkl, "rb"))  # the OBJECT, as the reference pickles it (:262-297)
    assert back.logZ == out.logZ and back.file_root == out.file_root and back.rundict == out.rundict
    assert hasattr(back, "datadict") and back.sampler == "UltraNest" and "evidence_b200.sampler" in back.sampler_impl
    # what the reference's post-processing reads for an UltraNest run (post_processing.py:85-88)
    import pandas as pd
    wp = pd.read_csv(os.path.join(out.base_dir, "run1/chains/weighted_post.txt"), sep=" ")
    assert list(wp.columns) == ["weight", "logl"] + model.parnames
    assert abs(wp["weight"].sum() - 1.0) < 1e-9 and np.all(np.diff(wp["logl"]) >= 0)


@pytest.fixture
def ultranest_double(monkeypatch):
    """The test double of the absent third-party package (tests/doubles/ultranest)."""
    import sys
    monkeypatch.syspath_prepend(os.path.join(os.path.dirname(os.path.abspath(__file__)), "doubles"))
    for name in [m for m in sys.modules if m == "ultranest" or m.startswith("ultranest.")]:
        monkeypatch.delitem(sys.modules, name)
    import ultranest
    assert ultranest.__version__.endswith("test-double")
    yield ultranest
    for name in [m for m in sys.modules if m == "ultranest" or m.startswith("ultranest.")]:
        sys.modules.pop(name, None)


@pytest.mark.parametrize("stepsampler,want_sizes", [("none", "region"), ("population-slice", "pop"),
                                                    ("region-slice", "one")])
def test_runner_ultranest_branch_through_the_double(tmp_path, ultranest_double, stepsampler, want_sizes):
    """The `which == "ultranest"` branch (evidence/ultranest/__init__.py:165-185 with the
    `vectorized` switch of :171 turned on): constructor keywords, step-sampler choice, run call,
    results, pickle and weighted_post, against UltraNest's calling convention."""
    seen = {}
    real_ctor = ultranest_double.ReactiveNestedSampler

    class Spy(real_ctor):
        def __init__(self, *a, **kw):
            seen["kw"] = kw
            seen["names"] = a[0]
            super().__init__(*a, **kw)
            seen["sampler"] = self
    ultranest_double.ReactiveNestedSampler = Spy
    model = GaussianModel(2)
    priordict = {p: priors.Uniform(-10, 10) for p in model.parnames}
    rundict = {"target": "gauss", "runid": "un", "save_dir": str(tmp_path), "nplanets": 0}
    out = runner.run(model, rundict, priordict, {"nlive": 100, "sampler": "auto", "nsteps": 4,
                                                 "ndraw_min": 64, "ndraw_max": 256,
                                                 "stepsampler": stepsampler})
    kw = seen["kw"]
    assert kw["vectorized"] is True and kw["ndraw_min"] == 64 and kw["ndraw_max"] == 256
    assert kw["num_test_samples"] == 100 and kw["num_bootstraps"] == 30  # :165-172
    assert list(kw["wrapped_params"]) == [False, False] and seen["names"] == model.parnames
    assert kw["log_dir"].endswith("ultraresults")
    sizes = np.array(seen["sampler"].call_sizes)
    if want_sizes == "region":     # region sampling: ndraw_min .. ndraw_max candidates per call
        assert sizes[0] == 100 and 1 <= sizes[1:].min() and sizes[1:].max() <= 256 and np.median(sizes[1:]) > 20
    elif want_sizes == "pop":      # population step sampler: up to popsize = ndraw_min walkers per call
        assert sizes[0] == 100 and sizes[1:].max() == 64 and sizes[1:].mean() > 8
    else:                          # the reference's RegionSliceSampler: one point per call
        assert sizes[0] == 100 and sizes[1:].max() == 1
    assert abs(out.logZ + 4.1536) < 0.5 and out.sampler == "UltraNest" and "test-double" in out.sampler_impl
    assert out.nlike == int(sizes.sum())
    back = pickle.load(open(os.path.join(os.path.dirname(out.base_dir), out.file_root + ".pkl"), "rb"))
    assert back.logZ == out.logZ
    assert os.path.exists(os.path.join(out.base_dir, "run1/chains/weighted_post.txt"))


def test_ultranest_double_rejects_wrong_shapes(ultranest_double):
    with pytest.raises(ValueError, match="loglikelihood"):
        ultranest_double.ReactiveNestedSampler(["a", "b"], lambda th: np.zeros((len(th), 1)),
                                               lambda u: u, vectorized=True)
    with pytest.raises(ValueError, match="transform"):
        ultranest_double.ReactiveNestedSampler(["a", "b"], lambda th: np.zeros(len(th)),
                                               lambda u: u[:, :1], vectorized=True)


def test_settings_defaults_match_the_reference():
    s = runner.set_ultrasettings({"target": "t", "runid": "r"}, None, 7, 0, "D",
                                 ["drift_lin", "drift_quad", "a"])
    assert (s["nlive"], s["nsteps"], s["dlogz"], s["frac_remain"], s["num_bootstraps"]) == \
        (175, 21, 0.5, 0.01, 30)  # :333-338
    assert "_d2_" in s["file_root"] and s["log_dir"].endswith("ultraresults")
    with pytest.raises(TypeError):
        runner.set_ultrasettings({"target": "t", "runid": "r"}, [1], 7, 0, "D", [])


def test_polychord_adapter_defaults_and_callbacks():
    from evidence_b200 import polychord as pc
    s = pc.default_settings(7)
    assert s["nlive"] == 175 and s["num_repeats"] == 35 and s["do_clustering"] is True
    assert s["precision_criterion"] == 0.001  # reference tests/test_polychord.py:47-50
    for bad in ({"nlive": 10.5}, {"num_repeats": 1.5}, {"do_clustering": 1},
                {"precision_criterion": 1}):
        with pytest.raises(TypeError):
            pc.default_settings(7, bad)
    model = GaussianModel(2)
    prior, loglike = pc.make_callbacks(model, {p: priors.Uniform(-10, 10) for p in model.parnames})
    assert np.allclose(prior(np.array([0.5, 0.75])), [0.0, 5.0])
    assert loglike(np.array([1.0, 1.0])) == (-1.0, [])


def test_polychord_sorted_priors_are_applied_per_group():
    from evidence_b200 import polychord as pc
    model = GaussianModel(3)
    shared = priors.make_prior("SortedUniform", 1.0, 100.0)  # one object for the whole group
    pd_ = {model.parnames[0]: shared, model.parnames[1]: priors.Uniform(0, 1), model.parnames[2]: shared}
    prior, _ = pc.make_callbacks(model, pd_)
    th = prior(np.array([0.3, 0.25, 0.81]))
    assert th[1] == 0.25 and 1.0 <= th[0] <= th[2] <= 100.0  # sorted within the group
    assert th[2] == pytest.approx(1.0 + 99.0 * 0.9) and th[0] == pytest.approx(1.0 + 99.0 * 0.3 * 0.9)


def test_speculative_evaluation_keeps_the_chain():
    """Evaluating the next m stepping-out positions and shrinkage candidates of every walker in one
    call is the sequential slice-sampling rule evaluated ahead of time: same chain bit for bit,
    ~3x fewer (larger) likelihood calls."""
    calls = {}

    def run(m):
        n = [0]

        def f(th):
            n[0] += 1
            return -0.5 * np.sum((th / 0.7) ** 2, axis=1)
        r = nested_sample(f, lambda u: -10 + 20 * u, 5, nlive=200, seed=9, nsteps=10, speculate=m)
        calls[m] = n[0]
        return r
    base = run(1)
    for m in (2, 4, 6):
        r = run(m)
        assert r.logz == base.logz and r.niter == base.niter
        assert np.array_equal(r.samples, base.samples)
        assert r.ncall >= base.ncall          # the discarded look-ahead evaluations are counted
    assert calls[4] < 0.45 * calls[1]


def test_device_resident_sampler_on_cpu_tensors():
    """evidence_b200.sampler_dev is device-agnostic torch code: driven here with CPU tensors and an
    analytic Gaussian likelihood (ln Z known), seeded and deterministic."""
    import math
    import torch
    from evidence_b200.sampler_dev import nested_sample_device
    ndim, sig = 4, 0.6

    def fused(U):
        th = -10 + 20 * U
        return th, -0.5 * ((th / sig) ** 2).sum(1)
    want = ndim * math.log(math.sqrt(2 * math.pi) * sig / 20)
    devs = []
    for seed in (1, 2, 3):
        r = nested_sample_device(fused, ndim, nlive=250, seed=seed, nsteps=10, device="cpu")
        assert abs(r.logz - want) < 4 * r.logzerr + 0.1, (seed, r.logz, want, r.logzerr)
        devs.append((r.logz - want) / r.logzerr)
        assert abs(np.std(r.samples, axis=0) / sig - 1).max() < 0.25
        assert abs(r.weights.sum() - 1) < 1e-12 and r.weighted_samples.shape[0] == len(r.weights)
    assert abs(np.mean(devs)) < 2.0
    a = nested_sample_device(fused, ndim, nlive=120, seed=5, nsteps=6, device="cpu")
    b = nested_sample_device(fused, ndim, nlive=120, seed=5, nsteps=6, device="cpu")
    assert a.logz == b.logz and np.array_equal(a.samples, b.samples)
    # the look-ahead depth does not change the distribution (here: not even the chain)
    c = nested_sample_device(fused, ndim, nlive=120, seed=5, nsteps=6, device="cpu", speculate=2)
    assert abs(c.logz - a.logz) < 3 * (a.logzerr + c.logzerr)


def test_likelihood_plateau_is_retired_as_a_whole():
    """70 % of the prior returns the model's invalid-Keplerian sentinel (-1e30,
    evidence/rvmodel/__init__.py:198-203): the tied points are retired together with X *= (n-g)/n
    (Fowlie et al. 2020) -- charging each its own 1/n biases ln Z by +0.17 here."""
    from evidence_b200.sampler import retire_groups
    a = 10.0 * np.sqrt(0.3)  # valid region: the central square holding 30 % of the prior volume

    def loglike(th):
        out = -0.5 * np.sum(th ** 2, axis=1)
        out[np.any(np.abs(th) > a, axis=1)] = -1e30
        return out
    want = np.log(2 * np.pi / 400.0)  # the Gaussian sits well inside the valid region
    errs, sig = [], []
    for seed in range(6):
        r = nested_sample(loglike, lambda u: -10 + 20 * u, 2, nlive=400, seed=seed, nsteps=8)
        errs.append(r.logz - want)
        sig.append(r.logzerr)
    assert abs(np.mean(errs)) < 0.08, (errs, sig)          # (the per-point 1/n rule gives +0.17)
    assert np.std(errs) < 2.5 * np.mean(sig) + 0.05
    # the bookkeeping itself: a plateau of g out of n points shrinks X by (n - g) / n
    dlogx, logw = retire_groups(np.array([-1e30] * 7 + [-3.0, -2.0]), 10)
    assert np.isclose(dlogx[:7].sum(), np.log(3 / 10)) and np.allclose(logw[:7], np.log(0.7 / 7))
    assert np.isclose(dlogx[7], -1 / 3) and np.isclose(dlogx[8], -1 / 2)
    with pytest.raises(ValueError):
        retire_groups(np.array([-1.0, -1.0]), 2)


@pytest.fixture
def polychord_double(monkeypatch):
    """The test double of the absent third-party package (tests/doubles/pypolychord)."""
    import sys
    monkeypatch.syspath_prepend(os.path.join(os.path.dirname(os.path.abspath(__file__)), "doubles"))
    for name in [m for m in sys.modules if m == "pypolychord" or m.startswith("pypolychord.")]:
        monkeypatch.delitem(sys.modules, name)
    import pypolychord
    assert pypolychord.__version__.endswith("test-double")
    yield pypolychord
    for name in [m for m in sys.modules if m == "pypolychord" or m.startswith("pypolychord.")]:
        sys.modules.pop(name, None)


@pytest.mark.parametrize("ndim,want", [(1, -2.0768), (2, -4.1536)])
def test_polychord_run_through_the_double(tmp_path, polychord_double, ndim, want):
    """The reference's own PolyChord integration tests (tests/test_polychord.py:75-151: unit Gaussian,
    U(-10, 10) priors, |ln Z - analytic| < 0.5, output files exist) on the adapter, with PolyChord's
    scalar calling convention reproduced by the test double."""
    from evidence_b200 import polychord as pc
    model = GaussianModel(ndim)
    calls = []
    orig = model.log_likelihood
    model.log_likelihood = lambda x: (calls.append(np.shape(x)), orig(x))[1]
    priordict = {p: priors.Uniform(-10, 10) for p in model.parnames}
    rundict = {"target": "gauss", "runid": f"poly{ndim}d", "save_dir": str(tmp_path), "nplanets": 0}
    out = pc.run(model, rundict, priordict, {"nlive": 25 * ndim + 50, "num_repeats": 5 * ndim})
    assert abs(out.logZ - want) < 0.5
    assert set(calls) == {(ndim,)} and len(calls) == out.nlike       # one point per call
    for attr in ("runtime", "rundict", "datadict", "fixedpardict", "model_name", "nlive", "nrepeats",
                 "isodate", "ncores", "parnames", "ndim", "sampler"):
        assert hasattr(out, attr), attr                               # :213-226 of the reference
    assert out.sampler == "PolyChord" and out.nrepeats == 5 * ndim
    assert out.file_root.startswith(f"gauss_poly{ndim}d_k0_nlive{25 * ndim + 50}_ncores1_polychord_")
    assert os.path.exists(os.path.join(out.base_dir, out.file_root + ".paramnames"))
    pkl = os.path.join(os.path.dirname(out.base_dir), out.file_root + ".pkl")
    assert pickle.load(open(pkl, "rb")).logZ == out.logZ
    s = pc.set_polysettings({"target": "t", "runid": "r"}, {"nlive": 10}, 3, 0, "D", ["a", "b", "c"])
    assert s.nlive == 10 and s.num_repeats == 15 and s.base_dir.endswith("polychains")

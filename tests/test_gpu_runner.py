"""
GPU tests: end-to-end evidence runs.  The same seeded sampler code is driven once by the device
likelihood + device prior transform and once by the CPU oracle; BASELINE.json's north_star asks
that ln Z agree within the reported uncertainty (they are in fact almost identical, because the two
likelihoods agree to ~1e-11 and the sampler is deterministic for a seed).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _case():
    from evidence_b200 import synth
    return synth.make_case(1, seed=4, n_epochs=96)


def test_lnz_device_vs_cpu_oracle():
    from evidence_b200.rvmodel import RVModel
    from evidence_b200.sampler import nested_sample
    from oracle import rv_oracle
    case = _case()
    model = RVModel(case.fixedpardict, case.datadict(), case.parnames)
    model.set_priors(case.priordict)
    kw = dict(nlive=120, seed=11, nsteps=21)  # 3 ndim slice moves per point: the reference's default
    dev = nested_sample(model.log_likelihood_batch, model.prior_transform_batch, case.ndim, **kw)

    # the fused u -> theta -> lnL call gives the identical run (same theta bits, same lnL bits)
    n0 = model.launch_count()
    fus = nested_sample(model.log_likelihood_batch, model.prior_transform_batch, case.ndim,
                        fused=model.transform_loglike_batch, **kw)
    assert fus.logz == dev.logz and fus.ncall == dev.ncall and np.array_equal(fus.samples, dev.samples)

    t, v, s, ids = case.arrays()
    desc = model.desc_bytes()

    def cpu_loglike(theta):
        return rv_oracle.c_loglike_batch(desc, t, v, s, ids, len(case.insts), theta)[0]
    cpu = nested_sample(cpu_loglike, case.transform, case.ndim, **kw)
    assert abs(dev.logz - cpu.logz) <= max(dev.logzerr, cpu.logzerr)
    assert abs(dev.logz - cpu.logz) < 0.05, (dev.logz, cpu.logz)  # same path, ulp-level lnL noise
    # the run found the injected planet
    per = np.median(dev.samples[:, case.parnames.index("planet1_period")])
    assert abs(per - case.truth["planet1_period"]) / case.truth["planet1_period"] < 0.02
    assert model.counters()["n_points"] == dev.ncall + fus.ncall
    model.close()


def test_runner_on_device_model(tmp_path):
    """run(model, rundict, priordict, settings): the reference's entry point on the device model."""
    from evidence_b200 import ultranest as runner
    from evidence_b200.rvmodel import RVModel
    case = _case()
    model = RVModel(case.fixedpardict, case.datadict(pandas=True), case.parnames)
    rundict = {"target": "synth", "runid": "c1", "save_dir": str(tmp_path), "nplanets": 1}
    out = runner.run(model, rundict, case.priordict,
                     {"nlive": 80, "sampler": "builtin", "seed": 2, "nsteps": 8})
    assert np.isfinite(out.logZ) and out.logZerr > 0 and out.nlike > 1000
    assert out.device_counters["n_points"] == out.nlike
    assert list(out.samples.columns) == model.parnames
    model.close()


def test_device_resident_sampler_matches_the_host_sampler():
    """The whole run on the GPU (sampler_dev): ln Z agrees with the numpy-bookkeeping sampler within
    the reported uncertainties, the injected planet is found, every evaluation went through the
    device (counters), and the runner accepts it as 'builtin_method': 'slice-device'."""
    import time
    from evidence_b200.rvmodel import RVModel
    from evidence_b200.sampler import nested_sample
    from evidence_b200.sampler_dev import nested_sample_device
    case = _case()
    model = RVModel(case.fixedpardict, case.datadict(), case.parnames)
    model.set_priors(case.priordict)
    kw = dict(nlive=200, nsteps=28)  # mixed chains: profiles/r2_lnz_scatter.txt
    t0 = time.perf_counter()
    host = nested_sample(model.log_likelihood_batch, model.prior_transform_batch, case.ndim,
                         fused=model.transform_loglike_batch, seed=3, **kw)
    t_host = time.perf_counter() - t0
    model.reset_counters()
    t0 = time.perf_counter()
    dev = nested_sample_device(lambda U: model.transform_loglike_device(U), case.ndim, seed=3, **kw)
    t_dev = time.perf_counter() - t0
    assert model.counters()["n_points"] == dev.ncall
    # two independent samplers (numpy generator / Philox on the device): ln Z within the REPORTED
    # uncertainties -- with mixed chains (nsteps >= 3 ndim) the run-to-run scatter is 0.2 against a
    # reported 0.35 (profiles/r2_lnz_scatter.txt; the r1 version of this test ran 12 steps and needed 8.0)
    assert abs(dev.logz - host.logz) <= 3.0 * np.hypot(dev.logzerr, host.logzerr), (dev.logz, host.logz)
    assert 0.1 < dev.logzerr_skilling < 1.0 and abs(dev.logzerr_skilling - host.logzerr_skilling) < 0.2
    assert dev.logzerr >= dev.logzerr_skilling and dev.logzerr_bootstrap > 0.05
    assert dev.method == "slice-device-native" and dev.unresolved_moves < 0.02 * dev.accepted_moves
    per = np.median(dev.samples[:, case.parnames.index("planet1_period")])
    assert abs(per - case.truth["planet1_period"]) / case.truth["planet1_period"] < 0.02
    print(f"host-bookkeeping {t_host:.2f} s, device-resident {t_dev:.2f} s")
    model.close()


def test_runner_with_the_device_resident_sampler(tmp_path):
    from evidence_b200 import ultranest as runner
    from evidence_b200.rvmodel import RVModel
    case = _case()
    model = RVModel(case.fixedpardict, case.datadict(pandas=True), case.parnames)
    rundict = {"target": "synth", "runid": "c1dev", "save_dir": str(tmp_path), "nplanets": 1}
    out = runner.run(model, rundict, case.priordict,
                     {"nlive": 60, "sampler": "builtin", "builtin_method": "slice-device",
                      "seed": 2, "nsteps": 6})
    assert np.isfinite(out.logZ) and out.logZerr > 0 and out.nlike > 1000
    assert out.device_counters["n_points"] == out.nlike
    assert list(out.samples.columns) == model.parnames
    model.close()


def test_native_slice_kernels_on_an_analytic_gaussian():
    """csrc/rvslice.cu (rvl_slice_phase) driving a torch likelihood on the GPU: ln Z of a Gaussian
    with known evidence, determinism for a seed, and agreement with the torch formulation."""
    import math
    import torch
    from evidence_b200.sampler_dev import nested_sample_device
    ndim, sig = 4, 0.6

    def fused(U):
        th = -10 + 20 * U
        return th, -0.5 * ((th / sig) ** 2).sum(1)
    want = ndim * math.log(math.sqrt(2 * math.pi) * sig / 20)
    devs = []
    for seed in (1, 2, 3, 4):
        r = nested_sample_device(fused, ndim, nlive=250, seed=seed, nsteps=10, device="cuda")
        assert r.method.endswith("native")
        assert abs(r.logz - want) < 4 * r.logzerr + 0.1, (seed, r.logz, want, r.logzerr)
        devs.append(r.logz - want)
        assert abs(np.std(r.samples, axis=0) / sig - 1).max() < 0.25
        assert r.unresolved_moves < 0.01 * r.accepted_moves
    assert abs(np.mean(devs)) < 0.25
    a = nested_sample_device(fused, ndim, nlive=120, seed=5, nsteps=6, device="cuda")
    b = nested_sample_device(fused, ndim, nlive=120, seed=5, nsteps=6, device="cuda")
    assert a.logz == b.logz and np.array_equal(a.samples, b.samples)
    t = nested_sample_device(fused, ndim, nlive=250, seed=1, nsteps=10, device="cuda", native=False)
    assert abs(t.logz - want) < 4 * t.logzerr + 0.1


def _with_ultranest_double():
    """Import the test double of the absent `ultranest` package (tests/doubles/ultranest)."""
    import os
    import sys
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "doubles")
    for name in [m for m in sys.modules if m == "ultranest" or m.startswith("ultranest.")]:
        del sys.modules[name]
    sys.path.insert(0, path)
    import ultranest
    assert ultranest.__version__.endswith("test-double")
    return path


@pytest.mark.parametrize("stepsampler", ["none", "population-slice"])
def test_ultranest_branch_on_the_device_model(tmp_path, stepsampler):
    """The `which == "ultranest"` branch of the runner (evidence/ultranest/__init__.py:165-185 with
    `vectorized=True`, :171) on the DEVICE model: UltraNest's calling convention (2-D batches of
    ndraw_min .. ndraw_max rows into the batched transform / likelihood) reproduced by the test
    double; the same double driven by the CPU oracle gives ln Z within the reported uncertainty."""
    import sys
    from evidence_b200 import ultranest as runner
    from evidence_b200.rvmodel import RVModel
    from oracle.rv_oracle import OracleRVModel
    path = _with_ultranest_double()
    try:
        case = _case()
        settings = {"nlive": 100, "sampler": "ultranest", "nsteps": 8, "ndraw_min": 256,
                    "ndraw_max": 4096, "stepsampler": stepsampler}
        model = RVModel(case.fixedpardict, case.datadict(pandas=True), case.parnames)
        out = runner.run(model, {"target": "synth", "runid": "un", "save_dir": str(tmp_path), "nplanets": 1},
                         case.priordict, dict(settings))
        assert out.sampler == "UltraNest" and "test-double" in out.sampler_impl
        assert out.device_counters["n_points"] >= out.nlike > 1000  # (+ the constructor's 100 test points)
        assert np.isfinite(out.logZ) and 0 < out.logZerr < 2.0
        per = np.median(out.samples["planet1_period"])
        assert abs(per - case.truth["planet1_period"]) / case.truth["planet1_period"] < 0.02
        if stepsampler == "population-slice":
            # same sampler, CPU reference path (scalar model protocol -> host callbacks); the
            # region-sampling run needs ~7e6 evaluations, minutes on the CPU: device only
            cpu_model = OracleRVModel(case.fixedpardict, case.datadict(), case.parnames)
            cpu_model.datadict, cpu_model.model_path = {}, None
            cpu = runner.run(cpu_model, {"target": "synth", "runid": "cpu", "save_dir": str(tmp_path),
                                         "nplanets": 1}, case.priordict, dict(settings))
            assert abs(out.logZ - cpu.logZ) <= 3.0 * np.hypot(out.logZerr, cpu.logZerr), (out.logZ, cpu.logZ)
        else:
            assert abs(out.logZ + 203.0) < 3.0  # (CPU runs of both step samplers: -203.08, -202.82)
        model.close()
    finally:
        sys.path.remove(path)
        for name in [m for m in sys.modules if m == "ultranest" or m.startswith("ultranest.")]:
            del sys.modules[name]


def test_polychord_adapter_on_the_device_model(tmp_path):
    """evidence.polychord.run's boundary on the DEVICE model: PolyChord's scalar callbacks (one point
    per call: a batch of 1 through the same C-ABI entry point), reproduced by the test double of the
    absent package; ln Z within the reported uncertainty of the vectorised sampler's."""
    import os
    import sys
    from evidence_b200 import polychord as pc
    from evidence_b200.rvmodel import RVModel
    from evidence_b200.sampler import nested_sample
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "doubles")
    sys.path.insert(0, path)
    try:
        case = _case()
        model = RVModel(case.fixedpardict, case.datadict(pandas=True), case.parnames)
        out = pc.run(model, {"target": "synth", "runid": "poly", "save_dir": str(tmp_path), "nplanets": 1},
                     case.priordict, {"nlive": 40, "num_repeats": 7})  # ~1e5 scalar device calls
        assert out.sampler == "PolyChord" and out.device_counters["n_points"] == out.nlike > 5000
        model.set_priors(case.priordict)
        ref = nested_sample(model.log_likelihood_batch, model.prior_transform_batch, case.ndim, nlive=200,
                            fused=model.transform_loglike_batch, seed=3, nsteps=28)
        assert abs(out.logZ - ref.logz) <= 3.0 * np.hypot(out.logZerr, ref.logzerr) + 1.0, (out.logZ, ref.logz)
        model.close()
    finally:
        sys.path.remove(path)
        for name in [m for m in sys.modules if m == "pypolychord" or m.startswith("pypolychord.")]:
            del sys.modules[name]

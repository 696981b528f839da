"""CPU tests: host-side priors against the reference's ppf values (tests/golden/priors.npz)."""
import numpy as np
import pytest

from _util import load_golden

from evidence_b200 import priors


def test_reference_pins():
    # tests/test_priors.py:11-15, 28 of the reference
    u = priors.Uniform(4, 6)
    assert u.pdf(5) == 0.5 and u.pdf(3) == 0.0
    assert u.ppf(0.5) == 5 and u.ppf(0.0) == 4 and u.ppf(1.0) == 6
    assert priors.Jeffreys(10, 100).pdf(10) == pytest.approx(0.043429448190325175, abs=1e-16)


def test_every_distribution_vs_reference():
    meta, z = load_golden("priors")
    q = z["q"]
    for spec, want in zip(meta["specs"], z["ppf"]):
        pr = priors.make_prior(spec["name"], *spec["pars"])
        got = np.asarray(pr.ppf(q), dtype=np.float64)
        ok = np.isfinite(want)
        assert ok.sum() >= len(q) - 6, spec
        assert np.allclose(got[ok], want[ok], rtol=1e-12, atol=1e-12), spec


def test_prior_constructor_walk():
    input_dict = {"planet1": {"k1": [0.0, 1, ["Jeffreys", 0.1, 100.]],
                              "period": [0.0, 1, ["UniformFrequency", 1, 100]],
                              "ecc": [0.1, 1, ["Beta", 0.867, 3.03]],
                              "epoch": [51050, 0]},
                  "hamilton": {"offset": [0., 1, ["Uniform", -10, 10]]}}
    pd_ = priors.prior_constructor(input_dict)
    assert sorted(pd_) == ["hamilton_offset", "planet1_ecc", "planet1_k1", "planet1_period"]
    with pytest.raises(priors.PriorError):
        priors.prior_constructor({"a": {"b": [0, 1, ["Nope", 1]]}})


def test_device_descriptors_pack_tables():
    prs = [priors.Uniform(0, 1), priors.Sine(0, 180), priors.Beta(0.867, 3.03)]
    descs, tables = priors.device_descriptors(prs)
    assert descs[0].kind == 0 and descs[1].kind == descs[2].kind == 7
    assert descs[1].table_offset == 0 and descs[2].table_offset == 2 * descs[1].table_len
    assert tables.size == 2 * descs[1].table_len + 3 * descs[2].table_len  # Beta ships slopes too
    assert descs[2].p[1] == 1.0 and descs[1].p[1] == 0.0
    cdf = tables[descs[2].table_offset: descs[2].table_offset + descs[2].table_len]
    assert np.all(np.diff(cdf) > 0)


def test_prior_constructor_reports_the_real_cause():
    # an unknown name keeps the reference's message (evidence/priors.py:500-503) ...
    with pytest.raises(priors.PriorError, match="Unknown type of prior"):
        priors.prior_constructor({"a": {"b": [0, 1, ["Nope", 1]]}})
    # ... a known prior with bad shape parameters says what is wrong with them
    with pytest.raises(priors.PriorError, match="a_b .Uniform.: Uniform needs xmin < xmax"):
        priors.prior_constructor({"a": {"b": [0, 1, ["Uniform", 3, 1]]}})


def test_polychord_sorted_priors_form_one_group():
    """evidence/polychord/__init__.py:137-160: all parameters with a SortedUniform prior are
    transformed together (prior_constructor builds one prior object PER parameter), so the periods
    come out non-decreasing whatever the cube says."""
    from evidence_b200 import polychord

    class FakeModel:
        parnames = sorted(["planet1_period", "planet2_period", "planet3_period", "planet1_k1",
                           "planet2_logk", "planet3_logk"])

        def log_likelihood(self, x):
            return -1.0

    input_dict = {f"planet{k}": {"period": [0.0, 1, ["SortedUniform", 1.0, 100.0]]} for k in (1, 2, 3)}
    input_dict["planet1"]["k1"] = [0.0, 1, ["Uniform", 0.0, 10.0]]
    input_dict["planet2"]["logk"] = [0.0, 1, ["SortedLogUniform", 0.1, 10.0]]
    input_dict["planet3"]["logk"] = [0.0, 1, ["SortedLogUniform", 0.1, 10.0]]
    priordict = priors.prior_constructor(input_dict)
    assert len({id(priordict[f"planet{k}_period"]) for k in (1, 2, 3)}) == 3  # separate objects
    prior, loglike = polychord.make_callbacks(FakeModel(), priordict)
    names = FakeModel.parnames
    rng = np.random.default_rng(0)
    for cube in [np.array([0.5, 0.9, 0.3, 0.1, 0.7, 0.2])] + list(rng.random((20, 6))):
        theta = prior(cube)
        per = [theta[names.index(f"planet{k}_period")] for k in (1, 2, 3)]
        logk = [theta[names.index(f"planet{k}_logk")] for k in (2, 3)]
        assert per[0] <= per[1] <= per[2] and 1.0 <= per[0] and per[2] <= 100.0
        assert logk[0] <= logk[1] and 0.1 <= logk[0] and logk[1] <= 10.0
        assert theta[names.index("planet1_k1")] == pytest.approx(10.0 * cube[names.index("planet1_k1")])
    # the reference applies forced identifiability to the group in parnames order
    cube = np.array([0.5, 0.9, 0.3, 0.1, 0.7, 0.2])
    idx = [names.index(f"planet{k}_period") for k in (1, 2, 3)]
    want = 1.0 + 99.0 * priors.forced_identifiability_transform(cube[idx])
    assert np.allclose(prior(cube)[idx], want)
    assert loglike(prior(cube)) == (-1.0, [])

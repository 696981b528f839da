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

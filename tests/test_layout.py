"""CPU tests: the model compiler reproduces the reference's name-driven structure rules."""
import pytest

from evidence_b200 import _abi
from evidence_b200.layout import ModelLayoutError, compile_model, structure_flags

A = ["hamilton_jitter", "hamilton_offset", "planet1_ecc", "planet1_k1", "planet1_ma0",
     "planet1_omega", "planet1_period"]


def test_sorted_slots_and_flags():
    d, names = compile_model(list(reversed(A)), {"planet1_epoch": 51050}, ["hamilton"], 50002.6)
    assert names == sorted(A)  # evidence/rvmodel/__init__.py:43
    assert (d.ndim, d.n_planets, d.n_inst) == (7, 1, 1)
    assert d.jitter_in_model == 1 and d.drift_in_model == 0 and d.n_linpar == 0
    pl = d.planet[0]
    assert pl.amp.slot == names.index("planet1_k1") and pl.amp_is_log == 0
    assert pl.e1.slot == names.index("planet1_ecc") and pl.ecc_mode == _abi.RVL_ECC_DIRECT
    assert pl.phase_mode == _abi.RVL_PHASE_MA0
    assert pl.epoch.slot == -1 and pl.epoch.value == 51050.0
    assert d.offset[0].slot == names.index("hamilton_offset")
    assert d.tol == 1e-4 and d.itmax == 10000


def test_structure_comes_from_free_names_only():
    # a jitter that is only FIXED is ignored; a planet whose k1 is fixed is not counted (:118-139)
    free = ["planet1_ecc", "planet1_k1", "planet1_ma0", "planet1_omega", "planet1_period"]
    fixed = {"planet1_epoch": 51050.0, "hamilton_offset": -2.0, "hamilton_jitter": 5.0}
    d, _ = compile_model(free, fixed, ["hamilton"], 0.0)
    assert d.jitter_in_model == 0
    assert structure_flags(["planet1_period", "hamilton_offset"])[0] == 0
    assert structure_flags(["planet1_logk1"])[0] == 1  # 'k1' in 'logk1'


def test_fixed_wins_on_clash():
    d, names = compile_model(A, {"planet1_epoch": 1.0, "planet1_k1": 30.0}, ["hamilton"], 0.0)
    assert d.planet[0].amp.slot == -1 and d.planet[0].amp.value == 30.0  # :178


def test_parametrisation_branches():
    free = ["hamilton_jitter", "planet1_logk1", "planet1_logperiod", "planet1_ml0",
            "planet1_secos", "planet1_sesin"]
    d, names = compile_model(free, {"planet1_epoch": 0.0, "hamilton_offset": 0.0}, ["hamilton"], 0.0)
    pl = d.planet[0]
    assert pl.amp_is_log == 1 and pl.period_is_log == 1
    assert pl.ecc_mode == _abi.RVL_ECC_SECOS_SESIN and pl.phase_mode == _abi.RVL_PHASE_ML0
    free = ["planet1_ecos", "planet1_esin", "planet1_k1", "planet1_ma0", "planet1_period"]
    d, _ = compile_model(free, {"planet1_epoch": 0.0, "hamilton_offset": 0.0}, ["hamilton"], 0.0)
    assert d.planet[0].ecc_mode == _abi.RVL_ECC_ECOS_ESIN


def test_drift_reference_time():
    free = ["drift_lin", "drift_quad", "hamilton_offset"]
    d, _ = compile_model(free, {}, ["hamilton"], 50002.5)
    assert d.drift_in_model == 1 and d.tref == 50002.5  # default time[0] (:259-260)
    assert d.drift[0].slot >= 0 and d.drift[1].slot >= 0 and d.drift[2].slot == -1
    d, _ = compile_model(free, {"drift_tref": 51050}, ["hamilton"], 50002.5)
    assert d.tref == 51050.0
    with pytest.raises(NotImplementedError):
        compile_model(free + ["drift_tref"], {}, ["hamilton"], 0.0)


def test_missing_parameters_raise_like_the_reference():
    with pytest.raises(KeyError):  # no offset for the instrument (:187)
        compile_model(["planet1_k1"], {}, ["hamilton"], 0.0)
    with pytest.raises(ModelLayoutError):  # eccentricity parametrisation (:445-447)
        compile_model(["planet1_k1", "planet1_period", "planet1_ma0", "hamilton_offset"],
                      {"planet1_epoch": 0.0}, ["hamilton"], 0.0)


def test_linear_parameter_smoother_vs_reference():
    """RVModel.linear_parameter (host helper) against the live reference's outputs
    (tests/golden/linpar_smoother.npz; the reference's 'box' kernel raises under numpy >= 2)."""
    import os
    import numpy as np
    from evidence_b200.rvmodel import RVModel
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "linpar_smoother.npz"))
    for key in z.files:
        if key in ("t", "ind"):
            continue
        kernel, ft = key.rsplit("_", 1)
        got = RVModel.linear_parameter(None, z["t"], z["ind"], kernel=None if kernel == "None" else kernel,
                                       timescale=0.3, filter_type=ft)
        assert np.max(np.abs(got - z[key])) < 1e-13, key
    box = RVModel.linear_parameter(None, z["t"], z["ind"], kernel="box", timescale=0.3)
    assert box.min() == -1.0 and box.max() == 1.0
    with pytest.raises(ValueError):
        RVModel.linear_parameter(None, z["t"], z["ind"], kernel="nope")

"""CPU tests: the configuration reader, on the checks of the reference's tests/test_config.py."""
import os
import textwrap

import numpy as np
import pytest

from evidence_b200 import config


@pytest.fixture()
def cfg(tmp_path):
    rng = np.random.default_rng(0)
    data = tmp_path / "star.rv"
    rows = ["rjd\tvrad\tsvrad", "---\t----\t-----"]
    rows += [f"{50000 + 3.1 * i:.6f}\t{rng.normal(0, 5):.2f}\t{rng.uniform(1, 2):.2f}" for i in range(40)]
    data.write_text("\n".join(rows) + "\n")
    mod = tmp_path / "config_star.py"
    mod.write_text(textwrap.dedent(f"""
        import numpy as np
        rundict = {{'target': 'star', 'runid': 'example', 'star_params': {{'star_mass': (1.11, 0.02)}},
                   'save_dir': r'{tmp_path}'}}
        datadict = {{'hamilton': {{'datafile': r'{data}', 'instrument': 'hamilton',
                                  'kwargs': {{'sep': '\\t', 'skiprows': (1,)}}}}}}
        planetdict1 = {{'k1': [0.0, 1, ['Jeffreys', 0.1, 100.]],
                       'period': [0.0, 1, ['UniformFrequency', 1, 100]],
                       'ecc': [0.1, 1, ['Beta', 0.867, 3.03]],
                       'omega': [0.1, 1, ['Uniform', 0., 2*np.pi]],
                       'ma0': [0.1, 1, ['Uniform', 0., 2*np.pi]],
                       'epoch': [51050, 0]}}
        hamiltondict = {{'offset': [0., 1, ['Uniform', -10, 10]], 'jitter': [0.75, 1, ['Uniform', 0., 50.]]}}
        input_dict = {{'planet1': planetdict1, 'hamilton': hamiltondict}}
        configdicts = [rundict, input_dict, datadict]
    """))
    return str(mod)


def test_nplanets_validation(cfg):
    with pytest.raises(TypeError):
        config.read_config(cfg, 0.5)
    with pytest.raises(ValueError):
        config.read_config(cfg, -6)


def test_single_planet_config(cfg):
    rundict, datadict, priordict, fixed = config.read_config(cfg)
    assert len(priordict) == 7 and len(fixed) == 1 and fixed["planet1_epoch"] == 51050
    assert len(datadict["hamilton"]["data"]) == 40
    assert rundict["star_params"]["star_mass"] == (1.11, 0.02) and "nplanets" not in rundict
    assert rundict["prior_names"]["planet1_k1"] == "Jeffreys: [0.1, 100.0]"


@pytest.mark.parametrize("n", [0, 1, 2, 3])
def test_planet_dicts_are_cloned_or_trimmed(cfg, n):
    rundict, _, priordict, fixed = config.read_config(cfg, n)
    assert sum("k1" in p for p in priordict) == n and rundict["nplanets"] == n
    assert len(fixed) == n


def test_config_feeds_the_model_compiler(cfg):
    from evidence_b200.layout import compile_model
    _, datadict, priordict, fixed = config.read_config(cfg, 2)
    d, names = compile_model(list(priordict), fixed, list(datadict), 50000.0)
    assert d.n_planets == 2 and d.ndim == 12 and names == sorted(priordict)

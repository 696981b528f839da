"""
Configuration reader: same entry point and return values as ``evidence.config.read_config``
(evidence/config.py:8-72), so existing config modules (``configdicts = [rundict, input_dict,
datadict]``, parameter entries ``[init, free_flag, [PriorName, *shape]]``) work unchanged.  The
priors it builds are ``evidence_b200.priors`` objects, which the device model can stage.
"""
import importlib.util
import os

from .priors import prior_constructor


def _load_module(configfile):
    name = os.path.splitext(os.path.basename(configfile))[0]
    spec = importlib.util.spec_from_file_location(name, configfile)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def read_config(configfile, nplanets=None):
    """
    Returns ``(rundict, datadict, priordict, fixedpardict)`` for the configuration module at
    ``configfile``; ``nplanets`` clones or trims the ``planetN`` dictionaries exactly like the
    reference (evidence/config.py:26-58).
    """
    c = _load_module(configfile)
    rundict, inputdict, datadict = map(dict.copy, c.configdicts)

    if nplanets is not None:
        if type(nplanets) is not int:
            raise TypeError("nplanets has to be an integer.")
        if nplanets < 0:
            raise ValueError("nplanets has to be positive.")
        n_dicts = sum("planet" in key for key in inputdict)
        if n_dicts > 1 and nplanets > n_dicts:
            raise ValueError("Not enough planet dictionaries for the requested number of planets.")
        if n_dicts == 1:
            # a single planet dictionary is the template for all of them; it is removed and
            # re-appended, so planets move to the end of the insertion order (:46-51)
            template = dict(inputdict["planet1"])
            del inputdict["planet1"]
            for k in range(1, nplanets + 1):
                inputdict[f"planet{k}"] = dict(template)
        elif n_dicts > 1:
            for k in range(nplanets + 1, n_dicts + 1):
                del inputdict[f"planet{k}"]
        rundict["nplanets"] = nplanets

    priordict = prior_constructor(inputdict)
    read_priors(inputdict, rundict)
    read_data(datadict)
    fixedpardict = get_fixedparvalues(inputdict)
    return rundict, datadict, priordict, fixedpardict


def get_parnames(inputdict):
    """(free, fixed) parameter names, evidence/config.py:75-87."""
    free, fixed = [], []
    for obj in inputdict:
        for par in inputdict[obj]:
            flag = inputdict[obj][par][1]
            if flag > 0:
                free.append(obj + "_" + par)
            elif flag == 0:
                fixed.append(obj + "_" + par)
    return free, fixed


def get_fixedparvalues(inputdict):
    """{name: value} of the parameters whose flag is 0, evidence/config.py:90-99."""
    return {obj + "_" + par: inputdict[obj][par][0]
            for obj in inputdict for par in inputdict[obj] if inputdict[obj][par][1] == 0}


def read_data(datadict):
    """Load every instrument's file with ``pandas.read_csv(**kwargs)``, evidence/config.py:102-114."""
    import pandas as pd
    for inst in datadict:
        entry = dict(datadict[inst])
        entry["data"] = pd.read_csv(entry["datafile"], **entry.get("kwargs", {}))
        datadict[inst] = entry


def read_priors(inputdict, rundict):
    """``rundict['prior_names'][name] = 'PriorName: [shape]'``, evidence/config.py:117-148."""
    names = {}
    for obj in inputdict:
        for par in inputdict[obj]:
            entry = inputdict[obj][par]
            if not isinstance(entry, list) or entry[1] == 0:
                continue
            names[obj + "_" + par] = f"{entry[2][0]}: {entry[2][1:]}"
    rundict["prior_names"] = names

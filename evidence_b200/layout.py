"""
Model compiler: turn the reference's name-driven model description into the flat slot table
(``rvl_model_desc``) the CUDA kernel consumes.

The reference decides the model structure from parameter NAMES on every likelihood call
(substring tests and dict lookups).  Here that happens once, on the host, with exactly the same
rules (citations relative to the reference checkout):

  * ``parnames`` are sorted (evidence/rvmodel/__init__.py:43) and theta columns follow that order;
  * structure flags come from the FREE names only (:118-139): ``nplanets`` = number of names
    containing ``'k1'``; drift / jitter / linpar present iff some free name contains
    ``'drift'`` / ``'jitter'`` / ``'linpar'`` -- a jitter that is only a fixed parameter is ignored;
  * values are looked up in ``pardict = free U fixed`` with fixed winning (:173-178);
  * planet parametrisation by presence, in modelk's branch order (:411-456).
"""
from . import _abi


class ModelLayoutError(KeyError):
    """A parameter the reference would look up (and raise KeyError on) is missing."""


def _lookup(name, slots, fixed):
    """(slot, value) of a pardict entry; fixed parameters win on a clash (:178)."""
    if name in fixed:
        return -1, float(fixed[name])
    if name in slots:
        return slots[name], 0.0
    raise ModelLayoutError(name)


def _has(name, slots, fixed):
    return name in fixed or name in slots


def _set(param, sv):
    param.slot, param.value = int(sv[0]), float(sv[1])


def structure_flags(parnames):
    """nplanets and the drift/linpar/jitter flags from free names (:118-139)."""
    nplanets = sum("k1" in p for p in parnames)
    return (nplanets, any("drift" in p for p in parnames),
            any("linpar" in p for p in parnames), any("jitter" in p for p in parnames))


def compile_model(parnames, fixedpardict, insts, time0, linpar_names=(), tol=1.0e-4,
                  itmax=10000):
    """
    Build the ``rvl_model_desc`` for a model with free parameters ``parnames`` (any order; they
    are sorted here like BaseModel does), fixed parameters ``fixedpardict``, instruments
    ``insts`` (datadict key order) and first time stamp ``time0`` (default drift reference,
    evidence/rvmodel/__init__.py:259-260).
    """
    names = sorted(parnames)
    if len(names) > _abi.RVL_MAX_DIM:
        raise ValueError(f"more than {_abi.RVL_MAX_DIM} free parameters")
    if len(set(names)) != len(names):
        raise ValueError("duplicate parameter names")
    slots = {n: i for i, n in enumerate(names)}
    fixed = dict(fixedpardict)
    nplanets, drift, linpar, jitter = structure_flags(names)
    if nplanets > _abi.RVL_MAX_PLANETS:
        raise ValueError(f"more than {_abi.RVL_MAX_PLANETS} planets")
    if not 1 <= len(insts) <= _abi.RVL_MAX_INST:
        raise ValueError(f"need 1..{_abi.RVL_MAX_INST} instruments")

    d = _abi.rvl_model_desc()
    d.abi_version = _abi.RVL_ABI_VERSION
    d.ndim = len(names)
    d.n_planets = nplanets
    d.n_inst = len(insts)
    d.jitter_in_model = int(jitter)
    d.drift_in_model = int(drift)
    d.itmax = int(itmax)
    d.tol = float(tol)

    for k in range(1, nplanets + 1):  # planets are numbered 1..nplanets (:371)
        pre = f"planet{k}_"
        pl = d.planet[k - 1]
        if _has(pre + "k1", slots, fixed):  # :412-415
            _set(pl.amp, _lookup(pre + "k1", slots, fixed))
            pl.amp_is_log = 0
        else:
            _set(pl.amp, _lookup(pre + "logk1", slots, fixed))
            pl.amp_is_log = 1
        if _has(pre + "period", slots, fixed):  # :417-420
            _set(pl.period, _lookup(pre + "period", slots, fixed))
            pl.period_is_log = 0
        else:
            _set(pl.period, _lookup(pre + "logperiod", slots, fixed))
            pl.period_is_log = 1
        if _has(pre + "secos", slots, fixed):  # :425-431
            pl.ecc_mode = _abi.RVL_ECC_SECOS_SESIN
            _set(pl.e1, _lookup(pre + "secos", slots, fixed))
            _set(pl.e2, _lookup(pre + "sesin", slots, fixed))
        elif _has(pre + "ecos", slots, fixed):  # :433-439
            pl.ecc_mode = _abi.RVL_ECC_ECOS_ESIN
            _set(pl.e1, _lookup(pre + "ecos", slots, fixed))
            _set(pl.e2, _lookup(pre + "esin", slots, fixed))
        else:  # :441-447
            pl.ecc_mode = _abi.RVL_ECC_DIRECT
            try:
                _set(pl.e1, _lookup(pre + "ecc", slots, fixed))
                _set(pl.e2, _lookup(pre + "omega", slots, fixed))
            except ModelLayoutError:
                raise ModelLayoutError("Something is wrong with the eccentricity parametrisation")
        if _has(pre + "ml0", slots, fixed):  # :449-454
            pl.phase_mode = _abi.RVL_PHASE_ML0
            _set(pl.phase, _lookup(pre + "ml0", slots, fixed))
        else:
            pl.phase_mode = _abi.RVL_PHASE_MA0
            _set(pl.phase, _lookup(pre + "ma0", slots, fixed))
        _set(pl.epoch, _lookup(pre + "epoch", slots, fixed))  # :456

    for i, inst in enumerate(insts):
        _set(d.offset[i], _lookup(f"{inst}_offset", slots, fixed))  # :187
        if jitter:
            _set(d.jitter[i], _lookup(f"{inst}_jitter", slots, fixed))  # :190
        else:
            _set(d.jitter[i], (-1, 0.0))

    d.tref = float(time0)
    if drift:  # :242-271
        for j, nm in enumerate(("lin", "quad", "cub", "quar")):
            if _has("drift_" + nm, slots, fixed):
                _set(d.drift[j], _lookup("drift_" + nm, slots, fixed))
            else:
                _set(d.drift[j], (-1, 0.0))
        if _has("drift_tref", slots, fixed):
            slot, val = _lookup("drift_tref", slots, fixed)
            if slot >= 0:
                raise NotImplementedError(
                    "a FREE drift_tref is not supported on the device path "
                    "(the drift time axis is staged once)")
            d.tref = val
    else:
        for j in range(4):
            _set(d.drift[j], (-1, 0.0))

    lin = list(linpar_names) if linpar else []
    if len(lin) > _abi.RVL_MAX_LINPAR:
        raise ValueError(f"more than {_abi.RVL_MAX_LINPAR} linear parameters")
    d.n_linpar = len(lin)
    for j, nm in enumerate(lin):  # :210-212
        _set(d.linpar[j], _lookup(f"linpar_{nm}", slots, fixed))
    return d, names

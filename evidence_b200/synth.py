"""
Synthetic RV data sets and parameter batches of the shapes BASELINE.json names
(SURVEY.md 8(d)).  Used by the tests, bench.py and the golden-vector generator; deterministic
for a given seed.

    case = make_case(2)                   # config 2: N=1000, K=2, 2 instruments, linear drift
    model = RVModel(case.fixedpardict, case.datadict(), case.parnames)
    theta = case.draw_theta(4096, seed=7) # prior draws, columns in sorted-parnames order
"""
import numpy as np

from . import priors as _priors

TWO_PI = 2 * np.pi
EPOCH = 52500.0

#            N      K  n_inst drift  ecc prior
CONFIGS = {
    1: dict(n_epochs=200, n_planets=1, n_inst=1, drift=0, ecc=("Beta", 0.867, 3.03)),
    2: dict(n_epochs=1000, n_planets=2, n_inst=2, drift=1, ecc=("Uniform", 0.0, 0.95)),
    3: dict(n_epochs=5000, n_planets=4, n_inst=3, drift=0, ecc=("Uniform", 0.0, 0.95)),
    5: dict(n_epochs=10000, n_planets=3, n_inst=3, drift=0, ecc=("Uniform", 0.0, 0.95)),
}


def _kepler_rv(t, K, P, e, w, M0, epoch):
    """Converged Keplerian signal for data generation (not the likelihood path)."""
    M = TWO_PI / P * (t - epoch) + M0
    E = M + e * np.sin(M)
    for _ in range(60):
        E = E - (E - e * np.sin(E) - M) / (1 - e * np.cos(E))
    nu = 2 * np.arctan2(np.sqrt(1 + e) * np.sin(E / 2), np.sqrt(1 - e) * np.cos(E / 2))
    return K * (np.cos(nu + w) + e * np.cos(w))


class SynthCase:
    def __init__(self, n_epochs, n_planets, n_inst, drift, ecc, seed, pure_noise=False):
        self.n_epochs, self.n_planets, self.n_inst, self.drift = n_epochs, n_planets, n_inst, drift
        rng = np.random.default_rng(seed)
        t = np.sort(rng.uniform(50000.0, 55000.0, n_epochs))
        inst = rng.integers(0, n_inst, n_epochs)
        err = rng.uniform(0.5, 2.0, n_epochs)
        self.insts = [f"inst{i}" for i in range(n_inst)]

        spec = {}
        for k in range(1, n_planets + 1):
            spec[f"planet{k}_k1"] = ("Uniform", 0.0, 20.0)
            spec[f"planet{k}_period"] = ("Jeffreys", 1.0, 1000.0)
            spec[f"planet{k}_ecc"] = tuple(ecc)
            spec[f"planet{k}_omega"] = ("Uniform", 0.0, TWO_PI)
            spec[f"planet{k}_ma0"] = ("Uniform", 0.0, TWO_PI)
        for name in self.insts:
            spec[f"{name}_offset"] = ("Uniform", -10.0, 10.0)
            spec[f"{name}_jitter"] = ("Uniform", 0.0, 10.0)
        for nm in ("lin", "quad", "cub", "quar")[:drift]:
            spec[f"drift_{nm}"] = ("Uniform", -1.0, 1.0)
        self.prior_spec = spec
        self.parnames = sorted(spec)
        self.priordict = {k: _priors.make_prior(*v) for k, v in spec.items()}
        self.fixedpardict = {f"planet{k}_epoch": EPOCH for k in range(1, n_planets + 1)}
        if drift:
            self.fixedpardict["drift_tref"] = EPOCH

        # hidden truth + noise
        truth = {p: float(self.priordict[p].ppf(rng.random())) for p in self.parnames}
        signal = np.zeros(n_epochs)
        if not pure_noise:
            for k in range(1, n_planets + 1):
                signal += _kepler_rv(t, truth[f"planet{k}_k1"], truth[f"planet{k}_period"],
                                     truth[f"planet{k}_ecc"], truth[f"planet{k}_omega"],
                                     truth[f"planet{k}_ma0"], EPOCH)
            for i, name in enumerate(self.insts):
                signal[inst == i] += truth[f"{name}_offset"]
            tt = (t - EPOCH) / 365.25
            for j, nm in enumerate(("lin", "quad", "cub", "quar")[:drift]):
                signal += truth[f"drift_{nm}"] * tt ** (j + 1)
            vrad = signal + rng.normal(0.0, np.sqrt(err ** 2 + 1.0))
        else:
            vrad = rng.normal(0.0, 5.0, n_epochs)
        self.truth = truth
        # instrument-major rows, like BaseModel's concatenation (evidence/rvmodel/__init__.py:50-55)
        self._tables = {}
        for i, name in enumerate(self.insts):
            m = inst == i
            self._tables[name] = {"rjd": t[m].copy(), "vrad": vrad[m].copy(),
                                  "svrad": err[m].copy()}

    @property
    def ndim(self):
        return len(self.parnames)

    def datadict(self, pandas=False):
        """A FRESH datadict (the model constructors add an inst_id column to what they get)."""
        out = {}
        for name, tab in self._tables.items():
            data = {k: v.copy() for k, v in tab.items()}
            if pandas:
                import pandas as pd
                data = pd.DataFrame(data)
            out[name] = {"data": data}
        return out

    def arrays(self):
        """(t, vrad, svrad, inst_id) concatenated instrument-major."""
        t = np.concatenate([self._tables[n]["rjd"] for n in self.insts])
        v = np.concatenate([self._tables[n]["vrad"] for n in self.insts])
        s = np.concatenate([self._tables[n]["svrad"] for n in self.insts])
        ids = np.concatenate([np.full(len(self._tables[n]["rjd"]), i, dtype=np.int32)
                              for i, n in enumerate(self.insts)])
        return t, v, s, ids

    def draw_unit(self, B, seed=0):
        return np.random.default_rng(seed).random((B, self.ndim))

    def transform(self, U):
        """Host prior transform (sorted-parnames columns)."""
        U = np.asarray(U, dtype=np.float64)
        out = np.empty_like(U)
        for i, p in enumerate(self.parnames):
            out[:, i] = self.priordict[p].ppf(U[:, i])
        return out

    def draw_theta(self, B, seed=0):
        return self.transform(self.draw_unit(B, seed))


def make_case(config, seed=None, pure_noise=False, **overrides):
    """One of BASELINE.json's shapes (1, 2, 3, 5); keyword overrides e.g. n_epochs=..."""
    kw = dict(CONFIGS[config])
    kw.update(overrides)
    return SynthCase(seed=config if seed is None else seed, pure_noise=pure_noise, **kw)

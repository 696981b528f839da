"""Test double of ``pypolychord.priors`` (published definitions; see the package docstring)."""
import numpy as np


def forced_indentifiability_transform(x):
    n = len(x)
    t = np.zeros(n)
    t[n - 1] = x[n - 1] ** (1.0 / n)
    for i in range(n - 2, -1, -1):
        t[i] = x[i] ** (1.0 / (i + 1)) * t[i + 1]
    return t


class UniformPrior:
    def __init__(self, a, b):
        self.a, self.b = a, b

    def __call__(self, x):
        return self.a + (self.b - self.a) * x


class SortedUniformPrior(UniformPrior):
    def __call__(self, x):
        return super().__call__(forced_indentifiability_transform(x))


class LogUniformPrior(UniformPrior):
    def __call__(self, x):
        return self.a * (self.b / self.a) ** x


class LogSortedUniformPrior(LogUniformPrior):
    def __call__(self, x):
        return super().__call__(forced_indentifiability_transform(x))

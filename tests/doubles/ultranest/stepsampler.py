"""Test double of ``ultranest.stepsampler`` (see the package docstring)."""


class RegionSliceSampler:
    """The reference's choice (evidence/ultranest/__init__.py:175): one walker, so the likelihood is
    called with ONE point at a time even in vectorised mode."""
    popsize = 1

    def __init__(self, nsteps, adaptive_nsteps=False, max_nsteps=1000, region_filter=False, log=False,
                 scale=1.0):
        self.nsteps = int(nsteps)
        self.adaptive_nsteps = adaptive_nsteps

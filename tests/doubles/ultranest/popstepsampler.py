"""Test double of ``ultranest.popstepsampler`` (see the package docstring)."""


def generate_region_oriented_direction(ui, region, scale=1):
    raise NotImplementedError("the double whitens with the live-point covariance itself")


def generate_unit_directions(ui, region, scale=1):
    raise NotImplementedError


class PopulationSliceSampler:
    """`popsize` walkers advanced together: `popsize` points per likelihood call."""

    def __init__(self, popsize, nsteps, generate_direction, scale=1.0, scale_adapt_factor=0.9, log=False,
                 logfile=None):
        self.popsize = int(popsize)
        self.nsteps = int(nsteps)
        self.generate_direction = generate_direction

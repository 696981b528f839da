"""
The reference's run script (evidence/examples/51Peg/run.py) with the two-import switch to the
B200 path: config -> model -> run().  Needs a GPU.

    python examples/synthetic/make_data.py
    python examples/synthetic/run.py [nplanets]
"""
import os
import sys
from pathlib import Path

here = Path(__file__).parent.absolute()
sys.path.insert(0, str(here.parent.parent))

from evidence_b200 import config, ultranest  # was: from evidence import config, ultranest  # noqa: E402
from evidence_b200.rvmodel import RVModel    # was: from evidence.rvmodel import RVModel    # noqa: E402

nplanets = int(sys.argv[1]) if len(sys.argv) > 1 else 1
rundict, datadict, priordict, fixedpardict = config.read_config(os.path.join(here, "config_synth.py"),
                                                                nplanets)
model = RVModel(fixedpardict, datadict, list(priordict.keys()))
output = ultranest.run(model, rundict, priordict, {"nlive": 200, "seed": 1, "ndraw_min": 4096})
print(f"k={nplanets}: ln Z = {output.logZ:.2f} +- {output.logZerr:.2f}  "
      f"({output.nlike} likelihood calls, {output.evals_per_second:.3g} lnL/s end to end, sampler {output.sampler})")

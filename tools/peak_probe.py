import sys; sys.path.insert(0,'.')
from evidence_b200 import synth
from evidence_b200.rvmodel import RVModel
case = synth.make_case(1)
m = RVModel(case.fixedpardict, case.datadict(), case.parnames)
print([m.fp64_peak_tflops() for _ in range(3)])

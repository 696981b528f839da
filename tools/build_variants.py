"""Build the kernel A/B variants of a round beside the product library (evidence_b200/variants/,
git-ignored like every .so; they travel to the GPU box).  usage: python tools/build_variants.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evidence_b200 import build

VARIANTS = {
    # round-1 arithmetic: FP64 |d|>tol compare, two-rounding E - e sin E, kepler_rv, no short last pass
    "r1": ["RVL_INT_TOL=0", "RVL_KRV2=0", "RVL_FINAL=0", "RVL_FMA_F=0"],
    "no_int_tol": ["RVL_INT_TOL=0"],
    "no_final": ["RVL_FINAL=0"],
    "no_fma_f": ["RVL_FMA_F=0"],
}
if __name__ == "__main__":
    d = os.path.join(os.path.dirname(build.OUT), "variants")
    os.makedirs(d, exist_ok=True)
    for name, defs in VARIANTS.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        print(build.build(defines=defs, out=os.path.join(d, f"librvlnl_{name}.so")))

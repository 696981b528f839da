"""Build the kernel A/B variants of a round beside the product library (evidence_b200/variants/,
git-ignored like every .so; they travel to the GPU box).  usage: python tools/build_variants.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from evidence_b200 import build

VARIANTS = {
    "nopin": ["RVL_PIN_KTAB=0"],   # constants re-loaded by the compiler where it likes
    "no_fma_f": ["RVL_FMA_F=0"],   # two-rounding E - e sin E
    "norot": ["RVL_ROT_FUSED=0", "RVL_ROT_CD=0", "RVL_ROT_SHEAR=0"],  # rotation as old value + small correction (6 instructions)
    "noshear": ["RVL_ROT_SHEAR=0"],  # short series: four-FMA rotation instead of three in-place shears
    "nocd": ["RVL_ROT_CD=0"],        # long series: s' = fma(c, sd, fma(-s, v, s)) (one register move per pass)
    "nopinwc": ["RVL_PIN_WC=0"],     # the warp's constant-block address re-derived per planet
}
if __name__ == "__main__":
    d = os.path.join(os.path.dirname(build.OUT), "variants")
    os.makedirs(d, exist_ok=True)
    for f in os.listdir(d):
        os.remove(os.path.join(d, f))
    for name, defs in VARIANTS.items():
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        print(build.build(defines=defs, out=os.path.join(d, f"librvlnl_{name}.so")))

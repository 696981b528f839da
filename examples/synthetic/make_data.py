"""Write a synthetic two-instrument RV data set (BASELINE.json config 2 shape) as .rv files in the
format of the reference's examples (tab separated `rjd vrad svrad`, one dashed line under the header)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from evidence_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main(n_epochs=400):
    case = synth.make_case(2, seed=7, n_epochs=n_epochs)
    for name, tab in case.datadict().items():
        d = tab["data"]
        with open(os.path.join(HERE, f"{name}.rv"), "w") as f:
            f.write("rjd\tvrad\tsvrad\n---\t----\t-----\n")
            for t, v, s in zip(d["rjd"], d["vrad"], d["svrad"]):
                f.write(f"{t:.6f}\t{v:.4f}\t{s:.4f}\n")
    print("truth:", {k: round(v, 4) for k, v in case.truth.items()})


if __name__ == "__main__":
    main()

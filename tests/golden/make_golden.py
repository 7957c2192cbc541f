"""Regenerates tests/golden/oracle_vectors.json: homogenised tensors of the CPU oracle for a few parity
cases at fixed macro points.  The reference itself cannot run in this image (DOLFINx/PETSc absent), so
these are REGRESSION vectors of the oracle (which is pinned on the reference's known answers in
tests/test_oracle_pins.py), not outputs of the reference.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cases as K  # noqa: E402

NAMES = ["p2_smooth_n16_c1", "p2_laminate_wavy_n32_c2", "p3_smooth_n8_c3", "p2_inclusion_n16", "p2_fulltensor_strat_n9",
         "p3_fulltensor_shear_n5", "e2_hooke_sin_strat_n7", "e3_hooke_smooth_shear_n3", "e3_fibre_rot_n4", "e3_fibre_rot_n8_c4",
         "e3_hooke_smooth_shear_n6", "e3_cubic_shear_n4"]  # fmt: skip


def main():
    out = {}
    for name in NAMES:
        case = K.BY_NAME[name]
        prog = K.program(case)
        mic = K.oracle_cell(case, prog)
        x = K.points(case, 2, seed=2024)
        out[name] = {"x": x.tolist(), "A_hom": [K.oracle_tensor(case, mic, xi).tolist() for xi in x]}
        print(name, "done")
    with open(os.path.join(HERE, "oracle_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

"""Regression pins for the ORACLE itself: products of the reference's shipped fixture with seeded vectors, evaluated
by the SciPy CSC product of the restated sparse(A) (the reference's own test oracle, test/test_symmetricblockmatrix.jl
:72-97) and stored as tests/golden/products_<name>.npz. They are oracle-generated, not reference-generated (Julia
cannot run here): they freeze today's agreed C = NumPy = CSC values so that a later change to any of the three shows
up, and they give the GPU tests fixed known answers on the real fixture.

    python tests/golden/make_products.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle_np as O  # noqa: E402


def main():
    for name in ("cuboid", "sphere"):
        A = O.load_golden_sbm(name)
        n = A.size[0]
        rng = np.random.default_rng(20261018)
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        y0 = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        S = O.sparse_sbm(A)
        out = {"x": x, "y0": y0}
        for op, M in (("N", S), ("T", S.T), ("C", S.conj().T)):
            out[f"y_{op}"] = M @ x
            out[f"y5_{op}"] = 1j * (M @ x) + 2j * y0          # mul!(y, A, x, im, 2im)
        np.savez_compressed(Path(__file__).resolve().parent / f"products_{name}.npz", **out)
        print(name, n, float(np.linalg.norm(out["y_N"])))


if __name__ == "__main__":
    main()

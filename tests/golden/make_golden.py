"""Generates tests/golden/*.npz from the reference's OWN compiled code (oracle/_ref, built by
oracle/build_ref.py from /root/reference).  Run where /root/reference exists:

    OPENBLAS_NUM_THREADS=1 python tests/golden/make_golden.py

Inputs are NOT stored: every case is rebuilt from implicit_b200.synthetic with the seed recorded in
the file, so a fixture is (recipe, expected outputs of the reference).  Outputs are float32.
"""
import os
import sys

os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import oracle  # noqa: E402
from implicit_b200 import synthetic  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

#: name -> recipe
CASES = {
    "chol_f16": dict(users=400, items=250, nnz=4000, factors=16, use_cg=False, iterations=2, seed=11, neg=0.0),
    "chol_f64": dict(users=300, items=200, nnz=6000, factors=64, use_cg=False, iterations=2, seed=12, neg=0.0),
    "chol_f64_neg": dict(users=300, items=200, nnz=6000, factors=64, use_cg=False, iterations=1, seed=13, neg=0.1),
    "chol_f40": dict(users=200, items=150, nnz=3000, factors=40, use_cg=False, iterations=1, seed=14, neg=0.0),
    "cg_f32": dict(users=400, items=250, nnz=5000, factors=32, use_cg=True, iterations=3, seed=15, neg=0.0),
    "cg_f128": dict(users=250, items=200, nnz=25000, factors=128, use_cg=True, iterations=3, seed=16, neg=0.05),
}


def build_case(rc):
    Cui = synthetic.power_law_csr(rc["users"], rc["items"], rc["nnz"], rc["seed"], rc["neg"])
    X0, Y0 = synthetic.initial_factors(rc["users"], rc["items"], rc["factors"], seed=42)
    return Cui, X0, Y0


def main():
    ref = oracle.get("ref")
    for name, rc in CASES.items():
        Cui, X, Y = build_case(rc)
        oracle.fit(Cui, X, Y, regularization=0.01, iterations=rc["iterations"], use_cg=rc["use_cg"], kind="ref")
        loss = ref.calculate_loss(Cui, X, Y, 0.01)
        # one more USER half from this (well conditioned) state: the tight per-half parity fixture.
        # X, Y above double as its inputs; Xh is the expected output.
        Xh = X.copy()
        if rc["use_cg"]:
            ref.least_squares_cg(Cui, Xh, Y, 0.01, cg_steps=3)
        else:
            ref.least_squares(Cui, Xh, Y, 0.01)
        ids, scores = ref.topk(Y, X[:64], 10, filter_query_items=Cui[:64], filter_items=np.array([0, 3, 7]))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), X=X, Y=Y, Xh=Xh, loss=np.float64(loss), topk_ids=ids,
                            topk_scores=scores, **{"recipe_" + k: np.asarray(v) for k, v in rc.items()})
        print(name, "loss", loss, X.shape, Y.shape)


if __name__ == "__main__":
    main()

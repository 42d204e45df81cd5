"""Writes the golden vectors under tests/golden/ by running the ORACLE (oracle/fm_oracle.hpp).

The reference itself cannot run here (it needs Eigen, which is neither installed nor fetchable)
and ships no golden vectors, so these files pin the oracle restatement against regressions and
give the GPU tests a fixture that does not need the oracle to be built.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from helpers import block_data, toy_matrix  # noqa: E402
from oracle import binding as oracle  # noqa: E402


def chain_dump(chain, n):
    out = dict(w0=[], w=[], V=[], alpha=[], mu_w=[], lambda_w=[], mu_V=[], lambda_V=[])
    for _ in range(n):
        chain.step()
        w0, w, V = chain.fm()
        h = chain.hyper()
        out["w0"].append(w0), out["w"].append(w), out["V"].append(V)
        for k in ("alpha", "mu_w", "lambda_w", "mu_V", "lambda_V"):
            out[k].append(h[k])
    return {k: np.asarray(v) for k, v in out.items()}


def main():
    X, y = toy_matrix()
    for dtype in ("f64", "f32"):
        np.savez(os.path.join(HERE, f"toy_c1_{dtype}.npz"),
                 **chain_dump(oracle.OracleChain(X, y, 4, dtype=dtype, seed=42, n_iter=10), 10))
    ycls = y * 2 - 1
    np.savez(os.path.join(HERE, "toy_c1_classifier_f64.npz"),
             **chain_dump(oracle.OracleChain(X, ycls, 4, dtype="f64", task="classification", seed=42, n_iter=10), 10))
    X_flat, main_, users, items, yb, group_shapes = block_data()
    np.savez(os.path.join(HERE, "block_f64.npz"),
             **chain_dump(oracle.OracleChain(main_, yb, 2, X_rel=[users, items], dtype="f64", fit_w0=False,
                                             group_shapes=group_shapes, n_iter=10), 10))


if __name__ == "__main__":
    main()

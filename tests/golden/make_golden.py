"""Generates tests/golden/*.npz by running the REFERENCE's own scheduler/containers/solver (oracle/_ref, built from
/root/reference/include unmodified) on small seeded instances.  Run here (the reference tree is not on the GPU box):

    python tests/golden/make_golden.py

Each fixture holds the inputs and, per iteration, what the reference computed: lower bound after every ComputePass,
state (theta, a, b) after the last pass, and labeling / primal cost of a forward and a backward rounding pass.
Potts and truncated-linear instances are fed to the reference as materialised dense tables (the reference has no
truncated-linear factor and its Potts factor must equal the dense one, test/potts_factor.cpp:34-66).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from checkers import RefLP, grid_edges, potts_table, ref_library, trunclin_table  # noqa: E402


def run_reference(lib, theta, ei, ej, tables, mode, iters):
    ref = RefLP(lib, ["-v", "0"])
    for t in theta:
        ref.AddUnaryFactor(t)
    for e in range(len(ei)):
        ref.AddPairwiseFactor(int(ei[e]), int(ej[e]), tables[e])
    ref.Begin()
    ref.set_reparametrization(mode)
    out = {"lb0": ref.LowerBound(), "lb": []}
    for it in range(iters):
        ref.ComputePass(it)
        out["lb"].append(ref.LowerBound())
    out["theta"] = ref.unaries()
    out["a"], out["b"] = ref.messages()
    ref.ComputeForwardPassAndPrimal(iters)
    out["labels_fwd"], out["primal_fwd"] = ref.labeling().astype(np.uint32), ref.EvaluatePrimal()
    ref.ComputeBackwardPassAndPrimal(iters)
    out["labels_bwd"], out["primal_bwd"] = ref.labeling().astype(np.uint32), ref.EvaluatePrimal()
    out["theta_end"] = ref.unaries()
    out["lb"] = np.array(out["lb"])
    return out


def main():
    lib = ref_library()
    assert lib is not None, "build oracle/_ref first (make -C oracle)"
    rng = np.random.default_rng(20261017)
    cases = {}
    # dense tables on a grid, all three reparametrisation modes
    H, W, L = 7, 6, 4
    ei, ej = grid_edges(H, W)
    theta, tables = rng.random((H * W, L)), rng.random((len(ei), L, L))
    for mode in ("anisotropic", "uniform", "damped_uniform"):
        cases[f"grid_dense_{mode}"] = dict(kind="dense", H=H, W=W, L=L, mode=mode, theta=theta, ei=ei, ej=ej, tables=tables)
    # Potts grid (BASELINE config 1 shape, cropped), lambda of mixed sign
    H, W, L = 12, 10, 8
    ei, ej = grid_edges(H, W)
    lam = rng.uniform(-0.3, 1.0, size=len(ei))
    cases["grid_potts"] = dict(kind="potts", H=H, W=W, L=L, mode="anisotropic", theta=rng.random((H * W, L)), ei=ei, ej=ej,
                               lam=lam, tables=np.stack([potts_table(L, x) for x in lam]))
    # truncated linear grid (BASELINE config 2 shape, cropped)
    H, W, L = 9, 11, 16
    ei, ej = grid_edges(H, W)
    cases["grid_trunclin"] = dict(kind="trunclin", H=H, W=W, L=L, mode="anisotropic", theta=rng.random((H * W, L)), ei=ei, ej=ej,
                                  slope=0.25, trunc=1.0, tables=np.stack([trunclin_table(L, 0.25, 1.0)] * len(ei)))
    # random sparse graph, dense 8x8 tables (BASELINE config 4 shape, tiny)
    n, L = 30, 8
    edges = set()
    while len(edges) < 45:
        i, j = rng.integers(0, n, size=2)
        if i != j:
            edges.add((int(min(i, j)), int(max(i, j))))
    edges = sorted(edges)
    ei, ej = np.array([e[0] for e in edges], dtype=np.uint64), np.array([e[1] for e in edges], dtype=np.uint64)
    cases["graph_dense"] = dict(kind="dense", H=0, W=0, L=L, mode="anisotropic", theta=rng.random((n, L)), ei=ei, ej=ej,
                                tables=rng.random((len(ei), L, L)))
    for name, c in cases.items():
        out = run_reference(lib, c["theta"], c["ei"], c["ej"], c["tables"], c["mode"], iters=6)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: v for k, v in c.items()}, **{"ref_" + k: v for k, v in out.items()})
        print(name, "LB", out["lb0"], "->", out["lb"][-1], "primal", out["primal_fwd"], out["primal_bwd"])


if __name__ == "__main__":
    main()

"""Golden vectors for the cluster-accuracy mapping, produced by executing the reference's own `linear_assignment`
(/root/reference/vae/trainer.py:40-67, verbatim) against a numpy stand-in for the tf names it uses (argmax, zeros_like,
unique_with_counts - unique values in order of first occurrence -, where, one_hot, squeeze).

    python scripts/make_reference_cluster_golden.py      # needs /root/reference (build container only)
Writes tests/golden/reference_cluster.json; tests/test_host.py checks splitvae_b200.trainer.linear_assignment against it.
"""
import json
import os
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/vae/trainer.py"


def unique_with_counts(x):
    vals, counts = [], []
    for v in np.asarray(x).tolist():
        if v in vals:
            counts[vals.index(v)] += 1
        else:
            vals.append(v)
            counts.append(1)
    return np.asarray(vals, np.int64), None, np.asarray(counts, np.int64)


def main():
    tf = types.SimpleNamespace(argmax=lambda x, axis=None: np.argmax(x, axis=axis), zeros_like=np.zeros_like,
                               unique_with_counts=unique_with_counts, where=np.where, squeeze=np.squeeze,
                               one_hot=lambda idx, depth: np.eye(depth)[np.asarray(idx)])
    text = open(REF).read().split("\n")
    a = next(i for i, l in enumerate(text) if l.startswith("def linear_assignment("))
    b = next(i for i in range(a + 1, len(text)) if text[i].startswith("def ") or text[i].startswith("@"))
    ns = {"tf": tf, "np": np}
    exec("\n".join(text[a:b]), ns)
    rng = np.random.default_rng(7)
    cases = []
    for n, num_class, num_cluster in ((40, 4, 6), (300, 10, 30), (64, 10, 30), (12, 3, 3)):
        lab = rng.integers(0, num_class, n)
        pred = np.round(rng.standard_normal((n, num_cluster)), 2)      # (rounded BEFORE the reference runs: keeps the fixture small)
        if n == 12:                                             # exact ties inside clusters: first-met class wins
            lab = np.array([2, 0, 0, 2, 1, 1, 0, 2, 1, 0, 2, 1])
            pred = np.eye(3)[np.array([0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2])] + 0.0
        labels = np.eye(num_class)[lab]
        out = ns["linear_assignment"](labels, pred)
        cases.append({"labels": lab.tolist(), "num_class": num_class, "pred": pred.tolist(), "assigned": np.argmax(out, axis=1).tolist(),
                      "accuracy": float(np.mean(np.argmax(out, axis=1) == lab))})
    path = os.path.join(ROOT, "tests", "golden", "reference_cluster.json")
    with open(path, "w") as f:
        json.dump({"source": f"{REF} lines {a + 1}-{b} executed verbatim", "cases": cases}, f)
    print("wrote", path, [c["accuracy"] for c in cases])


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""tests/golden/bucketing.json: batches formed by the REFERENCE's own bucketing code for seeded length lists.
datamodule/data_module.py cannot be imported here (it needs pytorch_lightning), so the two pure pieces --
`_batch_by_token_count` and the body of `CustomBucketDataset.__init__` -- are extracted from the unmodified file with
`ast` and executed.  TEST INFRASTRUCTURE ONLY; run in the build container."""
import ast
import json
import os
import random

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/datamodule/data_module.py"


def reference_functions():
    tree = ast.parse(open(SRC).read())
    ns = {"torch": torch, "random": random}
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "_batch_by_token_count"]
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "CustomBucketDataset"][0]
    cls.bases = []                                         # drop torch.utils.data.Dataset: only __init__'s arithmetic matters
    init = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "__init__"][0]
    init.body = [st for st in init.body if not (isinstance(st, ast.Expr) and isinstance(st.value, ast.Call)
                                                and getattr(getattr(st.value.func, "value", None), "func", None) is not None
                                                and getattr(st.value.func.value.func, "id", "") == "super")]
    mod = ast.Module(body=keep + [cls], type_ignores=[])
    ast.fix_missing_locations(mod)
    exec(compile(mod, SRC, "exec"), ns)
    return ns["CustomBucketDataset"]


def main():
    Bucket = reference_functions()
    cases = []
    rng = random.Random(7)
    for n, lo, hi, mf, nb in [(200, 20, 400, 1600, 50), (57, 5, 600, 1600, 50), (1000, 30, 400, 1600, 50), (13, 100, 100, 800, 4),
                              (300, 1, 75, 200, 10)]:
        lengths = [rng.randint(lo, hi) for _ in range(n)]
        ds = Bucket(list(range(n)), lengths, mf, nb)
        cases.append(dict(lengths=lengths, max_frames=mf, num_buckets=nb, batches=[[int(i) for i in b] for b in ds.batches]))
    with open(os.path.join(ROOT, "tests", "golden", "bucketing.json"), "w") as f:
        json.dump(cases, f)
    print(len(cases), "cases;", sum(len(c["batches"]) for c in cases), "batches")


if __name__ == "__main__":
    main()

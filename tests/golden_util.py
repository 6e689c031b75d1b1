"""Helpers shared by the oracle and CUDA parity tests: load a golden fixture."""
import json
import os
import random

import numpy as np
import torch

from oracle import kelpie_oracle as ko

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def seed_all(seed):
    np.random.seed(seed)
    torch.manual_seed(seed)
    random.seed(seed)


def load(kind):
    z = np.load(os.path.join(GOLDEN, f"{kind.lower()}_small.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    kg = ko.KG(z["train"], z["valid"], z["test"], int(z["n_ent"]), int(z["n_rel"]))
    kw = {}
    if kind == "TransE":
        kw["norm"] = meta["params"]["norm"]
    elif kind == "ComplEx":
        kw["init_scale"] = meta["params"]["init_scale"]
    else:
        names = ["conv_w", "conv_b", "fc_w", "fc_b"] + [
            f"bn{i}_{s}" for i in (1, 2, 3) for s in ("w", "b", "mean", "var")
        ]
        kw["conve"] = {n: torch.from_numpy(z["w_" + n].copy()) for n in names}
        p = meta["params"]
        kw["dropout"] = (p["input_dropout_rate"], p["feature_map_dropout_rate"], p["hidden_dropout_rate"])
    w = ko.Weights(kind, z["w_ent"].copy(), z["w_rel"].copy(), **kw)
    order = {int(k): [tuple(t) for t in v] for k, v in meta["fact_order"].items()}
    return z, meta, kg, w, order


def trace_of(z, tag):
    n = int(z[tag + "n"])
    return [(z[f"{tag}{i}_init"], z[f"{tag}{i}_final"], z[f"{tag}{i}_res"]) for i in range(n)]

"""Helpers shared by the test modules: load the golden states dumped from the reference
(tests/golden/*.json(.gz), made by tests/golden/make_goldens.py) into any engine."""
from __future__ import annotations

import gzip
import json
import os

import numpy as np

from examl_b200.model import ModelConstants
from examl_b200.phylo import Entry

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_json(name: str) -> dict:
    path = os.path.join(GOLDEN_DIR, name)
    if os.path.exists(path + ".gz"):
        with gzip.open(path + ".gz", "rt") as f:
            return json.load(f)
    with open(path) as f:
        return json.load(f)


def constants_of(part: dict) -> ModelConstants:
    return ModelConstants(part["states"], np.array(part["EIGN"]), np.array(part["EV"]), np.array(part["EI"]),
                          np.array(part["tipVector"]), np.array(part["gammaRates"]), np.array(part["frequencies"]))


def alignment_of(part: dict, mxtips: int):
    """(tips uint8 [mxtips, width], wgt int32 [width]) of one dumped partition."""
    width = part["width"]
    tips = np.zeros((mxtips, width), dtype=np.uint8)
    for i, hexrow in enumerate(part["tips"]):
        tips[i] = np.frombuffer(bytes.fromhex(hexrow), dtype=np.uint8)
    return tips, np.array(part["wgt"], dtype=np.int32)


def traversal_of(state: dict):
    return [Entry(e["tipCase"], e["p"], e["q"], e["r"], list(e["qz"]), list(e["rz"]))
            for e in state["traversal"]["entries"]]


def load_engine(make_engine, state: dict, alignment_state: dict | None = None, **kw):
    """Instantiate an engine (make_engine(mxtips, nparts, nbranches, **kw)) from a golden state.
    `alignment_state` supplies tips/weights when `state` was dumped without them."""
    src = alignment_state or state
    mxtips, nparts, nb = state["mxtips"], state["NumberOfModels"], state["numBranches"]
    eng = make_engine(mxtips, nparts, nb, **kw)
    if state.get("rateHetModel", 1) == 0:
        eng.set_rate_model(0, state.get("maxCategories", 25))
    for m, part in enumerate(src["partitions"]):
        tips, wgt = alignment_of(part, mxtips)
        eng.set_partition(m, part["states"], wgt, tips)
    eng.commit()
    set_models(eng, state)
    return eng


def set_models(eng, state: dict) -> None:
    for m, part in enumerate(state["partitions"]):
        if "EIGN_LG4" in part:
            eng.set_model_lg4(m, part["EIGN_LG4"], part["EV_LG4"], part["EI_LG4"], part["tipVector_LG4"],
                              part["gammaRates"], part["weights"])
        else:
            eng.set_model(m, constants_of(part))
        if state.get("rateHetModel", 1) == 0:
            eng.set_categories(m, part["perSiteRates"], part["rateCategory"])

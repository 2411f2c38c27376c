"""Checkpoint fixtures shared by the CPU and GPU tests: the reference's snapshot weights (data) and a writer that
rebuilds a checkpoint FILE with the reference's own pickle structure around them."""
import json
import os
import pickle

import numpy as np

from conftest import GOLDEN


def fixture():
    g = np.load(os.path.join(GOLDEN, "snapshot_breakout_77.npz"))
    return [g["W%d" % i] for i in range(5)], [g["S%d" % i] for i in range(5)], g["q_kat"]


def rebuild(skel, arrays):
    """Inverse of make_snapshot_fixture.skeleton(): arrays are consumed in traversal order."""
    if isinstance(skel, dict):
        if skel.get("__ndarray__"):
            a = next(arrays)
            assert list(a.shape) == skel["shape"] or skel["shape"][0] == 18, (a.shape, skel["shape"])
            return a
        if "__seq__" in skel:
            items = [rebuild(v, arrays) for v in skel["items"]]
            return tuple(items) if skel["__seq__"] == "tuple" else items
        return {k: rebuild(v, arrays) for k, v in skel.items()}
    return skel


def write_checkpoint(path, layout, ws, ss):
    """A checkpoint file with the reference's own structure (keys, nesting, type strings) around (ws, ss)."""
    meta = json.load(open(os.path.join(GOLDEN, "snapshot_layouts.json")))
    skel = meta["breakout_77" if layout == "pre-1.0" else "seaquest_178"]["skeleton"]
    order = []
    if layout == "pre-1.0":
        for w, s in zip(ws, ss):            # dict traversal order of the skeleton: params before states
            order += [w, s]
        d = rebuild(skel, iter(order))
    else:
        # json sorted the keys: inside a layer dict "params" precedes "states"
        for w, s in zip(ws, ss):
            order += [w, s]
        d = rebuild(skel, iter(order))
        d["model"]["config"]["layers"][-1]["config"]["nout"] = int(ws[4].shape[0])
    with open(path, "wb") as f:
        pickle.dump(d, f, protocol=2)
    return d



"""CPU half of the checkpoint tests: the committed snapshot fixture is what the reference ships (live check in the
build container), both rebuilt checkpoint layouts parse, and the numpy oracle reproduces the Q-value KAT of
SURVEY §8(c) from them.  The device half is tests/test_gpu_checkpoint.py."""
import json
import os
import pickle
import zlib

import numpy as np
import pytest

from ckpt_helpers import fixture, write_checkpoint
from conftest import GOLDEN, needs_reference
from oracle import dqn_oracle as O


@pytest.mark.parametrize("layout", ["pre-1.0", "neon-1.3.0"])
def test_rebuilt_checkpoints_parse_and_reproduce_the_kat(tmp_path, layout):
    ws, ss, q_kat = fixture()
    path = str(tmp_path / "c.pkl")
    d = write_checkpoint(path, layout, ws, ss)
    if layout == "neon-1.3.0":
        assert d["neon_version"] == "1.3.0+344372b" and len(d["model"]["config"]["layers"]) == 9
    w2, s2 = O.load_snapshot(path)
    assert all((a == b).all() for a, b in zip(w2, ws)) and all((a == b).all() for a, b in zip(s2, ss))
    states = np.random.RandomState(1234).randint(0, 256, (32, 4, 84, 84)).astype(np.uint8)
    q = O.forward(w2, states)
    assert np.allclose(q[0], [4.052785, 3.199721, 5.557730, 4.043888], atol=2e-5)
    assert (q == q_kat).all()


@needs_reference
def test_fixture_is_the_reference_snapshot_bit_for_bit():
    ws, ss, _ = fixture()
    rw, rs = O.load_snapshot("/root/reference/snapshots/breakout_77.pkl")
    for a, b in zip(ws + ss, list(rw) + list(rs)):
        assert a.dtype == np.float32 and (a == np.asarray(b)).all()
    meta = json.load(open(os.path.join(GOLDEN, "snapshot_layouts.json")))
    with open("/root/reference/snapshots/seaquest_178.pkl", "rb") as f:
        d = pickle.load(f, encoding="latin1")
    layers = [l for l in d["model"]["config"]["layers"] if "params" in l]
    sk = [l for l in meta["seaquest_178"]["skeleton"]["model"]["config"]["layers"]["items"] if "params" in l]
    for l, k in zip(layers, sk):
        assert (zlib.crc32(np.ascontiguousarray(l["params"]["W"]).tobytes()) & 0xffffffff) == k["params"]["W"]["crc32"]

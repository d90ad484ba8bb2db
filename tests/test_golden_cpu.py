"""The oracle reproduces the committed golden fixtures (tests/golden/*.json, made by tests/golden/make_golden.py)."""
import json
import os

import pytest

from tests.golden import make_golden as G

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_oracle_matches_golden(name):
    want = json.load(open(os.path.join(HERE, name + ".json")))
    make, rep = G.CASES[name]
    got = G.digest(make(), rep)
    assert len(got) == len(want["scenes"])
    for gs, ws in zip(got, want["scenes"]):
        assert [(g["id_M"], g["id_N"], g["faces"], g["vertices"], g["key_checksum"]) for g in gs] == \
               [(w["id_M"], w["id_N"], w["faces"], w["vertices"], w["key_checksum"]) for w in ws]
        for g, w in zip(gs, ws):
            assert g["total_area"] == pytest.approx(w["total_area"], rel=1e-12)
            assert g["sum_face_pressure"] == pytest.approx(w["sum_face_pressure"], rel=1e-12)
            assert g["centroid"] == pytest.approx(w["centroid"], rel=1e-12, abs=1e-15)

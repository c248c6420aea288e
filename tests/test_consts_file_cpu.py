"""`.cnts` container (src/stark/stark_constsPolsFile.js) round trip on the host: framing, section ids, u32 length fields."""
import struct

import numpy as np
import pytest

from pil2_stark_js_b200 import stark_consts_file as CF


def _consts(rng):
    w, h = 5, 8
    return {"fixedPolsEvals": rng.integers(0, 2**63, size=40, dtype=np.uint64),
            "constTree": {"width": w, "height": h, "elements": rng.integers(0, 2**63, size=w * h, dtype=np.uint64),
                          "nodes": rng.integers(0, 2**63, size=8 * h - 4, dtype=np.uint64)},
            "x_n": rng.integers(0, 2**63, size=4, dtype=np.uint64), "x_ext": rng.integers(0, 2**63, size=8, dtype=np.uint64)}


def test_cnts_round_trip_and_framing(tmp_path):
    c = _consts(np.random.default_rng(1))
    fn = tmp_path / "a.cnts"
    CF.writePilStarkConstsFile(c, fn)
    raw = fn.read_bytes()
    assert raw[:4] == b"cnts" and struct.unpack("<II", raw[4:12]) == (1, 5)
    sid, size = struct.unpack("<IQ", raw[12:24])                    # first section: fixed pols evals, u32 count + words
    assert sid == CF.CONSTS_PS_CONST_POLS_EVALS_SECTION and size == 4 + 8 * 40
    assert struct.unpack("<I", raw[24:28]) == (40,)
    back = CF.readPilStarkConstsFile(fn)
    for k in ("fixedPolsEvals", "x_n", "x_ext"):
        assert np.array_equal(back[k], c[k])
    t, bt = c["constTree"], back["constTree"]
    assert (bt["width"], bt["height"]) == (t["width"], t["height"])
    assert np.array_equal(bt["elements"], t["elements"]) and np.array_equal(bt["nodes"], t["nodes"])


def test_cnts_rejects_bad_files(tmp_path):
    fn = tmp_path / "b.cnts"
    fn.write_bytes(b"r1cs" + struct.pack("<II", 1, 0))
    with pytest.raises(ValueError, match="Invalid File format"):
        CF.readPilStarkConstsFile(fn)
    fn.write_bytes(b"cnts" + struct.pack("<II", 1, 0))
    with pytest.raises(ValueError, match="Missing section"):
        CF.readPilStarkConstsFile(fn)
    c = _consts(np.random.default_rng(2))
    CF.writePilStarkConstsFile(c, fn)
    fn.write_bytes(fn.read_bytes()[:-16])
    with pytest.raises(ValueError, match="truncated"):
        CF.readPilStarkConstsFile(fn)

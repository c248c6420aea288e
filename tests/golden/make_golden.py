#!/usr/bin/env python3
"""Extract the reference's committed golden proof into tests/golden/sm_all_proof.json.

Run in the build container only (reads /root/reference, which does not exist on the GPU box).  Sources:
  test/compressor/verifier.proof.zkin.json   a real GL proof of test/state_machines/sm_all (2^10 rows,
                                             blowup 2, 8 queries, FRI steps 11 -> 7 -> 3, standard hash)
  test/compressor/verifier.circom:802        rootC (constant-tree root) literal
The output keeps the numbers only (decimal strings -> ints) under this repo's own key names.
"""
import json, re, pathlib

REF = pathlib.Path("/root/reference/test/compressor")
OUT = pathlib.Path(__file__).resolve().parent / "sm_all_proof.json"


def to_int(x):
    if isinstance(x, list):
        return [to_int(v) for v in x]
    return int(x)


def main():
    z = json.loads((REF / "verifier.proof.zkin.json").read_text())
    circom = (REF / "verifier.circom").read_text()
    m = re.search(r"signal rootC\[4\] <== \[([0-9, ]+)\]", circom)
    root_c = [int(v) for v in m.group(1).split(",")]
    g = {
        "source": "pil2-stark-js test/compressor/verifier.proof.zkin.json + verifier.circom:802",
        "stark_struct": {"nBits": 10, "nBitsExt": 11, "nQueries": 8, "steps": [11, 7, 3]},
        "publics": to_int(z["publics"]),
        "roots": {"const": root_c, "stage1": to_int(z["root1"]), "stage2": to_int(z["root2"]),
                  "stage3": to_int(z["root3"]), "stageQ": to_int(z["root4"]),
                  "fri1": to_int(z["s1_root"]), "fri2": to_int(z["s2_root"])},
        "evals": to_int(z["evals"]),
        "final_pol": to_int(z["finalPol"]),
        "layer0": {k: {"rows": to_int(z["s0_vals" + s]), "siblings": to_int(z["s0_siblings" + s])}
                   for k, s in (("const", "C"), ("stage1", "1"), ("stage2", "2"), ("stage3", "3"), ("stageQ", "4"))},
        "fri1": {"rows": to_int(z["s1_vals"]), "siblings": to_int(z["s1_siblings"])},
        "fri2": {"rows": to_int(z["s2_vals"]), "siblings": to_int(z["s2_siblings"])},
    }
    OUT.write_text(json.dumps(g, separators=(",", ":")))
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()

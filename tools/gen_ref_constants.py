#!/usr/bin/env python
"""Generate tests/golden/ref_constants.json by running the REFERENCE's own Math library
(oracle/_ref/ref_math_probe, built by oracle/Makefile from /root/reference/Math where it lies) on the
fixture cameras of App/MainBehaviour.cpp:19-76 and SURVEY.md §8(d).

The file is committed: it is what pins `final184_b200.scene`'s numpy mirror (tests/test_constants.py)
and it is what every run uses as camera constants, so /root/reference is not needed at run time.
Floats are stored as C99 hex strings (exact).
"""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from final184_b200.scene import FIXTURE_NODES, probe_nodes  # noqa: E402

PROBE = os.path.join(REPO, "oracle", "_ref", "ref_math_probe")
OUT = os.path.join(REPO, "tests", "golden", "ref_constants.json")


def run(*args):
    out = subprocess.check_output([PROBE, *[repr(float(a)) if not isinstance(a, str) else a for a in args]]).decode()
    d = {}
    for line in out.strip().splitlines():
        k, v = line.split(":")
        d[k.strip()] = v.split()
    return d


def main():
    if not os.path.exists(PROBE):
        print("gen_ref_constants: oracle/_ref/ref_math_probe missing (run `make -C oracle ref`)", file=sys.stderr)
        return 1
    nodes = dict(FIXTURE_NODES)
    nodes.update(probe_nodes())
    out = {}
    for name, (tr, eu, spec) in nodes.items():
        d = run("view", *eu, *tr, 1.0, 1.0, 1.0)
        d.update(run(spec[0], *spec[1:]))
        d["_node"] = {"translation": tr, "euler_deg": eu, "projection": list(spec)}
        out[name] = d
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    json.dump(out, open(OUT, "w"), indent=0, sort_keys=True)
    print(f"gen_ref_constants: {len(out)} cameras → {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())

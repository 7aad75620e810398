#!/usr/bin/env python
"""Golden vectors from the reference's own shader text.  Needs oracle/_ref/libf184_refshaders.so (`make -C oracle ref`,
possible only where /root/reference exists); writes tests/golden/refshader_<case>.npz: the mode R voxel volume (sparse),
fragment count, and for two frames the RGBA16F outputs of lighting_indirect (second frame = temporal path), gtao_visibility,
gtao_blur, indirect_blurX, indirect_blurY, plus the SHA-256 of the inputs they were computed from."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import refshader as R                      # noqa: E402
from final184_b200 import api as A         # noqa: E402


def main():
    olib = A.Library(os.path.join(REPO, "oracle", "_build", "libf184_oracle.so"), "f184o_", product=False)
    for case in R.CASES:
        if not R.case_available(case):
            print(f"{case}: scene not staged, skipped")
            continue
        sc, cams, fis = R.case_inputs(case)
        out = R.run_reference_shaders(olib, sc, cams, fis)
        g = R.pack_golden(out, R.input_digest(sc, fis))
        np.savez_compressed(R.golden_path(case), **g)
        print(f"{case}: {int(out['fragments'])} fragments, {len(g['vox_index'])} voxels -> {R.golden_path(case)} "
              f"({os.path.getsize(R.golden_path(case)) / 1024:.0f} KB)")


if __name__ == "__main__":
    main()

"""The C++ host side above the C-ABI (include/f184_renderer.hpp): CVoxelizeRenderer + the CMaterial full-screen-pass protocol, with the
reference's method names, resource names and error behaviour.  tests/cpp/frame_driver.cpp writes the voxel/indirect section of
CMegaPipeline::Render() the way the reference does; here it is compiled, run for two frames on the pinned atrium case and held, bit for
bit, to the same frames driven through the ctypes binding — against the CPU oracle everywhere, against libf184.so on the GPU box."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

import refshader as R
from final184_b200 import api as A
from final184_b200 import scene as S

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("voxels", "indirect1", "ao1", "blur1", "lighting1", "color1", "taa1")


def split_by_material(sc):
    """One CPrimitive per material (stable triangle order inside each); returns the primitives and the same scene re-flattened in
    that draw order (mode R's store race resolves in draw order, so both drivers must draw in the same one)."""
    prims, pos, nrm, uv, idx, tmat = [], [], [], [], [], []
    base = 0
    for m in sorted(set(sc.tri_mat.tolist())):
        tris = sc.idx[sc.tri_mat == m]
        used, inv = np.unique(tris.reshape(-1), return_inverse=True)
        p = dict(material=m, pos=sc.pos[used], nrm=sc.nrm[used], uv=sc.uv[used], idx=inv.reshape(-1, 3).astype(np.uint32))
        prims.append(p)
        pos.append(p["pos"]); nrm.append(p["nrm"]); uv.append(p["uv"]); idx.append(p["idx"] + base); tmat.append(np.full(len(tris), m, np.uint16))
        base += len(used)
    T = sum(len(i) for i in idx)
    flat = S.Scene(pos=np.concatenate(pos), nrm=np.concatenate(nrm), uv=np.concatenate(uv), idx=np.concatenate(idx).astype(np.uint32),
                   tri_mat=np.concatenate(tmat), tri_model=np.zeros(T, np.uint16), model_mats=sc.model_mats[:1], mat_tex=sc.mat_tex,
                   mat_factor=sc.mat_factor, textures=sc.textures, name=sc.name + "_by_material")
    return prims, flat


def write_blob(path, sc, prims, cams, fis):
    W, H = fis[0]["depth"].shape[1], fis[0]["depth"].shape[0]
    with open(path, "wb") as f:
        f.write(struct.pack("<4I", R.N, W, H, R.SH))
        f.write(struct.pack("<I", len(sc.textures)))
        for t in sc.textures:
            t = np.ascontiguousarray(t, np.uint8)
            f.write(struct.pack("<2I", t.shape[1], t.shape[0])); f.write(t.tobytes())
        f.write(struct.pack("<I", len(sc.mat_tex)))
        for k in range(len(sc.mat_tex)):
            f.write(struct.pack("<i", int(sc.mat_tex[k]))); f.write(np.asarray(sc.mat_factor[k], np.float32).tobytes())
        f.write(struct.pack("<I", len(prims)))
        m34 = np.ascontiguousarray(np.asarray(sc.model_mats[0], np.float32)[:3, :4])             # tc::Matrix3x4, row-major
        for p in prims:
            f.write(struct.pack("<2H2I", p["material"], 0, len(p["pos"]), p["idx"].size))
            for a, dt in ((p["pos"], np.float32), (p["nrm"], np.float32), (p["uv"], np.float32), (p["idx"], np.uint32)):
                f.write(np.ascontiguousarray(a, dt).tobytes())
            f.write(m34.tobytes())
        f.write(bytes(A.view_constants_c(cams["voxel"])))
        k0 = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, 0, True)
        pl, dl = R.light_lists(k0)
        f.write(bytes(pl)); f.write(bytes(dl))
        f.write(struct.pack("<I", len(fis)))
        for i, fi in enumerate(fis):
            f.write(bytes(A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, i, i == 0)))
            for key in ("depth", "normals", "shadow", "albedo", "material"):
                f.write(np.ascontiguousarray(fi[key]).tobytes())


def read_outputs(path, W, H):
    raw = np.fromfile(path, np.uint8)
    n3 = R.N ** 3 * 4
    out, off = {"voxels": raw[:n3].view(np.uint16).reshape(R.N, R.N, R.N, 2)}, n3
    for k in KEYS[1:]:
        out[k] = raw[off:off + W * H * 8].view(np.uint16).reshape(H, W, 4)
        off += W * H * 8
    assert off == raw.size
    return out


def build_driver(tmp_path, oracle):
    exe = str(tmp_path / ("frame_driver_oracle" if oracle else "frame_driver"))
    libdir = os.path.join(REPO, "oracle", "_build") if oracle else os.path.join(REPO, "final184_b200", "csrc")
    lib = "f184_oracle" if oracle else "f184"
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(REPO, "include"), "-I", os.path.join(REPO, "tests", "cpp"),
           os.path.join(REPO, "tests", "cpp", "frame_driver.cpp"), "-o", exe, "-L", libdir, "-l" + lib, "-Wl,-rpath," + libdir]
    if oracle:
        cmd.insert(1, "-DF184_ORACLE")
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def run_case(tmp_path, oracle, check_lib):
    sc, cams, fis = R.case_inputs("atrium")
    prims, flat = split_by_material(sc)
    assert len(prims) >= 2
    blob, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    write_blob(blob, flat, prims, cams, fis)
    r = subprocess.run([build_driver(tmp_path, oracle), blob, outp], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    H, W = fis[0]["depth"].shape
    got = read_outputs(outp, W, H)
    want = R.run_library(check_lib, flat, cams, fis)
    assert R.compare(got, want, KEYS) == []
    assert (got["color1"][..., :3] != 0).mean() > 0.5 and got["voxels"].any()


def test_cpp_frame_through_the_reference_protocol_on_the_oracle(tmp_path, oracle_lib):
    run_case(tmp_path, True, oracle_lib)


@pytest.mark.gpu
def test_cpp_frame_through_the_reference_protocol_on_the_gpu(tmp_path, cuda_lib, oracle_lib):
    run_case(tmp_path, False, oracle_lib)


def test_create_failure_throws_like_the_reference(tmp_path):
    """f184::CVoxelGI turns a failed f184_create into an exception (the reference throws CRHIRuntimeError): a grid edge that is not a
    power of two is refused before any device work, so this runs without a GPU against libf184.so itself."""
    src = tmp_path / "t.cpp"
    src.write_text("""
#include "f184_renderer.hpp"
#include <cstdio>
int main()
{
    f184_config c{};
    c.mode = F184_MODE_NORTHSTAR; c.grid_n = 100; c.width = 8; c.height = 8; c.shadow_res = 16;
    try { f184::CVoxelGI g(c); } catch (const f184::CRuntimeError& e) { std::puts(e.what()); return 0; }
    return 1;
}
""")
    libdir = os.path.join(REPO, "final184_b200", "csrc")
    exe = str(tmp_path / "t")
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-I", os.path.join(REPO, "include"), str(src), "-o", exe, "-L", libdir, "-lf184", "-Wl,-rpath," + libdir],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "power of two" in r.stdout, r.stdout + r.stderr

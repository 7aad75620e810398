#!/usr/bin/env python
"""bench.py — the voxel-GI hot path of Final184 on B200, measured the way BASELINE.json names it:
ms/frame for voxelize(+normalise) + inject + six-direction mips + diffuse/specular cone trace.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--grid 512] [--width 3840] [--height 2160]

One "step" = one frame of the path over the Sponza fixture (App/MainBehaviour.cpp:19-76; the procedural
atrium of final184_b200/scene.py if the Sponza pack is not staged — config.workload says which).

  value          device time per frame, inputs resident in HBM, CUDA events on the context's stream, K frames
                 bracketed by barrier + synchronize, max over ranks
  e2e            the same frame through the C-ABI with HOST buffers: depth / normals / material / shadow copied
                 from pinned host memory and the traced image read back, every frame, inside the timed region
  roofline       the dominant kernel of the step (chosen from the per-stage device times accumulated over the same
                 timed region) against the bound that applies to it: the texture pipe for the cone tracer (trilinear fetch
                 rate, measured live by f184_microbench), HBM (MEASURED_PEAKS.json) for the volume stages; algorithmic bytes
                 are lower bounds (what must move, not what the kernel moves); roofline_stages lists every stage the same way
  cpu_baseline   the CPU oracle (oracle/, the "straight C++ transcription") on a bounded sample
  c4_scaling     BASELINE configs[3] (8 x tiled Sponza, 1024^3, 4K) at the same N: ms/frame, per-stage min/max over the ranks,
                 bytes and GB/s of the NVLink gather — the configuration north_star's scaling target is quoted on
  parity_vs_1gpu (N > 1) before the timed region every rank also computes the whole frame alone and compares: the gathered texture
                 storage and its own image rows must equal the one-GPU frame bit for bit
  spec_delta     (N = 1) the image difference between the amended cone tracer and SURVEY.md Appendix B as written
  --impl reference   the same metric from the CPU oracle alone (rank 0), each step a bounded sample

N > 1 (torchrun, one rank per GPU): see DESIGN.md "Multi-GPU".
The product path is libf184.so only; the oracle is loaded solely for the cpu_baseline / reference legs.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

METRIC = "ms/frame voxelize+inject+mip+cone-trace"
UNIT = "ms/frame"
ORACLE_SO = os.path.join(REPO, "oracle", "_build", "libf184_oracle.so")
ORACLE_NATIVE_SO = os.path.join(REPO, "oracle", "_build", "libf184_oracle_native.so")


def host_threads():
    """Threads the CPU legs may use: every core this process is allowed on."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def timing_oracle():
    """The copy of the oracle that is TIMED: same sources, -O3 -march=native, built on the box it runs on (a stamp
    of the CPU's model + flags and of the sources keeps a copy built elsewhere, or from older sources, from being reused).  Falls back to the checker build (-O2) if the compiler is
    not there.  -> (path, flags description)"""
    import glob
    import hashlib
    try:       # the CPU the copy was built for: model AND instruction-set flags (cloud CPUs share a generic model name)
        info = [l.strip() for l in open("/proc/cpuinfo") if l.startswith(("model name", "flags"))][:2]
    except Exception:
        info = ["unknown"]
    hsh = hashlib.sha1("\n".join(info).encode())
    for f in sorted(glob.glob(os.path.join(REPO, "oracle", "*.cpp")) + glob.glob(os.path.join(REPO, "oracle", "*.h")) + [os.path.join(REPO, "oracle", "Makefile")]
                    + glob.glob(os.path.join(REPO, "include", "*.h")) + glob.glob(os.path.join(REPO, "final184_b200", "csrc", "*.h"))):
        hsh.update(open(f, "rb").read())           # ... and the sources it was built from (a stale copy lacks newer entry points)
    model = hsh.hexdigest()
    stamp = ORACLE_NATIVE_SO + ".host"
    fresh = os.path.exists(ORACLE_NATIVE_SO) and os.path.exists(stamp) and open(stamp).read() == model
    if not fresh:
        if os.path.exists(ORACLE_NATIVE_SO):
            os.remove(ORACLE_NATIVE_SO)
        r = subprocess.run(["make", "-C", os.path.join(REPO, "oracle"), "native"], capture_output=True, text=True)
        if r.returncode == 0:
            open(stamp, "w").write(model)
            fresh = True
    if fresh:
        return ORACLE_NATIVE_SO, "-O3 -march=native -fopenmp -ffp-contract=off"
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle")], stdout=subprocess.DEVNULL)
    return ORACLE_SO, "-O2 -fopenmp -ffp-contract=off (native build unavailable)"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=100)
    p.add_argument("--warmup", type=int, default=10)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--grid", type=int, default=512)
    p.add_argument("--width", type=int, default=3840)
    p.add_argument("--height", type=int, default=2160)
    p.add_argument("--shadow", type=int, default=2048)
    p.add_argument("--workload", default="c3", choices=["c3", "c2", "c4"],
                   help="c3 (default, the headline): Sponza 512^3 / 3840x2160; c2: Sponza 256^3 / 1920x1080 (BASELINE configs[1]); "
                        "c4: 8 x tiled Sponza, 2.1 M triangles, 1024^3 / 3840x2160 (configs[3]).  The driver's bench line is c3.")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-overlap", action="store_true", help="F184_FLAG_NO_OVERLAP: every pass on one stream (A/B of the frame overlap)")
    p.add_argument("--schedule", default=None, choices=[None, "slab", "replicate"], help="multi-GPU schedule (default: slab)")
    p.add_argument("--cpu-budget-s", type=float, default=20.0, help="target CPU seconds of the cpu_baseline sample")
    p.add_argument("--no-c4", action="store_true", help="skip the c4_scaling block (8 x tiled Sponza at 1024^3 beside the headline workload)")
    p.add_argument("--no-extras", action="store_true", help="skip parity_vs_1gpu, spec_delta, c1_reference_mode and the secondary passes")
    a = p.parse_args()
    if a.workload == "c2":
        a.grid, a.width, a.height = 256, 1920, 1080
    elif a.workload == "c4":
        a.grid = 1024
    return a


# ------------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------------
def make_workload(args, rank=0, world=1):
    from final184_b200 import scene as S
    from final184_b200.fixture import frame_inputs
    sc = S.get_scene(prefer_sponza=True, seed=1, n_boxes=48, tex_size=256, subdiv=8)
    cams = {n: S.fixture_constants(n) for n in ("main", "shadow", "voxel")}
    if getattr(args, "workload", "c3") == "c4":          # 8 instances on a 2 x 2 x 2 lattice under the C4 voxel camera
        sc = S.tile_scene(sc, S.C4_OFFSETS)
        cams["voxel"] = S.fixture_constants("voxel_c4")
    if world > 1:
        # the G-buffer / shadow-map synthesiser is a CPU rasteriser: rank 0 renders (all cores) into the on-disk cache, the rest load it
        import torch.distributed as dist
        if rank == 0:
            frame_inputs(sc, cams["main"], cams["shadow"], args.width, args.height, args.shadow, 0)
        dist.barrier()
    fi = frame_inputs(sc, cams["main"], cams["shadow"], args.width, args.height, args.shadow, 0)
    name = ("8 x tiled Sponza" if getattr(args, "workload", "c3") == "c4" else ("Sponza" if sc.name == "sponza" else sc.name)) + f" {args.grid}^3 voxel GI (voxelize+normalise+inject+6-dir mips+" \
        f"6 diffuse/1 specular cones) at {args.width}x{args.height}, north-star mode"
    return sc, cams, fi, name


def mip_chain_bytes(n, b=4):
    """bytes the six-direction chain must move (read source once, write six outputs)."""
    total = b * n ** 3 + 6 * b * (n // 2) ** 3
    s = n // 2
    while s >= 2:
        total += 6 * b * s ** 3 + 6 * b * (s // 2) ** 3
        s //= 2
    return total


def algorithmic_bytes(stage, N, P, sc, counters):
    """Per-launch algorithmic bytes of each stage: LOWER bounds — what the stage must move whatever its data structure (DESIGN.md
    "Roofline"), so that traffic / algorithmic >= 1 and a fraction of the HBM peak cannot read above what the kernel moves.
    counters: fragments, bricks (listed 8^3 bricks), occupied (voxels)."""
    V, T = len(sc.pos), sc.n_tris
    occ, bricks = counters["occupied"], counters["bricks"]
    if stage == "voxelize":      # indexed triangle fetch + two 16-byte reductions per fragment
        return 32 * V + 12 * T + 32 * counters["fragments"]
    if stage == "normalise":     # per OCCUPIED voxel: read 32 B of sums, write 32 B of zeros (= next frame's clear), write 8 B albedo + normal
        return (32 + 32 + 8) * occ
    if stage == "inject":        # per occupied voxel: read 8 B (albedo, normal) + a 4 B shadow tap, write 4 B radiance
        return (8 + 4 + 4) * occ
    if stage == "mips":
        if counters.get("dense_mips"):
            return mip_chain_bytes(N)
        # sparse chain: read the occupied level-0 texels once; write levels 1-3 of the listed bricks once (6 x (64 + 8 + 1) texels);
        # then the dense tail: read level 3 once, write levels >= 4 once
        n3 = N // 8
        tail = 6 * 4 * n3 ** 3 + 6 * 4 * sum((n3 >> l) ** 3 for l in range(1, n3.bit_length()))
        return 4 * occ + bricks * 6 * 73 * 4 + tail
    if stage == "trace":         # per-pixel fixed I/O — depth 4, normal 8, material 4, history 8, out 8
        return 32 * P
    return 0


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx, self.rows, self.proc = device_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


# ------------------------------------------------------------------------------------------------------
# CPU legs (the oracle: test infrastructure, loaded ONLY here)
# ------------------------------------------------------------------------------------------------------
class CpuPath:
    """Bounded sample of one frame on the host cores through the oracle's mirror API.
    A frame = voxelize(all triangles) + inject + mips + trace(all rows).  The sample runs voxelize on one
    triangle chunk out of `vchunks` (rotating per step) and the trace on `bands` stratified bands covering
    1/`tfrac` of the rows, against a FULL volume built once at set-up; inject and mips run in full."""

    def __init__(self, args, sc, cams, fi, vchunks=1, tfrac=16, bands=2):
        import ctypes as C
        from final184_b200 import api as A
        self.A, self.args, self.sc, self.cams = A, args, sc, cams
        # torchrun exports OMP_NUM_THREADS=1 to its workers: say explicitly how many threads the CPU arm gets, and report what it got
        want = host_threads()
        os.environ["OMP_NUM_THREADS"] = str(want)
        so, self.build_flags = timing_oracle()
        self.lib = A.Library(so, "f184o_", product=False)
        self.lib.dll.f184o_omp_threads.argtypes = [C.c_int]
        self.lib.dll.f184o_omp_threads.restype = C.c_int
        self.cores = int(self.lib.dll.f184o_omp_threads(want))
        self.vchunks, self.tfrac, self.bands = vchunks, tfrac, bands
        self.k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], args.width, args.height, 0, True)
        mk = lambda: A.VoxelGI(args.grid, args.width, args.height, A.MODE_NORTHSTAR, shadow_res=args.shadow, lib=self.lib)
        self.full, self.part = mk(), mk()
        for c in (self.full, self.part):
            c.upload_scene(sc)
            for slot, key in ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_SHADOW, "shadow"), (A.SLOT_MATERIAL, "material")):
                c.upload(slot, fi[key])
        # full volume for the trace sample (set-up, untimed)
        self.full.voxelize(cams["voxel"]); self.full.inject(self.k); self.full.build_mips()
        self.step_no = 0

    def describe(self):
        return (f"oracle (C++/OpenMP, {self.build_flags}) on {self.cores} threads (measured inside a parallel region); per step: voxelize+normalise of triangle chunk k/{self.vchunks} (rotating, x{self.vchunks}), "
                f"inject + mips in full, trace of {self.bands} stratified bands = 1/{self.tfrac} of the rows (x{self.tfrac}) against the full volume")

    def step(self):
        """-> (estimated full-frame ms, wall ms of the sample, stage dict)"""
        A, a = self.A, self.args
        T = self.sc.n_tris
        ch = self.step_no % self.vchunks
        self.step_no += 1
        first = ch * T // self.vchunks
        count = (ch + 1) * T // self.vchunks - first
        t0 = time.perf_counter()
        self.part.set_triangle_range(first, count)
        self.part.voxelize(self.cams["voxel"])
        t1 = time.perf_counter()
        self.part.inject(self.k)
        t2 = time.perf_counter()
        self.part.build_mips()
        t3 = time.perf_counter()
        rows_per_band = max(1, a.height // (self.tfrac * self.bands))
        traced = 0
        for b in range(self.bands):
            y0 = b * a.height // self.bands + (self.step_no * rows_per_band) % max(1, a.height // self.bands - rows_per_band)
            self.full.set_trace_rows(y0, y0 + rows_per_band)
            self.full.trace_indirect(self.k)
            traced += rows_per_band
        t4 = time.perf_counter()
        st = {"voxelize+normalise": (t1 - t0) * 1e3 * self.vchunks, "inject": (t2 - t1) * 1e3, "mips": (t3 - t2) * 1e3,
              "trace": (t4 - t3) * 1e3 * a.height / traced}
        return sum(st.values()), (t4 - t0) * 1e3, st


REFSH_SO = os.path.join(REPO, "oracle", "_ref", "libf184_refshaders.so")


def c1_reference_mode(sc, cams, device, cpu_budget_s, gpu=True):
    """BASELINE configs[0], the only configuration the reference itself can run and the sizes its shaders hard-code:
    Sponza voxelized at 128^3 (centre-sample, last writer wins) + the 4 x 2 stochastic 60-step marches at 1280 x 720
    (+ GTAO, + the bilateral blur tail) — F184_MODE_REFERENCE.  GPU: CUDA stage times over 20 frames.  CPU, beside it:
    the reference's OWN shader text (oracle/_ref/libf184_refshaders.so: indirect.frag, gtao.frag, blur*.frag and main.lua's
    voxel stages compiled by g++ where they lie) on the host cores, voxel pass in full, screen passes on a bounded band of
    rows scaled to the frame.  Reported next to the headline; not part of it."""
    import ctypes as C
    from final184_b200 import api as A
    from final184_b200.fixture import frame_inputs
    N, W, H, SH = 128, 1280, 720, 2048
    fi = frame_inputs(sc, cams["main"], cams["shadow"], W, H, SH, 0)
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, 0, True)
    out = {"workload": f"{'Sponza' if sc.name == 'sponza' else sc.name} 128^3 centre-sample voxelization + 8 x 60-step march at 1280x720 "
                       "(+ GTAO, bilateral blur), reference-faithful mode"}
    if gpu:
        c = A.VoxelGI(N, W, H, A.MODE_REFERENCE, shadow_res=SH, device=device)
        c.upload_scene(sc)
        for slot, key in ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_SHADOW, "shadow"), (A.SLOT_MATERIAL, "material"),
                          (A.SLOT_ALBEDO, "albedo")):
            c.upload(slot, fi[key])
        frames = 20
        def frame():
            c.voxelize(cams["voxel"]); c.gtao(cams["main"]); c.trace_indirect(k); c.blur_indirect(k); c.lighting_deferred(k); c.composite(k)
        for _ in range(3):
            frame()
        c.sync()
        c.stage_time_reset(True)
        t0 = time.perf_counter()
        for _ in range(frames):
            frame()
        c.sync()
        wall = (time.perf_counter() - t0) * 1e3 / frames
        st = {}
        for s_ in range(A.STAGE_COUNT):
            tot, runs = c.stage_total_ms(s_)
            if runs:
                st[A.STAGE_NAMES[s_]] = round(tot / frames, 4)
        c.stage_time_reset(False)
        steps = c.counter(A.COUNTER_MARCH_STEPS)
        # ms_per_frame is the WALL time of a frame submitted from this Python driver (ctypes, ~25 launches, 12 event records per frame: at 2.6 ms
        # of device work the loop is sensitive to the host core it lands on — 2.65 ms on one box, 7.7 ms on another whose host was busy with
        # the other bench legs); device_ms_per_frame is the sum of the stages' CUDA-event times, the part that belongs to the kernels
        out["gpu"] = {"ms_per_frame": round(wall, 4), "device_ms_per_frame": round(sum(st.values()), 4), "stages_ms": st,
                      "fragments": c.counter(A.COUNTER_FRAGMENTS), "march_steps": steps,
                      "gmarch_steps_per_s": round(steps / (st["trace"] * 1e-3) / 1e9, 2)}
        c.close()
    if os.path.exists(REFSH_SO) and cpu_budget_s > 0:
        dll = C.CDLL(REFSH_SO)
        vp = C.c_void_p
        dll.refsh_indirect.argtypes = [C.POINTER(A.TraceConstantsC), vp, vp, vp, vp, vp, C.c_int, C.c_int, vp]
        dll.refsh_gtao.argtypes = [C.POINTER(A.ViewConstantsC), vp, vp, C.c_int, C.c_int, vp]
        dll.refsh_gtao_blur.argtypes = [vp, C.c_int, C.c_int, vp]
        dll.refsh_blur.argtypes = [C.c_int, C.POINTER(A.EngineMiscsC), vp, vp, C.c_int, C.c_int, vp]
        if not os.path.exists(ORACLE_SO):
            subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle")], stdout=subprocess.DEVNULL)
        olib = A.Library(ORACLE_SO, "f184o_", product=False)
        os.environ["OMP_NUM_THREADS"] = str(host_threads())
        olib.dll.f184o_omp_threads.argtypes = [C.c_int]
        olib.dll.f184o_omp_threads.restype = C.c_int
        omp_threads = int(olib.dll.f184o_omp_threads(host_threads()))
        hooks = olib.dll.f184o_debug_set_voxel_stage_hooks
        hooks.argtypes = [vp, vp, vp]
        o = A.VoxelGI(N, W, H, A.MODE_REFERENCE, shadow_res=SH, lib=olib)
        o.upload_scene(sc)
        hooks(o.h, C.cast(dll.refsh_voxel_gs, vp), C.cast(dll.refsh_voxel_ps, vp))
        t0 = time.perf_counter()
        o.voxelize(cams["voxel"])                      # rasteriser = the oracle's fixed-function stage, shaders = the reference's text
        t_vox = (time.perf_counter() - t0) * 1e3
        vox = o.readback(A.SLOT_VOXELS).copy()
        o.close()
        depth, normals, shadow = (np.ascontiguousarray(fi[k_]) for k_ in ("depth", "normals", "shadow"))
        hist, img = np.zeros((H, W, 4), np.uint16), np.zeros((H, W, 4), np.uint16)
        tmp = np.zeros((H, W, 4), np.uint16)
        ptr = lambda a: a.ctypes.data
        vc = A.view_constants_c(cams["main"])
        # a band of rows through the middle of the screen; grow it until the march alone has used ~half the budget
        rows, t_ind = 4, 0.0
        while True:
            y0 = H // 2 - rows // 2
            dll.refsh_set_rows(y0, y0 + rows)
            t0 = time.perf_counter()
            dll.refsh_indirect(C.byref(k), ptr(depth), ptr(normals), ptr(shadow), ptr(vox), ptr(hist), W, H, ptr(img))
            t_ind = time.perf_counter() - t0
            if t_ind > 0.35 * cpu_budget_s or rows >= H:
                break
            rows = min(H, rows * 2)
        t0 = time.perf_counter()
        dll.refsh_gtao(C.byref(vc), ptr(depth), ptr(normals), W, H, ptr(tmp))
        t_gtao = time.perf_counter() - t0
        t0 = time.perf_counter()
        dll.refsh_gtao_blur(ptr(tmp), W, H, ptr(hist))
        dll.refsh_blur(0, C.byref(k.miscs), ptr(img), ptr(depth), W, H, ptr(tmp))
        dll.refsh_blur(1, C.byref(k.miscs), ptr(tmp), ptr(depth), W, H, ptr(hist))
        t_blur = time.perf_counter() - t0
        dll.refsh_lighting_deferred.argtypes = [C.POINTER(A.ViewConstantsC), C.POINTER(A.ExtendedMatricesC), C.POINTER(A.LightListC),
                                                C.POINTER(A.LightListC), vp, vp, vp, vp, vp, C.c_int, C.c_int, vp]
        albedo, material = (np.ascontiguousarray(fi[k_]) for k_ in ("albedo", "material"))
        pl, dl = A.LightListC(), A.light_list_c([(tuple(k.sun.luminance), tuple(k.sun.position))])
        t0 = time.perf_counter()
        dll.refsh_lighting_deferred(C.byref(k.view), C.byref(k.ext), C.byref(pl), C.byref(dl), ptr(albedo), ptr(normals), ptr(depth),
                                    ptr(shadow), ptr(material), W, H, ptr(tmp))
        t_light = time.perf_counter() - t0
        dll.refsh_composite.argtypes = [C.POINTER(A.TraceConstantsC), vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp]
        col, taa_in, taa_out = (np.zeros((H, W, 4), np.uint16) for _ in range(3))
        t0 = time.perf_counter()
        dll.refsh_composite(C.byref(k), ptr(albedo), ptr(hist), ptr(depth), ptr(tmp), ptr(shadow), ptr(img), ptr(taa_in), W, H, ptr(col), ptr(taa_out))
        t_comp = time.perf_counter() - t0
        dll.refsh_set_rows(0, 1 << 30)
        sc_ = H / rows
        st = {"voxelize": round(t_vox, 2), "trace": round(t_ind * 1e3 * sc_, 1), "gtao": round(t_gtao * 1e3 * sc_, 1), "blur": round(t_blur * 1e3 * sc_, 1),
              "lighting": round(t_light * 1e3 * sc_, 1), "composite": round(t_comp * 1e3 * sc_, 1)}
        out["cpu_reference"] = {"ms_per_frame": round(sum(st.values()), 1), "stages_ms": st, "cores": omp_threads, "kind": "reference",
                                "sample": f"the reference's own GLSL compiled by g++ (oracle/_ref/libf184_refshaders.so), OpenMP on all cores; voxel pass in full, "
                                          f"screen passes on rows [{y0}, {y0 + rows}) of {H} scaled x{sc_:.1f}"}
    return out


def run_reference(args, rank, world):
    """--impl reference: the CPU restatement of the path, rank 0 only."""
    if rank != 0:
        return
    sc, cams, fi, wname = make_workload(args)
    per_step_budget = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    # size the sample from the budget: one chunk of voxelize (~6 s / vchunks at 512^3 on 16 cores) + mips + trace fraction
    # (512^3 / 4K on 16 cores: full voxelize ~4 s, inject+mips ~2 s, 1/16 of the trace ~3 s)
    small = per_step_budget < 8
    cpu = CpuPath(args, sc, cams, fi, vchunks=4 if small else 1, tfrac=32 if small else 16, bands=2)
    for _ in range(args.warmup):
        cpu.step()
    est, wall, stages = [], [], []
    for _ in range(args.steps):
        e, w, s = cpu.step()
        est.append(e); wall.append(w); stages.append(s)
    v = float(np.mean(est))
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": float(np.mean(wall)), "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "u8/i64/f32",
           "data": "synthetic", "config": {"workload": wname, "grid": args.grid, "width": args.width, "height": args.height},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": cpu.cores, "kind": "port", "sample": cpu.describe()},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "stages_ms": {k: float(np.mean([s[k] for s in stages])) for k in stages[0]}, "gpu_launches": 0}
    if sc.name == "sponza" and os.path.exists(REFSH_SO):
        out["c1_reference_mode"] = c1_reference_mode(sc, cams, 0, min(10.0, per_step_budget), gpu=False)
    emit(json.dumps(out))


# ------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------
INPUT_SLOTS = None


def _slots(A):
    return ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_MATERIAL, "material"), (A.SLOT_SHADOW, "shadow"))


class Arm:
    """One workload on this rank's GPU: the sharded context, its pinned host inputs, and the timed loops."""

    def __init__(self, A, torch, dist, N, W, H, SH, sc, cams, fi, rank, world, local_rank, stream, schedule=None, flags=0):
        from final184_b200.dist import ShardedVoxelGI
        self.A, self.torch, self.dist = A, torch, dist
        self.N, self.W, self.H, self.world, self.rank, self.local_rank, self.stream = N, W, H, world, rank, local_rank, stream
        self.sc, self.cams = sc, cams
        self.g = ShardedVoxelGI(grid_n=N, width=W, height=H, shadow_res=SH, device=local_rank, rank=rank, nranks=world, scene=sc,
                                voxel_cam=cams["voxel"], mode=schedule, flags=flags)
        self.g.ctx.set_stream(stream.cuda_stream)
        self.g.connect()
        self.k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, 0, True)
        self.pinned = {}
        for slot, key in _slots(A):
            self.pinned[slot] = torch.from_numpy(np.ascontiguousarray(fi[key]).view(np.uint8).reshape(-1)).pin_memory()
        self.rows = world > 1
        self.upload_inputs(shadow=True)
        self.g.ctx.sync()

    def upload_inputs(self, shadow=False):
        """the G-buffer of a frame (this rank's rows of it); the shadow map only when the light changes — it is static here"""
        A = self.A
        for slot, t in self.pinned.items():
            if slot == A.SLOT_SHADOW and not shadow:
                continue
            self.g.ctx.upload_ptr(slot, t.data_ptr(), t.numel(), rows=self.rows and slot != A.SLOT_SHADOW)

    def frame(self):
        self.g.frame(self.cams["voxel"], self.k)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, steps, warmup, clocks=False):
        """K frames, inputs resident in HBM: CUDA events on the pass stream, barrier + synchronize on both sides; per-stage device
        times accumulated by the library over the same region (no sync inside it)."""
        A, torch, g = self.A, self.torch, self.g
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):
                self.frame()
            g.ctx.stage_time_reset(True)
            l0 = g.ctx.counter(A.COUNTER_KERNEL_LAUNCHES)
            sampler = ClockSampler(self.local_rank) if clocks else None
            if sampler:
                sampler.start()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.barrier()
            ev0.record(self.stream)
            for _ in range(steps):
                self.frame()
            ev1.record(self.stream)
            self.barrier()
            ms_total = ev0.elapsed_time(ev1)
            clk = sampler.stop() if sampler else None
            launches = g.ctx.counter(A.COUNTER_KERNEL_LAUNCHES) - l0
            stage_ms = {}
            for s_ in range(A.STAGE_COUNT):
                tot, runs = g.ctx.stage_total_ms(s_)
                if runs:
                    stage_ms[A.STAGE_NAMES[s_]] = tot / steps          # per frame (a stage may run more than once in a frame)
            g.ctx.stage_time_reset(False)
            counters = {"fragments": g.ctx.counter(A.COUNTER_FRAGMENTS), "bricks": g.ctx.counter(A.COUNTER_BRICKS),
                        "occupied": g.ctx.counter(A.COUNTER_OCCUPIED), "cone_samples": g.ctx.counter(A.COUNTER_MARCH_STEPS),
                        "gather_bytes": g.ctx.counter(A.COUNTER_GATHER_BYTES)}
        return {"ms_total": ms_total, "stage_ms": stage_ms, "launches": launches, "counters": counters, "clocks": clk}

    def reduce(self, r, steps):
        """max of the time over the ranks, sums of the per-rank counters, per-stage min/max over the ranks (rank 0's view otherwise)"""
        torch, dist = self.torch, self.dist
        dev = f"cuda:{self.local_rank}"
        t = torch.tensor([r["ms_total"]], dtype=torch.float64, device=dev)
        c = r["counters"]
        cs = torch.tensor([float(c["cone_samples"]), float(r["launches"]), float(c["fragments"]), float(c["bricks"]), float(c["occupied"])], dtype=torch.float64, device=dev)
        stage_ranks, gather = None, None
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(cs, op=dist.ReduceOp.SUM)
            every = [None] * self.world
            dist.all_gather_object(every, (r["stage_ms"], c["gather_bytes"]))
            stage_ranks = {k_: [round(min(e[0].get(k_, 0.0) for e in every), 4), round(max(e[0].get(k_, 0.0) for e in every), 4)] for k_ in r["stage_ms"]}
            rates = [e[1] / (e[0]["exchange"] * 1e-3) / 1e9 for e in every if e[0].get("exchange")]
            gather = {"bytes_per_rank_min_max": [min(e[1] for e in every), max(e[1] for e in every)],
                      "gbs_per_rank_min_max": [round(min(rates), 1), round(max(rates), 1)] if rates else None,
                      "nvlink_reference_gbs": 770.0,
                      "note": "bytes each rank fetches from its peers per frame (F184_COUNTER_GATHER_BYTES) over the device time of its exchange stage "
                              "(level-0 test + peer loads + surface stores); reference: the measured peer copy of B200_PROFILING.md"}
        return {"ms_frame": float(t[0]) / steps, "cone_samples": float(cs[0]), "launches": int(cs[1]), "fragments": int(cs[2]), "bricks": int(cs[3]),
                "occupied": int(cs[4]), "stage_ranks": stage_ranks, "gather": gather}

    def parity_vs_1gpu(self, levels_from=1):
        """Every rank computes the whole frame alone (one context, the whole scene, every row) and compares it with what the sharded
        schedule left on this rank: the texture-side storage the tracer samples (all six directions of every level >= levels_from;
        level 0 travels only when a cone needs it, level 1 only for the bricks this rank's cones sample — their correctness shows in the
        rows) and this rank's image rows — equal bits or a named mismatch."""
        A, torch, dist, g = self.A, self.torch, self.dist, self.g
        one = A.VoxelGI(self.N, self.W, self.H, A.MODE_NORTHSTAR, shadow_res=g.ctx.cfg.shadow_res, device=self.local_rank, flags=A.FLAG_NO_OVERLAP)
        one.upload_scene(self.sc)
        for slot, t in self.pinned.items():
            one.upload_ptr(slot, t.data_ptr(), t.numel())
        with torch.cuda.stream(self.stream):
            self.frame()
        one.voxelize(self.cams["voxel"]); one.inject(self.k); one.build_mips(); one.trace_indirect(self.k)
        g.ctx.sync(); one.sync()
        bad = []
        m, lvl = self.N // 2, 0
        while m >= 1:
            if lvl + 1 >= levels_from:
                for d in range(6):
                    ga, oa = g.ctx.read_array(d, lvl, m), one.read_array(d, lvl, m)
                    if not np.array_equal(ga, oa):
                        bad.append(f"rank {self.rank}: texture array level {lvl + 1} direction {d}")
                        if os.environ.get("F184_PARITY_DEBUG"):
                            idx = np.argwhere((ga != oa).any(-1))
                            print(f"[parity debug] rank {self.rank} level {lvl + 1} dir {d}: {len(idx)} texels differ; first (z,y,x): {idx[:6].tolist()} "
                                  f"got {[ga[tuple(i)].tolist() for i in idx[:3]]} want {[oa[tuple(i)].tolist() for i in idx[:3]]}", file=sys.stderr, flush=True)
            m //= 2; lvl += 1
        mask = g.own_rows_mask()
        a, b = g.ctx.readback(A.SLOT_INDIRECT_OUT)[mask], one.readback(A.SLOT_INDIRECT_OUT)[mask]
        rows_ok = np.array_equal(a.view(np.uint16), b.view(np.uint16))
        one.close()
        every = [None] * self.world
        dist.all_gather_object(every, (bad, rows_ok))
        dist.barrier()
        vol_bad = [x for e in every for x in e[0]]
        return {"volumes": "bit-exact" if not vol_bad else f"MISMATCH ({len(vol_bad)} arrays): " + "; ".join(vol_bad[:3] + vol_bad[-3:]),
                "rows": "bit-exact" if all(e[1] for e in every) else "MISMATCH on rank(s) " + ",".join(str(i) for i, e in enumerate(every) if not e[1]),
                "checked": f"on every rank before the timed region: six-direction texture arrays of levels >= {levels_from} and the rank's own 8-row tile rows of the traced "
                           f"image against a one-GPU frame computed on the same device"}

    def close(self):
        self.g.close()


def c4_block(A, torch, dist, args, rank, world, local_rank, stream, steps):
    """BASELINE configs[3] — the configuration north_star's scaling target (>= 0.7 at 8 GPUs) is quoted on — at this N, beside the
    headline workload: 8 x tiled Sponza (2.1 M triangles) voxelized to 1024^3, traced at 3840 x 2160."""
    import copy
    a4 = copy.copy(args)
    a4.workload, a4.grid = "c4", 1024
    sc, cams, fi, wname = make_workload(a4, rank, world)
    arm = Arm(A, torch, dist, a4.grid, a4.width, a4.height, a4.shadow, sc, cams, fi, rank, world, local_rank, stream, schedule=args.schedule,
              flags=A.FLAG_NO_OVERLAP if args.no_overlap else 0)
    out = {"workload": wname, "n_gpus": world, "steps": steps}
    if world > 1 and not args.no_extras:
        free, _ = torch.cuda.mem_get_info()
        if free > 80e9:         # the one-GPU frame of the check needs a second 1024^3 context (~62 GB) on the same device
            out["parity_vs_1gpu"] = arm.parity_vs_1gpu(levels_from=2)
        else:
            out["parity_vs_1gpu"] = {"skipped": f"{free / 1e9:.0f} GB free on the device, the one-GPU reference frame needs ~62 GB more"}
    r = arm.timed(steps, 3)
    red = arm.reduce(r, steps)
    if rank == 0:
        out.update({"ms_per_frame": round(red["ms_frame"], 4), "stages_ms": {k_: round(v_, 4) for k_, v_ in r["stage_ms"].items()},
                    "stages_ms_min_max_over_ranks": red["stage_ranks"], "gather": red["gather"],
                    "gvoxel_per_s": round(a4.grid ** 3 / (red["ms_frame"] * 1e-3) / 1e9, 1),
                    "counters": {"fragments": red["fragments"], "bricks": red["bricks"], "occupied": red["occupied"], "cone_samples": int(red["cone_samples"])},
                    "note": "efficiency = ms_per_frame(N=1) / (N * ms_per_frame(N)): compare the lines of the scaling run; strong scaling, one frame split over the ranks"})
    arm.close()
    return out


def spec_delta(A, torch, dist, arm, args, fi, stream):
    """How far is the amended cone tracer (nearest mip level, one sample per voxel of the level) from SURVEY.md Appendix B
    as written (mip-linear sampling, half-diameter steps; F184_FLAG_SPEC_APPENDIX_B)?  Same volume, same G-buffer: the image difference,
    and the Appendix-B frame timed the same way as the headline (the frame pipeline, K frames) so its number sits beside `value`."""
    b = Arm(A, torch, dist, arm.N, arm.W, arm.H, arm.g.ctx.cfg.shadow_res, arm.sc, arm.cams, fi, 0, 1, arm.local_rank, stream,
            flags=A.FLAG_SPEC_APPENDIX_B | (A.FLAG_NO_OVERLAP if args.no_overlap else 0))
    steps = max(5, min(args.steps, 20))
    rb = b.timed(steps, 3)
    with torch.cuda.stream(stream):
        arm.frame(); b.frame()
    arm.g.ctx.sync(); b.g.ctx.sync()
    ib = b.g.ctx.readback(A.SLOT_INDIRECT_OUT).astype(np.float32)[..., :3]
    ia = arm.g.ctx.readback(A.SLOT_INDIRECT_OUT).astype(np.float32)[..., :3]
    out = {"rel_l2": float(np.linalg.norm(ia - ib) / max(np.linalg.norm(ib), 1e-30)), "max_abs": float(np.abs(ia - ib).max()),
           "mean_abs": float(np.abs(ia - ib).mean()), "mean_appendix_b": float(ib.mean()), "mean_amended": float(ia.mean()),
           "appendix_b_ms_per_frame": round(rb["ms_total"] / steps, 4), "appendix_b_stages_ms": {k_: round(v_, 4) for k_, v_ in rb["stage_ms"].items()},
           "appendix_b_trace_solo_ms": round(b.g.ctx.stage_ms(A.STAGE_TRACE), 4), "amended_trace_solo_ms": round(arm.g.ctx.stage_ms(A.STAGE_TRACE), 4),
           "appendix_b_cone_samples": b.g.ctx.counter(A.COUNTER_MARCH_STEPS), "amended_cone_samples": arm.g.ctx.counter(A.COUNTER_MARCH_STEPS),
           "note": "traced indirect radiance (rgb of the RGBA16F image, history reset) of the default tracer against the Appendix-B tracer on the same frame.  Both run at the "
                   "texture units' rate for the fetches their definition asks for (Appendix B: 1.35 x the samples, 16 texels per fetch instead of 8).  north_star's 1e-2 tolerance is "
                   "stated against the reference's shaders, which have no cone tracer: this number says how much the amendment changed the renderer, not which of the two is right; "
                   "`value` is the amended tracer, appendix_b_ms_per_frame the same frame with the tracer as surveyed"}
    b.close()
    return out


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from final184_b200 import api as A
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sc, cams, fi, wname = make_workload(args, rank, world)
    W, H, N = args.width, args.height, args.grid
    peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else None
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)") if peaks else (6650.0, "fallback (B200_PROFILING.md)")

    stream = torch.cuda.Stream(device=local_rank)
    arm = Arm(A, torch, dist, N, W, H, args.shadow, sc, cams, fi, rank, world, local_rank, stream, schedule=args.schedule,
              flags=A.FLAG_NO_OVERLAP if args.no_overlap else 0)
    g = arm.g
    parity = arm.parity_vs_1gpu(levels_from=2) if (world > 1 and not args.no_extras) else None

    # ---- timed region 1: device-resident inputs
    r = arm.timed(args.steps, args.warmup, clocks=True)
    red = arm.reduce(r, args.steps)
    stage_ms, counters, clk = r["stage_ms"], r["counters"], r["clocks"]
    comm_ms = stage_ms.get("exchange", 0.0) + stage_ms.get("barrier", 0.0) + stage_ms.get("apply", 0.0) + stage_ms.get("need", 0.0)
    ms_frame = red["ms_frame"]

    # ---- timed region 2: end to end with host buffers.  Every step uploads one frame's G-buffer (depth, normals, material) from
    # pinned host memory and reads one traced image back into pinned host memory; the shadow map travels when the light changes
    # (it is static in this workload: once, before the region).  The caller keeps ONE frame in flight, as a streaming consumer
    # does: it submits frame f (kernels), the inputs of frame f+1 (H2D on the copy stream: f184_upload_image double-buffers its
    # slots) and the read-back of frame f (device-side snapshot + D2H on the read-back stream), then waits for the image of
    # frame f-1 and consumes it.  All K images are on the host when the clock stops.
    out_info = g.ctx.image_info(A.SLOT_INDIRECT_OUT)
    full_d2h = int(out_info.size_bytes)
    out_hosts = [torch.empty(full_d2h, dtype=torch.uint8).pin_memory() for _ in range(2)]
    own_frac = float(g.own_rows_mask().mean()) if arm.rows else 1.0
    h2d = int(sum(int(t.numel()) * own_frac for slot, t in arm.pinned.items() if slot != A.SLOT_SHADOW))
    d2h = int(full_d2h * own_frac)
    with torch.cuda.stream(stream):
        arm.upload_inputs(shadow=True)
        for i in range(3):
            arm.frame(); arm.upload_inputs(); g.ctx.readback_async_ptr(A.SLOT_INDIRECT_OUT, out_hosts[i & 1].data_ptr(), full_d2h, rows=arm.rows)
        g.ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        arm.barrier()
        t_wall0 = time.perf_counter()
        e0.record(stream)
        checksum = 0
        for i in range(args.steps):
            arm.frame()                        # consumes the inputs uploaded one iteration ago
            arm.upload_inputs()                # next frame's G-buffer (H2D, copy stream)
            g.ctx.readback_async_ptr(A.SLOT_INDIRECT_OUT, out_hosts[i & 1].data_ptr(), full_d2h, rows=arm.rows)
            if i:
                g.ctx.readback_wait(1)         # the image of frame i-1 is on the host: the caller consumes it
                checksum += int(out_hosts[(i - 1) & 1][-8])
        g.ctx.readback_wait(0)
        checksum += int(out_hosts[(args.steps - 1) & 1][-8])
        g.ctx.sync()
        e1.record(stream)
        arm.barrier()
        e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t_wall0) * 1e3)
        # PCIe rates of this box (explains e2e: it cannot beat bytes / rate), measured with the same pinned buffers
        pcie = {}
        if rank == 0:
            big = max(arm.pinned.values(), key=lambda t: t.numel())
            slot_big = [s_ for s_, t in arm.pinned.items() if t is big][0]
            g.ctx.sync()
            t0 = time.perf_counter()
            for _ in range(4):
                g.ctx.upload_ptr(slot_big, big.data_ptr(), big.numel())
            g.ctx.image_info(slot_big); g.ctx.sync()
            torch.cuda.synchronize()
            pcie["h2d_gbs"] = 4 * big.numel() / (time.perf_counter() - t0) / 1e9
            t0 = time.perf_counter()
            for i in range(4):
                g.ctx.readback_async_ptr(A.SLOT_INDIRECT_OUT, out_hosts[i & 1].data_ptr(), full_d2h)
            g.ctx.sync()
            pcie["d2h_gbs"] = 4 * full_d2h / (time.perf_counter() - t0) / 1e9
        extra_ms, peaks_live = {}, {}
        if not args.no_extras:
            # secondary passes, reported beside the metric (not part of it): GTAO + its blur, the indirect blur tail, deferred lighting
            for _ in range(3):
                g.ctx.gtao(cams["main"]); g.ctx.blur_indirect(arm.k); g.ctx.lighting_deferred(arm.k)
            g.ctx.sync()
            extra_ms = {"gtao": g.ctx.stage_ms(A.STAGE_GTAO), "blur": g.ctx.stage_ms(A.STAGE_BLUR), "lighting_deferred": g.ctx.stage_ms(A.STAGE_LIGHTING)}
        # the two peaks MEASURED_PEAKS.json does not hold, measured live (rank 0, after the timed regions)
        if rank == 0:
            peaks_live = {"tex_trilinear_per_s": g.ctx.microbench(0), "red_v4_per_s": g.ctx.microbench(1)}
    e2 = torch.tensor([e2e_ms, float(h2d), float(d2h)], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        e2m = e2.clone()
        dist.all_reduce(e2m, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2, op=dist.ReduceOp.SUM)
        e2e_ms = float(e2m[0])
    h2d, d2h = int(e2[1]), int(e2[2])             # whole job: every rank's own G-buffer / image rows
    sd = spec_delta(A, torch, dist, arm, args, fi, stream) if (world == 1 and not args.no_extras and rank == 0) else None
    arm_mode, describe = g.mode, g.describe()
    arm.close()
    c4 = None
    if not args.no_c4 and args.workload == "c3":
        c4 = c4_block(A, torch, dist, args, rank, world, local_rank, stream, max(5, min(args.steps, 20)))

    if rank == 0:
        total_samples = red["cone_samples"]
        P = W * H
        agg = {"fragments": red["fragments"], "bricks": red["bricks"], "occupied": red["occupied"]}
        # roofline of every stage: algorithmic bytes (lower bounds; this rank's share of the whole-job counters) against the HBM peak
        rstages = {}
        for name, ms in stage_ms.items():
            share = dict(agg) if world == 1 else {k_: counters[k_] for k_ in agg}       # rank 0's own launch
            if name == "voxelize" and world > 1:
                share["fragments"] = counters["fragments"]
            b = algorithmic_bytes(name, N, P // world if name == "trace" else P, sc, share)
            if b and ms > 0:
                ach = b / (ms * 1e-3) / 1e9
                rstages[name] = {"ms": round(ms, 4), "algorithmic_bytes": int(b), "achieved": round(ach, 1), "frac": round(ach / hbm_peak, 4), "bound": "hbm"}
        # the cone tracer is bound by the texture pipe, not by HBM (ncu: l1tex__data_pipe_tex_wavefronts ~ 90 %, DRAM < 2 %): its roofline is
        # trilinear fetches / s against the rate f184_microbench measures on this GPU.  <= 3 fetches per cone-sample (1 at level 0, fewer
        # when a direction weight is 0), so `achieved` is an upper bound on the fetch rate by at most the axis-aligned share.
        tex_peak = peaks_live["tex_trilinear_per_s"] / 1e9
        tex_ach = 3.0 * (total_samples / world) / (stage_ms["trace"] * 1e-3) / 1e9
        red_peak = peaks_live["red_v4_per_s"] / 1e9
        other = {"trace_hbm": rstages.get("trace"),
                 "voxelize_red": {"unit": "G red.v4.f32/s", "achieved": 2.0 * counters["fragments"] / (stage_ms["voxelize"] * 1e-3) / 1e9, "peak": red_peak,
                                  "note": "peak: f184_microbench(1), 16-byte vector reductions with the voxelizer's locality (a warp's lanes inside one 8 KB brick, bricks scattered over 1 GiB)"}}
        other["voxelize_red"]["frac"] = other["voxelize_red"]["achieved"] / red_peak if red_peak else None
        if "trace" in rstages:
            rstages["trace"] = {"ms": round(stage_ms["trace"], 4), "bound": "texture", "unit": "G trilinear fetch/s", "achieved": round(tex_ach, 1),
                                "peak": round(tex_peak, 1), "frac": round(tex_ach / tex_peak, 4), "algorithmic_fetches": int(3 * total_samples / world)}
        # under the frame pipeline the stages of three frames share the SMs: a stage's event pair brackets kernels that run beside other
        # frames' kernels, so its elapsed time includes that sharing (the --no-overlap run and the ncu launch list give solo times)
        concurrent = [] if args.no_overlap else [n_ for n_ in rstages]
        for n_ in concurrent:
            rstages[n_]["concurrent"] = "shares the SMs with the other frames in flight (frame pipeline)"
        dom = max(rstages, key=lambda n_: rstages[n_]["ms"])
        traffic = None
        tp = os.path.join(REPO, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(f"{dom}@{N}")
        if dom == "trace":
            roofline = {"bound": "texture", "kernel": "trace (k_trace_n)", "achieved": rstages[dom]["achieved"], "peak": rstages[dom]["peak"], "unit": "G trilinear fetch/s",
                        "frac": rstages[dom]["frac"], "traffic": traffic,
                        "peak_source": "measured live: f184_microbench(0), trilinear RGBA8 3D fetches of an L1-resident volume (148 SM x 4 TEX x 1.965 GHz / 2 = 582 nominal)",
                        "note": "the contract's hbm|tensor does not apply: nothing on this path is a contraction and the tracer moves 32 B/pixel of compulsory HBM traffic "
                                "(other_bounds.trace_hbm); achieved counts 3 fetches per cone-sample (upper bound), inside the timed region where the kernel shares the SMs "
                                "with the next frames' build — its solo launch is in profiles/"}
        else:
            roofline = {"bound": "hbm", "kernel": dom, "achieved": rstages[dom]["achieved"], "peak": hbm_peak, "unit": "GB/s", "frac": rstages[dom]["frac"],
                        "traffic": traffic, "peak_source": peak_src, "note": ""}
        pipe = "" if args.no_overlap else ("; frame pipeline: accumulate of frame f+2, normalise/inject/mips" + ("/barriers/gather" if world > 1 and arm_mode == "slab" else "") +
                                           " of frame f+1 and the cone trace of frame f on three streams, texture-side volume double-buffered")
        out = {"metric": METRIC, "value": ms_frame, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_frame, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
               "dtype": "u8 volumes / i64 overlap tests / f32 shading", "data": "synthetic",
               "config": {"workload": wname, "grid": N, "width": W, "height": H, "shadow": args.shadow, "triangles": sc.n_tris,
                          "parallelism": describe + pipe,
                          "l2": "inputs larger than L2 (texture-side volume chain %.0f MB, accumulators %.0f MB; no flush)" % ((4 * N ** 3 + 24 * sum((N >> l) ** 3 for l in range(1, N.bit_length()))) / 1e6, 32 * N ** 3 / 1e6)},
               "gvoxel_per_s": N ** 3 / (ms_frame * 1e-3) / 1e9,
               "gcone_samples_per_s": total_samples / world / (stage_ms.get("trace", ms_frame) * 1e-3) / 1e9 * world,
               "stages_ms": {k_: round(v_, 4) for k_, v_ in stage_ms.items()}, "stages_concurrent": concurrent, "stages_ms_min_max_over_ranks": red["stage_ranks"],
               "comm_ms": comm_ms, "gather": red["gather"], "secondary_ms": extra_ms,
               "counters": {"fragments": red["fragments"], "bricks": red["bricks"], "occupied": red["occupied"], "cone_samples": int(total_samples)},
               "roofline": roofline, "roofline_stages": rstages, "other_bounds": other,
               "e2e": {"value": e2e_ms / args.steps, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "pcie": {k_: round(v_, 1) for k_, v_ in pcie.items()},
                       "note": "one frame in flight: H2D of frame f+1 and D2H of frame f-1 overlap the kernels of frame f; per step the G-buffer (depth, normals, material) goes up and the "
                               "traced image comes back; the shadow map (16.8 MB) is uploaded when the light changes — once here" +
                               ("" if world == 1 else "; each rank moves only the G-buffer / image rows it traces (f184_upload_image_rows)")},
               "gpu_launches": red["launches"], "clocks": clk}
        if parity is not None:
            out["parity_vs_1gpu"] = parity
        if sd is not None:
            out["spec_delta"] = sd
        if c4 is not None:
            out["c4_scaling"] = c4
        if world == 1 and not args.no_cpu_baseline:
            cpu = CpuPath(args, sc, cams, fi, vchunks=1, tfrac=16, bands=2)
            cpu.step()
            est, wall = [], []
            t0 = time.perf_counter()
            while (time.perf_counter() - t0 < args.cpu_budget_s and len(est) < 8) or not est:
                e, w, _ = cpu.step()
                est.append(e); wall.append(w)
            out["cpu_baseline"] = {"value": float(np.mean(est)), "unit": UNIT, "cores": cpu.cores, "kind": "port",
                                   "sample": cpu.describe() + f"; {len(est)} samples, {np.mean(wall) / 1e3:.1f} s each"}
        if world == 1 and args.workload == "c3" and args.grid == 512 and not args.no_extras:
            # configs[0] in the reference's own contract, beside the headline (GPU stage times always; the CPU leg with the cpu_baseline)
            out["c1_reference_mode"] = c1_reference_mode(sc, cams, local_rank, 0.0 if args.no_cpu_baseline else 10.0)
        emit(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: str):
    """The ONE JSON line goes to the real stdout; everything else any library prints (NCCL's version banner lands on
    stdout) was redirected to stderr at start-up."""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (line + "\n").encode())


def main():
    global _REAL_STDOUT
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        # convenience: `python bench.py --gpus N` relaunches itself under torchrun
        port = 29500 + (os.getpid() % 2000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

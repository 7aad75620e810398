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
                 timed region) against the measured HBM peak; roofline_stages lists every stage the same way
  cpu_baseline   the CPU oracle (oracle/, the "straight C++ transcription") on a bounded sample
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


def algorithmic_bytes(stage, args, sc, counters):
    """Per-launch algorithmic bytes of each stage (DESIGN.md "Roofline").  counters: fragments, bricks, occupied."""
    N, P = args.grid, args.width * args.height
    V, T = len(sc.pos), sc.n_tris
    if stage == "voxelize":      # indexed triangle fetch + two 16-byte reductions per fragment
        return 32 * V + 12 * T + 32 * counters["fragments"]
    if stage == "normalise":     # touched bricks: read+zero 32 B accumulators, write 8 B albedo+normal
        return (32 + 32 + 8) * 512 * counters["bricks"]
    if stage == "inject":        # listed bricks: read 8 B, write 4 B linear + 4 B array
        return (8 + 8) * 512 * counters["bricks"]
    if stage == "mips":
        if counters.get("dense_mips"):
            return mip_chain_bytes(N)
        # sparse path: per listed brick read 2 KB of level 0, write levels 1-3 (6 x (64 + 8 + 1) texels) twice (linear chain +
        # texture array); then the dense tail: read level 3 once, write levels >= 4 twice
        n3 = N // 8
        tail = 6 * 4 * n3 ** 3 + 2 * 6 * 4 * sum((n3 >> l) ** 3 for l in range(1, n3.bit_length()))
        return counters["bricks"] * (2048 + 2 * 6 * 73 * 4) + tail
    if stage == "trace":         # per-pixel fixed I/O (depth 4, normal 8, material 4, history 8, out 8) + one pass over the volume chain
        return 32 * P + 4 * N ** 3 + 4 * 6 * sum((N >> l) ** 3 for l in range(1, N.bit_length()))
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
        from final184_b200 import api as A
        self.A, self.args, self.sc, self.cams = A, args, sc, cams
        if not os.path.exists(ORACLE_SO):
            subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle")], stdout=subprocess.DEVNULL)
        self.lib = A.Library(ORACLE_SO, "f184o_", product=False)
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
        self.cores = os.cpu_count() or 1

    def describe(self):
        return (f"oracle (C++/OpenMP) on {self.cores} threads; per step: voxelize+normalise of triangle chunk k/{self.vchunks} (rotating, x{self.vchunks}), "
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
        out["gpu"] = {"ms_per_frame": round(sum(st.values()), 4), "wall_ms_per_frame": round(wall, 4), "stages_ms": st,
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
        out["cpu_reference"] = {"ms_per_frame": round(sum(st.values()), 1), "stages_ms": st, "cores": os.cpu_count() or 1, "kind": "reference",
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

    from final184_b200.dist import ShardedVoxelGI
    g = ShardedVoxelGI(grid_n=N, width=W, height=H, shadow_res=args.shadow, device=local_rank, rank=rank, nranks=world, scene=sc,
                       voxel_cam=cams["voxel"], mode=args.schedule, flags=A.FLAG_NO_OVERLAP if args.no_overlap else 0)
    stream = torch.cuda.Stream(device=local_rank)
    g.ctx.set_stream(stream.cuda_stream)
    g.connect()
    k = A.trace_constants_c(cams["main"], cams["shadow"], cams["voxel"], W, H, 0, True)
    slots = ((A.SLOT_DEPTH, "depth"), (A.SLOT_NORMALS, "normals"), (A.SLOT_MATERIAL, "material"), (A.SLOT_SHADOW, "shadow"))
    # pinned host copies of the per-frame inputs and of the result (e2e leg)
    pinned = {}
    for slot, key in slots:
        t = torch.from_numpy(np.ascontiguousarray(fi[key]).view(np.uint8).reshape(-1)).pin_memory()
        pinned[slot] = t
    out_info = g.ctx.image_info(A.SLOT_INDIRECT_OUT)
    out_host = torch.empty(out_info.size_bytes, dtype=torch.uint8).pin_memory()
    # a rank of a sharded frame moves only the G-buffer rows it traces (and reads back only those rows of the image); the shadow
    # map is needed whole by every rank
    rows = world > 1
    own_frac = float(g.own_rows_mask().mean()) if rows else 1.0
    h2d = int(sum(int(t.numel()) * (own_frac if slot != A.SLOT_SHADOW else 1.0) for slot, t in pinned.items()))
    d2h = int(out_host.numel() * own_frac)
    full_d2h = int(out_host.numel())

    def upload_inputs():
        for slot, t in pinned.items():
            g.ctx.upload_ptr(slot, t.data_ptr(), t.numel(), rows=rows and slot != A.SLOT_SHADOW)

    def frame():
        g.frame(cams["voxel"], k)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    upload_inputs()
    g.ctx.sync()
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            frame()
        # ---- timed region: device-resident inputs
        g.ctx.stage_time_reset(True)
        l0 = g.ctx.counter(A.COUNTER_KERNEL_LAUNCHES)
        clocks = ClockSampler(local_rank)
        clocks.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        for _ in range(args.steps):
            frame()
        ev1.record(stream)
        barrier()
        ms_total = ev0.elapsed_time(ev1)
        clk = clocks.stop()
        launches = g.ctx.counter(A.COUNTER_KERNEL_LAUNCHES) - l0
        stage_ms = {}
        for s in range(A.STAGE_COUNT):
            tot, runs = g.ctx.stage_total_ms(s)
            if runs:
                stage_ms[A.STAGE_NAMES[s]] = tot / args.steps          # per frame (a stage may run more than once in a frame)
        comm_ms = stage_ms.get("exchange", 0.0) + stage_ms.get("barrier", 0.0)
        g.ctx.stage_time_reset(False)
        counters = {"fragments": g.ctx.counter(A.COUNTER_FRAGMENTS), "bricks": g.ctx.counter(A.COUNTER_BRICKS),
                    "occupied": g.ctx.counter(A.COUNTER_OCCUPIED), "cone_samples": g.ctx.counter(A.COUNTER_MARCH_STEPS)}
        # ---- timed region: end to end with host buffers.  Every step uploads one frame's inputs from pinned host memory and
        # reads one traced image back into pinned host memory.  The caller keeps ONE frame in flight, as a streaming consumer
        # does: it submits frame f (kernels), the inputs of frame f+1 (H2D on the copy stream: f184_upload_image
        # double-buffers its slots) and the read-back of frame f (device-side snapshot + D2H on the read-back stream), then
        # waits for the image of frame f-1 and consumes it.  All K images are on the host when the clock stops.
        out_hosts = [out_host, torch.empty(out_info.size_bytes, dtype=torch.uint8).pin_memory()]
        upload_inputs()
        for i in range(3):
            frame(); upload_inputs(); g.ctx.readback_async_ptr(A.SLOT_INDIRECT_OUT, out_hosts[i & 1].data_ptr(), full_d2h, rows=rows)
        g.ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_wall0 = time.perf_counter()
        e0.record(stream)
        checksum = 0
        for i in range(args.steps):
            frame()                            # consumes the inputs uploaded one iteration ago
            upload_inputs()                    # next frame's inputs (H2D, copy stream)
            g.ctx.readback_async_ptr(A.SLOT_INDIRECT_OUT, out_hosts[i & 1].data_ptr(), full_d2h, rows=rows)
            if i:
                g.ctx.readback_wait(1)         # the image of frame i-1 is on the host: the caller consumes it
                checksum += int(out_hosts[(i - 1) & 1][-8])
        g.ctx.readback_wait(0)
        checksum += int(out_hosts[(args.steps - 1) & 1][-8])
        g.ctx.sync()
        e1.record(stream)
        barrier()
        e2e_ms_dev = e0.elapsed_time(e1)
        e2e_ms_wall = (time.perf_counter() - t_wall0) * 1e3
        e2e_ms = max(e2e_ms_dev, e2e_ms_wall)
        # PCIe rates of this box (explains e2e: it cannot beat bytes / rate), measured with the same pinned buffers
        pcie = {}
        if rank == 0:
            big = max(pinned.values(), key=lambda t: t.numel())
            slot_big = [s_ for s_, t in pinned.items() if t is big][0]
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
        # secondary pass, reported beside the metric (not part of it): GTAO + its blur, and the indirect blur tail
        for _ in range(3):
            g.ctx.gtao(cams["main"]); g.ctx.blur_indirect(k); g.ctx.lighting_deferred(k)
        g.ctx.sync()
        extra_ms = {"gtao": g.ctx.stage_ms(A.STAGE_GTAO), "blur": g.ctx.stage_ms(A.STAGE_BLUR), "lighting_deferred": g.ctx.stage_ms(A.STAGE_LIGHTING)}
        # the two peaks MEASURED_PEAKS.json does not hold, measured live (rank 0, after the timed regions)
        peaks_live = {"tex_trilinear_per_s": g.ctx.microbench(0), "red_v4_per_s": g.ctx.microbench(1)} if rank == 0 else {}

    stage_ranks = None
    if world > 1:
        every = [None] * world
        dist.all_gather_object(every, stage_ms)
        stage_ranks = {k_: [round(min(e.get(k_, 0.0) for e in every), 4), round(max(e.get(k_, 0.0) for e in every), 4)] for k_ in stage_ms}
    t = torch.tensor([ms_total, e2e_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    cs = torch.tensor([float(counters["cone_samples"]), float(launches), float(h2d), float(d2h)], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cs, op=dist.ReduceOp.SUM)
    ms_total, e2e_ms = float(t[0]), float(t[1])
    total_samples, launches_all = float(cs[0]), int(cs[1])
    h2d, d2h = int(cs[2]), int(cs[3])              # whole job: every rank's own G-buffer rows + its copy of the shadow map
    ms_frame = ms_total / args.steps
    if rank == 0:
        # roofline: dominant kernel = the stage with the largest mean device time in the timed region
        rstages = {}
        for name, ms in stage_ms.items():
            b = algorithmic_bytes(name, args, sc, counters)
            if name == "trace" and world > 1:
                b = b  # each rank reads the whole chain; per-rank pixels differ but the volume term dominates
            if b and ms > 0:
                ach = b / (ms * 1e-3) / 1e9
                rstages[name] = {"ms": round(ms, 4), "algorithmic_bytes": int(b), "achieved": round(ach, 1), "frac": round(ach / hbm_peak, 4)}
        # under the frame overlap the event pairs of voxelize / normalise bracket kernels that share the SMs with the previous
        # frame's cone trace: their elapsed times include that sharing and say nothing about the kernel alone, so they are
        # flagged and not candidates for the dominant kernel (the --no-overlap run and the ncu launch list give their solo times)
        concurrent = [] if args.no_overlap else (["voxelize", "normalise"] if world == 1 else (["voxelize"] if g.mode == "slab" else []))
        for n_ in concurrent:
            if n_ in rstages:
                rstages[n_]["concurrent_with"] = "trace of the previous frame"
        dom = max((n_ for n_ in rstages if n_ not in concurrent), key=lambda n_: rstages[n_]["ms"])
        traffic = None
        tp = os.path.join(REPO, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(f"{dom}@{N}")
        roofline = {"bound": "hbm", "kernel": dom, "achieved": rstages[dom]["achieved"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": rstages[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                    "note": "trace is bound by the texture units, not HBM: see tex_rate" if dom == "trace" else ""}
        # the bounds that actually apply to the two non-HBM kernels: texture fetch rate (trace), vector-atomic rate (voxelize)
        tex_fetches = 3.0 * total_samples / max(1, world)          # <= 3 trilinear fetches per cone-sample (fewer when a weight is 0)
        other = {"trace_tex": {"unit": "G trilinear fetch/s", "achieved": tex_fetches / (stage_ms["trace"] * 1e-3) / 1e9,
                               "peak": peaks_live["tex_trilinear_per_s"] / 1e9, "note": "upper bound on achieved: 3 fetches per cone-sample"},
                 "voxelize_red": {"unit": "G red.v4.f32/s", "achieved": 2.0 * counters["fragments"] / (stage_ms["voxelize"] * 1e-3) / 1e9,
                                  "peak": peaks_live["red_v4_per_s"] / 1e9}}
        for o in other.values():
            o["frac"] = o["achieved"] / o["peak"] if o["peak"] else None
        if dom == "trace":
            # the contract's roofline offers hbm|tensor; the cone tracer is bound by neither — carry the bound that applies along
            roofline["actual_bound"] = {"name": "texture pipe (trilinear RGBA8 3D fetches, measured by f184_microbench)", **other["trace_tex"]}
        out = {"metric": METRIC, "value": ms_frame, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_frame, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
               "dtype": "u8 volumes / i64 overlap tests / f32 shading", "data": "synthetic",
               "config": {"workload": wname, "grid": N, "width": W, "height": H, "shadow": args.shadow, "triangles": sc.n_tris,
                          "parallelism": g.describe() + ("" if args.no_overlap else ("; voxelize+normalise of frame f+1 overlap the cone trace of frame f (internal stream)" if world == 1 else "; the accumulation of frame f+1 (peer atomics) overlaps the gather + cone trace of frame f")), "l2": "inputs larger than L2 (volume chain %.0f MB + accumulators; no flush)" % (algorithmic_bytes("trace", args, sc, counters) / 1e6)},
               "gvoxel_per_s": N ** 3 / (ms_frame * 1e-3) / 1e9,
               "gcone_samples_per_s": total_samples / (stage_ms.get("trace", ms_frame) * 1e-3) / 1e9,
               "stages_ms": {k: round(v, 4) for k, v in stage_ms.items()}, "stages_concurrent": concurrent, "stages_ms_min_max_over_ranks": stage_ranks, "comm_ms": comm_ms, "secondary_ms": extra_ms, "counters": counters,
               "roofline": roofline, "roofline_stages": rstages, "other_bounds": other,
               "e2e": {"value": e2e_ms / args.steps, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "pcie": {k_: round(v_, 1) for k_, v_ in pcie.items()},
                       "note": "one frame in flight: H2D of frame f+1 and D2H of frame f-1 overlap the kernels of frame f" + ("" if world == 1 else "; each rank moves only the G-buffer / image rows it traces (f184_upload_image_rows), the shadow map whole")},
               "gpu_launches": launches_all, "clocks": clk}
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
        if world == 1 and args.workload == "c3" and args.grid == 512:
            # configs[0] in the reference's own contract, beside the headline (GPU stage times always; the CPU leg with the cpu_baseline)
            out["c1_reference_mode"] = c1_reference_mode(sc, cams, local_rank, 0.0 if args.no_cpu_baseline else 10.0)
        emit(json.dumps(out))
    g.close()
    if world > 1:
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

#!/usr/bin/env python
"""bench.py -- collision frames per BASELINE.json: triangle-pair tests/s and ms/frame on B200, the reference's CPU
path timed beside it.

A "step" is one collision frame of the hot path over one synthetic scene:
  entries -> broad (sort + sweep) -> pair matrices -> dual OBB-tree traversal (SAT) -> leaf triangle-triangle tests
  -> colliding entity pairs (+ end-of-frame NCCL gather when N > 1).

Default workload = BASELINE.json configs[2], the scene the north-star target is quoted on: a 163-node Sponza-shaped
static set (~289k triangles, Sponza's non-uniform node scale) against 100,000 dynamic 8,448-triangle bodies.
With N GPUs the broad-phase pair list is sharded by entity (strong scaling: the scene is fixed).

  value : tri-pair tests/s, whole job, inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e   : same metric through the public API (CollisionDetection.Reset / add_entries / ExecuteCollisionDetection /
          results) from HOST numpy buffers: pinned staging + H2D of every entry and D2H of the result inside the timer
  roofline / cpu_baseline : see DESIGN.md "Measurement"

`--impl reference` times the reference's own CPU code (oracle/_ref, the unmodified sources compiled in place; the
plain-C port when that library is absent) on a bounded sample of the same workload with all host threads, and the same
sample on ONE thread (the reference is single-threaded) beside it.

Other BASELINE configs, same JSON contract:  --workload c1  the reference's shipped Sponza + one 8,448-triangle sphere (assets in
oracle/_ref/assets), c2  4,096 tori, c4  one 10 M-triangle tree build (metric: triangles/s), c5  256 characters x 20,164 triangles
re-posed + refitted + collided per frame (metric: triangles/s).  --trees reference builds the trees in IMRCD_BUILD_REFERENCE mode:
the reference's own trees bit for bit, so the frame's hit set is the reference's by construction.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "triangle-pair tests/sec (collision frame: broad + mid + narrow)"
BUILD_METRIC = "OBB-tree build throughput (one mesh, build-only path)"
REPOSE_METRIC = "re-posed triangles/sec (triangle recompute + OBB-tree refit + collide per frame)"
C4_DESC = "C4: one synthetic 10 M-triangle mesh (displaced grid 3162 x 1581), GPU Morton build, build only"
C5_DESC = "C5: 256 skinned + morphed characters x 20,164 triangles (64 joints, 4 per vertex, 2 morph targets) re-posed every frame: vertices, triangles, batched refit, collide"
UNIT = "tri-pair tests/s"
ALG_BYTES_PER_TRI_TEST = 72.0      # two 36-B TrianglePosition
FLOP_PER_SAT = 1000.0              # 15-axis SAT + box transform + GetSurface, FMA disabled
ALG_BYTES_PER_SAT = 96.0           # two 48-B boxes


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return dict(hbm_gbs=float(d["hbm_gbs"]), sm_max_mhz=float(d.get("sm_max_mhz", 1965.0)), source="measured (MEASURED_PEAKS.json)")
        except Exception:
            pass
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback (B200_PROFILING.md)")


# --------------------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------------------
C1_ASSETS = os.path.join(ROOT, "oracle", "_ref", "assets")


def make_c1():
    """BASELINE configs[0] on the shipped assets: matrices and poses from tests/golden/c1_sponza.npz (made by the unmodified reference)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "c1_sponza.npz"))
    paths = (os.path.join(C1_ASSETS, "sponzaModel", "Sponza.gltf"), os.path.join(C1_ASSETS, "environmentTest", "environment.gltf"))
    if not all(os.path.exists(p) for p in paths):
        raise SystemExit("bench.py --workload c1: oracle/_ref/assets is absent (python tests/golden/make_golden_c1.py copies it from the reference checkout)")
    return z, paths


def make_workload(name: str, bodies: int):
    from inmyroom_vulkan_b200 import scenes
    if name == "c1":
        z, _ = make_c1()
        mats = np.concatenate([z["node_mat"], z["poses"][1][None]]).astype(np.float32)
        cb = np.zeros(164, np.uint8); cb[163] = 1
        scene = scenes.Scene([], np.concatenate([z["node_mesh"], [163]]).astype(np.uint32), mats, cb, np.arange(164, dtype=np.uint32), name="c1")
        desc = "C1: the reference's shipped Sponza (163 nodes, 263,911 triangles, non-uniform node scale) vs environment.gltf mesh 0 (8,448 triangles, scale 1.5), pose 1 of tests/golden/c1_sponza.npz"
        return scene, desc
    if name == "c3":
        body = scenes.uv_sphere(66, 65)
        scene = scenes.scene_static_vs_bodies(body, bodies, seed=2026, body_scale=(0.2, 0.5))
        desc = (f"C3: 163-node Sponza-shaped static set ({sum(m.n_tri for m in scene.meshes[:-1])} tris, non-uniform node scale) "
                f"vs {bodies} dynamic {body.n_tri}-tri bodies, scale 0.2-0.5, seed 2026")
    elif name == "c2":
        mesh = scenes.torus(100, 50)
        scene = scenes.scene_instances(mesh, bodies, seed=1234, neighbours=8.0)
        desc = f"C2: {bodies} random-pose instances of a {mesh.n_tri}-tri torus, all-pairs broad + narrow, seed 1234"
    else:
        raise SystemExit(f"unknown workload {name}")
    return scene, desc


def sample_scene(scene, name: str, n_bodies: int):
    """Bounded sample of the workload for the CPU legs: all static nodes + the first n_bodies bodies (c3),
    or the first n_bodies instances at the same density (c2)."""
    from inmyroom_vulkan_b200 import scenes
    if name == "c1":
        return scene
    if name == "c3":
        ns = len(scene.meshes) - 1
        keep = np.concatenate([np.arange(ns), ns + np.arange(min(n_bodies, scene.n_entries - ns))])
    else:
        # same spatial density: keep the instances inside a centred sub-cube holding ~n_bodies of them
        t = scene.matrices[:, 12:15]
        L = scene.meta["cube_side"]
        frac = min(1.0, n_bodies / scene.n_entries) ** (1.0 / 3.0)
        keep = np.nonzero((np.abs(t) <= 0.5 * L * frac).all(1))[0]
    return scenes.Scene(scene.meshes, scene.mesh_index[keep], np.ascontiguousarray(scene.matrices[keep]),
                        scene.should_callback[keep], scene.entities[keep], name=scene.name + f"[sample {len(keep)}]")


# --------------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py touches oracle/)
# --------------------------------------------------------------------------------------------------
def cpu_frame(orc, port, scene, trees, threads: int):
    """One frame of the reference's CPU path on `scene`: broad + (mid + narrow per pair)."""
    from oracle import bind
    entry_trees = [trees[m] for m in scene.mesh_index]
    t0 = time.perf_counter()
    if orc.kind == "reference" and scene.n_entries <= 65534:
        pairs, broad_s = orc.broad(scene.matrices, entry_trees, scene.should_callback)    # SweepAndPrune, timed inside the shim
    else:
        pairs, broad_s = port.broad(scene.matrices, entry_trees, scene.should_callback)
    r = bind.frame_pairs(orc, scene.matrices, entry_trees, pairs, threads=threads)
    r["broad_s"] = broad_s
    r["pairs"] = len(pairs)
    r["frame_s"] = broad_s + r["wall_s"]
    r["total_s"] = time.perf_counter() - t0
    return r


def load_cpu_checker():
    """The single doorway from measurement code to oracle/: bench.py's cpu_baseline and --impl reference legs, and the cpu_reference legs of
    scripts/build_bench.py (C4) and scripts/refit_bench.py (C5).  The product (inmyroom_vulkan_b200/) never comes through here."""
    from oracle import bind
    bind.build("port")
    port = bind.PortOracle()
    if os.path.isdir("/root/reference/inMyRoom_vulkan"):
        try:
            bind.build("ref")
        except Exception:
            pass
    orc = bind.RefOracle() if os.path.exists(bind.REF_SO) else port
    return orc, port


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    orc, port = load_cpu_checker()
    threads = host_cores()
    if args.workload in ("c4", "c5"):
        return run_reference_build_arm(args, orc, threads)
    full, desc = make_workload(args.workload, args.bodies)
    scene = sample_scene(full, args.workload, args.ref_sample)
    trees = cpu_trees(orc, scene, args.workload)
    for _ in range(args.warmup):
        cpu_frame(orc, port, scene, trees, threads)
    tests = 0; secs = 0.0; last = None
    for _ in range(args.steps):
        last = cpu_frame(orc, port, scene, trees, threads)
        tests += last["tri_tests"]; secs += last["frame_s"]
    value = tests / secs
    one = cpu_frame(orc, port, sample_scene(full, args.workload, max(args.ref_sample // 8, 500)), trees, 1)      # the reference as it runs: one thread
    sample = (f"{scene.n_entries} entries of the workload (all static nodes + first {args.ref_sample} bodies): "
              f"{last['pairs']} pairs, {last['tri_tests']} tri-pair tests per frame; broad {last['broad_s']*1e3:.1f} ms + "
              f"mid/narrow {last['wall_s']*1e3:.1f} ms wall on {threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "bodies": args.bodies, "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": orc.kind, "sample": sample},
        "single_thread": {"value": one["tri_tests"] / one["frame_s"], "unit": UNIT, "cores": 1,
                          "note": "the same code on one thread (the reference is single-threaded); this line's value uses all host threads over the pair loop, "
                                  "so a ratio against it moves with the box's core count, a ratio against this figure does not"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def cpu_trees(orc, scene, workload):
    """The CPU checker's trees of a workload's meshes (c1: from the shipped assets through the checker's own triangle assembly)."""
    if workload != "c1":
        return [orc.tree_build(m.positions, m.normals, m.vertex_ids) for m in scene.meshes]
    from inmyroom_vulkan_b200.gltf import GltfFile
    _, paths = make_c1()
    trees = []
    for path, meshes in ((paths[0], range(163)), (paths[1], [0])):
        with GltfFile(path) as g:
            for m in meshes:
                ps, ns, vs = [], [], []
                for pts, nrm, idx, mode, _ in g.primitives(m):
                    i = np.arange(len(pts), dtype=np.uint32) if idx is None else idx
                    p_, n_, v_ = orc.triangle_list(pts, nrm, i, mode)
                    ps.append(p_); ns.append(n_); vs.append(v_)
                trees.append(orc.tree_build(np.concatenate(ps), np.concatenate(ns), np.concatenate(vs)))
    return trees


def c4_mesh(n_million: float):
    from inmyroom_vulkan_b200 import scenes
    nx = int(round((n_million * 1e6) ** 0.5)); nz = nx // 2          # 2 * nx * nz = nx^2 triangles
    return scenes.grid_sheet(nx, nz, 1500.0, 750.0, bump=40.0), nx, nz


def run_reference_build_arm(args, orc, threads):
    """--impl reference for the build workloads: OBBtree::OBBtree (c4) / a rebuild per re-posed character (c5: the reference has no refit)
    on bounded samples, one tree per host thread."""
    from concurrent.futures import ThreadPoolExecutor
    from inmyroom_vulkan_b200 import scenes
    if args.workload == "c4":
        sub, _, _ = c4_mesh(0.1)
        meshes = [sub] * threads
        desc = C4_DESC
    else:
        ch = scenes.character()
        meshes = [ch.mesh] * threads
        desc = C5_DESC
    def build(m):
        orc.tree_build(m.positions, m.normals, m.vertex_ids)
    for _ in range(min(args.warmup, 1)):
        build(meshes[0])
    tri = 0; secs = 0.0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(build, meshes))
        secs += time.perf_counter() - t0; tri += sum(m.n_tri for m in meshes)
    t0 = time.perf_counter(); build(meshes[0]); one = meshes[0].n_tri / (time.perf_counter() - t0)
    value = tri / secs
    line = {"impl": "reference", "metric": BUILD_METRIC if args.workload == "c4" else REPOSE_METRIC, "value": value, "unit": "triangles/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 sums, f32 boxes" if args.workload == "c4" else "f32", "data": "synthetic", "config": {"workload": desc},
            "cpu_baseline": {"value": value, "unit": "triangles/s", "cores": threads, "kind": orc.kind,
                             "sample": f"{threads} trees of {meshes[0].n_tri} triangles per step, one per host thread (OBBtree::OBBtree, OBBtree.cpp:321)"},
            "single_thread": {"value": one, "unit": "triangles/s", "cores": 1},
            "e2e": {"value": value, "unit": "triangles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.path = tempfile.mktemp(prefix="imrcd_clocks_", suffix=".csv")
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for k, nm in enumerate(names):
                    if f[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# the GPU arm
# --------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from inmyroom_vulkan_b200.collision import CollisionDetection, Context, OBBtree
    from inmyroom_vulkan_b200 import parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    ctx = Context(local, stream.cuda_stream)

    from inmyroom_vulkan_b200.collision import IMRCD_BUILD_MORTON, IMRCD_BUILD_REFERENCE
    build_mode = IMRCD_BUILD_REFERENCE if args.trees == "reference" else IMRCD_BUILD_MORTON
    scene, desc = make_workload(args.workload, args.bodies)
    t0 = time.perf_counter()
    if args.workload == "c1":
        from inmyroom_vulkan_b200.gltf import load_gltf
        _, paths = make_c1()
        trees = load_gltf(ctx, paths[0], build_mode=build_mode) + load_gltf(ctx, paths[1], build_mode=build_mode)[:1]
        n_tri_total = sum(t.info()[0] for t in trees)
    else:
        trees = [OBBtree(ctx, m.positions, m.normals, m.vertex_ids, build_mode=build_mode) for m in scene.meshes]
        n_tri_total = sum(m.n_tri for m in scene.meshes)
    build_wall = time.perf_counter() - t0
    mesh_ids = np.array([trees[m].mesh_id for m in scene.mesh_index], np.uint32)
    cd = CollisionDetection(ctx=ctx)
    if world > 1:
        parallel.init_comm(ctx, rank, world)     # from here on every frame of this context ends with the library's own NCCL all-gather
    multi = world > 1

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    def l2_flush():
        with torch.cuda.stream(stream):
            flush.zero_()

    def barrier():
        stream.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def global_max(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def global_sum(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident loop: upload once, then K x run (+ gather) ----
    cd.Reset(); cd.add_entries(scene.matrices, mesh_ids, scene.should_callback, scene.entities); cd.upload()

    align = torch.zeros(1, device="cuda") if multi else None
    SPIN_CYCLES = 2_000_000

    def device_step():
        with torch.cuda.stream(stream):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            if multi:
                dist.all_reduce(align)              # not part of the frame: the ranks leave their L2 flushes at different times, and a step timed
                                                    # from an early rank's start would count its wait for the latest one inside the frame's collective
            torch.cuda._sleep(SPIN_CYCLES)          # ~1 ms of GPU spin in front of the timed region: the host gets the frame's graph enqueued while it
                                                    # runs, so that the events time the DEVICE and not how fast a (shared, busy) host gets a launch out
            e0.record(stream)
            cd.run_async()                          # the frame's kernels and, with N ranks, the end-of-frame all-gather right behind them
            e1.record(stream)                       # (both enqueued by the library on this stream); the host does not wait here
            if cd.finish():
                raise SystemExit("bench.py: a frame buffer overflowed inside the timed loop (capacities are settled by the warm-up)")
        return e0, e1

    cd.run()                                              # settle every capacity (frame buffers, gather blocks) before anything is timed
    clocks = ClockSampler(local); clocks.start()          # sampled from the warm-up to the end of the e2e loop
    for _ in range(args.warmup):
        l2_flush(); device_step()
    barrier()
    ev = []; st_acc = {}; launches = 0
    for _ in range(args.steps):
        l2_flush()
        ev.append(device_step())
        st = cd.stats()
        for k in ("ms_broad", "ms_pair_setup", "ms_traverse", "ms_narrow", "ms_reduce", "ms_total"):
            st_acc[k] = st_acc.get(k, 0.0) + st[k]
        launches += st["total_launches"]
    barrier()
    if os.environ.get("IMRCD_BENCH_DEBUG"):               # per-rank stage sums: where a multi-GPU step spends its time
        print(f"[rank {rank}] dev_ms/step {sum(a.elapsed_time(b) for a, b in ev) / args.steps:.3f} stages " +
              " ".join(f"{k}={v / args.steps:.3f}" for k, v in st_acc.items()) + f" pairs={cd.stats()['n_pairs']} hits={cd.stats()['n_hits']}", file=sys.stderr)
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    st = cd.stats()
    dev_ms_max = global_max(dev_ms)
    tests_total = global_sum(float(st["n_tri_tests"]))
    sat_total = global_sum(float(st["n_sat_tests"]))
    pairs_total = global_sum(float(st["n_pairs"]))
    hits_total = global_sum(float(st["n_hits"]))
    coll_total = global_sum(float(st["n_colliding"]))
    if multi and int(coll_total) != int(st["n_merged"]):
        raise SystemExit(f"bench.py: the merged set has {st['n_merged']} records, the ranks found {int(coll_total)}")
    value = tests_total * args.steps / (dev_ms_max * 1e-3)

    # warm-L2 figure (no flush), informational
    barrier()
    evw = [device_step() for _ in range(args.steps)]
    barrier()
    warm_ms = global_max(sum(a.elapsed_time(b) for a, b in evw)) / args.steps

    # ---- end-to-end loops through the public API ----
    # (1) headline: the frame's entries sit in PINNED host memory (the context's mapped staging, where an engine's
    #     AddCollisionDetectionEntry would write them); every step pays commit -> H2D -> kernels -> D2H of the result.
    # (2) secondary: the same from pageable numpy arrays through add_entries (one extra host copy into the staging).
    cd.Reset()
    views = cd.map_entries(scene.n_entries)
    views.current[:] = scene.matrices; views.mesh_ids[:] = mesh_ids; views.should_callback[:] = scene.should_callback; views.entities[:] = scene.entities

    def e2e_step_pinned():
        cd.Reset()
        cd.map_entries(scene.n_entries)             # same staging memory: the entries written above are still there
        cd.commit_entries(scene.n_entries, previous_valid=False)
        cd.ExecuteCollisionDetection()              # H2D + kernels (+ the collective) + D2H of the (merged) colliding pairs, one host wait
        return cd.results(want_hits=False)[0]

    def e2e_step_pageable():
        # the headline: the caller's entries live in ordinary host arrays and are WRITTEN into the library's staging inside the timer
        # (what AddCollisionDetectionEntry's push_back is to the reference); with N ranks each keeps, copies and uploads its share only
        cd.Reset()
        cd.add_entries(scene.matrices, mesh_ids, scene.should_callback, scene.entities, scene.previous)
        cd.ExecuteCollisionDetection()
        return cd.results(want_hits=False)[0]

    def time_e2e(step):
        for _ in range(args.warmup):
            l2_flush(); step()
        tot = 0.0; out = None
        for _ in range(args.steps):
            l2_flush()
            barrier()
            t0 = time.perf_counter()
            out = step()
            stream.synchronize()
            tot += time.perf_counter() - t0
        return global_max(tot), out

    e2e_s_max, res = time_e2e(e2e_step_pageable)
    if len(res) != int(coll_total):
        raise SystemExit(f"bench.py: the end-to-end step returned {len(res)} records, the device-resident frame {int(coll_total)}")
    e2e_pinned_s, _ = (time_e2e(e2e_step_pinned) if not multi else (None, None))
    clk = clocks.stop()
    e2e_value = tests_total * args.steps / e2e_s_max
    n_entries = scene.n_entries
    n_local = int(cd.stats()["n_entries_local"])
    h2d = n_local * (64 + 4 + 4 + 1 + (4 if multi else 0) + (64 if scene.previous is not None else 0))   # this rank's share; previous == current is not re-sent
    # bytes that cross to the host per step: the control block + the (merged) pair records, copied speculatively with a 25 % margin
    d2h = int(len(res) * 1.25 + 64) * 80 + 1280

    # ---- response stage: the same scene with every dynamic body moved since the last frame, so that the deltaVector of
    #      every colliding pair is computed (ShootUncollideRays.cpp:14-93); reported beside the headline, not inside it ----
    ns_static = len(scene.meshes) - 1 if args.workload == "c3" else (163 if args.workload == "c1" else 0)
    prev = scene.matrices.copy()
    prev[ns_static:, 12:15] += (np.random.default_rng(1).normal(size=(scene.n_entries - ns_static, 3)) * 0.02).astype(np.float32)
    cd.Reset(); cd.add_entries(scene.matrices, mesh_ids, scene.should_callback, scene.entities, prev); cd.upload()
    cd.run()                                              # settle the ray buffers
    for _ in range(args.warmup):
        l2_flush(); device_step()
    barrier()
    resp_ms = 0.0; moved_ms = 0.0
    for _ in range(args.steps):
        l2_flush(); device_step()
        stm = cd.stats()
        resp_ms += stm["ms_response"]; moved_ms += stm["ms_total"]; launches_moved = stm["total_launches"]
    barrier()
    response = {"ms_response": resp_ms / args.steps, "ms_frame_with_response": moved_ms / args.steps, "rays_shot": global_sum(float(stm["n_rays_shot"])),
                "responses": global_sum(float(stm["n_responses"])), "rays_per_s": stm["n_rays_shot"] / (resp_ms / args.steps * 1e-3) if resp_ms > 0 else 0.0,
                "scene": "every dynamic body translated by N(0, 0.02) since the previous frame (this rank's shard)"}

    # ---- roofline of the dominant kernel (stage times from CUDA events on the launching stream, this rank) ----
    peaks = measured_peaks()
    ms_trav = st_acc["ms_traverse"] / args.steps; ms_nar = st_acc["ms_narrow"] / args.steps
    fp32_peak_tflops = 148 * 128 * peaks["sm_max_mhz"] * 1e6 / 1e12      # one FP32 op per lane per clock with FMA disabled
    roof_narrow = {"kernel": "k_tritri", "bound": "hbm", "achieved": st["n_tri_tests"] * ALG_BYTES_PER_TRI_TEST / (ms_nar * 1e-3) / 1e9 if ms_nar > 0 else 0.0,
                   "peak": peaks["hbm_gbs"], "unit": "GB/s", "traffic": None, "ms": ms_nar,
                   "algorithmic": f"{ALG_BYTES_PER_TRI_TEST:.0f} B per tri-pair test x {st['n_tri_tests']} tests per launch"}
    roof_trav = {"kernel": "k_traverse", "bound": "fp32", "achieved": st["n_sat_tests"] * FLOP_PER_SAT / (ms_trav * 1e-3) / 1e12 if ms_trav > 0 else 0.0,
                 "peak": fp32_peak_tflops, "unit": "TFLOP/s", "traffic": None, "ms": ms_trav,
                 "algorithmic": f"{FLOP_PER_SAT:.0f} flop per SAT node-pair test x {st['n_sat_tests']} tests per launch (no FMA: 148 SM x 128 lanes x {peaks['sm_max_mhz']:.0f} MHz)",
                 "hbm_view_gbs": st["n_sat_tests"] * ALG_BYTES_PER_SAT / (ms_trav * 1e-3) / 1e9 if ms_trav > 0 else 0.0}
    for r in (roof_narrow, roof_trav):
        r["frac"] = r["achieved"] / r["peak"] if r["peak"] else None
        r["peak_source"] = peaks["source"] if r["bound"] == "hbm" else "derived from sm_max_mhz (" + peaks["source"] + ")"
    dominant, other = (roof_trav, roof_narrow) if ms_trav >= ms_nar else (roof_narrow, roof_trav)
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof) and world == 1:          # the ncu capture is of the whole frame on one GPU; a 1/N shard has no measured figure
        try:
            tr = json.load(open(prof))
            for r in (dominant, other):
                if r["kernel"] in tr:
                    r["traffic"] = tr[r["kernel"]].get("dram_bytes_per_launch")
                    r["traffic_note"] = tr[r["kernel"]].get("note")
        except Exception:
            pass

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "entries": n_entries, "triangles_in_trees": n_tri_total, "bodies": args.bodies,
                   "parallelism": f"frame sharded by entity over {world} GPUs (flagged entries replicated, the others dealt in blocks of 256; each rank uploads, sorts and sweeps its share only); end-of-frame merge inside the library: " + ("every rank's last kernel stores its block into every peer's buffer over NVLink and raises a flag (k_p2p_push / k_p2p_wait_compact), no collective call in a frame" if ctx.comm_transport() == 2 else "one ncclAllGather on the frame's stream") if world > 1 else "1 GPU",
                   "l2": "flushed between timed steps (256 MiB write); inputs (~35 MB) would otherwise stay L2-resident",
                   "timing": "CUDA events on the frame's stream around every step, max over ranks; a 1-ms GPU spin is enqueued in front of the first event so that the frame's graph is already queued when the timed region starts (device time, not host launch latency)",
                   "tree_build": "GPU Morton build (IMRCD_BUILD_MORTON): true PCA boxes, fewer tests for the same answer than the reference's trees (see same_work)" if args.trees == "morton"
                                 else "IMRCD_BUILD_REFERENCE: the reference's own trees bit for bit (OBBtree.cpp:321), so combos, tests and the hit set are the reference's by construction"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s_max / args.steps * 1e3,
                "inputs": "entries in ordinary (pageable) host arrays; the timer covers writing them into the library's pinned staging (add_entries), H2D, kernels"
                          + (", the NCCL merge" if multi else "") + " and D2H of the colliding pairs; bytes are per rank",
                "entries_kept_per_rank": n_local,
                "from_prefilled_pinned_staging": None if e2e_pinned_s is None else {"value": tests_total * args.steps / e2e_pinned_s, "ms_per_step": e2e_pinned_s / args.steps * 1e3}},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": dominant, "roofline_other": other, "response": response,
        "frame": {"pairs": pairs_total, "sat_tests": sat_total, "tri_tests": tests_total, "hits": hits_total, "colliding_pairs": coll_total,
                  "ms_broad": st_acc["ms_broad"] / args.steps, "ms_pair_setup": st_acc["ms_pair_setup"] / args.steps,
                  "ms_traverse": ms_trav, "ms_narrow": ms_nar, "ms_reduce": st_acc["ms_reduce"] / args.steps,
                  "ms_per_step_warm_l2": warm_ms, "sat_tests_per_s": sat_total * args.steps / (dev_ms_max * 1e-3),
                  "tree_build_wall_s": build_wall},
    }

    # ---- CPU baseline: the reference on a bounded sample, 1 thread (rank 0, N=1 only) ----
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        orc, port = load_cpu_checker()
        sample = sample_scene(scene, args.workload, args.cpu_sample)
        ctrees = cpu_trees(orc, sample, args.workload)
        # same-work throughput: the tests the REFERENCE performs on ITS OWN trees for this frame's pairs, over the GPU's frame time.
        # `value` counts the GPU's own tests: with Morton trees that is several times fewer for the same answer, so `value` is the smaller,
        # conservative figure, and it is the one every ratio against the reference's tests/s uses.
        try:
            from oracle import bind
            cd.Reset(); cd.add_entries(scene.matrices, mesh_ids, scene.should_callback, scene.entities); cd.ExecuteCollisionDetection()
            fr = bind.frame_pairs(orc, scene.matrices, [ctrees[m] for m in scene.mesh_index], cd.broad_pairs(), threads=host_cores())
            line["same_work"] = {"reference_tri_tests_per_frame": fr["tri_tests"], "gpu_tri_tests_per_frame": tests_total,
                                 "value": fr["tri_tests"] * args.steps / (dev_ms_max * 1e-3), "e2e_value": fr["tri_tests"] * args.steps / e2e_s_max, "unit": "reference tri-pair tests/s",
                                 "reference_colliding_pairs": fr["colliding"], "gpu_colliding_pairs": coll_total,
                                 "claim": "the >= 100x target is read on `value` / `e2e.value` (the GPU's own, smaller test count), not on this figure"}
        except Exception as e:
            line["same_work"] = {"error": str(e)[:200]}
        r = cpu_frame(orc, port, sample, ctrees, threads=1)
        # response stage on the CPU: the reference's per-pair code with and without movement on the first colliding pairs of the sample
        try:
            et = [ctrees[m] for m in sample.mesh_index]
            cpairs, _ = (orc.broad if (orc.kind != "reference" or sample.n_entries <= 65534) else port.broad)(sample.matrices, et, sample.should_callback)
            rngp = np.random.default_rng(1); t_moved = t_still = 0.0; n_rays_cpu = 0; n_used = 0
            for (i, j) in cpairs.tolist():
                rr = orc.pair(et[i], sample.matrices[i], et[j], sample.matrices[j])
                if not rr.colliding:
                    continue
                pi = sample.matrices[i].copy(); pj = sample.matrices[j].copy(); pj[12:15] += (rngp.normal(size=3) * 0.02).astype(np.float32)
                t0 = time.perf_counter(); orc.pair_delta(et[i], sample.matrices[i], pi, et[j], sample.matrices[j], pj); t1 = time.perf_counter()
                orc.pair_delta(et[i], sample.matrices[i], pi, et[j], sample.matrices[j], sample.matrices[j]); t2 = time.perf_counter()
                t_moved += t1 - t0; t_still += t2 - t1; n_rays_cpu += rr.rays_first + rr.rays_second; n_used += 1
                if n_used >= 150 or t_moved > 8.0:
                    break
            if n_used and t_moved > t_still:
                line["response"]["cpu_reference"] = {"rays_per_s": n_rays_cpu / (t_moved - t_still), "pairs": n_used, "rays": n_rays_cpu, "cores": 1, "kind": orc.kind}
        except Exception as e:          # the response figure is informational
            line["response"]["cpu_reference"] = {"error": str(e)[:200]}
        line["cpu_baseline"] = {"value": r["tri_tests"] / r["frame_s"], "unit": UNIT, "cores": 1, "kind": orc.kind,
                                "sample": (f"{sample.n_entries} entries (all static nodes + first {args.cpu_sample} bodies), {r['pairs']} pairs, "
                                           f"{r['tri_tests']} tri-pair tests; broad {r['broad_s']:.2f} s + mid {r['mid_s']:.2f} s + narrow {r['narrow_s']:.2f} s, "
                                           f"single thread of {host_cores()} host cores")}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        ctx.comm_destroy()                       # the library's communicator goes first (every rank), then torch's
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _single_gpu(args):
    import torch
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        if int(os.environ.get("RANK", "0")) != 0:
            return None, None, None              # these workloads do not shard: replicas only, rank 0 reports
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    from inmyroom_vulkan_b200.collision import Context
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream()
    return torch, stream, Context(0, stream.cuda_stream)


def run_build_arm(args):
    """BASELINE configs[3]: OBB-tree build throughput over one synthetic 10 M-triangle mesh.  A step = one build (Morton keys, sort, radix tree,
    records, the fit of every box) of the whole mesh; value: triangles already in HBM (device time of the build, CUDA events on the launching
    stream); e2e: imrcd_mesh_create from host arrays (H2D of positions, normals and vertex ids inside the timer)."""
    torch, stream, ctx = _single_gpu(args)
    if ctx is None:
        return 0
    from inmyroom_vulkan_b200.collision import OBBtree
    mesh, nx, nz = c4_mesh(args.triangles)
    n = mesh.n_tri
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    clocks = ClockSampler(0); clocks.start()
    dev, wall = [], []
    tree = None
    for r in range(args.warmup + args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()
        stream.synchronize()
        t0 = time.perf_counter()
        tree = OBBtree(ctx, mesh.positions, mesh.normals, mesh.vertex_ids)
        w = time.perf_counter() - t0
        if r >= args.warmup:
            dev.append(tree.build_ms()); wall.append(w)
    clk = clocks.stop()
    ms = float(np.mean(dev)); e2e_s = float(np.mean(wall))
    peaks = measured_peaks()
    achieved = n * 268.0 / (ms * 1e-3) / 1e9
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_build import check_boxes_contain, check_tree_structure       # size-independent properties of the tree just built
    flat = tree.export()
    lo, hi = check_tree_structure(flat, mesh)
    check_boxes_contain(flat, lo, hi, sample=300)
    line = {"metric": BUILD_METRIC, "value": n / (ms * 1e-3), "unit": "triangles/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 sums, f32 projections and boxes", "data": "synthetic",
            "config": {"workload": C4_DESC.replace("3162 x 1581", f"{nx} x {nz}"), "triangles": n, "l2": "flushed between timed steps (256 MiB write); the mesh (360 MB of positions) exceeds L2",
                       "parallelism": "1 GPU (one tree does not shard: replicas only)", "checked": f"structure + containment of 300 sampled boxes ok, {int(flat.nv)} tree vertices"},
            "e2e": {"value": n / e2e_s, "unit": "triangles/s", "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": n * 84, "d2h_bytes_per_step": 68,
                    "inputs": "positions, normals, vertex ids in pageable host arrays; imrcd_mesh_create copies them to HBM, builds, and returns the root box"},
            "gpu_launches": 30 * args.steps, "clocks": clk,
            "roofline": {"kernel": "whole build (k_bounds .. k_fit_boxes)", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                         "traffic": None, "ms": ms, "algorithmic": f"268 B per triangle (SURVEY 8d) x {n} triangles per launch sequence", "peak_source": peaks["source"]}}
    if not args.no_cpu_baseline:
        orc, _ = load_cpu_checker()
        sub, _, _ = c4_mesh(0.2)
        t0 = time.perf_counter(); orc.tree_build(sub.positions, sub.normals, sub.vertex_ids); dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": sub.n_tri / dt, "unit": "triangles/s", "cores": 1, "kind": orc.kind,
                                "sample": f"OBBtree::OBBtree over a {sub.n_tri}-triangle displaced grid ({dt:.2f} s), one thread of {host_cores()}"}
    print(json.dumps(line))
    return 0


def run_repose_arm(args):
    """BASELINE configs[4]: 256 skinned + morphed characters x 20,164 triangles re-posed per frame: triangle recompute + OBB-tree refit then
    collide.  A step = imrcd_meshes_repose (joint products, vertices, triangles) + imrcd_mesh_refit (every box of every character) + one
    collision frame.  value: joint matrices resident; e2e: the frame's joint matrices and morph weights (2 MB) start in host memory and the
    colliding pairs end there."""
    torch, stream, ctx = _single_gpu(args)
    if ctx is None:
        return 0
    from inmyroom_vulkan_b200 import scenes
    from inmyroom_vulkan_b200.collision import CollisionDetection, OBBtree, Skin, last_repose_ms, refit_meshes, repose_meshes
    ch = scenes.character()
    n_char = args.characters
    skin = Skin(ctx, ch.vertices, ch.joints, ch.weights)
    trees = [OBBtree(ctx, ch.mesh.positions, ch.mesh.normals, ch.mesh.vertex_ids) for _ in range(n_char)]
    for t in trees:
        t.bind_skin(skin)
    ground = scenes.grid_sheet(64, 64, 60.0, 60.0, bump=0.5)
    g_tree = OBBtree(ctx, ground.positions, ground.normals, ground.vertex_ids)
    sc = scenes.scene_instances(ch.mesh, n_char, seed=7, neighbours=6.0)
    mats = np.concatenate([sc.matrices, np.eye(4, dtype=np.float32).reshape(1, 16)])
    mesh_ids = np.array([t.mesh_id for t in trees] + [g_tree.mesh_id], np.uint32)
    cb = np.ones(n_char + 1, np.uint8); ents = np.arange(1, n_char + 2, dtype=np.uint32)
    cd = CollisionDetection(ctx=ctx)
    ib = np.stack([ch.inverse_bind] * n_char)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def frame_inputs(f):
        poses = [ch.pose(0.37 * k + 0.21 * f) for k in range(n_char)]
        return np.stack([p_[0] for p_ in poses]), np.stack([p_[1] for p_ in poses])

    inputs = [frame_inputs(f) for f in range(4)]
    n_tri = ch.mesh.n_tri * n_char
    clocks = ClockSampler(0); clocks.start()
    rows = []
    for r in range(args.warmup + args.steps):
        jm, mw = inputs[r % len(inputs)]
        with torch.cuda.stream(stream):
            flush.zero_()
        stream.synchronize()
        t0 = time.perf_counter()
        repose_meshes(ctx, trees, mw, jm, ib)
        refit_ms = refit_meshes(ctx)
        cd.Reset(); cd.add_entries(mats, mesh_ids, cb, ents); cd.ExecuteCollisionDetection()
        res = cd.results(want_hits=False)[0]
        wall = time.perf_counter() - t0
        st = cd.stats()
        if r >= args.warmup:
            rows.append(dict(repose_ms=last_repose_ms(ctx), refit_ms=refit_ms, collide_ms=st["ms_total"], wall_ms=wall * 1e3, colliding=len(res), hits=st["n_hits"], tri_tests=st["n_tri_tests"],
                             ms_broad=st["ms_broad"], ms_traverse=st["ms_traverse"], ms_narrow=st["ms_narrow"], ms_reduce=st["ms_reduce"], launches=st["total_launches"]))
    clk = clocks.stop()
    med = {k: float(np.mean([x[k] for x in rows])) for k in rows[0]}
    dev_ms = med["repose_ms"] + med["refit_ms"] + med["collide_ms"]
    peaks = measured_peaks()
    refit_gbs = n_tri * 72.0 / (med["refit_ms"] * 1e-3) / 1e9
    line = {"metric": REPOSE_METRIC, "value": n_tri / (dev_ms * 1e-3), "unit": "triangles/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (f64 covariance sums)", "data": "synthetic",
            "config": {"workload": C5_DESC.replace("256 ", f"{n_char} "), "triangles": n_tri, "l2": "flushed between timed steps (256 MiB write); the characters' triangles (330 MB) exceed L2",
                       "parallelism": "1 GPU (characters would shard by entity like any frame; not benched here)"},
            "e2e": {"value": n_tri / (med["wall_ms"] * 1e-3), "unit": "triangles/s", "ms_per_step": med["wall_ms"], "h2d_bytes_per_step": int(inputs[0][0].nbytes + inputs[0][1].nbytes + ib.nbytes + mats.nbytes + mesh_ids.nbytes * 2 + cb.nbytes),
                    "d2h_bytes_per_step": int(med["colliding"] * 1.25 + 64) * 80 + 1280,
                    "inputs": "joint matrices, inverse bind matrices, morph weights and the entries in pageable host arrays; the colliding pairs come back to the host"},
            "gpu_launches": int((3 + 9 + med["launches"]) * args.steps), "clocks": clk,
            "frame": {"ms_repose": med["repose_ms"], "ms_refit": med["refit_ms"], "ms_collide": med["collide_ms"], "colliding_pairs": med["colliding"], "hits": med["hits"], "tri_tests": med["tri_tests"],
                      "ms_broad": med["ms_broad"], "ms_traverse": med["ms_traverse"], "ms_narrow": med["ms_narrow"], "ms_reduce": med["ms_reduce"]},
            "roofline": {"kernel": "refit (k_fit_treelets .. k_fit_boxes)", "bound": "hbm", "achieved": refit_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": refit_gbs / peaks["hbm_gbs"],
                         "traffic": None, "ms": med["refit_ms"], "algorithmic": f"72 B per triangle (36 B read + 36 B of node boxes written, SURVEY 8d) x {n_tri} triangles", "peak_source": peaks["source"]}}
    if not args.no_cpu_baseline:
        orc, port = load_cpu_checker()
        k = 6
        jm, mw = inputs[0]
        t0 = time.perf_counter()
        for c in range(k):       # what the reference would have to do per frame: re-pose on the CPU (its GLSL pass restated) and REBUILD the tree (it has no refit)
            v = port.repose(ch.vertices, ch.vertices.shape[1] - 1, ch.joints, ch.weights, mw[c], jm[c], ch.inverse_bind)
            orc.tree_build(np.ascontiguousarray(v[:, :3][ch.mesh.vertex_ids].reshape(-1, 9)), ch.mesh.normals, ch.mesh.vertex_ids)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": ch.mesh.n_tri * k / dt, "unit": "triangles/s", "cores": 1, "kind": orc.kind,
                                "sample": f"{k} characters re-posed (port of the engine's GLSL pass) and their trees REBUILT (OBBtree::OBBtree; the reference has no refit), {dt:.2f} s on one thread; collide not included"}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--trees", default="morton", choices=["morton", "reference"], help="reference: IMRCD_BUILD_REFERENCE, the reference's own trees bit for bit")
    ap.add_argument("--triangles", type=float, default=10.0, help="c4: millions of triangles")
    ap.add_argument("--characters", type=int, default=256, help="c5")
    ap.add_argument("--bodies", type=int, default=None)
    ap.add_argument("--cpu-sample", type=int, default=30000, help="bodies in the cpu_baseline sample (about 10 s single-threaded)")
    ap.add_argument("--ref-sample", type=int, default=16000, help="bodies per step of the --impl reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.bodies is None:
        args.bodies = 100000 if args.workload == "c3" else (4096 if args.workload == "c2" else 0)
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload == "c4":
        return run_build_arm(args)
    if args.workload == "c5":
        return run_repose_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())

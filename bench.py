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
plain-C port when that library is absent) on a bounded sample of the same workload with all host threads.
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
def make_workload(name: str, bodies: int):
    from inmyroom_vulkan_b200 import scenes
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
    full, desc = make_workload(args.workload, args.bodies)
    scene = sample_scene(full, args.workload, args.ref_sample)
    trees = [orc.tree_build(m.positions, m.normals, m.vertex_ids) for m in scene.meshes]
    for _ in range(args.warmup):
        cpu_frame(orc, port, scene, trees, threads)
    tests = 0; secs = 0.0; last = None
    for _ in range(args.steps):
        last = cpu_frame(orc, port, scene, trees, threads)
        tests += last["tri_tests"]; secs += last["frame_s"]
    value = tests / secs
    sample = (f"{scene.n_entries} entries of the workload (all static nodes + first {args.ref_sample} bodies): "
              f"{last['pairs']} pairs, {last['tri_tests']} tri-pair tests per frame; broad {last['broad_s']*1e3:.1f} ms + "
              f"mid/narrow {last['wall_s']*1e3:.1f} ms wall on {threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "bodies": args.bodies, "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": orc.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
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

    scene, desc = make_workload(args.workload, args.bodies)
    t0 = time.perf_counter()
    trees = [OBBtree(ctx, m.positions, m.normals, m.vertex_ids) for m in scene.meshes]
    build_wall = time.perf_counter() - t0
    mesh_ids = np.array([trees[m].mesh_id for m in scene.mesh_index], np.uint32)
    cd = CollisionDetection(ctx=ctx)
    if world > 1:
        parallel.init_comm(ctx, rank, world)     # from here on every frame of this context ends with the library's own NCCL all-gather
    multi = world > 1
    n_tri_total = sum(m.n_tri for m in scene.meshes)

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

    def device_step():
        with torch.cuda.stream(stream):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            if multi:
                dist.all_reduce(align)              # not part of the frame: the ranks leave their L2 flushes at different times, and a step timed
                                                    # from an early rank's start would count its wait for the latest one inside the frame's collective
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
    ns_static = len(scene.meshes) - 1 if args.workload == "c3" else 0
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
                   "parallelism": f"frame sharded by entity over {world} GPUs (flagged entries replicated, the others dealt in blocks of 256; each rank uploads, sorts and sweeps its share only); one end-of-frame ncclAllGather inside the library" if world > 1 else "1 GPU",
                   "l2": "flushed between timed steps (256 MiB write); inputs (~35 MB) would otherwise stay L2-resident",
                   "tree_build": "GPU Morton build (IMRCD_BUILD_MORTON)"},
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
        ctrees = [orc.tree_build(m.positions, m.normals, m.vertex_ids) for m in sample.meshes]
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
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c2"])
    ap.add_argument("--bodies", type=int, default=None)
    ap.add_argument("--cpu-sample", type=int, default=30000, help="bodies in the cpu_baseline sample (about 10 s single-threaded)")
    ap.add_argument("--ref-sample", type=int, default=16000, help="bodies per step of the --impl reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.bodies is None:
        args.bodies = 100000 if args.workload == "c3" else 4096
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())

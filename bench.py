#!/usr/bin/env python
"""bench.py — batched hydroelastic ComputeContactSurfaces on N B200s (one process per GPU).

A "step" is one batched query (all scenes of this rank's shard): poses in -> packed ContactSurface data out.
Headline workload: BASELINE config 3 (ball_plate geometry batched, 16384 independent scenes per GPU, weak scaling;
at N > 1 the strong split -- 16384 scenes over all GPUs -- is measured as well). At N = 1 the other BASELINE configs
(C1 mesh_intersection_benchmark, C2 compliant_mesh_intersection_benchmark, C4 dense clutter, C5 high-resolution
stress) are measured in the same run and reported under "other_configs".

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...     # the CPU port of the reference path on the host cores (never maps
                                           # the product library)

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "contact_surfaces_per_sec"
UNIT = "surfaces/s"
CONFIGS = ["c1", "c2", "c2s", "c3", "c3h", "c4", "c5"]
# Aggregate device-to-host rate the GPU box's host side sustains with N GPUs copying at once (measured, GB/s).
D2H_CEILING_GBS = {1: 56.0, 8: 91.0}
DEFAULT_SCENES = {"c1": 1, "c2": 1, "c2s": 256, "c3": 16384, "c3h": 16384, "c4": 4096, "c5": 1024}
# The reference's own printed numbers for the two configs it can run (single thread, unnamed workstation: context,
# not a target): mesh_intersection_benchmark.cc:79 (args 3/4/0) and compliant_mesh_intersection_benchmark.cc:76 (3/3).
REFERENCE_PRINTED_MS = {"c1": (14.3, "geometry/benchmarking/mesh_intersection_benchmark.cc:79"),
                        "c2": (59.971, "geometry/benchmarking/compliant_mesh_intersection_benchmark.cc:76")}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=CONFIGS)
    ap.add_argument("--scenes", type=int, default=0, help="scenes per GPU (default: the config's)")
    ap.add_argument("--rep", default="", choices=["", "polygon", "triangle"], help="default: the config's")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the bounded CPU sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--soft-soft", default="", choices=["", "volume", "surface"],
                    help="soft-soft intersector (default: volume, the reference's ComputeContactSurfaces path)")
    ap.add_argument("--chunks", type=int, default=0, help="chunks of scenes per query (default: chosen by the library)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip C1 / C2 / C4 / C5 (they run at N = 1 only)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling leg (N > 1 only)")
    return ap.parse_args()


def make_workload(name: str, scenes: int, scene0: int, Mesh=None):
    from drake_b200 import workloads as W
    kw = {} if Mesh is None else {"Mesh": Mesh}
    if name == "c1":
        wl = W.config1(args=[(3, 4, 0)] * scenes, **kw)   # the finest case of the reference's table ...
        wl.normal_filter = False                          # ... run the way its benchmark runs it (no normal filter)
        return wl
    if name == "c2":
        return W.config2(resolutions=(3,), overlaps=(3,) * scenes, **kw)
    if name == "c3":
        return W.config3(scenes, scene0, **kw)
    if name == "c3h":
        return W.config3_half_space(scenes, scene0, **kw)
    if name == "c4":
        return W.config4(scenes, scene0, **kw)
    if name == "c5":
        return W.config5(scenes, scene0, **kw)
    return W.config2_spheres(scenes, scene0, 0.5, **kw)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """Samples SM clocks / throttle reasons during the timed region (NVML in a thread every ~3 ms; the timed
    region of the default run is only ~100 ms, too short for `nvidia-smi -lms`)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.index = index
        self.sm, self.mask = [], 0
        self.max_mhz = None
        self._stop = threading.Event()
        self.thread = None
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _sample(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                break
            time.sleep(0.003)

    def start(self):
        if self.h is None:
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        if self.h is None:
            return self._smi_fallback()
        self._stop.set()
        if self.thread:
            self.thread.join(timeout=2)
        reasons = sorted(n for bit, n in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "samples": len(self.sm), "reasons": reasons, "source": "nvml"}

    def _smi_fallback(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "samples": 1,
                    "reasons": [n for k, n in enumerate(names) if out[2 + k].strip().lower().startswith("active")],
                    "source": "nvidia-smi after the timed region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml and nvidia-smi unavailable"]}


# ------------------------------------------------------------------------------------------ the CPU port ---

def time_cpu_port(config: str, scenes: int, rep: int, budget_s: float, threads: int, obb: bool = True):
    """Times the CPU port of the reference path (reference-shaped BVH traversal + clipping + mesh construction) on a
    bounded sample of the config's scenes. obb: the reference's own hierarchy, Bvh<Obb> (oracle/ho_obb.h), else the
    axis-aligned Bvh<Aabb> of the same shape. Everything here is built by the oracle's own generators
    (oracle/ref_workloads.py): this leg never touches the product library."""
    from oracle import ho, ref_workloads as RW
    mode = ho.MODE_REFERENCE_OBB if obb else ho.MODE_REFERENCE_BVH
    wl = make_workload(config, min(scenes, 4096) if config in ("c4", "c5") else scenes, 0, Mesh=RW.RefMesh)
    sc, order = RW.scene_for(wl)
    poses = wl.poses[:, order, :]
    probe = min(len(poses), max(threads, 8))
    t, st = sc.batch_time(poses[:probe], rep, mode, threads)
    per_scene = max(t / probe, 1e-7)
    n = int(min(len(poses), max(probe, budget_s / per_scene)))
    # When the whole workload takes less than the budget, repeat it (bounded) so the sample is seconds, not a blink.
    passes = int(min(64, max(1, budget_s / max(per_scene * n, 1e-9)))) if n == len(poses) else 1
    t, surfaces, pairs = 0.0, 0, 0
    for _ in range(passes):
        tp, st = sc.batch_time(poses[:n], rep, mode, threads)
        t += tp
        surfaces += st["surfaces"]
        pairs += st["element_pairs"]
    return {"seconds": t, "scenes": n, "passes": passes, "stats": st, "surfaces_per_s": surfaces / t,
            "element_pairs_per_s": pairs / t, "ms_per_scene": 1e3 * t / (n * passes), "workload": wl}


def config_record(wl, scenes: int, rep_name: str, geoms: int):
    """The `config` object: identical keys in both arms (the driver compares them)."""
    return {"workload": wl.name, "description": wl.description, "scenes_per_gpu": scenes, "geometries_per_scene": geoms,
            "representation": rep_name,
            "l2": "256 MB L2 flush between steps; work lists and outputs stream through HBM, the instanced mesh assets "
                  "are L2-resident by design"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scenes = args.scenes or DEFAULT_SCENES[args.config]
    rep_name = args.rep or default_rep(args.config)
    rep = 1 if rep_name == "polygon" else 0
    threads = os.cpu_count() or 1
    # Each step is a bounded sample; K + W steps must finish within a few minutes.
    per_step_budget = min(args.cpu_seconds, 150.0 / max(1, args.steps + args.warmup))
    for _ in range(args.warmup):
        time_cpu_port(args.config, scenes, rep, per_step_budget, threads)
    vals, pairs, secs, r = [], [], [], None
    for _ in range(args.steps):
        r = time_cpu_port(args.config, scenes, rep, per_step_budget, threads)
        vals.append(r["surfaces_per_s"])
        pairs.append(r["element_pairs_per_s"])
        secs.append(r["seconds"])
    wl = r["workload"]
    v = float(np.mean(vals))
    sample = f"{r['passes']} x {r['scenes']} of {scenes} scenes of {wl.name} per step, all host threads"
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config_record(wl, scenes, rep_name, wl.geoms_per_scene),
           "element_pairs_per_sec": float(np.mean(pairs)),
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                            "bvh": "the reference's Bvh<Obb> (PCA + volume line search, median split, <= 1 tet / 3 tris per leaf), "
                                   "scalar BoxesOverlap",
                            "native": "oracle/_build/libhydro_oracle.so only: meshes, fields and BVHs come from the oracle's "
                                      "own generators (oracle/ref_workloads.py)"},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(json.dumps(out))


def default_rep(config: str) -> str:
    return "triangle" if config == "c1" else "polygon"


_REAL_STDOUT = None


def emit(line: str):
    """The ONE JSON line goes to the process's real stdout; everything else any library prints while the bench runs
    (e.g. NCCL's version banner) was sent to stderr by quiet_stdout()."""
    if _REAL_STDOUT is None:
        print(line, flush=True)
    else:
        os.write(_REAL_STDOUT, (line + "\n").encode())


def quiet_stdout():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


# ---------------------------------------------------------------------------------------- the GPU legs ---

class Dist:
    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            # Rank 0 prints ONE JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) off stdout.
            if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
                os.environ["NCCL_DEBUG"] = "WARN"
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_sum(self, max_vals, sum_vals):
        """(max over ranks of each of max_vals, sum over ranks of each of sum_vals)."""
        torch = self.torch
        a = torch.tensor(list(max_vals), dtype=torch.float64, device=self.dev)
        b = torch.tensor(list(sum_vals), dtype=torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(a, op=self.dist.ReduceOp.MAX)
            self.dist.all_reduce(b, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in a.tolist()], [float(x) for x in b.tolist()]

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


class Leg:
    """One workload registered on this rank's GPU: set-up (untimed), the device-timed loop, the end-to-end loops."""

    def __init__(self, D: Dist, config: str, scenes: int, scene0: int, scene0_b: int, rep_name: str,
                 soft_soft: str = "", chunks: int = 0):
        import drake_b200 as db
        from drake_b200 import engine
        torch = D.torch
        self.D, self.config, self.scenes, self.engine = D, config, scenes, engine
        self.rep_name = rep_name
        self.rep = engine.POLYGON if rep_name == "polygon" else engine.TRIANGLE
        self.wl = make_workload(config, scenes, scene0)
        wl_b = make_workload(config, scenes, scene0_b) if config in ("c3", "c3h", "c4", "c5", "c2s") else self.wl
        self.ctx = db.Context(D.local)
        if os.environ.get("DBH_BVH", "") == "host":   # A/B of the two BVH builders (default: the device-built LBVH)
            self.ctx.set_bvh_builder(engine.BVH_HOST_MEDIAN_SPLIT)
        # The soft-soft intersector: "volume" (default, what the reference's ComputeContactSurfaces runs) or "surface"
        # (its surface-BVH + BFS overload, field_intersection.cc:254-375; hydro_bfs.cuh).
        self.soft_soft = soft_soft or os.environ.get("DBH_SOFT_SOFT", "volume")
        if self.soft_soft == "surface":
            self.ctx.set_soft_soft_intersector(engine.INTERSECTOR_SURFACE_BFS)
        self.batch = self.wl.register(self.ctx)
        if chunks:
            self.batch.set_chunks(chunks)
        poses = [np.ascontiguousarray(w.poses.reshape(-1, 12)) for w in (self.wl, wl_b)]
        self.pose_sets = [torch.from_numpy(p).to(D.dev) for p in poses]
        self.pinned_poses = [torch.from_numpy(p).pin_memory() for p in poses]
        self.stats_local = torch.zeros((scenes, 6), dtype=torch.int64, device=D.dev)
        self.stats_all = torch.zeros((D.world * scenes, 6), dtype=torch.int64, device=D.dev) if D.world > 1 else None
        self.stream = torch.cuda.ExternalStream(self.batch.stream(), device=D.dev)
        self.batch.enable_timing(True)
        # Grow the device work lists until the workload fits (untimed).
        for k in (0, 1):
            self.batch.set_poses_device(self.pose_sets[k].data_ptr())
            self.batch.run(self.rep, engine.STRICT, self.stats_local.data_ptr())

    def step(self, k: int, flush, timed_events=None):
        torch, D = self.D.torch, self.D
        with torch.cuda.stream(self.stream):
            flush.zero_()  # L2 flush between steps, outside the timed events
            if timed_events is not None:
                timed_events[0].record(self.stream)
            self.batch.set_poses_device(self.pose_sets[k % 2].data_ptr())
            self.batch.query(self.rep, self.engine.STRICT, self.stats_local.data_ptr())
            if D.world > 1:
                # The only collective of the path: per-scene statistics over NVLink.
                D.dist.all_gather_into_tensor(self.stats_all, self.stats_local)
            if timed_events is not None:
                timed_events[1].record(self.stream)
        self.batch.sync()

    def device_timed(self, steps: int, warmup: int, flush, sampler=None):
        torch, D = self.D.torch, self.D
        for k in range(warmup):
            self.step(k, flush)
        stage_ms, step_stats = {}, []
        events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        D.barrier()
        if sampler is not None:
            sampler.start()
        wall0 = time.perf_counter()
        for k in range(steps):
            self.step(k, flush, events[k])
            step_stats.append(self.batch.stats())
            for name, ms in self.batch.stage_times_ms().items():
                stage_ms[name] = stage_ms.get(name, 0.0) + ms
        D.barrier()
        wall = time.perf_counter() - wall0
        clocks = sampler.stop() if sampler is not None else None
        dev_ms = sum(a.elapsed_time(b) for a, b in events)
        tot = [sum(s[k] for s in step_stats) for k in ("surfaces", "bv_pair_tests", "faces")]
        tot.append(sum(s["tet_tri_pairs"] + s["tet_tet_pairs"] + s["half_space_elements"] for s in step_stats))
        (max_ms,), (surfaces, bvtests, faces, pairs) = D.max_sum([dev_ms], tot)
        sec = max_ms / 1e3
        return {"ms_per_step": max_ms / steps, "value": surfaces / sec, "element_pairs_per_sec": pairs / sec,
                "bv_pair_tests_per_sec": bvtests / sec, "faces_per_sec": faces / sec,
                "surfaces_per_step": surfaces / steps, "element_pairs_per_step": pairs / steps,
                "wall_ms_per_step_incl_flush_and_sync": 1e3 * wall / steps,
                "stage_ms": {k: v / steps for k, v in stage_ms.items()},
                "stats_mean": {k: float(np.mean([s[k] for s in step_stats])) for k in step_stats[0]},
                "launches_per_step": self.batch.last_launch_count, "clocks": clocks}

    # -- end to end through the C-ABI with HOST buffers: H2D poses + D2H of what the variant returns, every step.
    def e2e(self, steps: int, warmup: int, variants=("full", "faces_only", "spatial_forces", "discrete_pairs")):
        engine, torch, D = self.engine, self.D.torch, self.D
        batch, rep = self.batch, self.rep
        sizes = batch.sizes()
        cap = engine.ResultSizes()
        for f, _ in engine.ResultSizes._fields_:
            setattr(cap, f, getattr(sizes, f))
        for f in ("num_surfaces", "num_faces", "num_vertices", "num_polygons", "face_data_ints"):
            setattr(cap, f, int(getattr(sizes, f) * 1.25) + 1024)

        def pinned(shape, dtype):
            n = int(np.prod(shape))
            tt = torch.empty(max(n, 1), dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
            return tt.numpy()[:n].reshape(shape)

        # Two batches over the same assets, alternated: while one batch's results travel D2H (asynchronous copies into
        # pinned buffers on its own stream), the other batch's poses travel H2D and its query runs. Every step still
        # includes the H2D copy of its poses and the D2H read of its complete result.
        batches = [batch, self.wl.register(self.ctx, assets=self.wl.assets)]
        for k in (1, 0):
            batches[1].set_poses_host_ptr(self.pinned_poses[k].data_ptr())
            batches[1].run(rep, engine.STRICT, 0)   # grows the second batch's work lists (untimed)
        G = self.wl.geoms_per_scene
        out = {}
        for variant in variants:
            if variant in ("full", "faces_only"):
                fields = None if variant == "full" else engine.PackedResults.FACE_FIELDS
                hosts = [engine.PackedResults(sizes, allocator=pinned, capacity=cap, fields=fields) for _ in batches]
            elif variant == "spatial_forces":
                compliant = np.array([getattr(self.wl.meshes[m], "vertices_per_element", 0) == 4 or
                                      getattr(self.wl.meshes[m], "compliant", False) for m in self.wl.geom_mesh])
                E = np.tile(np.where(compliant, 3e4, np.inf), self.scenes)
                vel = np.random.default_rng(1234 + D.rank).uniform(-0.1, 0.1, size=(self.scenes * G, 6))
                for b in batches:
                    b.set_contact_properties(E, np.ones(self.scenes * G), np.full(self.scenes * G, 0.5))
                    b.set_body_kinematics(None, vel)
                hosts = [pinned((cap.num_surfaces, 18), np.float64) for _ in batches]
            else:
                hosts = [pinned((cap.num_faces, 9), np.float64) for _ in batches]

            def run(n_steps: int):
                surf, nbytes, pending = 0, 0, None
                for k in range(n_steps):
                    b, h = batches[k % 2], hosts[k % 2]
                    b.set_poses_host_ptr(self.pinned_poses[k % 2].data_ptr())
                    b.query(rep, engine.STRICT, 0)
                    if variant == "spatial_forces":
                        b.compute_spatial_forces(1e-4)
                    elif variant == "discrete_pairs":
                        b.compute_discrete_pairs()
                    if pending is not None:
                        pending.sync()          # the previous step's D2H copies
                    b.sync()                    # this step's query: result sizes are now known
                    if variant in ("full", "faces_only"):
                        r = b.copy_results(h, wait=False)
                        surf += r.num_surfaces
                        nbytes += r.nbytes()
                    elif variant == "spatial_forces":
                        ns = b.copy_spatial_forces(h, wait=False)
                        surf += ns
                        nbytes += ns * 18 * 8
                    else:
                        view = b.copy_discrete_pairs(h, wait=False)
                        surf += int(b.sizes().num_surfaces)
                        nbytes += view.size * 8
                    pending = b
                if pending is not None:
                    pending.sync()
                return surf, nbytes

            run(max(2, min(4, warmup)))
            D.barrier()
            t0 = time.perf_counter()
            surf, nbytes = run(steps)
            D.barrier()
            sec = time.perf_counter() - t0
            (sec_max,), (surf_all,) = D.max_sum([sec], [float(surf)])
            out[variant] = {"value": surf_all / sec_max, "unit": UNIT, "ms_per_step": 1e3 * sec_max / steps,
                            "d2h_bytes_per_step": int(nbytes / steps)}
            del hosts
        return out, int(self.pinned_poses[0].numel() * 8)


E2E_WHAT = {
    "full": "the complete packed ContactSurface data of every surface (vertices, pressures, face data, per-face area / "
            "normal / centroid / pressure / gradients, per-surface totals)",
    "faces_only": "per-face area, normal, centroid, pressure at centroid, grad_eM, grad_eN; per-surface ids, offsets, "
                  "total area, centroid (what MultibodyPlant's discrete update reads, SURVEY 3.6)",
    "spatial_forces": "one spatial force per surface computed on the device (k_wrench = HydroelasticTractionCalculator, "
                      "hydroelastic_traction_calculator.cc:24-83): F_Ac_W, F_Ao_W, F_Bo_W, 18 doubles per surface",
    "discrete_pairs": "the per-face data of DiscreteUpdateManager::AppendDiscreteContactPairsForHydroelasticContact computed "
                      "on the device (k_discrete, discrete_update_manager.cc:896-1036): p_WC, nhat_BA_W, phi0, fn0, stiffness, "
                      "9 doubles per face"}


def ncu_profile(config: str):
    """Per-kernel DRAM bytes / instructions of the committed `ncu --set full` capture of this config at its default size
    (profiles/*_kernels_<config>.json, written by tools/ncu_kernels.py), or None."""
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", f"*_kernels_{config}.json")))
    if not cands:
        return None, None
    try:
        return json.load(open(cands[-1])), os.path.relpath(cands[-1], ROOT)
    except Exception:
        return None, None


def roofline_for(leg: Leg, res: dict, config: str, default_size: bool):
    """The dominant kernel group by event time, its algorithmic bytes, and beside it what the
    committed ncu capture says really bounds it: DRAM bytes moved and issue slots used."""
    st, ms = res["stats_mean"], res["stage_ms"]
    n_poly, n_vert = st["polygons"], st["vertices"] if leg.rep == 1 else st["vertices"] - st["polygons"]
    cand_tri, cand_tet = st["tet_tri_pairs"], st["tet_tet_pairs"]
    leaf_tet = st["leaf_pairs_tet_tet"]
    leaf_tri = st["leaf_pairs"] - leaf_tet
    bv_tet = st["bv_pair_tests_tet_tet"]
    bv_tri = st["bv_pair_tests"] - bv_tet
    surv_tri, surv_tet = st["tet_tri_full_clips"], st["tet_tet_full_clips"]
    share_tri = cand_tri / max(1.0, cand_tri + cand_tet + st["half_space_elements"])
    soup = n_poly * 40.0 + n_vert * 32.0          # staged polygons: header 16 + normal 24 + 32 B per vertex
    out_bytes = n_vert * 32.0 + n_poly * 108.0    # final layout (32 n + 96 / 120, + face_data)
    # Per kernel group: name -> (event ms, algorithmic bytes, kernels)
    groups = {
        "mid_phase_tet_tri": (ms["cull_tet_tri"] + ms["traverse"] * (bv_tri / max(1.0, bv_tri + bv_tet)),
                              bv_tri * 64.0 + leaf_tri * 176.0 + surv_tri * 16.0,
                              "k_small_pairs<tri> (or k_traverse + k_filter<tri>): 64 B per box test, 176 B of fp32 cull "
                              "records per leaf pair, 16 B per survivor"),
        "mid_phase_tet_tet": (ms["cull_tet_tet"] + ms["traverse"] * (bv_tet / max(1.0, bv_tri + bv_tet)),
                              bv_tet * 64.0 + leaf_tet * 256.0 + surv_tet * 16.0,
                              "k_small_pairs<tet> (or k_traverse + k_filter<tet>): 64 B per box test, 256 B per leaf pair, "
                              "16 B per survivor"),
        "exact_clip_tet_tri": (ms["clip_tet_tri"], surv_tri * (16.0 + 128 + 256 + 96) + soup * share_tri,
                               "k_clip<tri>: survivor entry + TetRec + TetAux + TriRec in, staged polygon out"),
        "exact_clip_tet_tet": (ms["clip_tet_tet"], surv_tet * (16.0 + 2 * (128 + 256)) + soup * (1 - share_tri),
                               "k_clip<tet>: survivor entry + 2 x (TetRec + TetAux) in, staged polygon out")
        if leg.soft_soft != "surface" else
        # The surface-seeded intersector: every queued pair reads its entry, two TetRec lines and the second line of
        # two TetAux records, two neighbour records, probes the visited set once per intersected face (<= 8 x 8 B) and
        # writes on average one new queue entry; the staged polygon leaves as in k_clip.
                              (ms["clip_tet_tet"], surv_tet * (16.0 + 2 * (128 + 128) + 2 * 16 + 8 * 8 + 16) + soup * (1 - share_tri),
                               "k_bfs_seed + k_bfs_wave (surface-seeded soft-soft intersector, hydro_bfs_kernels.cuh): queue "
                               "entry + 2 x (TetRec + half a TetAux) + neighbours + visited-set probes in, staged polygon + "
                               "new queue entries out"),
        "pack": (ms["pack"], soup + out_bytes, "k_pack: staged polygon in, final layout out"),
    }
    if st["half_space_elements"]:
        groups["half_space"] = (ms["half_space"], st["half_space_elements"] * 168.0, "k_plane_setup + k_plane<count>")
    dominant = max(groups, key=lambda k: groups[k][0])
    d_ms, d_bytes, d_kernels = groups[dominant]
    peak, peak_kind = measured_peaks()
    achieved = d_bytes / (d_ms * 1e-3) / 1e9 if d_ms > 0 else 0.0
    prof, prof_src = ncu_profile(config) if default_size else (None, None)
    traffic, issue = None, None
    # A capture taken on fewer scenes of the same config (ncu replays every launch ~40 times: the full-size clutter and
    # stress configs are captured on a slice) is scaled by the scene ratio: per-scene work is statistically uniform.
    k = (leg.scenes / prof["scenes"]) if prof and prof.get("scenes") else 1.0
    if prof and dominant in prof.get("groups", {}):
        g = dict(prof["groups"][dominant])
        for key in ("dram_bytes", "warp_instructions"):
            if g.get(key) is not None:
                g[key] = g[key] * k
        traffic = g.get("dram_bytes")
        if g.get("warp_instructions") and d_ms > 0:
            # Issue-slot roofline: warp instructions of the capture / live event time, against 4 schedulers x SMs x clock.
            sm_clock = (res.get("clocks") or {}).get("sm_mhz") or 1965.0
            sms = prof.get("sms", 148)
            peak_issue = 4.0 * sms * sm_clock * 1e6
            rate = g["warp_instructions"] / (d_ms * 1e-3)
            issue = {"warp_instructions_per_launch": g["warp_instructions"], "achieved_ginst_s": rate / 1e9,
                     "peak_ginst_s": peak_issue / 1e9, "frac": rate / peak_issue,
                     "ncu_issue_active_pct": g.get("issue_active_pct"), "ncu_active_lanes": g.get("active_lanes"),
                     "ncu_warps_active_pct": g.get("warps_active_pct")}
    whole_alg = st["algorithmic_bytes"]
    compulsory = None
    if prof and prof.get("compulsory_dram_bytes"):
        compulsory = {"bytes": prof["compulsory_dram_bytes"] * k, "staging_bytes": (prof.get("staging_dram_bytes") or 0) * k,
                      "measured_dram_bytes_whole_query": (prof.get("dram_bytes_whole_query") or 0) * k,
                      "scaled_from_scenes": prof.get("scenes") if k != 1.0 else None,
                      "what": "bytes: what any implementation must move per query (poses in + the packed result out); "
                              "staging_bytes: this pipeline's own intermediate lists, each written once and read once "
                              "(survivor entries, staged polygons, leaf lists of the per-level path); measured: "
                              "dram__bytes_read + write over the query's kernels in the committed ncu capture"}
    return {"bound": "hbm", "kernel": d_kernels, "group": dominant, "achieved": achieved, "peak": peak,
            "peak_source": peak_kind, "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": traffic,
            "traffic_source": prof_src, "algorithmic_bytes_per_launch": d_bytes, "kernel_ms": d_ms,
            "kernel_ms_breakdown": res["stage_ms"],
            "groups": {k: {"ms": v[0], "algorithmic_bytes": v[1], "frac": (v[1] / (v[0] * 1e-3) / 1e9 / peak) if v[0] > 0 else None}
                       for k, v in groups.items()},
            "issue_slots": issue, "compulsory_dram": compulsory,
            "note": "achieved = algorithmic bytes of the group (DESIGN.md 4) / its CUDA-event time. The instanced mesh assets "
                    "are cache-resident, so DRAM traffic is far below the algorithmic bytes: the group is bound by "
                    "instruction issue, which `issue_slots` measures (profiles/README.md)",
            "whole_step_frac": (whole_alg / (res["ms_per_step"] * 1e-3) / 1e9) / peak if peak else None}


def cpu_baseline_record(config: str, scenes: int, rep: int, seconds: float):
    threads = os.cpu_count() or 1
    r = time_cpu_port(config, scenes, rep, seconds, threads)
    r1 = time_cpu_port(config, scenes, rep, min(5.0, seconds), 1)
    ra = time_cpu_port(config, scenes, rep, min(5.0, seconds), threads, obb=False)
    return {"value": r["surfaces_per_s"], "unit": UNIT, "cores": threads, "kind": "port",
            "bvh": "Bvh<Obb> as in the reference (oracle/ho_obb.h)",
            "sample": f"{r['passes']} x {r['scenes']} of {scenes} scenes of {r['workload'].name}, {r['seconds']:.1f} s",
            "element_pairs_per_sec": r["element_pairs_per_s"], "ms_per_scene_all_cores": r["ms_per_scene"],
            "bv_tests_per_scene": r["stats"]["bv_tests"] / max(1, r["scenes"]),
            "element_pairs_per_scene": r["stats"]["element_pairs"] / max(1, r["scenes"]),
            "single_thread": {"value": r1["surfaces_per_s"], "scenes": r1["scenes"], "ms_per_scene": r1["ms_per_scene"],
                              "element_pairs_per_sec": r1["element_pairs_per_s"]},
            "aabb_hierarchy": {"value": ra["surfaces_per_s"], "cores": threads, "ms_per_scene_all_cores": ra["ms_per_scene"],
                               "bv_tests_per_scene": ra["stats"]["bv_tests"] / max(1, ra["scenes"]),
                               "element_pairs_per_scene": ra["stats"]["element_pairs"] / max(1, ra["scenes"]),
                               "what": "the same port over Bvh<Aabb> (looser boxes under rotation)"}}


def other_config_line(D: Dist, config: str, flush, cpu_seconds: float, with_cpu: bool, soft_soft: str = "",
                      chunks: int = 0):
    """One BASELINE config measured the same way as the headline, with fewer steps."""
    scenes = DEFAULT_SCENES[config]
    rep_name = default_rep(config)
    t0 = time.perf_counter()
    leg = Leg(D, config, scenes, 0, scenes, rep_name, soft_soft, chunks)
    steps, warmup = (5, 3) if config in ("c4", "c5") else (20, 5)
    res = leg.device_timed(steps, warmup, flush)
    line = {"workload": leg.wl.name, "description": leg.wl.description, "scenes": scenes, "representation": rep_name,
            "steps": steps, "warmup": warmup, "ms_per_step": res["ms_per_step"], "value": res["value"], "unit": UNIT,
            "element_pairs_per_sec": res["element_pairs_per_sec"], "bv_pair_tests_per_sec": res["bv_pair_tests_per_sec"],
            "faces_per_sec": res["faces_per_sec"], "surfaces_per_step": res["surfaces_per_step"],
            "element_pairs_per_step": res["element_pairs_per_step"], "gpu_launches_per_step": res["launches_per_step"],
            "chunks": leg.batch.num_chunks, "roofline": roofline_for(leg, res, config, leg.soft_soft != "surface"),
            "soft_soft_intersector": leg.soft_soft}
    if leg.soft_soft == "surface":
        st = leg.batch.stats()
        line["bfs_pairs_per_step"], line["bfs_waves_per_step"] = st["bfs_pairs"], st["bfs_waves"]
    sizes = leg.batch.sizes()
    result_bytes = int(sizes.num_faces) * 150 + int(sizes.num_vertices) * 36
    if result_bytes < (3 << 30):
        e2e, h2d = leg.e2e(steps, 2, variants=("full",))
        line["e2e"] = dict(e2e["full"], h2d_bytes_per_step=h2d, what=E2E_WHAT["full"])
    else:
        line["e2e"] = None
        line["e2e_skipped"] = f"the complete result of one query is ~{result_bytes / 2**30:.1f} GiB: not staged on the host here"
    if config in REFERENCE_PRINTED_MS:
        ms, src = REFERENCE_PRINTED_MS[config]
        line["reference_printed"] = {"ms": ms, "source": src,
                                     "note": "the reference's own example output (single thread, unnamed workstation, timed "
                                             "region = its narrow phase only): context, not a measurement of this run"}
    del leg
    if with_cpu:
        line["cpu_baseline"] = cpu_baseline_record(config, scenes, 1 if rep_name == "polygon" else 0, cpu_seconds)
    line["bench_seconds"] = time.perf_counter() - t0
    return line


def main():
    args = parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    D = Dist()
    torch = D.torch
    rep_name = args.rep or default_rep(args.config)
    scenes = args.scenes or DEFAULT_SCENES[args.config]
    # Bind this rank to the CPUs (NUMA node) next to its GPU before any pinned host buffer is allocated: the
    # end-to-end leg moves ~0.6 GB per step per GPU over PCIe, and with N ranks the copies otherwise cross sockets.
    numa = "unbound"
    all_cpus = os.sched_getaffinity(0)
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(D.local)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        numa = f"{len(os.sched_getaffinity(0))} cpus near gpu {D.local}"
    except Exception as ex:  # affinity is an optimisation only
        numa = f"unbound ({type(ex).__name__})"

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=D.dev)  # > 126 MB L2
    # ---- headline: weak scaling, `scenes` per GPU.
    leg = Leg(D, args.config, scenes, D.rank * scenes, (D.world + D.rank) * scenes, rep_name, args.soft_soft, args.chunks)
    res = leg.device_timed(args.steps, args.warmup, flush, ClockSampler(D.local) if D.rank == 0 else None)
    e2e = None
    if not args.no_e2e:
        variants, h2d = leg.e2e(args.steps, args.warmup)
        e2e = dict(variants["full"], h2d_bytes_per_step=h2d, what=E2E_WHAT["full"],
                   pipelining="two batches alternated: D2H of step k overlaps H2D + query of step k+1", host_affinity=numa)
        for k in ("faces_only", "spatial_forces", "discrete_pairs"):
            e2e[k] = dict(variants[k], what=E2E_WHAT[k])
        # How close the complete copy-out runs to what the box's host side can take (plain pinned cudaMemcpyAsync from
        # every GPU at once, nothing of this repository on the path: tools/d2h_ceiling.py, profiles/r2y_d2h_ceiling_8gpu.json).
        ceiling = D2H_CEILING_GBS.get(D.world)
        agg = D.world * e2e["d2h_bytes_per_step"] / (e2e["ms_per_step"] * 1e-3) / 1e9
        e2e["d2h_aggregate_gbs"] = agg
        e2e["d2h_ceiling_gbs"] = ceiling
        e2e["d2h_frac_of_ceiling"] = agg / ceiling if ceiling else None
        if D.world > 1:
            e2e["note"] = ("all ranks copy out concurrently: the box's host side sustains a fixed aggregate D2H rate "
                           "(profiles/README.md, tools/d2h_ceiling.py), so the complete copy-out does not scale with N; "
                           "the device-side consumers (spatial_forces, discrete_pairs) do")
    G = leg.wl.geoms_per_scene
    config = config_record(leg.wl, scenes, rep_name, G)
    roofline = roofline_for(leg, res, args.config, scenes == DEFAULT_SCENES[args.config] and rep_name == default_rep(args.config)) \
        if D.rank == 0 else None
    if leg.soft_soft == "surface":   # an experiment switch: the default line (volume intersector) keeps its keys
        st = leg.batch.stats()
        config["soft_soft_intersector"] = "surface"
        config["chunks"] = leg.batch.num_chunks
        config["bfs_pairs_per_step"], config["bfs_waves_per_step"] = st["bfs_pairs"], st["bfs_waves"]
    del leg

    # ---- strong scaling: the config's scene count split over all GPUs (BASELINE config 3: "16384 independent scenes,
    # sharded 1/2/4/8 GPUs").
    strong = None
    if D.world > 1 and not args.no_strong:
        from drake_b200 import sharding
        total = DEFAULT_SCENES[args.config]
        bounds = sharding.shard_bounds(total, D.world)
        mine = int(bounds[D.rank + 1] - bounds[D.rank])
        sleg = Leg(D, args.config, mine, int(bounds[D.rank]), total + int(bounds[D.rank]), rep_name)
        sres = sleg.device_timed(args.steps, args.warmup, flush)
        strong = {"scenes_total": total, "scenes_per_gpu": mine, "ms_per_step": sres["ms_per_step"], "value": sres["value"],
                  "unit": UNIT, "element_pairs_per_sec": sres["element_pairs_per_sec"],
                  "gpu_launches_per_step": sres["launches_per_step"], "stage_ms": sres["stage_ms"]}
        if not args.no_e2e:
            sv, _ = sleg.e2e(args.steps, args.warmup, variants=("full", "discrete_pairs"))
            strong["e2e"] = {k: v for k, v in sv.items()}
        del sleg

    if D.rank != 0:
        D.close()
        return

    cpu_baseline = None
    other = None
    if D.world == 1:
        os.sched_setaffinity(0, all_cpus)   # the CPU legs get every host core again
        if not args.no_cpu_baseline:
            cpu_baseline = cpu_baseline_record(args.config, scenes, 1 if rep_name == "polygon" else 0, args.cpu_seconds)
        if not args.no_other_configs and args.config == "c3" and not args.scenes:
            other = {}
            for cfg in ("c1", "c2", "c4", "c5"):
                try:
                    other[cfg] = other_config_line(D, cfg, flush, min(6.0, args.cpu_seconds), not args.no_cpu_baseline)
                except Exception as ex:  # a failing side config must not take the headline with it
                    other[cfg] = {"error": f"{type(ex).__name__}: {ex}"}
            # Config 5 again through the surface-seeded soft-soft intersector (the reference's second IntersectFields
            # overload): work proportional to the contact surface instead of the overlap volume. One chunk: its
            # lists are two orders of magnitude shorter, and the waves of 1024 scenes together fill the machine.
            try:
                other["c5_surface_bfs"] = other_config_line(D, "c5", flush, 0.0, False, "surface", 1)
            except Exception as ex:
                other["c5_surface_bfs"] = {"error": f"{type(ex).__name__}: {ex}"}
    out = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": D.world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
           "element_pairs_per_sec": res["element_pairs_per_sec"], "bv_pair_tests_per_sec": res["bv_pair_tests_per_sec"],
           "faces_per_sec": res["faces_per_sec"], "surfaces_per_step": res["surfaces_per_step"],
           "element_pairs_per_step": res["element_pairs_per_step"],
           "wall_ms_per_step_incl_flush_and_sync": res["wall_ms_per_step_incl_flush_and_sync"],
           "e2e": e2e, "gpu_launches": res["launches_per_step"] * args.steps, "clocks": res["clocks"], "roofline": roofline,
           "cpu_baseline": cpu_baseline, "strong_scaling": strong, "other_configs": other}
    emit(json.dumps(out))
    D.close()


if __name__ == "__main__":
    main()

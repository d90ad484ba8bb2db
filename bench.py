#!/usr/bin/env python
"""bench.py — batched hydroelastic ComputeContactSurfaces on N B200s (one process per GPU).

A "step" is one batched query (all scenes of this rank's shard): poses in ->
packed ContactSurface data out. Default workload: BASELINE config 3
(ball_plate geometry batched, 16384 independent scenes per GPU, weak scaling).

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...     # the CPU port of the reference path on the host cores

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=["c2s", "c3", "c3h", "c4", "c5"])
    ap.add_argument("--scenes", type=int, default=0, help="scenes per GPU (default: the config's)")
    ap.add_argument("--rep", default="polygon", choices=["polygon", "triangle"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the bounded CPU sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--forces", action="store_true",
                    help="also time the device-side consumer (k_wrench: one spatial force per surface, 18 doubles out "
                         "per surface). Off by default: parity-green on the B200 but not yet timed there (DESIGN.md 4b)")
    return ap.parse_args()


DEFAULT_SCENES = {"c2s": 256, "c3": 16384, "c3h": 16384, "c4": 4096, "c5": 16}


def make_workload(name: str, scenes: int, scene0: int):
    from drake_b200 import workloads as W
    if name == "c3":
        return W.config3(scenes, scene0)
    if name == "c3h":
        return W.config3_half_space(scenes, scene0)
    if name == "c4":
        return W.config4(scenes, scene0)
    if name == "c5":
        return W.config5(scenes, scene0)
    return W.config2_spheres(scenes, scene0, 0.5)


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


def oracle_scene(workload):
    """Builds the oracle's scene from the SAME arrays the GPU registered (cpu_baseline / reference arm)."""
    from tests import parity
    return parity.oracle_scene_for(workload)


def time_cpu_port(workload, rep: int, budget_s: float, threads: int):
    """Times the CPU port of the reference path (reference-shaped Bvh<Aabb> traversal + clipping +
    mesh construction) on a bounded sample of the workload's scenes."""
    from oracle import ho
    sc, order = oracle_scene(workload)
    poses = workload.poses[:, order, :]
    probe = min(len(poses), max(threads, 8))
    t, st = sc.batch_time(poses[:probe], rep, ho.MODE_REFERENCE_BVH, threads)
    per_scene = max(t / probe, 1e-7)
    n = int(min(len(poses), max(probe, budget_s / per_scene)))
    # When the whole workload takes less than the budget, repeat it (bounded) so the sample is seconds, not a blink.
    passes = int(min(32, max(1, budget_s / max(per_scene * n, 1e-9)))) if n == len(poses) else 1
    t, surfaces, pairs = 0.0, 0, 0
    for _ in range(passes):
        tp, st = sc.batch_time(poses[:n], rep, ho.MODE_REFERENCE_BVH, threads)
        t += tp
        surfaces += st["surfaces"]
        pairs += st["element_pairs"]
    return {"seconds": t, "scenes": n, "passes": passes, "stats": st, "surfaces_per_s": surfaces / t,
            "element_pairs_per_s": pairs / t}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from drake_b200 import engine
    rep = engine.POLYGON if args.rep == "polygon" else engine.TRIANGLE
    scenes = args.scenes or DEFAULT_SCENES[args.config]
    wl = make_workload(args.config, scenes, 0)
    threads = os.cpu_count() or 1
    # Each step is a bounded sample; K + W steps must finish within a few minutes.
    per_step_budget = min(args.cpu_seconds, 150.0 / max(1, args.steps + args.warmup))
    for _ in range(args.warmup):
        time_cpu_port(wl, rep, per_step_budget, threads)
    vals, pairs, secs, nscenes, npasses = [], [], [], 0, 1
    for _ in range(args.steps):
        r = time_cpu_port(wl, rep, per_step_budget, threads)
        vals.append(r["surfaces_per_s"])
        pairs.append(r["element_pairs_per_s"])
        secs.append(r["seconds"])
        nscenes = r["scenes"]
        npasses = r["passes"]
    v = float(np.mean(vals))
    sample = f"{npasses} x {nscenes} of {scenes} scenes of {wl.name} per step, all host threads"
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": wl.name, "scenes_per_gpu": scenes, "representation": args.rep,
                      "description": wl.description},
           "element_pairs_per_sec": float(np.mean(pairs)),
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(json.dumps(out))


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


def main():
    args = parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist

    import drake_b200 as db
    from drake_b200 import engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # Rank 0 prints ONE JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) off stdout.
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    rep = engine.POLYGON if args.rep == "polygon" else engine.TRIANGLE
    scenes = args.scenes or DEFAULT_SCENES[args.config]
    # Bind this rank to the CPUs (NUMA node) next to its GPU before any pinned host buffer is allocated: the
    # end-to-end leg moves ~0.6 GB per step per GPU over PCIe, and with N ranks the copies otherwise cross sockets.
    numa = "unbound"
    all_cpus = os.sched_getaffinity(0)
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        numa = f"{len(os.sched_getaffinity(0))} cpus near gpu {local}"
    except Exception as ex:  # affinity is an optimisation only
        numa = f"unbound ({type(ex).__name__})"

    # ---- set-up (untimed): meshes, assets, BVHs, scene batch, two pose sets resident in HBM.
    wl = make_workload(args.config, scenes, rank * scenes)
    wl_b = make_workload(args.config, scenes, (world + rank) * scenes)
    ctx = db.Context(local)
    batch = wl.register(ctx)
    G = wl.geoms_per_scene
    pose_sets = [torch.from_numpy(np.ascontiguousarray(w.poses.reshape(-1, 12))).to(dev) for w in (wl, wl_b)]
    pinned_poses = [torch.from_numpy(np.ascontiguousarray(w.poses.reshape(-1, 12))).pin_memory() for w in (wl, wl_b)]
    stats_local = torch.zeros((scenes, 6), dtype=torch.int64, device=dev)
    stats_all = torch.zeros((world * scenes, 6), dtype=torch.int64, device=dev) if world > 1 else None
    stream = torch.cuda.ExternalStream(batch.stream(), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    batch.enable_timing(True)

    def step(k: int, timed_events=None):
        with torch.cuda.stream(stream):
            flush.zero_()  # L2 flush between steps, outside the timed events
            if timed_events is not None:
                timed_events[0].record(stream)
            batch.set_poses_device(pose_sets[k % 2].data_ptr())
            batch.query(rep, engine.STRICT, stats_local.data_ptr())
            if world > 1:
                # The only collective of the path: per-scene statistics over NVLink.
                dist.all_gather_into_tensor(stats_all, stats_local)
            if timed_events is not None:
                timed_events[1].record(stream)
        batch.sync()

    # Grow device work lists until the workload fits (untimed).
    batch.set_poses_device(pose_sets[0].data_ptr())
    batch.run(rep, engine.STRICT, stats_local.data_ptr())
    batch.set_poses_device(pose_sets[1].data_ptr())
    batch.run(rep, engine.STRICT, stats_local.data_ptr())
    for k in range(args.warmup):
        step(k)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    kernel_ms = {}
    step_stats = []
    events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    if rank == 0:
        sampler.start()
    wall0 = time.perf_counter()
    for k in range(args.steps):
        step(k, events[k])
        step_stats.append(batch.stats())
        for name, ms in batch.kernel_times_ms().items():
            kernel_ms[name] = kernel_ms.get(name, 0.0) + ms
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = sum(a.elapsed_time(b) for a, b in events)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([sum(s["surfaces"] for s in step_stats), sum(s["tet_tri_pairs"] + s["tet_tet_pairs"] + s["half_space_elements"] for s in step_stats),
                        sum(s["bv_pair_tests"] for s in step_stats), sum(s["faces"] for s in step_stats)],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    max_ms = float(t.item())
    surfaces, pairs, bvtests, faces = [float(x) for x in tot.tolist()]
    sec = max_ms / 1e3
    value = surfaces / sec

    # ---- end-to-end through the C-ABI with HOST buffers (H2D poses, D2H packed surfaces).
    e2e = None
    if not args.no_e2e:
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

        # Two batches over the same assets, alternated: while one batch's packed results travel D2H
        # (asynchronous copies into pinned buffers on its own stream), the other batch's poses travel
        # H2D and its query runs. Every step still includes the H2D copy of its poses and the D2H
        # read of its complete result.
        batches = [batch, wl.register(ctx, assets=wl.assets)]
        hosts = [engine.PackedResults(sizes, allocator=pinned, capacity=cap) for _ in batches]
        batches[1].set_poses_host_ptr(pinned_poses[1].data_ptr())
        batches[1].run(rep, engine.STRICT, 0)   # grows the second batch's work lists (untimed)
        batches[1].set_poses_host_ptr(pinned_poses[0].data_ptr())
        batches[1].run(rep, engine.STRICT, 0)

        def e2e_run(n_steps: int):
            surf, nbytes, pending = 0, 0, None
            for k in range(n_steps):
                b, h = batches[k % 2], hosts[k % 2]
                b.set_poses_host_ptr(pinned_poses[k % 2].data_ptr())
                b.query(rep, engine.STRICT, 0)
                if pending is not None:
                    pending.sync()          # the previous step's D2H copies
                b.sync()                    # this step's query: result sizes are now known
                r = b.copy_results(h, wait=False)
                surf += r.num_surfaces
                nbytes += r.nbytes()
                pending = b
            if pending is not None:
                pending.sync()
            return surf, nbytes

        e2e_run(max(2, min(4, args.warmup)))
        barrier()
        e0 = time.perf_counter()
        e_surf, d2h = e2e_run(args.steps)
        barrier()
        e_sec = time.perf_counter() - e0
        # The same loop copying out only what MultibodyPlant's discrete update reads from a ContactSurface
        # (per-face area / normal / centroid / pressure / gradients, per-surface totals): reported
        # beside the complete copy-out, never instead of it.
        full_hosts = hosts
        hosts = [engine.PackedResults(sizes, allocator=pinned, capacity=cap, fields=engine.PackedResults.FACE_FIELDS)
                 for _ in batches]
        e2e_run(2)
        barrier()
        f0 = time.perf_counter()
        f_surf, f_d2h = e2e_run(args.steps)
        barrier()
        f_sec = time.perf_counter() - f0
        hosts = full_hosts
        forces_e2e = None
        if args.forces:
            # Same loop, but what leaves the GPU is HydroelasticTractionCalculator's output: synthetic
            # contact properties (compliant E = 3e4 Pa, rigid = inf; dissipation 1 s/m; friction 0.5) and body velocities.
            G = wl.geoms_per_scene
            compliant = np.array([getattr(wl.meshes[m], "vertices_per_element", 0) == 4 or getattr(wl.meshes[m], "compliant", False)
                                  for m in wl.geom_mesh])
            E = np.tile(np.where(compliant, 3e4, np.inf), scenes)
            vel = np.random.default_rng(1234 + rank).uniform(-0.1, 0.1, size=(scenes * G, 6))
            for b in batches:
                b.set_contact_properties(E, np.ones(scenes * G), np.full(scenes * G, 0.5))
                b.set_body_kinematics(None, vel)

            def forces_run(n_steps: int):
                surf = 0
                for k in range(n_steps):
                    b = batches[k % 2]
                    b.set_poses_host_ptr(pinned_poses[k % 2].data_ptr())
                    b.query(rep, engine.STRICT, 0)
                    surf += len(b.spatial_forces(1e-4)[0])   # waits, then copies 3 x [num_surfaces, 6] doubles
                return surf

            forces_run(2)
            barrier()
            g0 = time.perf_counter()
            g_surf = forces_run(args.steps)
            barrier()
            g_sec = time.perf_counter() - g0
            forces_e2e = {"value": g_surf / g_sec, "unit": UNIT, "ms_per_step": 1e3 * g_sec / args.steps,
                          "d2h_bytes_per_step": int(g_surf / args.steps) * 18 * 8,
                          "what": "host poses in, one spatial force per surface (F_Ac_W, F_Ao_W, F_Bo_W) out; this rank only"}
        ft = torch.tensor([f_sec, float(f_surf)], dtype=torch.float64, device=dev)
        if world > 1:
            fmax = ft.clone()
            dist.all_reduce(fmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(ft, op=dist.ReduceOp.SUM)
            ft[0] = fmax[0]
        faces_only = {"value": float(ft[1].item()) / float(ft[0].item()), "unit": UNIT,
                      "d2h_bytes_per_step": int(f_d2h / args.steps), "ms_per_step": 1e3 * float(ft[0].item()) / args.steps,
                      "arrays": "per-face area, normal, centroid, pressure at centroid, grad_eM, grad_eN; per-surface ids, "
                                "offsets, total area, centroid"}
        et = torch.tensor([e_sec], dtype=torch.float64, device=dev)
        es = torch.tensor([float(e_surf)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
            dist.all_reduce(es, op=dist.ReduceOp.SUM)
        e2e = {"value": float(es.item()) / float(et.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(pinned_poses[0].numel() * 8), "d2h_bytes_per_step": int(d2h / args.steps),
               "ms_per_step": 1e3 * float(et.item()) / args.steps,
               "pipelining": "two batches alternated: D2H of step k overlaps H2D + query of step k+1",
               "host_affinity": numa, "faces_only": faces_only}
        if forces_e2e is not None:
            e2e["spatial_forces"] = forces_e2e

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (algorithmic bytes, SURVEY.md §8d / DESIGN.md).
    K = args.steps
    avg = {k: v / K for k, v in kernel_ms.items()}
    s_mean = {k: float(np.mean([s[k] for s in step_stats])) for k in step_stats[0]}
    alg = {"traverse": s_mean["bv_pair_tests"] * 64.0,
           "clip_tet_tri": s_mean["tet_tri_pairs"] * 260.0,
           "clip_tet_tet": s_mean["tet_tet_pairs"] * 392.0}
    # Polygon output bytes (32 n + 96 / + 120) are attributed to the clip kernels in proportion
    # to their candidate counts.
    out_bytes = (s_mean["vertices"] if rep == engine.POLYGON else s_mean["vertices"] - s_mean["polygons"]) * 32.0 + \
        s_mean["polygons"] * 108.0
    # Half-space pairs (k_plane count pass, timed in the same slot as the tet-tet clip): 168 B per sliced
    # tetrahedron (the bench workloads have no compliant half space, whose triangles would count 108 B).
    plane_bytes = s_mean["half_space_elements"] * 168.0
    alg["clip_tet_tet"] += plane_bytes
    cand = max(1.0, s_mean["tet_tri_pairs"] + s_mean["tet_tet_pairs"] + s_mean["half_space_elements"])
    alg["clip_tet_tri"] += out_bytes * s_mean["tet_tri_pairs"] / cand
    alg["clip_tet_tet"] += out_bytes * (s_mean["tet_tet_pairs"] + s_mean["half_space_elements"]) / cand
    dominant = max(alg, key=lambda k: avg.get(k, 0.0))
    peak, peak_kind = measured_peaks()
    achieved = alg[dominant] / (avg[dominant] * 1e-3) / 1e9 if avg[dominant] > 0 else 0.0
    # DRAM traffic of the same stage from the committed ncu --set full capture of this workload at this size
    # (profiles/*_traffic_<config>.json, written by tools/ncu_traffic.py); null when no capture matches.
    traffic, traffic_src = None, None
    if scenes == DEFAULT_SCENES[args.config] and args.rep == "polygon":
        import glob
        cands = sorted(glob.glob(os.path.join(ROOT, "profiles", f"*_traffic_{args.config}.json")))
        if cands:
            try:
                tj = json.load(open(cands[-1]))
                traffic = tj["stages"][dominant]["dram_bytes"]
                traffic_src = os.path.relpath(cands[-1], ROOT)
            except Exception:
                traffic = None
    roofline = {"bound": "hbm", "kernel": {"traverse": "k_traverse (all levels of one query)",
                                           "clip_tet_tri": "k_filter<false> + k_clip<false>",
                                           "clip_tet_tet": "k_filter<true> + k_clip<true>" if not plane_bytes else
                                           "k_filter<true> + k_clip<true> + k_plane_setup + k_plane<count>"}[dominant],
                "achieved": achieved, "peak": peak, "peak_source": peak_kind, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": alg[dominant], "kernel_ms": avg[dominant],
                "kernel_ms_breakdown": avg,
                "note": "achieved = SURVEY 8(d) algorithmic bytes of the stage / its CUDA-event time; the instanced mesh "
                        "assets are cache-resident, so DRAM traffic is far below the algorithmic bytes and the stage is "
                        "bound by issue / latency, not HBM (profiles/README.md)",
                "whole_step_frac": (s_mean["algorithmic_bytes"] / (max_ms / K * 1e-3) / 1e9) / peak}

    launches_per_step = batch.last_launch_count  # counted by the library at every kernel launch of a query
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)   # the CPU baseline gets every host core again
        threads = os.cpu_count() or 1
        r = time_cpu_port(wl, rep, args.cpu_seconds, threads)
        r1 = time_cpu_port(wl, rep, min(5.0, args.cpu_seconds), 1)
        cpu_baseline = {"value": r["surfaces_per_s"], "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{r['passes']} x {r['scenes']} of {scenes} scenes of {wl.name}, {r['seconds']:.1f} s",
                        "element_pairs_per_sec": r["element_pairs_per_s"],
                        "single_thread": {"value": r1["surfaces_per_s"], "scenes": r1["scenes"],
                                          "element_pairs_per_sec": r1["element_pairs_per_s"]}}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
           "ms_per_step": max_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": wl.name, "description": wl.description, "scenes_per_gpu": scenes,
                      "geometries_per_scene": G, "representation": args.rep,
                      "l2": "256 MB L2 flush between steps; work lists and outputs stream through HBM, the "
                            "instanced mesh assets are L2-resident by design"},
           "element_pairs_per_sec": pairs / sec, "bv_pair_tests_per_sec": bvtests / sec, "faces_per_sec": faces / sec,
           "surfaces_per_step": surfaces / K, "element_pairs_per_step": pairs / K,
           "wall_ms_per_step_incl_flush_and_sync": 1e3 * wall / K,
           "e2e": e2e, "gpu_launches": launches_per_step * K, "clocks": clocks, "roofline": roofline,
           "cpu_baseline": cpu_baseline}
    emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

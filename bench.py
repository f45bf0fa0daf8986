#!/usr/bin/env python
"""bench.py — FLOAM per-frame odometry throughput on B200 (BASELINE.json metric, configs[1]).

A "step" is ONE FRAME of the hot path (deskew-free configs[1]: featureExtraction + scan-to-map odometry + keyframe map update)
of a synthetic HDL-64-shaped sequence (64 rings x 1875 azimuths, ~118k returns/scan, procedurally generated street scene).
Before the timed region the sequence is pre-rolled (map-seeding frame + 11 frames in which the reference's outer-iteration count
decays 11 -> 2) and W warm-up frames are run; then EXACTLY K frames are timed.

  value     frames/s, scans already resident in HBM, frames enqueued back to back (floam_replay_staged: the pose-independent half of
            frame k+1 overlaps the solve / map update of frame k), CUDA events on the context's streams
  e2e       frames/s through the public C-ABI call a node would make (floam_process_submit / floam_process_wait) with HOST scans in
            pinned memory: H2D upload of every scan and D2H of every pose inside the timed region
  roofline  the kernel class with the largest share of the frame, timed live with CUDA event pairs recorded as nodes of the frame
            graphs (floam_set_kernel_timing), pair overhead calibrated with an empty kernel and subtracted
  cpu_baseline  the reference's own classes (oracle/_ref: the reference sources compiled unmodified against stand-in third-party
            headers; kind "reference") on one host thread over a bounded sample of the same sequence; the oracle port ("port") only when
            that build is not present
  multi_sequence  BASELINE.json configs[4] per GPU: S independent sequences replayed concurrently on ONE GPU (one context + one host
            thread each); one sequence alone leaves the device front-end about half idle

N > 1 (torchrun): every rank replays its own independent sequence(s) on its own GPU (replicas, no collective on the data path);
value = total frames / max-over-ranks time.   --impl reference: the reference classes on the host cores (3 pipelined threads per
sequence like the reference's three ROS nodes), rank 0 only.
"""
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

METRIC = "odometry frames/sec at HDL-64 scan"
UNIT = "frames/s"
PREROLL = 12            # frame 0 seeds the map (optimization_count = 12); 11 updates bring the outer count down to 2
SEED_BASE = 0           # sequence s uses generator seed SEED_BASE + s
SENSOR = "hdl64"
ODOM = dict(min_distance=2.0, max_distance=60.0, map_resolution=0.4, loss="cauchy")


def shard_sequences(n_sequences, rank, world):
    """Sequence ids replayed by `rank`: round-robin, no data exchanged between ranks."""
    return [s for s in range(n_sequences) if s % world == rank]


def reduce_over_ranks(frames, seconds, device="cuda"):
    """(sum of frames, max of seconds) over ranks; identity when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return frames, seconds
    f = torch.tensor([float(frames)], dtype=torch.float64, device=device)
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    dist.all_reduce(f, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(round(f.item())), t.item()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic_bytes(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the newest committed `ncu --set full` capture
    (profiles/*_ncu_full_top_kernels.csv), or None when the kernel was not captured."""
    import csv
    import glob
    found = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_full_top_kernels.csv")))
    if not found:
        return None
    path = found[-1]
    try:
        with open(path) as f:
            rows = list(csv.reader(f))
        hdr, units = rows[0], rows[1]
        ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        vals = [float(r[ir].replace(",", "")) * scale.get(units[ir], 1.0) + float(r[iw].replace(",", "")) * scale.get(units[iw], 1.0)
                for r in rows[2:] if r[ik].replace("_kernel", "") == kernel]
        return float(np.mean(vals)) if vals else None
    except Exception:
        return None


def algorithmic_bytes(kernel, st):
    """Compulsory bytes one launch of `kernel` moves (DESIGN.md 'Kernels and rooflines'; SURVEY.md 8d per-unit figures x the units
    of the frame). st: per-frame averages — N scan points, F features, Q downsampled queries, C correspondences, M map points."""
    N, F, Q, C, M = st["N"], st["F"], st["Q"], st["C"], st["M"]
    table = {
        "ring_count": 32 * N,                         # reads the scan once
        "ring_scatter": 32 * N + 36 * N,              # reads the scan, writes the gated ring-bucketed copy + source index
        "sector": 16 * N + 4 * F,                     # xyz of every ring point once, one id per classified point
        "feature_gather": 68 * F,                     # 32 B in + 32 B out + 4 B source index per feature
        "assoc_eval": 16 * Q + 20 * Q + 80 * Q + 80 * C,   # queries, 5 ids in, 5 neighbours gathered, fit parameters out
        "lm_cluster": 4 * 88 * C,                     # <= 4 step attempts x (point 24 B + fit parameters <= 64 B) per correspondence
        "assoc_knn": 16 * Q + 16 * M + 40 * Q,        # queries, map cells touched once, 5 ids + 5 distances out
        "radix_scatter": 16 * st["sort_n"],           # 8 B key/value in, 8 B out per element and pass
        "radix_hist": 4 * st["sort_n"],
        "voxel_reduce": 24 * st["sort_n"] + 16 * st["sort_out"],
        "voxel_keys": 16 * st["sort_n"] + 8 * st["sort_n"],
        "voxel_bbox": 16 * st["sort_n"],
        "voxel_rank": 4 * st["sort_n"] + 4 * st["sort_out"],   # sorted keys in, run starts out
        "scan_add": 8 * M, "unpack_pc2": 22 * N + 32 * N,
        "grid_scatter": 32 * M, "grid_count": 16 * M, "grid_bbox": 16 * M,
        "crop_flags": 16 * M, "crop_scatter": 32 * M,
    }
    return table.get(kernel)


class StdoutToStderr:
    """The reference's classes print to std::cout (e.g. "Use loss function: ...", src/odomEstimationClass.cpp:24); the bench's stdout
    carries exactly one JSON line, so C-level stdout is pointed at stderr while they run."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.saved)


def cpu_backend():
    """(module, kind): oracle/_ref (the reference's own class sources, compiled unmodified) when it is here, else the oracle port."""
    try:
        from oracle import pyref
        if pyref.available():
            return pyref, "reference"
    except Exception:
        pass
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle, "port"


def build_sequence(rank_seed, frames):
    from floam_b200 import synth
    seq = synth.Sequence(SENSOR, seed=SEED_BASE + rank_seed)
    scans, off = seq.scans(0, frames)
    return seq, scans, off


def run_reference(args, rank):
    """--impl reference: the oracle port of the reference classes on the host cores, pipelined over three threads like the
    reference's three ROS nodes (laserProcessing -> odomEstimation -> laserMapping). Bounded sample; rank 0 only."""
    if rank != 0:
        return 0
    import queue
    from floam_b200 import synth
    po, kind = cpu_backend()
    K = max(1, min(args.steps, 120)); W = max(0, min(args.warmup, 10))
    frames = PREROLL + W + K
    S = max(1, args.gpus)   # the N-GPU workload is N independent sequences (configs[4]): the host runs as many pipelines side by side
    spans = [None] * S

    def pipeline(si):
        seq, scans, off = build_sequence(si, frames)
        nl = seq.num_lines
        q1, q2 = queue.Queue(maxsize=4), queue.Queue(maxsize=4)
        stamps = {}

        filtered = {}

        def node_features():
            for f in range(frames):
                e, s, _, _, _ = po.feature_extract(scans[off[f]:off[f + 1]], nl, ODOM["min_distance"], ODOM["max_distance"])
                filtered[f] = np.concatenate([e, s])[::4]   # /velodyne_points_filtered = edge + surf (src/laserProcessingNode.cpp:139-145), every 4th point
                q1.put((f, e, s))
            q1.put(None)

        def node_odom():
            od = po.Odom(num_lines=nl, map_resolution=ODOM["map_resolution"], loss=ODOM["loss"])
            while True:
                item = q1.get()
                if item is None:
                    break
                f, e, s = item
                if f == 0:
                    od.init_map(synth.to_xyzi(e), synth.to_xyzi(s)); T = np.eye(4)
                else:
                    od.update(e, s, False); T = od.get()[0]
                stamps[f] = time.perf_counter()
                q2.put((f, T))
            q2.put(None)

        def node_mapping():
            mp = po.Mapping(map_resolution=ODOM["map_resolution"])
            while True:
                item = q2.get()
                if item is None:
                    break
                f, T = item
                mp.update(synth.to_xyzi(filtered.pop(f)), T)   # a quarter of the filtered cloud keeps the mapping node off the critical path

        th = [threading.Thread(target=t) for t in (node_features, node_odom, node_mapping)]
        [t.start() for t in th]
        [t.join() for t in th]
        spans[si] = (stamps[PREROLL + W - 1], stamps[frames - 1])

    pipes = [threading.Thread(target=pipeline, args=(si,)) for si in range(S)]
    with StdoutToStderr():
        [t.start() for t in pipes]
        [t.join() for t in pipes]
    t0 = min(sp[0] for sp in spans); t1 = max(sp[1] for sp in spans)
    fps = S * K / (t1 - t0)
    cores = min(3 * S, os.cpu_count() or 1)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * S / fps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(),
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": "frames %d..%d of %d sequence(s); %s; 3 pipelined host threads per sequence like the reference's three nodes "
                                       "(featureExtraction | odometry | LaserMapping, the last fed every 4th point so that it never becomes the "
                                       "bottleneck: this favours the CPU arm); %d host cores" % (
                                           PREROLL + W, frames - 1, S,
                                           "the reference's own class sources compiled unmodified (oracle/_ref; PCL / Eigen / Ceres internals are the "
                                           "restated stand-ins)" if kind == "reference" else "oracle port of the reference classes", os.cpu_count() or 1)},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def workload_config():
    return {"workload": "configs[1]: synthetic HDL-64E KITTI-shaped sequence (64x1875 rays, ~118k returns/scan), full odometry, deskew off",
            "sensor": SENSOR, "map_resolution": ODOM["map_resolution"], "loss": ODOM["loss"], "min_dis": ODOM["min_distance"],
            "max_dis": ODOM["max_distance"], "preroll_frames": PREROLL, "parallelism": "replicas",
            "cache": "every frame is a new 3.8 MB scan; the working set (scan + local maps) is L2-resident by nature of the path"}


def cpu_baseline(scans, off, num_lines, n_frames):
    """The reference's classes on ONE host thread (they are single-threaded), bounded sample: PREROLL + n_frames frames.  kind
    "reference" = oracle/_ref (reference sources compiled unmodified); the per-stage split comes from the oracle port, whose stages can
    be timed from inside (the reference class offers no hooks)."""
    po, kind = cpu_backend()
    from oracle import pyoracle as port
    port.build()
    last = min(len(off) - 1, PREROLL + n_frames)
    kw = dict(min_dis=ODOM["min_distance"], max_dis=ODOM["max_distance"], map_resolution=ODOM["map_resolution"], loss=ODOM["loss"], deskew=False)
    with StdoutToStderr():
        sec, poses, ms, q = po.replay_sequence(scans[:off[last]], off[:last + 1], num_lines, **kw)
    steady = ms[PREROLL:]
    fps = 1e3 / float(np.mean(steady)) if len(steady) else 0.0
    short = min(last, PREROLL + 40)
    stages = port.replay_sequence_stages(scans[:off[short]], off[:short + 1], num_lines, PREROLL, **kw)[4]
    return {"value": fps, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "frames %d..%d of the same sequence (after the same pre-roll), featureExtraction + odometry, one thread" % (PREROLL, last - 1),
            "p50_ms": float(np.percentile(steady, 50)) if len(steady) else None,
            "stage_ms_per_frame": {k: round(v, 3) for k, v in stages.items()},
            "stage_ms_source": "oracle port, frames %d..%d" % (PREROLL, short - 1)}, poses


def run_multi_sequence(capi, device, rank, world, S, skip, n_timed, prm, solo_poses):
    """BASELINE.json configs[4] on ONE GPU: S independent sequences (seeds rank, rank + world, ...: shard_sequences) replayed concurrently,
    one context and one host thread each; contexts share nothing.  Returns (frames, wall seconds, details).  Sequence `rank` is the one
    the single-sequence legs replayed: its poses must come out identical whatever else runs on the device."""
    import torch
    seeds = shard_sequences(S * world, rank, world)
    ctxs = []
    for sd in seeds:
        seq, scans, off = build_sequence(sd, skip + n_timed)
        c = capi.Context(device=device, **prm)
        c.stage_scans(scans, off)
        del scans          # the host copy is not needed once the frames sit in HBM
        c.replay_staged(0, skip)
        ctxs.append(c)
    poses = [None] * S; ms = [0.0] * S
    barrier = threading.Barrier(S + 1)

    def worker(i):
        barrier.wait()
        poses[i], ms[i] = ctxs[i].replay_staged(skip, n_timed)

    th = [threading.Thread(target=worker, args=(i,)) for i in range(S)]
    [t.start() for t in th]
    torch.cuda.synchronize()
    barrier.wait(); t0 = time.perf_counter()
    [t.join() for t in th]
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    same = bool(np.array_equal(poses[0], solo_poses[:n_timed])) if solo_poses is not None else None
    for c in ctxs:
        c.close()
    return S * n_timed, wall, {"per_sequence_device_ms_per_frame": [round(m / n_timed, 4) for m in ms], "first_sequence_identical_to_solo_replay": same}


def replica_identity(rank, poses):
    """1 = this rank's first poses have the committed single-GPU CRC of its seed (tests/golden/replica_pose_crc.json, tools/replica_crc.py),
    0 = they differ, -1 = no committed CRC for this seed.  Reduced over ranks with MIN."""
    import zlib
    import torch
    import torch.distributed as dist
    flag = -1
    try:
        with open(os.path.join(ROOT, "tests", "golden", "replica_pose_crc.json")) as f:
            ref = json.load(f)
        n = int(ref["frames"])
        want = ref["seeds"].get(str(SEED_BASE + rank))
        if want is not None and len(poses) >= n:
            flag = 1 if (zlib.crc32(np.ascontiguousarray(poses[:n]).tobytes()) & 0xffffffff) == int(want) else 0
    except (OSError, ValueError, KeyError):
        flag = -1
    if dist.is_available() and dist.is_initialized():
        t = torch.tensor([flag], dtype=torch.int32, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        flag = int(t.item())
    return {1: True, 0: False}.get(flag)


def mapping_leg(capi, device, seq, scans, off, poses7, prm, n_frames=30):
    """LaserMappingClass::updateCurrentPointsToMap / getMap (src/laserMappingClass.cpp:148-200) at HDL-64 size, fed like the
    mapping node: the filtered cloud (edge + surf features) and the odometry pose of the same frame.  Host buffers in, host wall clock
    (upload + kernels + sync) per call, beside the reference class on one host thread with identical inputs."""
    from floam_b200 import synth
    po, kind = cpu_backend()
    p = dict(prm); p["max_global_map_points"] = 1 << 22
    ctx = capi.Context(device=device, **p)
    ref = po.Mapping(map_resolution=ODOM["map_resolution"])
    upd, get, cpu_upd, cpu_get, sizes = [], [], [], [], []
    first = PREROLL
    for f in range(first, first + n_frames):
        e, sf = ctx.feature_extract(scans[off[f]:off[f + 1]])
        cloud = synth.to_xyzi(np.concatenate([e, sf]))
        T = synth.pose7_to_matrix(poses7[f])
        t0 = time.perf_counter(); ctx.mapping_update(cloud, T); t1 = time.perf_counter()
        m = ctx.mapping_get_map(); t2 = time.perf_counter()
        with StdoutToStderr():
            c0 = time.perf_counter(); ref.update(cloud, T); c1 = time.perf_counter()
            rm = ref.get_map(); c2 = time.perf_counter()
        upd.append((t1 - t0) * 1e3); get.append((t2 - t1) * 1e3); cpu_upd.append((c1 - c0) * 1e3); cpu_get.append((c2 - c1) * 1e3)
        sizes.append((len(cloud), len(m), len(rm)))
    ctx.close()
    # getMap() after every frame (src/laserMappingNode.cpp:87) on a drive that leaves cells behind: the same clouds placed 25 m apart along
    # the trajectory (1 km over 40 frames).  Full download against floam_mapping_get_changed_cells (what the shim's getMap() uses).
    ctx = capi.Context(device=device, **p)
    full_ms, inc_ms, full_pts, inc_pts = [], [], [], []
    for f in range(first, first + 40):
        e, sf = ctx.feature_extract(scans[off[f]:off[f + 1]])
        cloud = synth.to_xyzi(np.concatenate([e, sf]))
        T = seq.pose(2.5 * (f - first))
        ctx.mapping_update(cloud, T)
        t0 = time.perf_counter(); pc, _cells = ctx.mapping_get_changed_cells(); t1 = time.perf_counter()
        m = ctx.mapping_get_map(); t2 = time.perf_counter()
        inc_ms.append((t1 - t0) * 1e3); full_ms.append((t2 - t1) * 1e3); inc_pts.append(len(pc)); full_pts.append(len(m))
    ctx.close()
    incremental = {"frames": 40, "spacing_m": 25.0, "map_points_end": full_pts[-1], "changed_points_p50": float(np.percentile(inc_pts[10:], 50)),
                   "get_changed_cells_ms_p50": float(np.percentile(inc_ms[10:], 50)), "get_map_ms_last": float(np.mean(full_ms[-5:])),
                   "get_changed_cells_ms_last": float(np.mean(inc_ms[-5:])),
                   "note": "every cell of the current 5x5x5 block counts as changed, like the reference re-filters all 125 of them: the hand-out is bounded by the block, the full download grows with the map"}
    return {"frames": n_frames, "incremental_get_map": incremental, "points_per_update": float(np.mean([s_[0] for s_ in sizes])), "map_points_end": sizes[-1][1],
            "update_ms_p50": float(np.percentile(upd[3:], 50)), "get_map_ms_p50": float(np.percentile(get[3:], 50)),
            "cpu_update_ms_p50": float(np.percentile(cpu_upd[3:], 50)), "cpu_get_map_ms_p50": float(np.percentile(cpu_get[3:], 50)), "cpu_kind": kind,
            "map_sizes_equal_to_cpu": bool(all(s_[1] == s_[2] for s_ in sizes)),
            "note": "floam_mapping_update / floam_mapping_get_map with host clouds (H2D of the cloud and D2H of the whole map inside the timings)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--cpu-frames", type=int, default=150, help="frames of the bounded cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--timing-frames", type=int, default=60, help="frames of the per-kernel timing pass (roofline leg)")
    ap.add_argument("--sequences-per-gpu", type=int, default=4, help="concurrent sequences per GPU of the multi_sequence leg (configs[4]); 0 = skip")
    ap.add_argument("--multi-frames", type=int, default=100, help="timed frames per sequence of the multi_sequence leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch
    import torch.distributed as dist
    from floam_b200 import capi
    K = max(1, args.steps); W = max(3, args.warmup)
    K = min(K, 2500)   # bounds host + pinned + device copies of the sequence (3.8 MB per frame each)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    frames = PREROLL + W + K
    t_gen = time.time()
    seq, scans, off = build_sequence(rank, frames)
    t_gen = time.time() - t_gen
    prm = dict(num_lines=seq.num_lines, max_scan_points=seq.max_points + 1024, max_map_points=1 << 21, max_global_map_points=0, max_grid_cells=1 << 23, **ODOM)

    # ---- device-resident replay: the headline `value` ----
    ctx = capi.Context(device=local, **prm)
    ctx.stage_scans(scans, off)
    poses_pre, _ = ctx.replay_staged(0, PREROLL)
    poses_warm, _ = ctx.replay_staged(PREROLL, W)
    map_points_start = list(ctx.odom_map_sizes())
    sampler = ClockSampler(local); sampler.start()
    time.sleep(0.3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ctx.launch_count(reset=True)
    tl0 = ctx.debug_fetch(capi.DBG_TIMELINE, np.int64)
    wall0 = time.time()
    poses_dev, ms_dev = ctx.replay_staged(PREROLL + W, K)
    torch.cuda.synchronize()
    wall1 = time.time()
    launches = ctx.launch_count()
    # the pose-dependent chain of the timed frames from the stamps its kernels leave on the device (globaltimer ns, sums kept in the state):
    # this path is bound by that chain, not by bandwidth, so this is the breakdown the roofline fraction cannot give
    tl1 = ctx.debug_fetch(capi.DBG_TIMELINE, np.int64)
    n_tl = max(1, int(tl1[3] - tl0[3]))
    chain = {"frames": n_tl, "pose_dependent_half_us": float(tl1[0] - tl0[0]) / n_tl / 1e3, "solve_part_us": float(tl1[2] - tl0[2]) / n_tl / 1e3,
             "map_update_part_us": float((tl1[0] - tl0[0]) - (tl1[2] - tl0[2])) / n_tl / 1e3,
             "rest_of_frame_us": 1e3 * ms_dev / K - float(tl1[0] - tl0[0]) / n_tl / 1e3,
             "note": "device globaltimer stamps (predict start, write-back, start of the last kernels of the map update); rest = this rank's frame time minus the half: "
                     "its last kernels, the branch join, mail_state, the graph-to-graph boundary"}
    if world > 1:
        dist.barrier()
    total_frames, max_s = reduce_over_ranks(K, ms_dev * 1e-3)
    value = total_frames / max_s
    # replicas must be replicas: this rank's sequence (seed = rank) against the CRC of the same sequence replayed on ONE GPU (committed)
    identity = replica_identity(rank, np.concatenate([poses_pre, poses_warm, poses_dev]))
    d = ctx.debug()
    ne_map, ns_map = ctx.odom_map_sizes()
    ctx.close()

    # ---- end to end through the public call with host scans in pinned memory ----
    ctx2 = capi.Context(device=local, **prm)
    # the whole sequence sits in page-locked host memory, as a capture driver would leave it; every frame's scan is uploaded
    # from there inside the timed region (floam_process_submit) and its pose read back (floam_process_wait)
    pinned = capi.PinnedBuffer(len(scans))
    pinned.array[:] = scans

    def run_e2e(f0, f1, lat=None):
        poses = np.zeros((f1 - f0, 7)); pending = []
        for f in range(f0, f1):
            if len(pending) == 3:   # the API keeps up to three frames in flight: upload + FRONT of k+2 run under the BACK of k and k+1
                g = pending.pop(0); poses[g - f0] = ctx2.process_wait()
                if lat is not None:
                    lat.append(ctx2.last_frame_ms())
            ctx2.process_submit(pinned.array[off[f]:off[f + 1]])
            pending.append(f)
        for g in pending:
            poses[g - f0] = ctx2.process_wait()
            if lat is not None:
                lat.append(ctx2.last_frame_ms())
        return poses
    run_e2e(0, PREROLL + W)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    lat = []
    e0 = time.time()
    poses_e2e = run_e2e(PREROLL + W, frames, lat)
    torch.cuda.synchronize()
    e1 = time.time()
    if world > 1:
        dist.barrier()
    sampler.stop()
    e2e_frames, e2e_s = reduce_over_ranks(K, e1 - e0, device="cuda")
    e2e_value = e2e_frames / e2e_s
    h2d = float(np.mean(np.diff(off)[PREROLL + W:frames])) * 32 + 4
    ctx2.close()
    identical = bool(np.array_equal(poses_dev, poses_e2e))

    # ---- the same end-to-end call fed with the raw sensor_msgs/PointCloud2 bytes (22 B per point for the Velodyne XYZIRT layout
    #      instead of the 32-byte PCL struct): floam_process_submit_pc2 re-packs on the device, replacing pcl::fromROSMsg ----
    K2 = min(K, 200)
    first_pc2 = PREROLL   # the W warm-up frames go through the same call (its first use allocates the raw-message buffers)
    lay = [capi.pc2_layout(int(off[f + 1] - off[f]), 22) for f in range(first_pc2, PREROLL + W + K2)]
    raw_off = np.zeros(len(lay) + 1, np.int64); raw_off[1:] = np.cumsum([L.row_step * L.height for L in lay])
    raw_pin = capi.PinnedBuffer((int(raw_off[-1]) + 31) // 32 + 1)
    raw_all = raw_pin.array.view(np.uint8)
    for k in range(len(lay)):
        f = first_pc2 + k
        raw_all[raw_off[k]:raw_off[k + 1]] = capi.pack_pointcloud2(pinned.array[off[f]:off[f + 1]], lay[k])
    ctx5 = capi.Context(device=local, **prm)
    ctx5.stage_scans(scans[:off[PREROLL]], off[:PREROLL + 1])
    ctx5.replay_staged(0, PREROLL)

    def run_pc2(k0, k1):
        poses = np.zeros((k1 - k0, 7)); pending = []
        for k in range(k0, k1):
            if len(pending) == 3:
                g = pending.pop(0); poses[g - k0] = ctx5.process_wait()
            ctx5.process_submit_pc2(raw_all[raw_off[k]:raw_off[k + 1]], lay[k])
            pending.append(k)
        for g in pending:
            poses[g - k0] = ctx5.process_wait()
        return poses
    run_pc2(0, W)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    p0 = time.time()
    poses_pc2 = run_pc2(W, W + K2)
    torch.cuda.synchronize()
    p1 = time.time()
    pc2_frames, pc2_s = reduce_over_ranks(K2, p1 - p0, device="cuda")
    pc2 = {"value": pc2_frames / pc2_s, "unit": UNIT, "h2d_bytes_per_step": float(raw_off[-1] - raw_off[W]) / K2, "d2h_bytes_per_step": 56 + 8, "steps": K2,
           "poses_identical_to_device_replay": bool(np.array_equal(poses_pc2, poses_dev[:K2])),
           "call": "floam_process_submit_pc2 / floam_process_wait, raw PointCloud2 bytes in pinned host memory, three frames in flight"}
    ctx5.close()
    raw_pin.close()
    pinned.close()

    # ---- configs[4] per GPU: several sequences on one device ----
    multi = None
    if args.sequences_per_gpu > 0:
        n_ms = max(10, min(args.multi_frames, K))
        if world > 1:
            dist.barrier()
        mf, mw, detail = run_multi_sequence(capi, local, rank, world, args.sequences_per_gpu, PREROLL + W, n_ms, prm, poses_dev)
        tot_f, max_w = reduce_over_ranks(mf, mw, device="cuda")
        multi = {"value": tot_f / max_w, "unit": UNIT, "sequences_per_gpu": args.sequences_per_gpu, "sequences": args.sequences_per_gpu * world,
                 "frames_per_sequence": n_ms, "timing": "host wall clock around the concurrent replays, max over ranks", **detail}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline leg: per-kernel-class device time, kernels launched one by one with CUDA event pairs ----
    TF = max(5, min(args.timing_frames, K))
    ctx3 = capi.Context(device=local, **prm)
    ctx3.stage_scans(scans[:off[PREROLL + W + TF]], off[:PREROLL + W + TF + 1])
    ctx3.replay_staged(0, PREROLL + W)
    ctx3.set_kernel_timing(True)
    stats = {"N": 0.0, "F": 0.0, "Q": 0.0, "C": 0.0, "M": 0.0}
    for f in range(PREROLL + W, PREROLL + W + TF):
        ctx3.process_staged(f)
        dd = ctx3.debug_fetch(capi.DBG_SCALARS, np.int32)
        e_n, s_n = ctx3.odom_map_sizes()
        stats["N"] += float(off[f + 1] - off[f]); stats["Q"] += float(dd[2] + dd[3]); stats["C"] += float(dd[4]); stats["M"] += float(e_n + s_n)
        stats["F"] += float(len(ctx3.debug_fetch(capi.DBG_FEATURE_SRC_EDGE, np.int32)) + len(ctx3.debug_fetch(capi.DBG_FEATURE_SRC_SURF, np.int32)))
    timing = ctx3.kernel_timing()
    ctx3.set_kernel_timing(False)
    ctx3.close()
    # ---- single-frame latency: one scan in flight at a time (host scan in, pose out), what a live 10 Hz sensor would see ----
    NL = min(50, K)
    ctx4 = capi.Context(device=local, **prm)
    ctx4.stage_scans(scans[:off[frames - NL]], off[:frames - NL + 1])
    ctx4.replay_staged(0, frames - NL)
    single = []
    one = capi.PinnedBuffer(int(np.max(np.diff(off))))
    for f in range(frames - NL, frames):
        n_f = int(off[f + 1] - off[f])
        one.array[:n_f] = scans[off[f]:off[f + 1]]      # the capture driver's buffer: filling it is not part of the latency
        t_s = time.perf_counter()
        ctx4.process_scan(one.array[:n_f])
        single.append((time.perf_counter() - t_s) * 1e3)
    ctx4.close()
    one.close()
    mapping = None
    if not args.no_cpu_baseline and world == 1:
        mapping = mapping_leg(capi, local, seq, scans, off, np.concatenate([poses_pre, poses_warm, poses_dev]), prm)
    for k in stats:
        stats[k] /= TF
    # every kernel is bracketed by an event pair inside the frame graph; the pair itself costs a few microseconds, measured by an
    # empty kernel launched the same way (slot "noop") and subtracted
    noop = timing.pop("noop", None)
    overhead_us = (noop[0] / noop[1] * 1e3) if noop else 0.0
    kt = {name: (max(v[0] / v[1] * 1e3 - overhead_us, 0.3), v[1] / TF) for name, v in timing.items()}   # (us per launch, launches per frame)
    total_us = sum(u * n for u, n in kt.values())
    shares = sorted(((name, u * n / total_us, u, n) for name, (u, n) in kt.items()), key=lambda x: -x[1])
    top = shares[0][0]
    peak, peak_kind = measured_peak_gbs()
    # radix / voxel kernels run for several clouds per frame; their per-launch element count is the mean over those clouds
    stats["sort_n"] = (stats["F"] + stats["M"] + stats["Q"]) / 4.0
    stats["sort_out"] = (stats["Q"] + stats["M"]) / 4.0
    ab = algorithmic_bytes(top, stats)
    top_us = kt[top][0]
    roofline = {"bound": "hbm", "kernel": top, "achieved": (ab / (top_us * 1e-6) / 1e9) if ab else None, "peak": peak, "peak_kind": peak_kind,
                "unit": "GB/s", "frac": (ab / (top_us * 1e-6) / 1e9 / peak) if ab else None, "traffic": ncu_traffic_bytes(top),
                "algorithmic_bytes_per_launch": ab, "avg_launch_us": top_us, "event_pair_overhead_us": overhead_us,
                "share_of_frame": shares[0][1], "launches_per_frame": shares[0][3],
                "sum_of_kernel_us_per_frame": total_us,
                "frame_algorithmic_bytes": frame_bytes(stats),
                "frame_hbm_frac": frame_bytes(stats) * value / world / 1e9 / peak,
                "kernel_shares": [{"kernel": n, "share": round(s, 4), "avg_us": round(u, 2), "launches_per_frame": round(l, 2)} for n, s, u, l in shares[:14]],
                "note": "kernel durations = CUDA event pairs recorded as nodes of the frame graph, minus the pair overhead measured with an empty kernel"}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu, poses_cpu = cpu_baseline(scans, off, seq.num_lines, min(args.cpu_frames, W + K))
        n_cmp = min(len(poses_cpu), PREROLL + W + K)
        ours = np.concatenate([poses_pre, poses_warm, poses_dev])[:n_cmp]
        cpu["max_pose_diff_vs_gpu"] = float(np.abs(ours - poses_cpu[:n_cmp]).max())

    clocks = sampler.summary(wall0, e1)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1e3 * max_s / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 geometry / f64 solve", "data": "synthetic",
            "config": workload_config(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 56 + 8,
                    "poses_identical_to_device_replay": identical,
                    "call": "floam_process_submit / floam_process_wait, 32-byte PointXYZIRT scans in pinned host memory, three frames in flight"},
            "e2e_pointcloud2": pc2, "multi_sequence": multi, "laser_mapping": mapping,
            "replicas_identical_to_1gpu_run": identity,
            "gpu_launches": int(launches), "launches_per_frame": launches / K,
            "p50_ms_per_frame": float(np.percentile(lat, 50)), "p99_ms_per_frame": float(np.percentile(lat, 99)),
            "single_frame_latency_ms": {"p50": float(np.percentile(single, 50)), "p99": float(np.percentile(single, 99)), "frames": NL,
                                        "note": "floam_process_scan with nothing else in flight: upload + FRONT + BACK + pose read-back, host wall clock"},
            "knn_queries_per_s": float(stats["Q"] * 2 * value / world),
            "roofline": roofline, "chain": chain, "cpu_baseline": cpu, "clocks": clocks,
            "frame_stats": {k: round(v, 1) for k, v in stats.items()}, "map_points_timed_region": {"start": map_points_start, "end": [ne_map, ns_map]},
            "last_frame": {"n_corr": d["n_corr"], "outer_iterations": d["outer_iterations"], "keyframe": d["keyframe"]},
            "wall_s_timed_region": wall1 - wall0, "gen_s": t_gen}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def frame_bytes(st):
    """SURVEY.md 8d B_frame with the live per-frame averages (n_out = 2, P_lm = 2 x 5 passes, every frame a keyframe)."""
    N, F, Q, C, M = st["N"], st["F"], st["Q"], st["C"], st["M"]
    return 32 * N + 32 * F + 16 * F + 16 * Q + 2 * (16 * Q + 16 * M + 20 * Q + 80 * C) + 10 * 88 * C + 2 * 16 * (M + Q)


if __name__ == "__main__":
    sys.exit(main())

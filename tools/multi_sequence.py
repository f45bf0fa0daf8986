"""BASELINE.json configs[4] on ONE GPU: S independent HDL-64 sequences replayed concurrently, one context + one host thread each
(contexts share nothing; the per-frame path is latency-bound, so several sequences interleave on the same device).
Usage: python tools/multi_sequence.py [S=8] [frames=212]   -> one JSON line"""
import json
import sys
import threading
import time

import numpy as np

sys.path.insert(0, ".")
from floam_b200 import capi, synth


def run(S, frames, device=0):
    seqs = []
    for s in range(S):
        seq = synth.Sequence("hdl64", seed=s)
        scans, off = seq.scans(0, frames)
        seqs.append((seq, scans, off))
    ctxs = []
    for seq, scans, off in seqs:
        c = capi.Context(device=device, num_lines=64, loss="cauchy", max_scan_points=seq.max_points + 1024, max_map_points=1 << 21,
                         max_global_map_points=0, max_grid_cells=1 << 23)
        c.stage_scans(scans, off)
        ctxs.append(c)
    for c in ctxs:   # pre-roll + graph capture, one after the other
        c.replay_staged(0, 12)
        c.replay_staged(12, 20)
    poses = [None] * S; ms = [0.0] * S
    barrier = threading.Barrier(S + 1)

    def worker(i):
        barrier.wait()
        poses[i], ms[i] = ctxs[i].replay_staged(32, frames - 32)

    th = [threading.Thread(target=worker, args=(i,)) for i in range(S)]
    [t.start() for t in th]
    barrier.wait(); t0 = time.perf_counter()
    [t.join() for t in th]
    wall = time.perf_counter() - t0
    # single-sequence reference on the same frames: sequence 0 alone
    c0 = capi.Context(device=device, num_lines=64, loss="cauchy", max_scan_points=seqs[0][0].max_points + 1024, max_map_points=1 << 21,
                      max_global_map_points=0, max_grid_cells=1 << 23)
    c0.stage_scans(seqs[0][1], seqs[0][2]); c0.replay_staged(0, 32)
    p_alone, ms_alone = c0.replay_staged(32, frames - 32)
    out = {"sequences": S, "frames_per_sequence": frames - 32, "aggregate_frames_per_s_wall": S * (frames - 32) / wall,
           "per_sequence_device_ms_per_frame": [round(m / (frames - 32), 4) for m in ms],
           "single_sequence_frames_per_s": (frames - 32) / (ms_alone * 1e-3),
           "sequence0_identical_to_solo_run": bool(np.array_equal(poses[0], p_alone))}
    for c in ctxs + [c0]:
        c.close()
    return out


if __name__ == "__main__":
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 212
    print(json.dumps(run(S, frames)))

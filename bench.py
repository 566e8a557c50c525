#!/usr/bin/env python
"""Benchmark of the PCGCv2 hot path: encode + decode of a synthetic vox10 cloud.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = ``--depth`` (default 2) point-cloud frames per GPU, each through one full ``Codec.encode`` +
``Codec.decode`` (BASELINE.json config 2 stand-in: ``synthetic_vox10``, 795 124 occupied voxels, r3
checkpoint, rho = 1), kept in flight together by ``pcgcv2_b200.pipeline.FramePipeline`` (one host thread +
CUDA stream per frame, so the sequential host range coder of one frame overlaps the kernels of the
other); ``config.serial_ms_per_frame`` is the one-frame-at-a-time latency measured in the same run.  With N
ranks every rank codes its own frames (seed = rank, radii jittered +-10 %: config 3) -- the path is
embarrassingly per-cloud, so the only collective is an all-gather of per-rank counters.

Prints ONE JSON line (rank 0).  ``value`` = Mpoints/s with the input voxels already resident in
HBM; ``e2e`` = the same through the public API with HOST buffers (pinned int32 coordinates in,
decoded coordinates copied back).  The stream every timed frame produces and consumes has all FOUR parts of
coder.py:169-170: the stride-8 coordinates are coded in-process by the octree coder (own format; the reference
spawns tmc3 for this, which ``config.reference_equivalent_wall_ms`` includes) and decoded from the stream again.

``--impl reference`` times the CPU restatement of the reference path (``oracle/``: MinkowskiEngine
and torchac are not installable here, so this is a "port", not the reference's own binaries) on a
bounded sample of the same workload with all host threads.
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
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "Mpoints/sec encode+decode vox10"
UNIT = "Mpoints/s"
CPU_SAMPLE_SCALE = 0.5          # cpu baseline: the same generator on a 512^3 grid (~1/4 of the voxels)


def load_weights(name="r3"):
    import torch
    z = np.load(os.path.join(ROOT, "tests", "golden", f"ckpt_{name}.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.1] or [r for _, r in self.rows]
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if len(r) > 3 + i and r[3 + i].lower() == "active"})
        pw = [float(r[2]) for r in rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(rows), "reasons": reasons}


# ------------------------------------------------------------------------------- ours
def k3_algorithmic_bytes(n, pairs, cin, cout):
    """every feature row read once, every output row written once, every (in,out)
    int32 pair read once, weights once."""
    return 4 * (n * cin + n * cout) + 8 * pairs + 4 * 27 * cin * cout


def pass_algorithmic_bytes(n, p):
    """SURVEY.md section 8(d) summed over every layer of one encode + decode.  ``n`` / ``p``: rows and k=3 pairs of the seven
    coordinate sets {"L0".."L3": analysis levels, "U2","U1","U0": synthesis candidate sets 8 N3 / 8 N2 / 8 N1}.
    Per conv layer 4 (N_in Cin + N_out Cout) + 8 P + 4 K Cin Cout; kernel-map build 8 N + 24 N + 8 P per set; prune
    4 C (N_in + N_kept) + 8 (N_in + N_kept) + N_in; top-k 8 N_in.  Channel plan: pcc_model.py:11-12, autoencoder.py:7-57."""
    conv = lambda nin, nout, pairs, cin, cout, k: 4 * (nin * cin + nout * cout) + 8 * pairs + 4 * k * cin * cout
    k3 = lambda s, cin, cout: conv(n[s], n[s], p[s], cin, cout, 27)
    k1 = lambda s, cin, cout: conv(n[s], n[s], n[s], cin, cout, 1)
    irn = lambda s, c: k3(s, c, c // 4) + k3(s, c // 4, c // 2) + k1(s, c, c // 4) + k3(s, c // 4, c // 4) + k1(s, c // 4, c // 2)
    total = k3("L0", 1, 16)
    for lo, hi, cin, cout, nxt in (("L0", "L1", 16, 32, 32), ("L1", "L2", 32, 64, 64), ("L2", "L3", 64, 32, 8)):
        total += conv(n[lo], n[hi], n[lo], cin, cout, 8) + 3 * irn(hi, cout) + k3(hi, cout, nxt)
    for src, up, kept, cin, cout in (("L3", "U2", "L2", 8, 64), ("L2", "U1", "L1", 64, 32), ("L1", "U0", "L0", 32, 16)):
        total += conv(n[src], n[up], n[up], cin, cout, 8) + k3(up, cout, cout) + 3 * irn(up, cout) + k3(up, cout, 1)
        total += 8 * n[up] + 4 * cout * (n[up] + n[kept]) + 8 * (n[up] + n[kept]) + n[up]
    total += sum(32 * n[s] + 8 * p[s] for s in n)
    return total


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu --set full
    capture (profiles/r02_roofline_traffic.json, refreshed per round), per launch; None if the capture is absent."""
    try:
        return int(json.load(open(os.path.join(ROOT, "profiles", "r02_roofline_traffic.json")))["dram_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        return None


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from pcgcv2_b200 import _lib, ops, synth
    from pcgcv2_b200 import dist as pdist
    from pcgcv2_b200.pipeline import FramePipeline

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")

    pts = synth.synthetic_vox10(seed=rank, jitter=0.1 if world > 1 else 0.0)      # rank 0 @ N=1: 795 124 voxels
    n0 = len(pts)
    depth = max(1, args.depth)
    pipe = FramePipeline(load_weights("r3"), device=dev, depth=depth)             # `depth` frames in flight on this GPU
    codec = pipe.codecs[0]
    host_coords = torch.from_numpy(pts).pin_memory()
    dev_coords = host_coords.to(dev)

    def step_device():                                       # inputs resident in HBM, result left on the device
        return pipe.roundtrip([dev_coords] * depth, to_host=False)[-1]

    def step_e2e():                                          # public API with HOST buffers: H2D and D2H inside
        return pipe.roundtrip([host_coords] * depth, to_host=True, copy=False)[-1]

    def step_serial():                                       # one frame at a time on one stream (latency view)
        st = codec.encode(dev_coords)
        return st, codec.decode(st, to_host=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0, t0 = _lib.launch_count(), time.time()
        start.record()
        for _ in range(steps):
            last = fn()
        end.record()
        barrier()
        ms = pdist.max_over_ranks(start.elapsed_time(end), dev)
        return ms, last, _lib.launch_count() - launches0, (t0, time.time())

    for _ in range(args.warmup):
        st, out = step_device()
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
        step_serial()
    assert out.shape[0] == n0, "decode did not return N0 voxels"

    # roofline probe: the dominant kernel = k3 conv 16->16 on the finest decoder set (8*N1 rows)
    probe_name = "decoder.conv2"
    codec.record = {}
    step_serial()
    _, probe_keys, _ = codec.record[probe_name]
    codec.record = None
    _, npairs = ops.kernel_map_k3(probe_keys, ops.HashTable(probe_keys), count_pairs=True)
    probe_n, probe_pairs = int(probe_keys.shape[0]), int(npairs.item())
    del probe_keys

    sampler = ClockSampler(local_rank) if rank == 0 else None
    codec.probe = {probe_name: []}                           # events on worker 0's stream, inside the timed region
    ms_total, (st, out), launches, (t0, t1) = timed(step_device, args.steps)
    probe_ms = [a.elapsed_time(b) for a, b in codec.probe[probe_name]]
    codec.probe = {}
    clocks = sampler.stop(t0, t1) if sampler else None
    ms_e2e, _, _, _ = timed(step_e2e, args.steps)
    codec.probe = {probe_name: []}                           # the same kernel with nothing else on the GPU
    ms_serial, _, _, _ = timed(step_serial, args.steps)
    probe_ms_serial = [a.elapsed_time(b) for a, b in codec.probe[probe_name]]
    codec.probe = {}
    pipe.close()

    # the path's only collective: per-rank counters
    counters = pdist.gather_counters(torch.tensor([n0 * depth, st.bits() * depth, out.shape[0]], dtype=torch.int64, device=dev))
    total_pts, total_bits = int(counters[:, 0].sum()), int(counters[:, 1].sum())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_step = ms_total / args.steps
    value = pdist.aggregate_throughput(counters[:, 0], ms_step)
    e2e_value = pdist.aggregate_throughput(counters[:, 0], ms_e2e / args.steps)
    # the dominant kernel's duration: CUDA events on its own stream in the serial timed region (nothing else on the GPU);
    # inside the pipelined region the same events also span the time slices of the other frame's kernels
    kern_ms, kern_ms_pipelined = float(np.mean(probe_ms_serial)), float(np.mean(probe_ms))
    alg = k3_algorithmic_bytes(probe_n, probe_pairs, 16, 16)
    achieved = alg / (kern_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "synthetic_vox10(seed=rank) full 3-scale encode+decode, r3 weights, rho=1 "
                               "(stand-in for longdress_vox10_1300.ply)",
                   "points_per_frame": n0, "frames_per_step": world * depth,
                   "parallelism": f"frames sharded over {world} GPU(s), {depth} frame(s) in flight per GPU (one host thread + "
                                  "CUDA stream each: the host range coder of one frame overlaps the kernels of the other)",
                   "serial_ms_per_frame": round(ms_serial / args.steps, 3), "host_cpus": len(os.sched_getaffinity(0)),
                   "arithmetic": "k=3 / k=2 layers: operands split into f16 hi + f16 lo (22 significand bits), three tensor-core "
                                 "products per term pair, f32 accumulation; remaining layers fp32 / 3xTF32",
                   "bpp_features": round(total_bits / total_pts, 5),
                   "coords_side_channel": "raw int32 hand-over (tmc3 subprocess out of scope)",
                   "l2": "per-step traffic (~8.6 GB algorithmic, >1 GB live) exceeds the 126 MB L2; no flush needed"},
        # H2D: input voxels + (decode side) bottleneck coordinates and int16 symbols;
        # D2H: decoded voxels + (encode side) bottleneck coordinates, int16 symbols and the uint16 table
        "e2e": {"value": round(e2e_value, 3), "unit": UNIT,
                "h2d_bytes_per_step": depth * int(host_coords.numel() * 4 + st.coords.size * 4 + st.coords.shape[0] * 8 * 2),
                "d2h_bytes_per_step": depth * int(out.shape[0] * 3 * 4 + st.coords.size * 4 + st.coords.shape[0] * 8 * 2)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "conv_k3_octet_h2_kernel<16,16> (decoder.conv2: k=3 conv 16->16 on the "
                                               "finest decoder set; pre-split f16 hi/lo features, mma.sync m16n8k16, "
                                               "4x4x4 halo per octet staged in shared memory by cp.async)",
                     "rows": probe_n, "pairs": probe_pairs, "algorithmic_bytes": alg,
                     "kernel_ms": round(kern_ms, 4), "kernel_ms_in_pipelined_region": round(kern_ms_pipelined, 4),
                     "timed_in": f"{args.steps} one-frame-at-a-time steps inside bench.py (CUDA events on the launching stream); "
                                 "with 2 frames in flight the events also cover kernels of the other stream sharing the SMs",
                     "achieved": round(achieved, 1), "peak": hbm_peak,
                     "peak_source": peak_src, "unit": "GB/s", "frac": round(achieved / hbm_peak, 4),
                     "traffic": ncu_traffic_bytes()},
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------- cpu arms
def cpu_pass(sd, pts):
    from oracle import codec_ref
    from util import with_batch
    t = time.time()
    st = codec_ref.encode(sd, with_batch(pts))
    dec, _ = codec_ref.decode(sd, st)
    assert len(dec) == len(pts)
    return time.time() - t, st


def cpu_baseline(steps):
    """the oracle (a port of the reference's CPU algorithm: per-offset gather -> mm -> index_add) on the
    box's host cores, on a bounded sample of the workload."""
    import torch
    from pcgcv2_b200 import synth
    torch.set_flush_denormal(True)                          # 46 % of the r3 weights are denormals
    sd = load_weights("r3")
    pts = synth.synthetic_vox10(seed=0, scale=CPU_SAMPLE_SCALE)
    secs = [cpu_pass(sd, pts)[0] for _ in range(steps)]
    return {"value": round(len(pts) / float(np.mean(secs)) / 1e6, 5), "unit": UNIT, "cores": torch.get_num_threads(),
            "kind": "port",
            "sample": f"synthetic_vox10(seed=0, scale={CPU_SAMPLE_SCALE}): {len(pts)} voxels, full encode+decode, "
                      f"{steps} pass(es), {np.mean(secs):.1f} s each, flush-denormal on"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    import torch
    from pcgcv2_b200 import synth
    torch.set_flush_denormal(True)
    sd = load_weights("r3")
    pts = synth.synthetic_vox10(seed=0, scale=CPU_SAMPLE_SCALE)
    for _ in range(args.warmup):
        cpu_pass(sd, pts)
    secs = [cpu_pass(sd, pts)[0] for _ in range(args.steps)]
    sec = float(np.mean(secs))
    value = round(len(pts) / sec / 1e6, 5)
    sample = (f"synthetic_vox10(seed=0, scale={CPU_SAMPLE_SCALE}): {len(pts)} voxels per step (bounded sample of the "
              f"795 124-voxel workload), full encode+decode, flush-denormal on")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(sec * 1e3, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "synthetic_vox10 full 3-scale encode+decode, r3 weights, rho=1 (CPU port of the "
                                   "reference path; MinkowskiEngine/torchac are not installable offline)",
                       "points_per_frame": len(pts)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--depth", type=int, default=2, help="frames in flight per GPU (1 = one frame at a time)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        if args.steps == 20 and args.warmup == 3:           # defaults sized for the GPU arm; keep the CPU arm bounded
            args.steps, args.warmup = 2, 1
        run_reference(args, rank, world)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

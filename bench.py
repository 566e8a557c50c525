#!/usr/bin/env python
"""Benchmark of the PCGCv2 hot path: encode + decode of a synthetic vox10 cloud.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = ``--depth`` (default 4; 3 on hosts with fewer than 6 cores per rank) point-cloud frames per GPU, each through one full ``Codec.encode`` +
``Codec.decode`` (BASELINE.json config 2 stand-in: ``synthetic_vox10``, 795 124 occupied voxels, r3
checkpoint, rho = 1), kept in flight together by ``pcgcv2_b200.pipeline.FramePipeline`` (one host thread +
CUDA stream per frame, so the sequential host range coder of one frame overlaps the kernels of the
other); ``config.serial_ms_per_frame`` is the one-frame-at-a-time latency measured in the same run.  With N
ranks every rank codes the same workload (weak scaling; ``--config3-frames`` switches to config 3's four jittered
clouds) -- the path is embarrassingly per-cloud, so the only collective is an all-gather of per-rank counters.

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
    convs = k3("L0", 1, 16)
    other = sum(32 * n[s] + 8 * p[s] for s in n)                       # kernel-map builds
    for lo, hi, cin, cout, nxt in (("L0", "L1", 16, 32, 32), ("L1", "L2", 32, 64, 64), ("L2", "L3", 64, 32, 8)):
        convs += conv(n[lo], n[hi], n[lo], cin, cout, 8) + 3 * irn(hi, cout) + k3(hi, cout, nxt)
    for src, up, kept, cin, cout in (("L3", "U2", "L2", 8, 64), ("L2", "U1", "L1", 64, 32), ("L1", "U0", "L0", 32, 16)):
        convs += conv(n[src], n[up], n[up], cin, cout, 8) + k3(up, cout, cout) + 3 * irn(up, cout) + k3(up, cout, 1)
        other += 8 * n[up] + 4 * cout * (n[up] + n[kept]) + 8 * (n[up] + n[kept]) + n[up]      # top-k + prune
    return convs, convs + other                                         # (the 106 convolutions: SURVEY Appendix C's 8.57 GB; everything)


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu --set full
    capture (profiles/r02_roofline_traffic.json, refreshed per round), per launch; None if the capture is absent."""
    try:
        return int(json.load(open(os.path.join(ROOT, "profiles", "r02_roofline_traffic.json")))["dram_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        return None


def pin_rank_threads(local_rank, local_world):
    """give every rank of this node its own slice of the host cores: each rank runs `depth` frame threads, `depth` coordinate-coder
    threads and the main thread, and eight ranks on one shared core set contend (round 1: 0.92 scaling efficiency at 8 GPUs with
    no collective on the data path)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // max(local_world, 1)
        if local_world > 1 and per >= 2:
            mine = cores[local_rank * per:(local_rank + 1) * per]
            os.sched_setaffinity(0, mine)
            return len(mine)
        return len(cores)
    except (AttributeError, OSError):
        return os.cpu_count() or 1


def reference_equivalent_wall(pts, sd, res=1024):
    """SURVEY 8(d) "reference-equivalent wall": the reference's OWN, unchanged coder.py (installed under oracle/_ref by build())
    over the drop-in shims -- ASCII PLY read, ME.SparseTensor, Coder.encode (files + tmc3 subprocess), Coder.decode
    (tmc3 + files) -- i.e. what `python coder.py` prints as Enc Time + Dec Time (coder.py:155-162), on this GPU."""
    import tempfile
    import torch
    from oracle import refbin
    if not refbin.reference_sources_installed():
        return None
    from pcgcv2_b200 import ops
    coder = refbin.load_reference_coder()
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as tmp:
        ply = os.path.join(tmp, "in.ply")
        ops.ply_write_ascii(ply, pts)
        os.makedirs(os.path.join(tmp, "output"))
        model = coder.PCCModel().to(coder.device)
        model.load_state_dict(refbin.reference_state_dict(sd))
        c = coder.Coder(model=model, filename=os.path.join(tmp, "output", "in"))
        out = {}
        for rep in range(2):                                     # first pass warms the shim's weight packing
            torch.cuda.synchronize()
            t0 = time.time()
            x = coder.load_sparse_tensor(ply, coder.device)
            torch.cuda.synchronize()
            t1 = time.time()
            c.encode(x)
            torch.cuda.synchronize()
            t2 = time.time()
            dec = c.decode(rho=1)
            torch.cuda.synchronize()
            t3 = time.time()
            out = {"load_ply_ms": round(1e3 * (t1 - t0), 1), "encode_ms": round(1e3 * (t2 - t1), 1), "decode_ms": round(1e3 * (t3 - t2), 1),
                   "encode_plus_decode_ms": round(1e3 * (t3 - t1), 1), "decoded_points": int(len(dec)),
                   "bits": 8 * sum(os.path.getsize(os.path.join(tmp, "output", "in" + p)) for p in ("_C.bin", "_F.bin", "_H.bin", "_num_points.bin"))}
        return out


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from pcgcv2_b200 import _lib, ops, synth
    from pcgcv2_b200 import dist as pdist
    from pcgcv2_b200.pipeline import FramePipeline

    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
    host_cores = pin_rank_threads(local_rank, local_world)
    torch.set_num_threads(1)                                  # the frame threads are the parallelism; no intra-op pools per rank
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    stdout_fd = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries the one JSON line only:
        sys.stdout.flush()                                          # NCCL prints its version banner to fd 1 whatever the debug
        stdout_fd = os.dup(1)                                       # file says, so fd 1 points at stderr until the line is printed
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")

    # Every slot of every rank codes BASELINE config 2's stand-in, synthetic_vox10(0), 795 124 voxels: per-GPU work is the same
    # at every N (weak scaling), and config 3's four distinct clouds are parity-test cases (tests/test_fullsize_gpu.py).
    # --config3-frames: the four jittered clouds instead (seeds 0-3, radii +-10 %: 690-840 k voxels); the `depth` frames a rank
    # keeps in flight are clouds (rank + j) mod 4, so with depth 4 every rank codes all four each step and the ranks still
    # carry equal work (round 1 gave rank r cloud r only: the max-over-ranks time then measured the largest cloud).
    depth = args.depth if args.depth >= 1 else (4 if host_cores >= 6 else 3)
    args.same_frames = not args.config3_frames
    if world > 1 and not args.same_frames:
        clouds = [synth.synthetic_vox10(seed=s, jitter=0.1) for s in range(4)]
        frames_np = [clouds[(rank + j) % 4] for j in range(depth)]
    else:
        frames_np = [synth.synthetic_vox10(seed=0)] * depth
    pts = frames_np[0]
    n0 = len(pts)
    step_points = sum(len(f) for f in frames_np)
    sd = load_weights("r3")
    pipe = FramePipeline(sd, device=dev, depth=depth, coord_bits=10)   # `depth` frames in flight on this GPU; vox10: --res 1024
    codec = pipe.codecs[0]
    host_frames = [torch.from_numpy(f).pin_memory() for f in frames_np]
    dev_frames = [h.to(dev) for h in host_frames]
    host_coords, dev_coords = host_frames[0], dev_frames[0]

    # k steps = k * depth frames handed to the pipeline in ONE call: the frames of consecutive steps follow each other through the
    # workers without a drain / refill of the pipeline at every step boundary (the steady state of a stream of frames)
    def step_device(k=1):                                    # inputs resident in HBM, result left on the device
        return pipe.roundtrip(dev_frames * k, to_host=False)[0]

    def step_e2e(k=1):                                       # public API with HOST buffers: H2D and D2H inside
        return pipe.roundtrip(host_frames * k, to_host=True, copy=False)[0]

    def step_serial():                                       # one frame at a time on one stream (latency view)
        st = codec.encode(dev_coords)
        return st, codec.decode(st, to_host=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, stream_steps=False):
        barrier()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0, t0 = _lib.launch_count(), time.time()
        start.record()
        if stream_steps:
            last = fn(steps)
        else:
            for _ in range(steps):
                last = fn()
        end.record()
        barrier()
        own = start.elapsed_time(end)
        return pdist.max_over_ranks(own, dev), last, _lib.launch_count() - launches0, (t0, time.time()), own

    for _ in range(args.warmup):
        st, out = step_device()
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
        step_serial()
    assert out.shape[0] == n0, "decode did not return N0 voxels"

    # roofline probes: (1) the dominant kernel = k3 conv 16->16 on the finest decoder set (8 N1 rows); (2) the widest layer,
    # 64->64 on the 8 N3 candidate set, which runs on the tcgen05 / TMA kernel
    probes = {"decoder.conv2": (16, 16), "decoder.conv0": (64, 64)}
    codec.record = {}
    step_serial()
    sets = {"L0": "encoder.conv0", "L1": "encoder.conv1", "L2": "encoder.conv2", "L3": "encoder.conv3",
            "U2": "decoder.conv0", "U1": "decoder.conv1", "U0": "decoder.conv2"}
    rows, pairs = {}, {}
    for tag, layer in sets.items():                           # rows and k=3 pairs of the seven coordinate sets (outside every timed region)
        keys = codec.record[layer][1]
        _, npairs = ops.kernel_map_k3(keys, ops.HashTable(keys), count_pairs=True)
        rows[tag], pairs[tag] = int(keys.shape[0]), int(npairs.item())
    codec.record = None
    alg_convs, alg_all = pass_algorithmic_bytes(rows, pairs)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    codec.probe = {name: [] for name in probes}               # events on worker 0's stream, inside the timed region
    ms_total, (st, out), launches, (t0, t1), ms_own = timed(step_device, args.steps, stream_steps=not args.step_barrier)
    probe_ms = {name: [a.elapsed_time(b) for a, b in ev] for name, ev in codec.probe.items()}
    codec.probe = {}
    clocks = sampler.stop(t0, t1) if sampler else None
    ms_e2e, _, _, _, ms_e2e_own = timed(step_e2e, args.steps, stream_steps=not args.step_barrier)
    codec.probe = {name: [] for name in probes}               # the same kernels with nothing else on the GPU
    ms_serial, _, _, _, _ = timed(step_serial, args.steps)
    probe_ms_serial = {name: [a.elapsed_time(b) for a, b in ev] for name, ev in codec.probe.items()}
    codec.probe = {}
    pipe.close()

    # the path's only collective: per-rank counters
    # (bits are those of slot 0's cloud, scaled to the step's points: the other slots' streams are not kept)
    scale_bits = step_points / n0
    mine = [step_points, int(st.bits() * scale_bits), out.shape[0], int(8 * len(st.F) * scale_bits), int(8 * len(st.C or b"") * scale_bits),
            int(round(1e3 * ms_own / args.steps)), int(round(1e3 * ms_e2e_own / args.steps)), step_points]
    counters = pdist.gather_counters(torch.tensor(mine, dtype=torch.int64, device=dev))
    total_pts, total_bits = int(counters[:, 0].sum()), int(counters[:, 1].sum())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_step = ms_total / args.steps
    value = pdist.aggregate_throughput(counters[:, 0], ms_step)
    e2e_value = pdist.aggregate_throughput(counters[:, 0], ms_e2e / args.steps)

    def kernel_line(name, tag, cin, cout, what):
        # duration: CUDA events on the kernel's own stream in the serial timed region (nothing else on the GPU); inside the
        # pipelined region the same events also span the time slices of the other frame's kernels
        ms, ms_pipe = float(np.mean(probe_ms_serial[name])), float(np.mean(probe_ms[name]))
        alg = k3_algorithmic_bytes(rows[tag], pairs[tag], cin, cout)
        ach = alg / (ms * 1e-3) / 1e9
        return {"kernel": what, "rows": rows[tag], "pairs": pairs[tag], "algorithmic_bytes": alg, "kernel_ms": round(ms, 4),
                "kernel_ms_in_pipelined_region": round(ms_pipe, 4), "achieved": round(ach, 1), "unit": "GB/s",
                "frac": round(ach / hbm_peak, 4), "tflops": round(2 * pairs[tag] * cin * cout / (ms * 1e-3) / 1e12, 1)}

    dom = kernel_line("decoder.conv2", "U0", 16, 16,
                      "conv_k3_octet_h2_kernel<16,16> (decoder.conv2: k=3 conv 16->16 on the finest decoder set; pre-split f16 hi/lo "
                      "features, mma.sync m16n8k16, 4x4x4 halo per octet staged in shared memory by cp.async)")
    wide = kernel_line("decoder.conv0", "U2", 64, 64,
                       "wide::conv_k3_wide_kernel<64,64> (decoder.conv0: k=3 conv 64->64 on the 8 N3 candidate set; tcgen05.mma kind::f16, "
                       "accumulators in tensor memory, weight tiles by TMA bulk copy, h2 rows gathered into the SWIZZLE_128B operand)")
    frame_ms = ms_step / depth
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (emulated on the tensor cores: operands f16 hi + f16 lo = 22 significand bits, f32 accumulation)",
        "data": "synthetic",
        "config": {"workload": ("synthetic_vox10(seed=0) full 3-scale encode+decode, r3 weights, rho=1 (stand-in for longdress_vox10_1300.ply)"
                                if (world == 1 or args.same_frames) else
                                "config 3: the four jittered synthetic_vox10 clouds (seeds 0-3), full 3-scale encode+decode, r3 weights, rho=1; "
                                "rank r keeps clouds (r + j) mod 4, j < depth, in flight, i.e. every rank codes all four each step"),
                   "points_per_frame": n0 if (world == 1 or args.same_frames) else [len(c) for c in clouds],
                   "frames_per_step": world * depth,
                   "parallelism": f"frames sharded over {world} GPU(s), {depth} frame(s) in flight per GPU (one host thread + "
                                  "CUDA stream each: the host range coder of one frame overlaps the kernels of the other)",
                   "timed_region": ("every step's frames are awaited before the next step is submitted" if args.step_barrier else
                                    f"the {args.steps} x {depth} frames per GPU of the timed steps are submitted as one stream of frames: no pipeline "
                                    "drain between steps; barrier + synchronize on both sides of the region"),
                   "serial_ms_per_frame": round(ms_serial / args.steps, 3), "host_cpus": len(os.sched_getaffinity(0)),
                   "host_cores_per_rank": host_cores,
                   "arithmetic": "k=3 / k=2 layers: operands split into f16 hi + f16 lo (22 significand bits), products on the tensor "
                                 "cores (tcgen05.mma for 64->64, mma.sync elsewhere), f32 accumulation; remaining layers fp32 / 3xTF32",
                   "bpp_total": round(total_bits / total_pts, 5),
                   "bpp_features": round(int(counters[:, 3].sum()) / total_pts, 5),
                   "bpp_coords": round(int(counters[:, 4].sum()) / total_pts, 5),
                   "coords_side_channel": "in-process octree coder (own format, ~1.5 bits per bottleneck point; tmc3 gives ~1.0: parity runs "
                                          "use Tmc3CoordinateCoder), coded and decoded inside every timed frame on a side thread",
                   "coord_bits": "10 (the --res=1024 of coder.py:196 handed to Codec: radix sorts run over 30 key bits; inputs beyond it are "
                                 "detected on the device and re-coded at full width)",
                   "cdf_table": "built on the host once per symbol range and cached: after warm-up no table work is left in the timed region",
                   "per_rank": {"ms_per_step": [round(v / 1e3, 3) for v in counters[:, 5].tolist()],
                                "e2e_ms_per_step": [round(v / 1e3, 3) for v in counters[:, 6].tolist()],
                                "points_per_step": counters[:, 7].tolist()},
                   "l2": "per-step traffic (~8.6 GB algorithmic, >1 GB live) exceeds the 126 MB L2; no flush needed"},
        # H2D: input voxels + (decode side) bottleneck coordinates and int16 symbols;
        # D2H: decoded voxels + (encode side) bottleneck coordinates and int16 symbols
        "e2e": {"value": round(e2e_value, 3), "unit": UNIT,
                "h2d_bytes_per_step": int(step_points * 12 + depth * (st.coords.size * 4 + st.coords.shape[0] * 8 * 2)),
                "d2h_bytes_per_step": int(step_points * 12 + depth * (st.coords.size * 4 + st.coords.shape[0] * 8 * 2))},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", **dom, "peak": hbm_peak, "peak_source": peak_src,
                     "timed_in": f"{args.steps} one-frame-at-a-time steps inside bench.py (CUDA events on the launching stream); "
                                 f"with {depth} frames in flight the events also cover kernels of the other streams sharing the SMs",
                     "traffic": ncu_traffic_bytes(),
                     "wide_layer": wide,
                     "whole_pass": {"algorithmic_bytes_convs": alg_convs, "algorithmic_bytes_all": alg_all,
                                    "frame_ms_pipelined": round(frame_ms, 3),
                                    "frac": round(alg_convs / (frame_ms * 1e-3) / 1e9 / hbm_peak, 4),
                                    "note": f"106 convolutions' SURVEY 8(d) bytes / wall time per frame with {depth} frames in flight (upper bound on "
                                            "GPU-busy time) / HBM peak"}},
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["config"]["reference_equivalent_wall_ms"] = reference_equivalent_wall(pts, sd)
        except Exception as e:                                 # the reference install is optional on the box
            line["config"]["reference_equivalent_wall_ms"] = {"error": f"{type(e).__name__}: {e}"[:200]}
        line["cpu_baseline"] = cpu_baseline(1, full_size=True)
    if stdout_fd is not None:
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
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


def cpu_baseline(steps, full_size=False):
    """the oracle (a port of the reference's CPU algorithm: per-offset gather -> mm -> index_add) on the
    box's host cores: one pass over the FULL 795 124-voxel cloud (the GPU arm's own workload) when ``full_size``, with the
    half-scale sample the ``--impl reference`` arm steps over beside it."""
    import torch
    from pcgcv2_b200 import synth
    torch.set_flush_denormal(True)                          # 46 % of the r3 weights are denormals
    torch.set_num_threads(len(os.sched_getaffinity(0)))     # all host threads the oracle's torch ops can use
    sd = load_weights("r3")
    pts = synth.synthetic_vox10(seed=0, scale=CPU_SAMPLE_SCALE)
    secs = [cpu_pass(sd, pts)[0] for _ in range(steps)]
    half = round(len(pts) / float(np.mean(secs)) / 1e6, 5)
    out = {"value": half, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
           "sample": f"synthetic_vox10(seed=0, scale={CPU_SAMPLE_SCALE}): {len(pts)} voxels, full encode+decode, "
                     f"{steps} pass(es), {np.mean(secs):.1f} s each, flush-denormal on"}
    if full_size:
        full = synth.synthetic_vox10(seed=0)
        sec = cpu_pass(sd, full)[0]
        out.update({"value": round(len(full) / sec / 1e6, 5), "half_scale_value": half,
                    "sample": f"synthetic_vox10(seed=0): the full {len(full)}-voxel cloud of the GPU arm, one encode+decode pass, {sec:.1f} s, "
                              f"flush-denormal on (half-scale sample of the --impl reference arm: {out['sample']})"})
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    import torch
    from pcgcv2_b200 import synth
    torch.set_flush_denormal(True)
    torch.set_num_threads(len(os.sched_getaffinity(0)))     # all host threads, whatever OMP_NUM_THREADS torchrun exported
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
    ap.add_argument("--depth", type=int, default=0,
                    help="frames in flight per GPU (1 = one frame at a time); 0 = auto: 4 with 6 or more host cores per rank, else 3 "
                         "(measured on B200 with the streamed timed region: 16 cores 3 -> 129, 4 -> 136, 6 -> 134 Mpoints/s; 4 cores "
                         "3 -> 130, 4 -> 127, 6 -> 125; profiles/r02_scaling_host_experiments.txt)")
    ap.add_argument("--workload", default="codec", choices=["codec", "train"],
                    help="codec = BASELINE's headline (default); train = the config-5 training step (tools/bench_train.py)")
    ap.add_argument("--batch", type=int, default=32, help="--workload train: samples per rank and step")
    ap.add_argument("--step-barrier", action="store_true",
                    help="wait for every step's frames before submitting the next step's (drains the pipeline at each step boundary)")
    ap.add_argument("--config3-frames", action="store_true",
                    help="N > 1: code config 3's four jittered clouds (every rank all four, rotated) instead of config 2's cloud in every slot")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        if args.steps == 20 and args.warmup == 3:           # defaults sized for the GPU arm; keep the CPU arm bounded
            args.steps, args.warmup = 2, 1
        run_reference(args, rank, world)
    elif args.workload == "train":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_train
        bench_train.run(args, rank, world, local_rank)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

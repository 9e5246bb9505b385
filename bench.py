#!/usr/bin/env python
"""bench.py — BEV frames/s of the point-cloud -> BEV front end (voxelize + PFN + scatter) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path over one batch of synthetic frames (BASELINE.json configs[1]: KITTI-shaped
frames, batch 16, x 0..80 / y +-40 @ 0.1 m -> 800x800 canvas, [128,128,128] PFN, fp32 forward, eval-mode BN).
N > 1: one process per GPU (torchrun), frames shard by rank (weak scaling: every rank runs its own batch of 16;
`--scaling strong`: the workload's batch is split over the ranks, BASELINE configs[2]), no data-path collective;
barrier + synchronize on both sides, device time via CUDA events, max over ranks.

Keys beyond the base contract: `roofline` (dominant kernel, algorithmic bytes / event time vs MEASURED_PEAKS.json, and
`step_frac` = the whole step against the fused-path HBM floor), `kernels` (per-kernel breakdown), `cpu_baseline`
(oracle port timed on this box's host cores at 1 / 6 / all threads), `e2e` (same metric through the C-ABI pipelined
entry with HOST points: pinned host points -> H2D -> K1,K2,K3 -> D2H of the per-frame pillar counts), `train` (N > 1 or
--train: forward + backward + NCCL gradient allreduce of the encoder incl. LayerNorm, the north star's only collective).
`--impl reference` times the oracle port (the reference's CPU path cannot be installed: mmcv/mmdet3d absent).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bev_frames_per_sec_voxelize_pfn_scatter"
UNIT = "frames/s"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json copy bandwidth)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class NvmlSampler:
    """SM clock / throttle reasons through NVML every ~5 ms during the timed region (a step is ~2 ms, so nvidia-smi's
    200 ms period sees one or two samples of a default run); falls back to ClockSampler when NVML is unavailable."""
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index: int):
        self.index, self.ok, self.samples, self.stop_flag = index, False, [], False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:  # noqa: BLE001
            self.fallback = ClockSampler(index)

    def start(self):
        if not self.ok:
            return self.fallback.start()
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.samples.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                                     int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)),
                                     nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def stop(self):
        if not self.ok:
            return self.fallback.stop()
        self.stop_flag = True
        self.thread.join(timeout=1)
        sm = [s[0] for s in self.samples]
        reasons = sorted(n for n, b in self.BITS.items() if any(s[1] & b for s in self.samples))
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": self.max, "reasons": reasons, "samples": len(sm),
                "power_w_max": max((s[2] for s in self.samples), default=None), "source": "nvml"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(name: str, rank: int, batch=None, n=None):
    from mask_bev_b200.synthetic import CONFIGS, encoder_kwargs, gen_batch
    cfg = CONFIGS[name]
    B = cfg["batch"] if batch is None else batch
    frames = gen_batch(name, batch=B, n=n, first_frame=rank * B)
    return cfg, encoder_kwargs(name), frames


def make_encoder(kwargs, device):
    """Random-init weights of the configured architecture (no checkpoints offline): oracle-side randomisation so the
    oracle and the product share them."""
    import torch
    import mask_bev_b200 as M
    from oracle import oracle as O
    enc = M.MaskBevEncoder(**kwargs)
    pfn = O.make_pfn_oracle(in_channels=kwargs["pc_point_dim"], feat_channels=kwargs["feat_channels"],
                            with_distance=True, voxel_size=[kwargs["voxel_size_x"], kwargs["voxel_size_y"], kwargs["voxel_size_z"]],
                            point_cloud_range=[kwargs["x_range"][0], kwargs["y_range"][0], kwargs["z_range"][0],
                                               kwargs["x_range"][1], kwargs["y_range"][1], kwargs["z_range"][1]])
    O.randomise_pfn(pfn, seed=0)
    enc._voxel_encoder.load_state_dict(pfn.state_dict())
    return enc.to(device).eval(), pfn.eval()


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's CPU path (serial C voxelizer + dense torch-CPU PFN + scatter)
# ------------------------------------------------------------------------------------------------------
def cpu_path_once(frames, kwargs, pfn):
    import torch
    from oracle import oracle as O
    geo = O.encoder_geometry(kwargs["x_range"], kwargs["y_range"], kwargs["z_range"], kwargs["voxel_size_x"],
                             kwargs["voxel_size_y"], kwargs["voxel_size_z"])
    vs, ns, cs = [], [], []
    for b, fr in enumerate(frames):
        f, _ = O.filter_in_range_c(fr, kwargs["x_range"], kwargs["y_range"], kwargs["z_range"])
        v, c, n, _ = O.hard_voxelize_c(f, geo["voxel_size"], geo["point_cloud_range"], kwargs["max_num_points"],
                                       250000, want_kept=False)
        vs.append(v); ns.append(n)
        cs.append(np.concatenate([np.full((len(c), 1), b, np.int32), c], 1))
    voxels, nump, coors = np.concatenate(vs), np.concatenate(ns), np.concatenate(cs)
    with torch.no_grad():
        feats = pfn(torch.from_numpy(voxels), torch.from_numpy(nump), torch.from_numpy(coors)).numpy()
    return O.scatter_np(feats, coors, len(frames), geo["ny"], geo["nx"])


def time_cpu(frames, kwargs, pfn, budget_s=20.0, min_reps=2, max_reps=10):
    cpu_path_once(frames[:1], kwargs, pfn)  # warm-up (page-in, thread pools)
    ts = []
    t_all = time.perf_counter()
    while len(ts) < max_reps and (len(ts) < min_reps or time.perf_counter() - t_all < budget_s):
        t0 = time.perf_counter()
        cpu_path_once(frames, kwargs, pfn)
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), len(ts)


def run_reference_arm(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build_c_oracle()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = build_workload(args.workload, 0, batch=1)[0]
    B = cfg["batch"] if args.batch is None else args.batch
    _, kwargs, frames = build_workload(args.workload, 0, batch=B)
    _, pfn = make_encoder_cpu_only(kwargs)
    cpu_path_once(frames[:1], kwargs, pfn)
    t0 = time.perf_counter()
    cpu_path_once(frames[:1], kwargs, pfn)
    t_frame = time.perf_counter() - t0
    # bounded sample: as many frames of the B-frame batch per step as keep the whole --steps/--warmup run near 2 minutes
    sample_frames = int(max(1, min(B, 120.0 / (t_frame * (args.steps + max(args.warmup, 1))))))
    for _ in range(max(args.warmup, 1)):
        cpu_path_once(frames[:sample_frames], kwargs, pfn)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_path_once(frames[:sample_frames], kwargs, pfn)
    dt = time.perf_counter() - t0
    fps = sample_frames * args.steps / dt
    sample = (f"{sample_frames} frames of the {B}-frame batch per step (frames/s is batch-normalised); serial C voxelizer "
              f"(mmcv's CPU kernel is serial) + dense torch-CPU PFN over all P*T slots + scatter, torch threads={cores}")
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, cfg, B), "sample_frames_per_step": sample_frames,
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "oracle port: the reference's own CPU path (mmcv==2.0.0 / mmdet3d==1.1.0) is not installable offline"}
    print(json.dumps(line), flush=True)


def make_encoder_cpu_only(kwargs):
    from oracle import oracle as O
    pfn = O.make_pfn_oracle(in_channels=kwargs["pc_point_dim"], feat_channels=kwargs["feat_channels"],
                            with_distance=True, voxel_size=[kwargs["voxel_size_x"], kwargs["voxel_size_y"], kwargs["voxel_size_z"]],
                            point_cloud_range=[kwargs["x_range"][0], kwargs["y_range"][0], kwargs["z_range"][0],
                                               kwargs["x_range"][1], kwargs["y_range"][1], kwargs["z_range"][1]])
    O.randomise_pfn(pfn, seed=0)
    return None, pfn.eval()


def workload_config(name, cfg, batch_per_gpu):
    return {"workload": name, "frames_per_gpu_per_step": batch_per_gpu, "points_per_frame": cfg["n"],
            "point_feats": cfg["C"], "x_range": list(cfg["x_range"]), "y_range": list(cfg["y_range"]),
            "voxel_size": cfg["voxel_size"], "max_num_points": cfg["T"], "pfn_feat_channels": list(cfg["feat_channels"]),
            "bn_mode": "eval", "l2_policy": "per-step working set (canvas) is >> the 126 MB L2: no flush needed"}


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def ev_time(fn, iters, stream_sync):
    import torch
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()  # untimed: first launch of a kernel pays CUDA's lazy module load and any lazily sized workspace
    stream_sync()
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    e.synchronize()
    return s.elapsed_time(e) / iters  # ms


def time_cpu_threads(frames, kwargs, pfn, threads, nframes, budget_s):
    import torch
    torch.set_num_threads(threads)
    t, reps = time_cpu(frames[:nframes], kwargs, pfn, budget_s=budget_s, min_reps=1, max_reps=5)
    return {"value": nframes / t, "frames": nframes, "reps": reps}


def train_block(args, enc, kwargs, rank, world, dev, barrier):
    """Forward + backward of MaskBevEncoder.forward (K1, K2, K3+LayerNorm and their backward kernels, train-mode BN)
    on this rank's frames, then the gradient allreduce of mask_bev_b200.data_parallel over NCCL — the only collective
    of the north star (train_mask_bev.py:92-96 strategy='ddp'). Timed with CUDA events, MAX over ranks."""
    import torch
    import torch.distributed as dist
    import mask_bev_b200 as M
    from mask_bev_b200.data_parallel import FrontEndDataParallel, gradient_bytes
    from mask_bev_b200.synthetic import gen_batch
    tb = args.train_batch
    torch.manual_seed(0)  # identical initial weights on every rank
    tenc = M.MaskBevEncoder(**kwargs).to(dev)
    tenc._voxel_encoder.load_state_dict(enc._voxel_encoder.state_dict())
    tenc.train()
    frames = [torch.from_numpy(f).to(dev) for f in gen_batch(args.workload, batch=tb, first_frame=10_000 + rank * tb)]
    out = {}
    g = None
    for overlap in (False, True):
        dp = FrontEndDataParallel(tenc, overlap=overlap)

        def step(reduce=True):
            nonlocal g
            tenc.zero_grad()  # set_to_none=True, torch's and Lightning's default: the fresh LayerNorm gradients become .grad without a 655 MB zero + add pass
            y = tenc(frames)
            if g is None:
                g = torch.randn_like(y)
            y.backward(g)
            return dp.reduce_gradients() if reduce else None

        def timed(n, reduce):
            barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(n):
                rep = step(reduce)
            e.record()
            barrier()
            t = torch.tensor([s.elapsed_time(e) / n], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t), rep

        for _ in range(2):
            step()
        n = max(2, min(args.steps, 5))
        if not overlap:
            out["fwd_bwd_ms"], _ = timed(n, False)
        ms, rep = timed(n, True)
        out["step_overlapped_ms" if overlap else "step_serial_ms"] = ms
        dp.close()
    nbytes = gradient_bytes(tenc)
    out.update({"what": "encoder training step (train-mode BN; row-space PFN forward whose rows the backward reuses, fused "
                        "scatter+LayerNorm fwd/bwd; zero_grad() = set_to_none) + gradient allreduce; "
                        "overlapped = the LayerNorm-gradient allreduce is issued from a post-accumulate hook and runs "
                        "under the PFN backward",
                "frames_per_gpu": tb, "allreduce_bytes_per_rank": nbytes, "collectives_per_step": rep.collectives if rep else 0,
                "frames_per_s": world * tb / (out["step_overlapped_ms"] * 1e-3)})
    if world > 1:  # the collective alone, back to back, for its bus bandwidth
        big = [p.grad for p in tenc.parameters() if p.grad is not None and p.grad.numel() * 4 >= (1 << 20)]
        for _ in range(2):
            for t in big:
                dist.all_reduce(t)
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(3):
            for t in big:
                dist.all_reduce(t)
        e.record()
        barrier()
        t_ar = torch.tensor([s.elapsed_time(e) / 3], dtype=torch.float64, device=dev)
        dist.all_reduce(t_ar, op=dist.ReduceOp.MAX)
        by = sum(t.numel() * 4 for t in big)
        out.update({"allreduce_alone_ms": float(t_ar), "allreduce_bus_gbs": 2.0 * (world - 1) / world * by / float(t_ar) / 1e6,
                    "allreduce_exposed_ms": out["step_overlapped_ms"] - out["fwd_bwd_ms"],
                    "nvlink_reference": "8-rank all-reduce bus bandwidth 725 GB/s at 1 GiB (B200_PROFILING.md)"})
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from mask_bev_b200 import _lib
    from mask_bev_b200.runtime import FusedEncoderRunner

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    saved_stdout = None
    if world > 1:
        # NCCL writes its version banner to stdout when the first communicator is built; rank 0 must print ONE
        # JSON line, so everything before that line goes to stderr (fd 1 -> fd 2 until the final print)
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    from mask_bev_b200.synthetic import CONFIGS, encoder_kwargs, gen_batch
    cfg = CONFIGS[args.workload]
    kwargs = encoder_kwargs(args.workload)
    strong = args.scaling == "strong"
    if strong:  # BASELINE configs[2]: ONE batch of the workload's size split over the ranks (frame i -> rank i mod N)
        Bg = cfg["batch"] if args.batch is None else args.batch
        frames = [f for i, f in enumerate(gen_batch(args.workload, batch=Bg)) if i % world == rank]
    else:
        frames = gen_batch(args.workload, batch=args.batch, first_frame=rank * (cfg["batch"] if args.batch is None else args.batch))
    B = len(frames)
    if B == 0:
        raise SystemExit(f"bench.py: rank {rank} owns no frame (batch smaller than the number of GPUs)")
    enc, pfn_cpu = make_encoder(kwargs, dev)
    runner = FusedEncoderRunner(enc, [len(f) for f in frames], dev, scatter_ctas_per_sm=args.scatter_ctas)
    host = torch.from_numpy(np.concatenate(frames, 0)).pin_memory()
    runner.set_points(host)
    counts_host = torch.empty((B + 1,), dtype=torch.int32).pin_memory()
    sync = lambda: torch.cuda.synchronize(dev)  # noqa: E731

    def barrier():
        if world > 1:
            dist.barrier()
        sync()

    # ---- device-resident throughput (`value`): a stream of batches through mbev_encode_batch_pipelined (K1 of step
    # i+2 on a prep stream, K2 of step i+1 on a PFN stream, K3 of step i on the current stream);
    # `serial_ms_per_step` = mbev_encode_batch, everything on one stream ----
    step_fn = runner.run_device if args.serial else runner.run_pipelined
    for _ in range(args.warmup):
        runner.run_device()
    barrier()
    s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(args.steps):
        runner.run_device()
    e0.record()
    barrier()
    ms_serial = s0.elapsed_time(e0)
    for _ in range(args.warmup):
        step_fn()
    barrier()
    l0 = _lib.launch_count()
    sampler = NvmlSampler(local)
    if rank == 0:
        sampler.start()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        step_fn()
    e.record()
    barrier()
    ms = s.elapsed_time(e)
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    # ---- end to end through the host entry (`e2e`): pinned host points in, per-frame pillar counts out, every
    # step; pipelined = the H2D copy and K1 of step i+2 overlap K2 / K3 of the steps before ----
    host2 = host.clone().pin_memory()  # alternate two host batches so that no step can reuse a stale device copy

    def e2e_loop(fn, base_of):
        for i in range(max(1, args.warmup // 2)):
            fn(host2 if i & 1 else host)
        barrier()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        for i in range(args.steps):
            fn(host2 if i & 1 else host)
            counts_host.copy_(base_of(), non_blocking=True)  # ordered on the current stream, after this step's K3
        e2.record()
        barrier()
        return s2.elapsed_time(e2)

    ms2s = e2e_loop(runner.run_host, lambda: runner.pillar_base)
    ms2 = ms2s if args.serial else e2e_loop(runner.run_pipelined, lambda: runner.last_pillar_base)
    t = torch.tensor([ms, ms2, ms2s, ms_serial], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms2, ms2s, ms_serial = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    nfr = torch.tensor([B], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(nfr)
    frames_all = int(nfr.item())  # frames all ranks process per step

    train = None
    if (world > 1 or args.train) and not args.no_train:
        runner_state = None
        try:
            train = train_block(args, enc, kwargs, rank, world, dev, barrier)
        except Exception as ex:  # noqa: BLE001  (never lose the headline line to the auxiliary block)
            train = {"error": f"{type(ex).__name__}: {ex}"[:300]}
        del runner_state
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel breakdown and roofline (rank 0) ----
    peak, peak_src = _peaks()
    iters = max(3, min(args.steps, 10))
    runner.run_device(); sync()
    base = runner.pillar_base.cpu().numpy()
    P = int(base[-1])
    nk = int(runner.num_points[:P].sum().item())
    N = runner.total
    C, T, G, Co = runner.C, runner.geo.max_points, runner.ny * runner.nx, runner.c_out
    t_vox = ev_time(runner.run_voxelize, iters, sync)
    t_pfn = ev_time(runner.run_pfn, iters, sync)
    t_sc = ev_time(runner.run_scatter, iters, sync)
    # a plain fill of the same canvas: what a write-only stream reaches on this GPU (the roofline `peak` below is the
    # driver's measured COPY bandwidth, which a pure write stream can exceed)
    t_fill = ev_time(lambda: runner.canvas.zero_(), iters, sync)
    runner.run_scatter()
    stream_ok = bool(runner.lib.mbev_scatter_stream_supported(Co, runner.ny, runner.nx, ctypes.c_void_p(runner.canvas.data_ptr())))
    t_st = {c: ev_time(lambda c=c: runner.run_scatter_stream(c), iters, sync) for c in (1, 2, 4, 8)} if stream_ok else {}
    t_nhwc = ev_time(runner.run_scatter_nhwc, iters, sync) if Co % 4 == 0 else None
    runner.run_scatter(); sync()
    # algorithmic bytes per launch (per frame x frames in the batch; DESIGN.md "Roofline accounting")
    by_vox = N * C * 4 + N * 4 + P * 20
    by_pfn = nk * (C * 4 + 4) + P * 20 + P * Co * 4
    by_sc = P * Co * 4 + P * 16 + B * G * Co * 4
    by_floor = N * C * 4 + B * G * Co * 4   # fused-path floor: read the points once, write the canvas once
    units = [l.units for l in enc._voxel_encoder.pfn_layers]
    ins = [l.linear.in_features for l in enc._voxel_encoder.pfn_layers]
    mac_row = sum(i * u for i, u in zip(ins, units))
    fl_pfn = 2.0 * mac_row * (nk + P)

    def hbm(by, t):
        return {"ms": t, "alg_bytes": by, "gbs": by / t / 1e6, "frac_hbm": by / t / 1e6 / peak}

    kernels = {
        "K1_voxelize": hbm(by_vox, t_vox),
        "K2_pfn": {**hbm(by_pfn, t_pfn), "alg_tflops_upstream_equiv": fl_pfn / t_pfn / 1e9, "fp32_fma_peak_tflops": 74.4,
                   "path": "tcgen05 3xTF32" if runner.lib.mbev_pfn_path(ctypes.byref(runner.params), T) == 2 else "fp32 FMA"},
        "K3_scatter": {**hbm(by_sc, t_sc), "kernel": "k_scatter_run (registers -> st.global.cs.v4; tasks ordered frame / 8-plane chunk / run: sequential DRAM streams)"},
    }
    for c, tt in t_st.items():
        kernels[f"K3_scatter_stream_{c}cta"] = {**hbm(by_sc, tt), "kernel": f"k_scatter_bulk (TMA bulk stores), {c} x 148 CTAs of 128 threads"}
    if t_nhwc is not None:
        kernels["K3_scatter_channels_last"] = {**hbm(by_sc, t_nhwc), "kernel": "k_scatter_nhwc ((B, ny, nx, C) canvas)"}
    # BASELINE config 4 flavour (not the headline, which is fp32): same path with a bfloat16 canvas
    bf16 = None
    if (G & 3) == 0:
        t_sc16 = ev_time(runner.run_scatter_bf16, iters, sync)
        t_step16 = ev_time(runner.run_device_bf16, iters, sync)
        by_sc16 = P * Co * 4 + P * 16 + B * G * Co * 2
        bf16 = {"K3_scatter_bf16_ms": t_sc16, "alg_bytes": by_sc16, "gbs": by_sc16 / t_sc16 / 1e6,
                "frac_hbm": by_sc16 / t_sc16 / 1e6 / peak, "fp32_pfn_ms_per_step": t_step16,
                "fp32_pfn_frames_per_s": B / (t_step16 * 1e-3),
                "note": "K1 -> K2 -> K3 with a bf16 canvas on one stream; fp32_pfn_*: K2 in fp32 (3xTF32), the canvas "
                        "rounded on the way out; bf16_pfn_*: K2 with gemm_path = MBEV_GEMM_TCGEN05_BF16 as well"}
        net = enc._voxel_encoder
        net.gemm_path = "tcgen05_bf16"
        try:
            runner.refresh_params()
            if runner.lib.mbev_pfn_path(ctypes.byref(runner.params), T) == 3:
                t_pfn16 = ev_time(runner.run_pfn, iters, sync)
                t_all16 = ev_time(runner.run_device_bf16, iters, sync)
                bf16.update({"K2_pfn_bf16_ms": t_pfn16, "bf16_pfn_ms_per_step": t_all16,
                             "bf16_pfn_frames_per_s": B / (t_all16 * 1e-3)})
        finally:
            net.gemm_path = "auto"
            runner.refresh_params()
            runner.run_pfn(); sync()   # feats back to the fp32 path's for the timings below
        runner.canvas_bf16 = None
    # SURVEY §8 f1 ("next" row, not part of the headline metric): the LayerNorm that follows the scatter in
    # MaskBevEncoder.forward, fused into the scatter, next to torch's own LayerNorm on the finished canvas
    layernorm = None
    if not args.no_layernorm:
        from mask_bev_b200 import functional as F_
        ln = enc._layer_norm
        out_ln = torch.empty_like(runner.canvas)

        def fused_ln(walk):
            return F_.scatter_layernorm_forward(runner.feats, runner.cell_table, runner.pillar_base, B, runner.ny,
                                                runner.nx, ln.weight, ln.bias, ln.eps, out=out_ln, walk=walk)
        if fused_ln(_lib.LN_WALK_RUNS) is not None:
            t_ln = ev_time(lambda: fused_ln(_lib.LN_WALK_RUNS), iters, sync)
            t_lnf = ev_time(lambda: fused_ln(_lib.LN_WALK_FRAMES), iters, sync) if Co % 4 == 0 else None
            with torch.no_grad():
                t_torch = ev_time(lambda: ln(runner.canvas), 2, sync)
            by_ln = B * G * Co * 4 + 2 * G * Co * 4 + P * Co * 4 + B * G * 4
            layernorm = {"K3+LN_walk_runs_ms": t_ln, "K3+LN_walk_frames_ms": t_lnf, "alg_bytes": by_ln,
                         "default_walk": "frames" if F_.LN_WALK_DEFAULT == _lib.LN_WALK_FRAMES else "runs",
                         "K3_then_torch_layernorm_ms": t_sc + t_torch, "torch_layernorm_ms": t_torch,
                         "note": "mbev_scatter_layernorm_forward vs K3 followed by nn.LayerNorm([C,ny,nx]) (mask_bev_encoders.py:75,92)"}
            t_best = min(x for x in (t_ln, t_lnf) if x is not None)
            layernorm.update({"K3+LN_fused_ms": t_best, "gbs": by_ln / t_best / 1e6, "frac_hbm": by_ln / t_best / 1e6 / peak})
            # its backward (training): out_ln stands in for the incoming gradient (any dense tensor of that shape)
            if F_.scatter_layernorm_backward_supported(B, Co, runner.ny, runner.nx):
                stats_ln = fused_ln(_lib.LN_WALK_RUNS)[1]

                def fused_ln_bwd():
                    return F_.scatter_layernorm_backward(out_ln, runner.feats, runner.cell_table, runner.coors,
                                                         runner.pillar_base[B:], ln.weight, stats_ln)
                t_lnb = ev_time(fused_ln_bwd, iters, sync)
                by_lnb = B * G * Co * 4 + 3 * G * Co * 4 + B * G * 4 + 5 * P * Co * 4
                layernorm.update({"K3+LN_backward_ms": t_lnb, "backward_alg_bytes": by_lnb,
                                  "backward_gbs": by_lnb / t_lnb / 1e6, "backward_frac_hbm": by_lnb / t_lnb / 1e6 / peak,
                                  "backward_note": "mbev_scatter_layernorm_backward: dy read once, dweight / dbias "
                                                   "written once, dfeats in pillar space"})
        del out_ln
    # SURVEY §8 f2 ("next" row): the first consumer of the pseudo image, Swin's patch embedding (swin.py:578-586,
    # 745-746; patch 4, embed_dims 192 as in configs/training), fed from the pillars — next to the canvas route
    # (fused K3 + LayerNorm, then torch's Conv2d + LayerNorm on the 5 GB image)
    patch_embed = None
    if not args.no_layernorm and not args.no_patch_embed:
        import mask_bev_b200 as M
        if _lib.load().mbev_patch_embed_supported(B, Co, runner.ny, runner.nx, 4, 192):
            pe = M.PillarPatchEmbed(in_channels=Co, embed_dims=192, kernel_size=4, stride=4, norm_cfg=dict(type="LN")).to(dev)
            ln = enc._layer_norm

            def f2():
                return pe.forward_pillars(runner.feats, runner.coors, runner.cell_table, runner.pillar_base, B, runner.ny,
                                          runner.nx, ln)
            with torch.no_grad():
                tok, (Hp, Wp) = f2()
                t_f2 = ev_time(f2, iters, sync)
                by_f2 = P * Co * 4 * 2 + P * 16 + B * G * 4 + B * Hp * Wp * 192 * 4  # feats + ln weight rows, coors, table, tokens
                patch_embed = {"f2_tokens_ms": t_f2, "alg_bytes": by_f2, "gbs": by_f2 / t_f2 / 1e6,
                               "frac_hbm": by_f2 / t_f2 / 1e6 / peak, "tokens_shape": [B, Hp * Wp, 192],
                               "front_end_to_tokens_ms": t_vox + t_pfn + t_f2,
                               "note": "K1 + K2 + mbev_patch_embed_forward: LayerNorm([C,ny,nx]) + Conv2d(k=s=4) + LayerNorm(E) "
                                       "from the pillars, the pseudo image is never written"}
                try:  # the canvas route on the same GPU: our fused K3+LN, then torch / cuDNN for the convolution
                    from mask_bev_b200 import functional as F_
                    x = torch.empty_like(runner.canvas)
                    F_.scatter_layernorm_forward(runner.feats, runner.cell_table, runner.pillar_base, B, runner.ny, runner.nx,
                                                 ln.weight, ln.bias, ln.eps, out=x)

                    pad = (0, Wp * 4 - runner.nx, 0, Hp * 4 - runner.ny)  # PatchEmbed's corner padding

                    def dense():
                        xp = torch.nn.functional.pad(x, pad) if (pad[1] or pad[3]) else x
                        return pe.norm(pe.projection(xp).flatten(2).transpose(1, 2))
                    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                        ref = dense()
                        t_dense = ev_time(dense, 3, sync)
                    err = float((tok - ref).abs().max() / ref.abs().max())
                    patch_embed.update({"canvas_route_conv_ms": t_dense, "rel_err_vs_canvas_route": err,
                                        "canvas_route_note": "torch Conv2d (fp32, no TF32) + LayerNorm on the LayerNorm-ed canvas; "
                                                             "add K3+LN_fused_ms for the whole canvas route"})
                    del x, ref
                except Exception as ex:  # noqa: BLE001
                    patch_embed["canvas_route_error"] = str(ex)[:200]
            del tok
    # the K3 form the timed step used: its stand-alone launch time is the roofline line; step_frac is the whole step
    use_stream = (not args.serial) and stream_ok and args.scatter_ctas > 0
    k3_key = f"K3_scatter_stream_{args.scatter_ctas}cta" if use_stream and args.scatter_ctas in t_st else "K3_scatter"
    k3 = kernels[k3_key]
    dom = max(("K1_voxelize", "K2_pfn", k3_key), key=lambda k: kernels[k]["ms"])
    step_ms = ms / args.steps
    roof = {"kernel": f"{k3_key} ({k3['kernel']})", "bound": "hbm", "achieved": k3["gbs"], "peak": peak, "unit": "GB/s",
            "frac": k3["frac_hbm"], "traffic": None, "peak_source": peak_src, "launch_ms": k3["ms"],
            "dominant_by_time": dom, "share_of_serial_step": k3["ms"] / (t_vox + t_pfn + k3["ms"]),
            "timed": "stand-alone launches (CUDA events on the launching stream, inputs resident)"
                     + ("; inside the pipelined step this kernel shares the SMs with K2 of the next batch" if use_stream else ""),
            "fill_gbs": runner.canvas.numel() * 4 / t_fill / 1e6, "frac_of_fill": k3["gbs"] / (runner.canvas.numel() * 4 / t_fill / 1e6),
            "fill_note": "torch fill kernel on the same canvas, timed live: the write-only ceiling (frac > 1 means above the COPY peak, not above the hardware)",
            "step_floor_bytes": by_floor, "step_frac": by_floor / step_ms / 1e6 / peak,
            "step_frac_note": "fused-path HBM floor (points read once + canvas written once) / ms_per_step / peak",
            "serial_step_frac": by_floor / (ms_serial / args.steps) / 1e6 / peak}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            tj = json.load(open(prof))
            roof["traffic"] = tj.get(k3_key + "_dram_bytes_per_launch", tj.get("K3_scatter_dram_bytes_per_launch"))
        except Exception:  # noqa: BLE001
            pass

    # ---- CPU baseline: oracle port on this box's host cores, bounded samples at 1 / 6 (the reference's
    # OMP_NUM_THREADS, train_mask_bev.py:14) / all threads ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build_c_oracle()
        cores = os.cpu_count() or 1
        by_threads = {}
        for th, nf, bud in ((1, 1, 4.0), (min(6, cores), 1, 4.0), (cores, 2, 10.0)):
            by_threads[str(th)] = time_cpu_threads(frames, kwargs, pfn_cpu, th, min(nf, B), bud)
        allc = by_threads[str(cores)]
        cpu = {"value": allc["value"], "unit": UNIT, "cores": cores, "kind": "port", "by_threads": by_threads,
               "sample": f"{allc['frames']} frames of this workload x {allc['reps']} reps (median) at {cores} threads "
                         f"(1 frame at 1 and 6 threads); serial C voxelizer + dense torch-CPU PFN + scatter"}
    fps = frames_all * args.steps / (ms * 1e-3)
    fps2 = frames_all * args.steps / (ms2 * 1e-3)
    wcfg = workload_config(args.workload, cfg, B)
    if strong:
        wcfg["global_batch"] = frames_all
    line = {"metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "serial_ms_per_step": ms_serial / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": wcfg,
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": fps2, "unit": UNIT, "h2d_bytes_per_step": int(host.numel() * 4),
                    "d2h_bytes_per_step": int(counts_host.numel() * 4),
                    "serial_value": frames_all * args.steps / (ms2s * 1e-3),
                    "what": "mbev_encode_batch_pipelined with HOST points, every step: pinned host points -> H2D + K1 on "
                            "a prep stream -> K2 on a PFN stream -> K3 on the caller's stream (two buffer sets: three "
                            "batches in flight) -> canvas in HBM, where the reference's consumer (LayerNorm / Swin) "
                            "reads it + D2H of the per-frame pillar counts (the canvas itself, 5.2 GB per step, is NOT "
                            "copied back); serial_value = same through mbev_encode_batch_host on one stream"},
            "roofline": roof, "kernels": kernels, "layernorm_f1": layernorm, "patch_embed_f2": patch_embed, "bf16_canvas": bf16, "cpu_baseline": cpu,
            "train": train, "pillars_per_step": P, "kept_points_per_step": nk, "points_per_step": N}
    if saved_stdout is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)  # teardown chatter, if any, stays off stdout as well
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="kitti_b16")
    ap.add_argument("--batch", type=int, default=None, help="frames per GPU per step (default: the workload's own)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: ONE batch of the workload's size split over the ranks (BASELINE configs[2])")
    ap.add_argument("--serial", action="store_true", help="value / e2e on one stream (no three-stage pipeline)")
    ap.add_argument("--scatter-ctas", type=int, default=0,
                    help="K3 of the pipelined step: CTAs per SM of the TMA-engine scatter (0: the register scatter)")
    ap.add_argument("--train", action="store_true", help="also time the encoder's training step (always on when N > 1)")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--train-batch", type=int, default=4, help="frames per GPU of the training step (semantic_kitti/01:28)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-layernorm", action="store_true", help="skip the K3+LayerNorm (SURVEY f1) timing")
    ap.add_argument("--no-patch-embed", action="store_true", help="skip the pillar patch embedding (SURVEY f2) timing")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

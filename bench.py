#!/usr/bin/env python
"""bench.py — BEV frames/s of the point-cloud -> BEV front end (voxelize + PFN + scatter) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path over one batch of synthetic frames (BASELINE.json configs[1]: KITTI-shaped
frames, batch 16, x 0..80 / y +-40 @ 0.1 m -> 800x800 canvas, [128,128,128] PFN, fp32 forward, eval-mode BN).
N > 1: one process per GPU (torchrun), frames shard by rank (weak scaling: every rank runs its own batch of 16),
no data-path collective; barrier + synchronize on both sides, device time via CUDA events, max over ranks.

Keys beyond the base contract: `roofline` (dominant kernel, algorithmic bytes / event time vs MEASURED_PEAKS.json),
`kernels` (per-kernel breakdown), `cpu_baseline` (oracle port timed on this box's host cores), `e2e` (same metric
through the C-ABI host entry: pinned host points -> H2D -> K1,K2,K3 -> D2H of the per-frame pillar counts).
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
    sample_frames = 2
    cfg, kwargs, frames = build_workload(args.workload, 0, batch=sample_frames)
    _, pfn = make_encoder_cpu_only(kwargs)
    for _ in range(max(args.warmup, 1)):
        cpu_path_once(frames[:1], kwargs, pfn)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_path_once(frames, kwargs, pfn)
    dt = time.perf_counter() - t0
    fps = sample_frames * args.steps / dt
    sample = (f"{sample_frames} frames of the {cfg['batch']}-frame batch per step; serial C voxelizer (mmcv's CPU kernel "
              f"is serial) + dense torch-CPU PFN over all P*T slots + scatter, torch threads={cores}")
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, cfg, sample_frames),
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
    cfg, kwargs, frames = build_workload(args.workload, rank, batch=args.batch)
    B = len(frames)
    enc, pfn_cpu = make_encoder(kwargs, dev)
    runner = FusedEncoderRunner(enc, [len(f) for f in frames], dev, overlap=args.overlap)
    host = torch.from_numpy(np.concatenate(frames, 0)).pin_memory()
    runner.points_dev.copy_(host)
    counts_host = torch.empty((B + 1,), dtype=torch.int32).pin_memory()
    sync = lambda: torch.cuda.synchronize(dev)  # noqa: E731

    def barrier():
        if world > 1:
            dist.barrier()
        sync()

    # ---- device-resident throughput (`value`): a stream of batches through mbev_encode_batch_pipelined (K1 of
    # step i+1 on a prep stream under K2 / K3 of step i); `serial_ms_per_step` = mbev_encode_batch, one stream ----
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
    # step; pipelined = the H2D copy of step i+1 overlaps the kernels of step i (two device point buffers) ----
    host2 = host.clone().pin_memory()  # alternate two host batches so that no step can reuse a stale device copy

    def e2e_loop(fn):
        for i in range(max(1, args.warmup // 2)):
            fn(host2 if i & 1 else host)
        barrier()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        for i in range(args.steps):
            fn(host2 if i & 1 else host)
            counts_host.copy_(runner.pillar_base, non_blocking=True)
        e2.record()
        barrier()
        return s2.elapsed_time(e2)

    ms2s = e2e_loop(runner.run_host)
    if args.serial:
        ms2 = e2e_loop(runner.run_host_pipelined)
    else:
        def e2e_pipe(h):
            runner.run_pipelined(h)
            runner.pillar_base = runner.last_pillar_base  # the D2H below reads the counts of the batch just enqueued
        base0 = runner.pillar_base
        ms2 = e2e_loop(e2e_pipe)
        runner.pillar_base = base0
    t = torch.tensor([ms, ms2, ms2s, ms_serial], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms2, ms2s, ms_serial = float(t[0]), float(t[1]), float(t[2]), float(t[3])
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
    fused_ok = bool(runner.lib.mbev_pfn_scatter_supported(ctypes.byref(runner.params), T, B, runner.ny, runner.nx,
                                                          ctypes.c_void_p(runner.canvas.data_ptr())))
    t_fused = ev_time(runner.run_pfn_scatter, iters, sync) if fused_ok else None
    split_ok = bool(runner.lib.mbev_scatter_split_supported(runner.ny, runner.nx, ctypes.c_void_p(runner.canvas.data_ptr())))
    t_sc2 = ev_time(runner.run_scatter_split, iters, sync) if split_ok else None
    t_fill = ev_time(runner.run_fill_empty, iters, sync) if split_ok else None
    t_occ = ev_time(runner.run_scatter_occupied, iters, sync) if split_ok else None
    # algorithmic bytes per launch (per frame x frames in the batch; DESIGN.md "Roofline accounting")
    by_vox = N * C * 4 + N * 4 + P * 20
    by_pfn = nk * (C * 4 + 4) + P * 20 + P * Co * 4
    by_sc = P * Co * 4 + P * 16 + B * G * Co * 4
    units = [l.units for l in enc._voxel_encoder.pfn_layers]
    ins = [l.linear.in_features for l in enc._voxel_encoder.pfn_layers]
    mac_row = sum(i * u for i, u in zip(ins, units))
    fl_pfn = 2.0 * mac_row * (nk + P)
    kernels = {
        "K1_voxelize": {"ms": t_vox, "alg_bytes": by_vox, "gbs": by_vox / t_vox / 1e6, "frac_hbm": by_vox / t_vox / 1e6 / peak},
        "K2_pfn": {"ms": t_pfn, "alg_bytes": by_pfn, "gbs": by_pfn / t_pfn / 1e6, "frac_hbm": by_pfn / t_pfn / 1e6 / peak,
                   "alg_tflops_upstream_equiv": fl_pfn / t_pfn / 1e9, "fp32_fma_peak_tflops": 74.4},
        "K3_scatter": {"ms": t_sc, "alg_bytes": by_sc, "gbs": by_sc / t_sc / 1e6, "frac_hbm": by_sc / t_sc / 1e6 / peak},
    }
    if t_sc2 is not None:
        kernels["K3ab_fill_empty+scatter_occupied"] = {"ms": t_sc2, "alg_bytes": by_sc, "gbs": by_sc / t_sc2 / 1e6,
                                                       "frac_hbm": by_sc / t_sc2 / 1e6 / peak,
                                                       "note": "two-kernel form of K3 used by the fused path; timed back to "
                                                               "back on one stream here, K3a overlaps K2 in the step",
                                                       "K3a_fill_empty_ms": t_fill, "K3b_scatter_occupied_ms": t_occ}
    if t_fused is not None:
        kernels["K2+K3_fused_kernel"] = {"ms": t_fused, "alg_bytes": by_pfn + by_sc - 2 * P * Co * 4,
                                         "gbs": (by_pfn + by_sc - 2 * P * Co * 4) / t_fused / 1e6,
                                         "frac_hbm": (by_pfn + by_sc - 2 * P * Co * 4) / t_fused / 1e6 / peak,
                                         "default": bool(runner.lib.mbev_pfn_scatter_default()),
                                         "note": "single kernel (PFN in cell order + canvas writer warps); opt-in with "
                                                 "MBEV_FUSED_CANVAS=1"}
    # BASELINE config 4 flavour (not the headline, which is fp32): same path with a bfloat16 canvas
    bf16 = None
    if (G & 3) == 0:
        t_sc16 = ev_time(runner.run_scatter_bf16, iters, sync)
        t_step16 = ev_time(runner.run_device_bf16, iters, sync)
        by_sc16 = P * Co * 4 + P * 16 + B * G * Co * 2
        bf16 = {"K3_scatter_bf16_ms": t_sc16, "alg_bytes": by_sc16, "gbs": by_sc16 / t_sc16 / 1e6,
                "frac_hbm": by_sc16 / t_sc16 / 1e6 / peak, "ms_per_step": t_step16,
                "frames_per_s": B / (t_step16 * 1e-3),
                "note": "fp32 PFN (3xTF32), canvas rounded to bf16 on the way out (mbev_scatter_forward_bf16); one GPU"}
        runner.canvas_bf16 = None
    # SURVEY §8 f1 ("next" row, not part of the headline metric): the LayerNorm that follows the scatter in
    # MaskBevEncoder.forward, fused into the scatter, next to torch's own LayerNorm on the finished canvas
    layernorm = None
    if not args.no_layernorm:
        from mask_bev_b200 import functional as F_
        ln = enc._layer_norm
        out_ln = torch.empty_like(runner.canvas)

        def fused_ln():
            return F_.scatter_layernorm_forward(runner.feats, runner.cell_table, runner.pillar_base, B, runner.ny,
                                                runner.nx, ln.weight, ln.bias, ln.eps, out=out_ln)
        if fused_ln() is not None:
            t_ln = ev_time(fused_ln, iters, sync)
            with torch.no_grad():
                t_torch = ev_time(lambda: ln(runner.canvas), 2, sync)
            by_ln = B * G * Co * 4 + 2 * G * Co * 4 + P * Co * 4 + B * G * 4
            layernorm = {"K3+LN_fused_ms": t_ln, "alg_bytes": by_ln, "gbs": by_ln / t_ln / 1e6,
                         "frac_hbm": by_ln / t_ln / 1e6 / peak, "K3_then_torch_layernorm_ms": t_sc + t_torch,
                         "torch_layernorm_ms": t_torch,
                         "note": "mbev_scatter_layernorm_forward vs K3 followed by nn.LayerNorm([C,ny,nx]) (mask_bev_encoders.py:75,92)"}
            # its backward (training): out_ln stands in for the incoming gradient (any dense tensor of that shape)
            if F_.scatter_layernorm_backward_supported(B, Co, runner.ny, runner.nx):
                stats_ln = fused_ln()[1]

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
    dom = max(("K1_voxelize", "K2_pfn", "K3_scatter"), key=lambda k: kernels[k]["ms"])
    roof = {"kernel": "K3_scatter (k_scatter_run<2>)", "bound": "hbm", "achieved": kernels["K3_scatter"]["gbs"], "peak": peak,
            "unit": "GB/s", "frac": kernels["K3_scatter"]["frac_hbm"], "traffic": None, "peak_source": peak_src,
            "launch_ms": t_sc, "dominant_by_time": dom,
            "share_of_step": t_sc / (t_vox + t_pfn + t_sc)}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            roof["traffic"] = json.load(open(prof)).get("K3_scatter_dram_bytes_per_launch")
        except Exception:  # noqa: BLE001
            pass

    # ---- CPU baseline: oracle port on this box's host cores, bounded sample ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build_c_oracle()
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        nfr = 2
        tcpu, reps = time_cpu(frames[:nfr], kwargs, pfn_cpu, budget_s=15.0)
        cpu = {"value": nfr / tcpu, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{nfr} frames of this workload x {reps} reps (median); serial C voxelizer + dense torch-CPU PFN "
                         f"(threads={cores}) + scatter"}
    fps = world * B * args.steps / (ms * 1e-3)
    fps2 = world * B * args.steps / (ms2 * 1e-3)
    line = {"metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "serial_ms_per_step": ms_serial / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args.workload, cfg, B),
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": fps2, "unit": UNIT, "h2d_bytes_per_step": int(host.numel() * 4),
                    "d2h_bytes_per_step": int(counts_host.numel() * 4),
                    "serial_value": world * B * args.steps / (ms2s * 1e-3),
                    "what": "mbev_encode_batch_pipelined, every step: pinned host points -> H2D + K1 on a prep stream "
                            "(two buffer sets: overlaps K2 / K3 of the previous step) -> K2,K3 -> canvas in HBM "
                            "(where the reference's consumer reads it) + D2H of per-frame pillar counts; serial_value "
                            "= same through mbev_encode_batch_host (copy and kernels on one stream)"},
            "roofline": roof, "kernels": kernels, "layernorm_f1": layernorm, "bf16_canvas": bf16, "cpu_baseline": cpu,
            "pillars_per_step": P, "kept_points_per_step": nk, "points_per_step": N}
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
    ap.add_argument("--serial", action="store_true", help="value / e2e without the K1-under-K3 pipeline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-layernorm", action="store_true", help="skip the K3+LayerNorm (SURVEY f1) timing")
    ap.add_argument("--overlap", action="store_true", help="two streams: K3a zero-fill under K2, then K3b (default: one stream, one-pass K3)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

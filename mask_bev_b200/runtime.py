"""Serving-style runner for the fused path: every buffer preallocated once, one C-ABI call per batch
(``mbev_encode_batch`` / ``mbev_encode_batch_host`` on one stream, ``mbev_encode_batch_pipelined`` for a stream of
batches), no host synchronisation, no allocator traffic. This is what bench.py times; ``MaskBevEncoder.forward`` is the
autograd-aware equivalent.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from . import _lib
from . import functional as F_
from ._lib import check, ptr


class FusedEncoderRunner:
    """`scatter_ctas_per_sm`: K3 of the pipelined entry — 0 (default) = the register scatter on a machine-filling grid;
    N >= 1 = the small-footprint scatter (k_scatter_bulk) with N CTAs per SM. Measured on B200 (kitti_b16,
    profiles/r2_corun_probes.txt): 0 is the fastest — a co-resident K2 and K3 slow each other down by more than the
    overlap gains."""

    def __init__(self, encoder, frame_sizes: Sequence[int], device: torch.device, scatter_ctas_per_sm: int = 0):
        self.enc = encoder
        self.device = torch.device(device)
        self.lib = _lib.load()
        self.sizes = [int(s) for s in frame_sizes]
        self.B = len(self.sizes)
        C = encoder._voxel_encoder.raw_in_channels
        self.C = C
        self.geo = encoder._voxel_layer._geometry(C, strict_filter=True)
        self.off, self.total = F_._offsets(self.sizes)
        self.cap = F_.pillar_capacity(self.geo, self.sizes)
        self.ny, self.nx = encoder._num_voxel_y, encoder._num_voxel_x
        self.c_out = encoder._out_features
        self.scatter_ctas_per_sm = int(scatter_ctas_per_sm)
        dev = self.device
        T = self.geo.max_points
        cells = self.ny * self.nx
        with torch.cuda.device(dev):
            self.points_dev = torch.empty((self.total, C), dtype=torch.float32, device=dev)
            self.cell_table = torch.empty((self.B, cells), dtype=torch.int32, device=dev)
            self.coors = torch.empty((self.cap, 4), dtype=torch.int32, device=dev)
            self.num_points = torch.empty((self.cap,), dtype=torch.int32, device=dev)
            self.kept_idx = torch.empty((self.cap, T), dtype=torch.int32, device=dev)
            self.pillar_base = torch.zeros((self.B + 1,), dtype=torch.int32, device=dev)
            self.feats = torch.empty((self.cap, self.c_out), dtype=torch.float32, device=dev)
            self.canvas = torch.empty((self.B, self.c_out, self.ny, self.nx), dtype=torch.float32, device=dev)
            self.refresh_params()
            nbytes = ctypes.c_size_t()
            check(self.lib.mbev_encode_batch_workspace_bytes(ctypes.byref(self.geo), ctypes.byref(self.params), self.B,
                                                             self.total, self.cap, ctypes.byref(nbytes)),
                  "encode_batch_workspace_bytes")
            self.ws = torch.empty(max(nbytes.value, 16), dtype=torch.uint8, device=dev)
        self._sets = None
        self._points_dirty = True

    def refresh_params(self) -> None:
        """Re-fold the eval-mode BatchNorm after a weight update."""
        net = self.enc._voxel_encoder
        self._weights = [F_._f32c(l.linear.weight) for l in net.pfn_layers]
        self._scales, self._shifts, self._ss, _ = net._folded()
        self.params = F_._pfn_struct(net._config(), self._weights, self._scales, self._shifts)

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def run_device(self, points: Optional[torch.Tensor] = None) -> torch.Tensor:
        """points: (sum N_i, C) float32 on the device (defaults to the runner's own resident copy)."""
        p = self.points_dev if points is None else points
        with torch.cuda.device(self.device):
            check(self.lib.mbev_encode_batch(ptr(p), self.off, self.B, ctypes.byref(self.geo), ctypes.byref(self.params),
                                             ptr(self.cell_table), ptr(self.coors), ptr(self.num_points),
                                             ptr(self.kept_idx), ptr(self.pillar_base), self.cap, ptr(self.feats),
                                             ptr(self.canvas), ptr(self.ws), self.ws.numel(), self._stream()),
                  "encode_batch")
        return self.canvas

    def run_host(self, points_host: torch.Tensor) -> torch.Tensor:
        """points_host: (sum N_i, C) float32 HOST tensor (pinned for an asynchronous copy)."""
        with torch.cuda.device(self.device):
            check(self.lib.mbev_encode_batch_host(ptr(points_host), ptr(self.points_dev), self.off, self.B,
                                                  ctypes.byref(self.geo), ctypes.byref(self.params),
                                                  ptr(self.cell_table), ptr(self.coors), ptr(self.num_points),
                                                  ptr(self.kept_idx), ptr(self.pillar_base), self.cap, ptr(self.feats),
                                                  ptr(self.canvas), ptr(self.ws), self.ws.numel(), self._stream()),
                  "encode_batch_host")
        return self.canvas

    def _pipe_init(self):
        dev = self.device
        nb, pb = ctypes.c_size_t(), ctypes.c_size_t()
        check(self.lib.mbev_voxelize_workspace_bytes(ctypes.byref(self.geo), self.B, self.total, ctypes.byref(nb)), "ws")
        check(self.lib.mbev_pfn_workspace_bytes(ctypes.byref(self.params), self.geo.max_points, self.cap, 0,
                                                ctypes.byref(pb)), "pfn ws")
        self._sets = []
        with torch.cuda.device(dev):
            for k in range(2):
                ev = []
                for _ in range(3):  # ready (K1 done), feats (K2 done), consumed (K3 done)
                    e = ctypes.c_void_p()
                    check(self.lib.mbev_event_create(ctypes.byref(e)), "event_create")
                    ev.append(e)
                first = k == 0
                self._sets.append(dict(
                    points=torch.empty_like(self.points_dev),
                    cell_table=self.cell_table if first else torch.empty_like(self.cell_table),
                    coors=self.coors if first else torch.empty_like(self.coors),
                    num_points=self.num_points if first else torch.empty_like(self.num_points),
                    kept_idx=self.kept_idx if first else torch.empty_like(self.kept_idx),
                    pillar_base=self.pillar_base if first else torch.zeros_like(self.pillar_base),
                    feats=self.feats if first else torch.empty_like(self.feats),
                    vox_ws=torch.empty(max(nb.value, 16), dtype=torch.uint8, device=dev), ev=ev))
            self._pfn_ws = torch.empty(max(pb.value, 16), dtype=torch.uint8, device=dev)
            self._prep = torch.cuda.Stream(device=dev)
            self._pfn = torch.cuda.Stream(device=dev)
            self._resident = torch.cuda.Event()
        self._i = 0
        self._last = self._sets[0]

    def run_pipelined(self, points_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Stream-of-batches form (mbev_encode_batch_pipelined): [H2D +] K1 of this batch on a prep stream, K2 on a
        PFN stream, K3 on the current stream, over two alternating buffer sets — K1 of batch i+2, K2 of batch i+1 and
        K3 of batch i overlap. Without `points_host` the runner's resident `points_dev` is the input: write it with
        `set_points` (or call `mark_points_updated` after writing it on the current stream) and not while batches are
        in flight.
        `last_pillar_base` / `last_feats` belong to the batch just enqueued."""
        if self._sets is None:
            self._pipe_init()
        s = self._sets[self._i & 1]
        self._i += 1
        with torch.cuda.device(self.device):
            if points_host is None and self._points_dirty:
                # writes to points_dev enqueued on the current stream so far happen before K1 reads it. Once per
                # update only: the current stream also carries K3 of the previous batch, and ordering K1 after it on
                # every call would serialise the pipeline.
                self._resident.record(torch.cuda.current_stream(self.device))
                self._prep.wait_event(self._resident)
                self._points_dirty = False
            pts_dev = s["points"] if points_host is not None else self.points_dev
            check(self.lib.mbev_encode_batch_pipelined(
                ptr(points_host), ptr(pts_dev), self.off, self.B, ctypes.byref(self.geo), ctypes.byref(self.params),
                ptr(s["cell_table"]), ptr(s["coors"]), ptr(s["num_points"]), ptr(s["kept_idx"]), ptr(s["pillar_base"]),
                self.cap, ptr(s["feats"]), ptr(self.canvas), ptr(s["vox_ws"]), s["vox_ws"].numel(), ptr(self._pfn_ws),
                self._pfn_ws.numel(), self.scatter_ctas_per_sm, self._stream(),
                ctypes.c_void_p(self._prep.cuda_stream), ctypes.c_void_p(self._pfn.cuda_stream),
                s["ev"][0], s["ev"][1], s["ev"][2]), "encode_batch_pipelined")
        self._last = s
        return self.canvas

    def set_points(self, points: torch.Tensor) -> None:
        """Copy a (sum N_i, C) float32 batch (host or device) into the resident `points_dev` on the current stream."""
        self.points_dev.copy_(points, non_blocking=True)
        self._points_dirty = True

    def mark_points_updated(self) -> None:
        self._points_dirty = True

    @property
    def last_pillar_base(self) -> torch.Tensor:
        return self._last["pillar_base"] if self._sets is not None else self.pillar_base

    @property
    def last_feats(self) -> torch.Tensor:
        return self._last["feats"] if self._sets is not None else self.feats

    def close(self):
        for st in self._sets or []:
            for e in st["ev"]:
                self.lib.mbev_event_destroy(e)
        self._sets = None

    # stage-by-stage entry points (per-kernel timing in bench.py)
    def run_voxelize(self):
        nb = ctypes.c_size_t()
        check(self.lib.mbev_voxelize_workspace_bytes(ctypes.byref(self.geo), self.B, self.total, ctypes.byref(nb)), "ws")
        with torch.cuda.device(self.device):
            check(self.lib.mbev_voxelize(ptr(self.points_dev), self.off, self.B, ctypes.byref(self.geo),
                                         ptr(self.cell_table), ptr(self.coors), ptr(self.num_points), ptr(self.kept_idx),
                                         ptr(self.pillar_base), self.cap, ptr(self.ws), nb.value, self._stream()),
                  "voxelize")

    def run_pfn(self):
        with torch.cuda.device(self.device):
            check(self.lib.mbev_pfn_forward(ptr(self.points_dev), self.C, ptr(self.kept_idx), ptr(self.num_points),
                                            ptr(self.coors), ptr(self.pillar_base[self.B:]), self.cap,
                                            self.geo.max_points, ctypes.byref(self.params), ptr(self.feats), ptr(self.ws),
                                            self.ws.numel(), self._stream()), "pfn_forward")

    def run_scatter(self):
        with torch.cuda.device(self.device):
            check(self.lib.mbev_scatter_forward(ptr(self.feats), ptr(self.cell_table), self.B, self.c_out, self.ny,
                                                self.nx, ptr(self.canvas), self._stream()), "scatter_forward")

    def run_scatter_stream(self, ctas_per_sm: int = 1):
        """K3 through the TMA engine (the pipelined entry's form)."""
        with torch.cuda.device(self.device):
            check(self.lib.mbev_scatter_forward_stream(ptr(self.feats), ptr(self.cell_table), self.B, self.c_out,
                                                       self.ny, self.nx, ptr(self.canvas), int(ctas_per_sm),
                                                       self._stream()), "scatter_forward_stream")

    def run_scatter_nhwc(self):
        """K3 into the channels-last layout (the canvas buffer reinterpreted as (B, ny, nx, C))."""
        with torch.cuda.device(self.device):
            check(self.lib.mbev_scatter_forward_nhwc(ptr(self.feats), ptr(self.cell_table), self.B, self.c_out, self.ny,
                                                     self.nx, ptr(self.canvas), self._stream()), "scatter_forward_nhwc")

    def run_scatter_bf16(self):
        """K3 into a bfloat16 canvas (allocated on first use)."""
        if getattr(self, "canvas_bf16", None) is None:
            self.canvas_bf16 = torch.empty((self.B, self.c_out, self.ny, self.nx), dtype=torch.bfloat16, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.mbev_scatter_forward_bf16(ptr(self.feats), ptr(self.cell_table), self.B, self.c_out, self.ny,
                                                     self.nx, ptr(self.canvas_bf16), self._stream()),
                  "scatter_forward_bf16")

    def run_device_bf16(self):
        """K1 -> K2 -> K3(bf16 canvas) on the resident points: three C-ABI calls, one stream, no host sync."""
        self.run_voxelize()
        self.run_pfn()
        self.run_scatter_bf16()
        return self.canvas_bf16

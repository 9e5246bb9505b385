"""Serving-style runner for the fused path: every buffer preallocated once, one C-ABI call per batch
(``mbev_encode_batch`` for device-resident points, ``mbev_encode_batch_host`` for pinned host points), no host
synchronisation, no allocator traffic. This is what bench.py times; ``MaskBevEncoder.forward`` is the
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
    def __init__(self, encoder, frame_sizes: Sequence[int], device: torch.device, overlap: bool = False):
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
        dev = self.device
        T = self.geo.max_points
        cells = self.ny * self.nx
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
        # overlap=True: second stream for the zero-fill of the canvas (K3a runs under K2, K3b after it). Measured on
        # B200 (kitti_b16): 2.35 ms/step against 2.26 ms for the single-stream one-pass scatter — the streaming stores
        # back up the LSU that K2's shuffles and shared-memory traffic also use — so the default stays single-stream.
        self.aux = torch.cuda.Stream(device=dev) if overlap else None

    def _aux(self):
        return ctypes.c_void_p(self.aux.cuda_stream) if self.aux is not None else ctypes.c_void_p(None)

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
        check(self.lib.mbev_encode_batch(ptr(p), self.off, self.B, ctypes.byref(self.geo), ctypes.byref(self.params),
                                         ptr(self.cell_table), ptr(self.coors), ptr(self.num_points),
                                         ptr(self.kept_idx), ptr(self.pillar_base), self.cap, ptr(self.feats),
                                         ptr(self.canvas), ptr(self.ws), self.ws.numel(), self._stream(), self._aux()),
              "encode_batch")
        return self.canvas

    def run_host(self, points_host: torch.Tensor) -> torch.Tensor:
        """points_host: (sum N_i, C) float32 HOST tensor (pinned for an asynchronous copy)."""
        check(self.lib.mbev_encode_batch_host(ptr(points_host), ptr(self.points_dev), self.off, self.B,
                                              ctypes.byref(self.geo), ctypes.byref(self.params),
                                              ptr(self.cell_table), ptr(self.coors), ptr(self.num_points),
                                              ptr(self.kept_idx), ptr(self.pillar_base), self.cap, ptr(self.feats),
                                              ptr(self.canvas), ptr(self.ws), self.ws.numel(), self._stream(),
                                              self._aux()),
              "encode_batch_host")
        return self.canvas

    def _pipe_init(self):
        self._pipe_points = [self.points_dev, torch.empty_like(self.points_dev)]
        self._pipe_copy = torch.cuda.Stream(device=self.device)
        self._pipe_events = []
        for _ in range(4):  # (copied, consumed) x 2 buffers
            e = ctypes.c_void_p()
            check(self.lib.mbev_event_create(ctypes.byref(e)), "event_create")
            self._pipe_events.append(e)
        self._pipe_i = 0

    def run_host_pipelined(self, points_host: torch.Tensor) -> torch.Tensor:
        """Like run_host for a stream of batches: two device point buffers used alternately, the H2D copy on its
        own stream, so that the copy of batch i+1 overlaps K1..K3 of batch i (mbev_encode_batch_host_async)."""
        if getattr(self, "_pipe_points", None) is None:
            self._pipe_init()
        k = self._pipe_i & 1
        self._pipe_i += 1
        check(self.lib.mbev_encode_batch_host_async(ptr(points_host), ptr(self._pipe_points[k]), self.off, self.B,
                                                    ctypes.byref(self.geo), ctypes.byref(self.params),
                                                    ptr(self.cell_table), ptr(self.coors), ptr(self.num_points),
                                                    ptr(self.kept_idx), ptr(self.pillar_base), self.cap,
                                                    ptr(self.feats), ptr(self.canvas), ptr(self.ws), self.ws.numel(),
                                                    self._stream(), self._aux(),
                                                    ctypes.c_void_p(self._pipe_copy.cuda_stream),
                                                    self._pipe_events[2 * k], self._pipe_events[2 * k + 1]),
              "encode_batch_host_async")
        return self.canvas

    def _pipe2_init(self):
        dev, T = self.device, self.geo.max_points
        cells = self.ny * self.nx
        nb = ctypes.c_size_t()
        check(self.lib.mbev_voxelize_workspace_bytes(ctypes.byref(self.geo), self.B, self.total, ctypes.byref(nb)), "ws")
        self._p2_sets = []
        for k in range(2):
            ev = []
            for _ in range(2):
                e = ctypes.c_void_p()
                check(self.lib.mbev_event_create(ctypes.byref(e)), "event_create")
                ev.append(e)
            self._p2_sets.append(dict(
                points=torch.empty_like(self.points_dev),
                cell_table=self.cell_table if k == 0 else torch.empty_like(self.cell_table),
                coors=self.coors if k == 0 else torch.empty_like(self.coors),
                num_points=self.num_points if k == 0 else torch.empty_like(self.num_points),
                kept_idx=self.kept_idx if k == 0 else torch.empty_like(self.kept_idx),
                pillar_base=self.pillar_base if k == 0 else torch.zeros_like(self.pillar_base),
                vox_ws=torch.empty(max(nb.value, 16), dtype=torch.uint8, device=dev), ev=ev))
        self._p2_prep = torch.cuda.Stream(device=dev)
        self._p2_i = 0
        self._p2_last = self._p2_sets[0]

    def run_pipelined(self, points_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Stream-of-batches form (mbev_encode_batch_pipelined): the H2D copy (when `points_host` is given; otherwise
        the runner's resident points) and K1 of this batch run on a prep stream into one of two buffer sets and
        overlap K2 / K3 of the previous batch. `last_pillar_base` is the pillar_base of the batch just enqueued."""
        if getattr(self, "_p2_sets", None) is None:
            self._pipe2_init()
        s = self._p2_sets[self._p2_i & 1]
        self._p2_i += 1
        pts_dev = s["points"] if points_host is not None else self.points_dev
        check(self.lib.mbev_encode_batch_pipelined(ptr(points_host), ptr(pts_dev), self.off, self.B,
                                                   ctypes.byref(self.geo), ctypes.byref(self.params),
                                                   ptr(s["cell_table"]), ptr(s["coors"]), ptr(s["num_points"]),
                                                   ptr(s["kept_idx"]), ptr(s["pillar_base"]), self.cap,
                                                   ptr(self.feats), ptr(self.canvas), ptr(s["vox_ws"]),
                                                   s["vox_ws"].numel(), ptr(self.ws), self.ws.numel(), self._stream(),
                                                   ctypes.c_void_p(self._p2_prep.cuda_stream), s["ev"][0], s["ev"][1]),
              "encode_batch_pipelined")
        self._p2_last = s
        return self.canvas

    @property
    def last_pillar_base(self) -> torch.Tensor:
        return self._p2_last["pillar_base"] if getattr(self, "_p2_sets", None) is not None else self.pillar_base

    def close(self):
        for st in getattr(self, "_p2_sets", None) or []:
            for e in st["ev"]:
                self.lib.mbev_event_destroy(e)
        self._p2_sets = None
        for e in getattr(self, "_pipe_events", []):
            self.lib.mbev_event_destroy(e)
        self._pipe_events = []

    # stage-by-stage entry points (per-kernel timing in bench.py)
    def run_voxelize(self):
        nb = ctypes.c_size_t()
        check(self.lib.mbev_voxelize_workspace_bytes(ctypes.byref(self.geo), self.B, self.total, ctypes.byref(nb)), "ws")
        check(self.lib.mbev_voxelize(ptr(self.points_dev), self.off, self.B, ctypes.byref(self.geo),
                                     ptr(self.cell_table), ptr(self.coors), ptr(self.num_points), ptr(self.kept_idx),
                                     ptr(self.pillar_base), self.cap, ptr(self.ws), nb.value, self._stream()), "voxelize")

    def run_pfn(self):
        check(self.lib.mbev_pfn_forward(ptr(self.points_dev), self.C, ptr(self.kept_idx), ptr(self.num_points),
                                        ptr(self.coors), ptr(self.pillar_base[self.B:]), self.cap, self.geo.max_points,
                                        ctypes.byref(self.params), ptr(self.feats), ptr(self.ws), self.ws.numel(),
                                        self._stream()), "pfn_forward")

    def run_pfn_scatter(self):
        """K2 + K3 as the single fused kernel (after run_voxelize)."""
        nb = ctypes.c_size_t()
        check(self.lib.mbev_pfn_scatter_workspace_bytes(ctypes.byref(self.params), self.geo.max_points, self.cap,
                                                        self.B, self.ny, self.nx, ctypes.byref(nb)), "ws")
        if getattr(self, "_ws_fused", None) is None or self._ws_fused.numel() < nb.value:
            self._ws_fused = torch.empty(max(nb.value, 16), dtype=torch.uint8, device=self.device)
        check(self.lib.mbev_pfn_scatter_forward(ptr(self.points_dev), self.C, ptr(self.kept_idx), ptr(self.num_points),
                                                ptr(self.coors), self.cap, self.geo.max_points,
                                                ctypes.byref(self.params), ptr(self.cell_table), self.B, self.ny,
                                                self.nx, ptr(self.feats), ptr(self.canvas), ptr(self._ws_fused),
                                                self._ws_fused.numel(), self._stream()), "pfn_scatter_forward")

    def run_fill_empty(self):
        check(self.lib.mbev_scatter_fill_empty(ptr(self.cell_table), self.B, self.c_out, self.ny, self.nx,
                                               ptr(self.canvas), self._stream()), "scatter_fill_empty")

    def run_scatter_occupied(self):
        check(self.lib.mbev_scatter_occupied(ptr(self.feats), ptr(self.coors), ptr(self.pillar_base[self.B:]), self.cap,
                                             ptr(self.cell_table), self.B, self.c_out, self.ny, self.nx,
                                             ptr(self.canvas), self._stream()), "scatter_occupied")

    def run_scatter_split(self):
        """K3a (zero-fill of the empty sectors) + K3b (occupied sectors), back to back on one stream."""
        check(self.lib.mbev_scatter_fill_empty(ptr(self.cell_table), self.B, self.c_out, self.ny, self.nx,
                                               ptr(self.canvas), self._stream()), "scatter_fill_empty")
        check(self.lib.mbev_scatter_occupied(ptr(self.feats), ptr(self.coors), ptr(self.pillar_base[self.B:]), self.cap,
                                             ptr(self.cell_table), self.B, self.c_out, self.ny, self.nx,
                                             ptr(self.canvas), self._stream()), "scatter_occupied")

    def run_scatter_bf16(self):
        """K3 into a bfloat16 canvas (allocated on first use)."""
        if getattr(self, "canvas_bf16", None) is None:
            self.canvas_bf16 = torch.empty((self.B, self.c_out, self.ny, self.nx), dtype=torch.bfloat16, device=self.device)
        check(self.lib.mbev_scatter_forward_bf16(ptr(self.feats), ptr(self.cell_table), self.B, self.c_out, self.ny,
                                                 self.nx, ptr(self.canvas_bf16), self._stream()), "scatter_forward_bf16")

    def run_device_bf16(self):
        """K1 -> K2 -> K3(bf16 canvas) on the resident points: three C-ABI calls, one stream, no host sync."""
        self.run_voxelize()
        self.run_pfn()
        self.run_scatter_bf16()
        return self.canvas_bf16

    def run_scatter(self):
        check(self.lib.mbev_scatter_forward(ptr(self.feats), ptr(self.cell_table), self.B, self.c_out, self.ny,
                                            self.nx, ptr(self.canvas), self._stream()), "scatter_forward")

"""`PillarFeatureNet` / `PFNLayer` — drop-in for mmdet3d==1.1.0's pillar encoder as MaskBEV builds and calls it
(/root/reference/mask_bev/models/encoders/mask_bev_encoders.py:70-72, :119-120). Same constructor, forward
signature and state-dict keys (``pfn_layers.{l}.linear.weight``, ``pfn_layers.{l}.norm.{weight,bias,
running_mean,running_var,num_batches_tracked}``) so reference checkpoints load unchanged; the arithmetic runs
in K2 (csrc/pfn.cu, csrc/pfn_bwd.cu). ``nn.Linear`` / ``nn.BatchNorm1d`` objects are parameter holders only —
their forward is never called. There is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn

from . import functional as F_
from ._lib import MAX_UNITS, MbevError


class PFNLayer(nn.Module):
    """Parameter container with upstream's layout: Linear(in, units, bias=False) + BN1d(units)."""

    def __init__(self, in_channels: int, out_channels: int, norm_cfg: Optional[dict] = None,
                 last_layer: bool = False, mode: str = 'max'):
        super().__init__()
        norm_cfg = dict(type='BN1d', eps=1e-3, momentum=0.01) if norm_cfg is None else norm_cfg
        if norm_cfg.get('type', 'BN1d') not in ('BN1d', 'BN'):
            raise MbevError(f"norm {norm_cfg.get('type')} is not on the MaskBEV path (BN1d only)")
        if mode != 'max':
            raise MbevError("PFNLayer mode 'avg' is not on the MaskBEV path (max only)")
        self.name = 'PFNLayer'
        self.last_vfe = last_layer
        if not self.last_vfe:
            out_channels = out_channels // 2
        self.units = out_channels
        self.norm = nn.BatchNorm1d(self.units, eps=norm_cfg.get('eps', 1e-3), momentum=norm_cfg.get('momentum', 0.01))
        self.linear = nn.Linear(in_channels, self.units, bias=False)
        self.mode = mode


class _PfnFunction(torch.autograd.Function):
    """feats = PFN(rows; W, gamma, beta). Gradients flow to the parameters only."""

    @staticmethod
    def forward(ctx, net, rows, kept_idx, num_points, coors, npil_dev, capacity, T, *params):
        L = len(net.pfn_layers)
        weights, gammas, betas = params[:L], params[L:2 * L], params[2 * L:]
        cfg = net._config()
        training = net.training
        if cfg.gemm_path == 3 and (training or not isinstance(ctx, _NullCtx)):
            raise MbevError("gemm_path='tcgen05_bf16' is the inference path (eval mode, no autograd): batch statistics and "
                            "the backward need the fp32-accurate kernels")
        if not isinstance(ctx, _NullCtx) and cfg.gemm_path == 0 and not net.autograd_tensor_cores:
            # opt-out: forward on the fp32 FMA pipe, bit-consistent with what K2' recomputes (round 1's default).
            cfg.gemm_path = 1
        # compact row bound without a host sync: every real row is a stored point (<= rows.shape[0] of them) and
        # there is at most one virtual row per pillar
        rows_capacity = int(min(capacity * (T + 1), rows.shape[0] + capacity))
        ctx.saved_rows = None
        if training and not isinstance(ctx, _NullCtx) and net.train_rows:
            # a training step: the forward runs ONCE, in row space, and keeps its rows for the backward
            feats, scale_shift, batch_stats, ctx.saved_rows = F_.pfn_forward_train_rows(
                rows, kept_idx, num_points, coors, npil_dev, capacity, T, cfg, weights, gammas, betas, rows_capacity)
            net._update_running_stats(batch_stats, npil_dev, T)
        elif training:
            feats, scale_shift, batch_stats = F_.pfn_forward_train(rows, kept_idx, num_points, coors, npil_dev,
                                                                  capacity, T, cfg, weights, gammas, betas)
            net._update_running_stats(batch_stats, npil_dev, T)
        else:
            scales, shifts, scale_shift, batch_stats = net._folded()
            feats = F_.pfn_forward_eval(rows, kept_idx, num_points, coors, npil_dev, capacity, T, cfg, weights,
                                        scales, shifts)
        ctx.net, ctx.cfg, ctx.training, ctx.capacity, ctx.T = net, cfg, training, capacity, T
        ctx.save_for_backward(rows, kept_idx if kept_idx is not None else torch.empty(0, device=rows.device),
                              num_points, coors, npil_dev, scale_shift, batch_stats, *weights)
        ctx.has_kept = kept_idx is not None
        ctx.rows_capacity = rows_capacity
        return feats

    @staticmethod
    def backward(ctx, dfeats):
        saved = ctx.saved_tensors
        rows, kept_idx, num_points, coors, npil_dev, scale_shift, batch_stats = saved[:7]
        L = len(ctx.cfg.units)
        weights = saved[7:7 + L]
        if ctx.saved_rows is not None:
            dws, dgs, dbs = F_.pfn_backward_rows(ctx.saved_rows, npil_dev, ctx.capacity, ctx.T, ctx.cfg, weights, dfeats,
                                                 ctx.rows_capacity)
            ctx.saved_rows = None  # the rows (GBs at full batch sizes) are released with the step
            return (None,) * 8 + tuple(dws) + tuple(dgs) + tuple(dbs)
        dws, dgs, dbs = F_.pfn_backward(rows, kept_idx if ctx.has_kept else None, num_points, coors, npil_dev,
                                        ctx.capacity, ctx.T, ctx.cfg, weights, scale_shift, batch_stats,
                                        ctx.training, dfeats, ctx.rows_capacity)
        return (None,) * 8 + tuple(dws) + tuple(dgs) + tuple(dbs)


class PillarFeatureNet(nn.Module):
    def __init__(self, in_channels: int = 4, feat_channels: tuple = (64,), with_distance: bool = False,
                 with_cluster_center: bool = True, with_voxel_center: bool = True,
                 voxel_size: Tuple[float] = (0.2, 0.2, 4),
                 point_cloud_range: Tuple[float] = (0, -40, -3, 70.4, 40, 1),
                 norm_cfg: Optional[dict] = None, mode: str = 'max', legacy: bool = True,
                 voxel_center_dims: int = 3):
        super().__init__()
        assert len(feat_channels) > 0
        self.legacy = legacy
        self.raw_in_channels = in_channels
        if with_cluster_center:
            in_channels += 3
        if with_voxel_center:
            in_channels += voxel_center_dims
        if with_distance:
            in_channels += 1
        self._with_distance = with_distance
        self._with_cluster_center = with_cluster_center
        self._with_voxel_center = with_voxel_center
        self._voxel_center_dims = voxel_center_dims
        self.fp16_enabled = False
        self.in_channels = in_channels
        chans = [in_channels] + list(feat_channels)
        self.pfn_layers = nn.ModuleList(
            [PFNLayer(chans[i], chans[i + 1], norm_cfg=norm_cfg, last_layer=(i >= len(chans) - 2), mode=mode)
             for i in range(len(chans) - 1)])
        for layer in self.pfn_layers:
            if layer.units > MAX_UNITS or layer.units % 4:
                raise MbevError(f"PFNLayer.units={layer.units}: the fused kernel needs units % 4 == 0 and <= {MAX_UNITS}")
        self.vx, self.vy, self.vz = voxel_size[0], voxel_size[1], voxel_size[2]
        self.x_offset = self.vx / 2 + point_cloud_range[0]
        self.y_offset = self.vy / 2 + point_cloud_range[1]
        self.z_offset = self.vz / 2 + point_cloud_range[2]
        self.point_cloud_range = point_cloud_range
        # forward Linear layers: 'auto' (tcgen05 3xTF32 when the stack fits, else fp32 FMA), 'fma', 'tcgen05',
        # 'tcgen05_bf16' (inference only: layers >= 1 as single-pass bf16 MMAs, 1e-2 tolerance class)
        self.gemm_path = "auto"
        # forward under autograd on the tensor cores as well. K2' recomputes the activations on the fp32 FMA pipe and, in
        # train mode, takes the BatchNorm batch statistics from ITS OWN rows (k_row_stats), so the backward stays
        # self-consistent although the two passes differ by ~1e-6 (with the forward's statistics reused, round 1 measured
        # train-mode gradients at 2x torch's own fp32 error). False: FMA forward, bit-consistent with the recompute.
        self.autograd_tensor_cores = True
        # train mode under autograd: run the forward once in compact row space (fp32 FMA rows — the ones BatchNorm's
        # backward has to differentiate) and keep the rows for the backward, instead of the tensor-core forward with its L
        # statistics passes followed by K2''s recompute of the same rows. False: the recompute pair (no memory held
        # between forward and backward).
        self.train_rows = True

    # -- helpers ------------------------------------------------------------------------------------
    def _config(self) -> F_.PfnConfig:
        return F_.PfnConfig(
            in_channels=self.raw_in_channels, units=[l.units for l in self.pfn_layers],
            in_dims=[l.linear.in_features for l in self.pfn_layers],
            with_cluster_center=self._with_cluster_center, with_voxel_center=self._with_voxel_center,
            with_distance=self._with_distance, legacy=self.legacy, voxel_center_dims=self._voxel_center_dims,
            vx=self.vx, vy=self.vy, vz=self.vz, x_offset=self.x_offset, y_offset=self.y_offset,
            z_offset=self.z_offset, eps=self.pfn_layers[0].norm.eps,
            gemm_path={"auto": 0, "fma": 1, "tcgen05": 2, "tcgen05_bf16": 3}[self.gemm_path])

    def _param_list(self):
        ls = self.pfn_layers
        return [l.linear.weight for l in ls] + [l.norm.weight for l in ls] + [l.norm.bias for l in ls]

    @torch.no_grad()
    def _folded(self):
        """Eval-mode BatchNorm folded to y*scale + shift (tiny elementwise prep on the device)."""
        L = len(self.pfn_layers)
        dev = self.pfn_layers[0].linear.weight.device
        scale_shift = torch.zeros((L, 2, MAX_UNITS), dtype=torch.float32, device=dev)
        batch_stats = torch.zeros((L, 2, MAX_UNITS), dtype=torch.float32, device=dev)
        scales, shifts = [], []
        for l, layer in enumerate(self.pfn_layers):
            bn, U = layer.norm, layer.units
            sc = bn.weight.float() * torch.rsqrt(bn.running_var.float() + bn.eps)
            sh = bn.bias.float() - bn.running_mean.float() * sc
            scale_shift[l, 0, :U], scale_shift[l, 1, :U] = sc, sh
            batch_stats[l, 0, :U], batch_stats[l, 1, :U] = bn.running_mean.float(), bn.running_var.float()
            scales.append(scale_shift[l, 0, :U])
            shifts.append(scale_shift[l, 1, :U])
        return scales, shifts, scale_shift, batch_stats

    @torch.no_grad()
    def _update_running_stats(self, batch_stats: torch.Tensor, npil_dev: torch.Tensor, T: int) -> None:
        """BatchNorm1d bookkeeping: momentum update with the unbiased variance, M = P*T slots. No host sync. A step
        without any pillar (every point filtered out) leaves the buffers untouched (nn.BatchNorm1d raises on an
        empty input; silently decaying the statistics towards zero would be worse than either)."""
        M = npil_dev.to(torch.float32).reshape(()) * float(T)   # (npil_dev is a 1-element tensor: 0-dim from here on)
        live = (M > 0).to(torch.float32)
        unbias = M / torch.clamp(M - 1.0, min=1.0)
        bns = [(l, layer.norm, layer.units) for l, layer in enumerate(self.pfn_layers) if layer.norm.track_running_stats]
        if not bns:
            return
        moms = {bn.momentum for _, bn, _ in bns}
        if len(moms) == 1 and None not in moms and all(bn.running_mean.dtype == torch.float32 for _, bn, _ in bns):
            # one momentum for all layers (the configured case): the same arithmetic, running = running * (1 - mom) +
            # new * mom, as a handful of multi-tensor launches instead of ~14 small ones per layer
            mom = live * moms.pop()
            new = batch_stats * mom                      # (L, 2, MAX_UNITS): mean * mom, var * mom
            new[:, 1] *= unbias                          # unbiased variance
            running = [t for _, bn, _ in bns for t in (bn.running_mean, bn.running_var)]
            update = [new[l, j, :U] for l, _, U in bns for j in (0, 1)]
            torch._foreach_mul_(running, 1.0 - mom)
            torch._foreach_add_(running, update)
            torch._foreach_add_([bn.num_batches_tracked for _, bn, _ in bns],
                                live.to(bns[0][1].num_batches_tracked.dtype).reshape(()))
            return
        for l, bn, U in bns:
            bn.num_batches_tracked += live.to(bn.num_batches_tracked.dtype).reshape(())
            if bn.momentum is not None:
                mom = live * bn.momentum
            else:
                mom = live / torch.clamp(bn.num_batches_tracked.to(torch.float32), min=1.0)
            bn.running_mean.mul_(1.0 - mom).add_(batch_stats[l, 0, :U].to(bn.running_mean.dtype) * mom)
            bn.running_var.mul_(1.0 - mom).add_((batch_stats[l, 1, :U] * unbias).to(bn.running_var.dtype) * mom)

    def apply_rows(self, rows, kept_idx, num_points, coors, npil_dev, capacity: int, T: int) -> torch.Tensor:
        """Shared entry for the dense (module-level) and sparse (fused) row sources."""
        params = self._param_list()
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if needs_grad:
            return _PfnFunction.apply(self, rows, kept_idx, num_points, coors, npil_dev, capacity, T, *params)
        with torch.no_grad():
            return _PfnFunction.forward(_NullCtx(), self, rows, kept_idx, num_points, coors, npil_dev, capacity, T,
                                        *params)

    # -- upstream forward -----------------------------------------------------------------------------
    def forward(self, features: torch.Tensor, num_points: torch.Tensor, coors: torch.Tensor, *args, **kwargs):
        """features (P, T, C) float32 zero-padded, num_points (P,), coors (P, 4) int (b, z, y, x) -> (P, C_out).
        Upstream's legacy mode also overwrites features[:, :, :3] in place; nobody reads it afterwards
        (mask_bev_encoders.py:90-91) and it is not reproduced."""
        if features.dtype != torch.float32:
            raise MbevError(f"features must be float32, got {features.dtype}")
        P, T, C = features.shape
        if C != self.raw_in_channels:
            raise MbevError(f"features have {C} channels, PillarFeatureNet was built for {self.raw_in_channels}")
        dev = features.device
        if P == 0:
            return features.new_zeros((0, self.pfn_layers[-1].units))
        rows = features.contiguous().view(P * T, C)
        npil = torch.full((1,), P, dtype=torch.int32, device=dev)
        feats = self.apply_rows(rows, None, num_points.to(torch.int32).contiguous(),
                                coors.to(torch.int32).contiguous(), npil, P, T)
        return feats


class _NullCtx:
    def save_for_backward(self, *a):
        pass

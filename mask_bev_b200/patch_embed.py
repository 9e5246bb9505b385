"""Swin patch embedding that consumes pillars.

``PillarPatchEmbed`` mirrors the constructor arguments and state-dict keys of mmdet's ``PatchEmbed`` as the reference
builds it (/root/reference/mask_bev/models/networks/swin/swin.py:578-586: ``conv_type='Conv2d'``, ``kernel_size =
stride = patch_size``, ``padding='corner'``, ``norm_cfg=dict(type='LN')`` when ``patch_norm``) — ``projection.weight``,
``projection.bias``, ``norm.weight``, ``norm.bias`` load unchanged. Instead of the dense ``forward(x)`` it offers
``forward_pillars``: the encoder's pillar features go straight to the (B, Hp*Wp, E) tokens, with the encoder's
``nn.LayerNorm([C, ny, nx])`` (mask_bev_encoders.py:75, 92) folded in algebraically; the 5 GB pseudo image is never
written. CUDA only: there is no CPU path (csrc/patch_embed.cu)."""
from __future__ import annotations

import ctypes
import math
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as tF

from . import _lib
from ._lib import MbevError, check, ptr


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class PillarPatchEmbed(nn.Module):
    def __init__(self, in_channels: int = 3, embed_dims: int = 768, conv_type: str = 'Conv2d', kernel_size: int = 16,
                 stride: Optional[int] = None, padding='corner', dilation: int = 1, bias: bool = True,
                 norm_cfg: Optional[dict] = None, input_size=None, init_cfg=None):
        super().__init__()
        stride = kernel_size if stride is None else stride
        if conv_type not in (None, 'Conv2d') or stride != kernel_size or dilation != 1 or padding != 'corner':
            raise MbevError("PillarPatchEmbed covers the non-overlapping Conv2d patch embedding of swin.py:576-586 "
                            "(kernel_size == stride, dilation 1, padding='corner')")
        self.embed_dims = embed_dims
        self.in_channels = in_channels
        self.patch = int(kernel_size)
        self.projection = nn.Conv2d(in_channels, embed_dims, kernel_size=self.patch, stride=self.patch, bias=bias)
        if norm_cfg is not None:
            if norm_cfg.get('type', 'LN') != 'LN':
                raise MbevError(f"norm_cfg {norm_cfg}: only LayerNorm (type='LN') follows the patch embedding in swin.py")
            self.norm = nn.LayerNorm(embed_dims, eps=norm_cfg.get('eps', 1e-5))
        else:
            self.norm = None
        self._cache_key = None
        self._cache = None

    def out_size(self, ny: int, nx: int) -> Tuple[int, int]:
        return (math.ceil(ny / self.patch), math.ceil(nx / self.patch))

    # parameter-only work, redone only when a parameter changed (optimizer steps bump ``_version``)
    def _prepared(self, ln: nn.LayerNorm, ny: int, nx: int):
        w, b = self.projection.weight, self.projection.bias
        key = (w._version, w.data_ptr(), None if b is None else (b._version, b.data_ptr()), ln.weight._version,
               ln.weight.data_ptr(), ln.bias._version, ln.bias.data_ptr(), ny, nx, str(w.device))
        if key == self._cache_key:
            return self._cache
        lib = _lib.load()
        E, C, ps = self.embed_dims, self.in_channels, self.patch
        Hp, Wp = self.out_size(ny, nx)
        with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            pad = (0, Wp * ps - nx, 0, Hp * ps - ny)  # corner padding: zeros at the bottom / right, after the LayerNorm
            lw = tF.pad(ln.weight.detach().float()[None], pad)
            lb = tF.pad(ln.bias.detach().float()[None], pad)
            wf = w.detach().float()
            # channels-last (Hp*Wp, E) parameter images: P0 = conv(ln_bias) + bias, P1 = conv(ln_weight)
            p0 = tF.conv2d(lb, wf, None if b is None else b.detach().float(), stride=ps)[0].permute(1, 2, 0).reshape(Hp * Wp, E).contiguous()
            p1 = tF.conv2d(lw, wf, None, stride=ps)[0].permute(1, 2, 0).reshape(Hp * Wp, E).contiguous()
            lnw_cl = ln.weight.detach().float().permute(1, 2, 0).reshape(ny * nx, C).contiguous()
            w_img = torch.empty(ps * ps * 2 * E * C, dtype=torch.float32, device=w.device)
            with torch.cuda.device(w.device):
                check(lib.mbev_patch_embed_prepare_weights(ptr(wf.contiguous()), E, C, ps, ptr(w_img), _stream()),
                      "patch_embed_prepare_weights")
        self._cache_key, self._cache = key, (lnw_cl, w_img, p0, p1)
        return self._cache

    def forward_pillars(self, feats: torch.Tensor, coors: torch.Tensor, cell_table: torch.Tensor,
                        pillar_base: torch.Tensor, batch: int, ny: int, nx: int, layer_norm: nn.LayerNorm):
        """feats (cap, C), coors (cap, 4), cell_table (B, ny*nx), pillar_base (B+1) from K1 / K2; layer_norm = the
        encoder's nn.LayerNorm([C, ny, nx]). Returns (tokens (B, Hp*Wp, E), (Hp, Wp)) like PatchEmbed.forward."""
        if not feats.is_cuda:
            raise MbevError("PillarPatchEmbed has no CPU path: voxel_features must be a CUDA tensor")
        if torch.is_grad_enabled() and (feats.requires_grad or self.projection.weight.requires_grad):
            raise MbevError("PillarPatchEmbed.forward_pillars is forward-only: call under torch.no_grad()")
        lib = _lib.load()
        C, E, ps = self.in_channels, self.embed_dims, self.patch
        if feats.shape[1] != C or tuple(layer_norm.normalized_shape) != (C, ny, nx):
            raise MbevError(f"expected {C} feature channels and LayerNorm([{C}, {ny}, {nx}])")
        if not lib.mbev_patch_embed_supported(batch, C, ny, nx, ps, E):
            raise MbevError(f"patch embedding on pillars does not cover C={C}, E={E}, patch={ps} "
                            "(C in {32, 64, 128}, E % 32 == 0, E <= 256, weights within shared memory)")
        dev = feats.device
        feats = feats.detach().float().contiguous()
        lnw_cl, w_img, p0, p1 = self._prepared(layer_norm, ny, nx)
        Hp, Wp = self.out_size(ny, nx)
        cap = feats.shape[0]
        tokens = torch.empty((batch, Hp * Wp, E), dtype=torch.float32, device=dev)
        stats = torch.empty((batch, 2), dtype=torch.float32, device=dev)
        nbytes = ctypes.c_size_t()
        check(lib.mbev_patch_embed_workspace_bytes(batch, cap, ps, E, ctypes.byref(nbytes)), "patch_embed_workspace_bytes")
        ws = torch.empty(max(nbytes.value, 16), dtype=torch.uint8, device=dev)
        nw = None if self.norm is None else self.norm.weight.detach().float().contiguous()
        nb = None if self.norm is None else self.norm.bias.detach().float().contiguous()
        with torch.cuda.device(dev):
            check(lib.mbev_patch_embed_forward(ptr(feats), ptr(coors), ptr(cell_table), ptr(pillar_base), cap, batch, C,
                                               ny, nx, ps, E, ptr(lnw_cl), float(layer_norm.eps), ptr(w_img), ptr(p0),
                                               ptr(p1), ptr(nw), ptr(nb),
                                               float(self.norm.eps) if self.norm is not None else 0.0, ptr(tokens),
                                               ptr(stats), ptr(ws), ws.numel(), _stream()), "patch_embed_forward")
        return tokens, (Hp, Wp)

    def forward(self, x):
        raise MbevError("PillarPatchEmbed consumes pillars: use MaskBevEncoder.forward_patch_tokens / forward_pillars "
                        "(the dense pseudo image is what this module exists to avoid)")

"""CPU: the C-ABI library loads and exports every symbol include/mask_bev_b200.h declares (no compute calls), the
host-side mirror keeps the reference's interface and state-dict keys, CPU inputs are rejected loudly, and the
reference's own encoder file imports unchanged against the shims."""
import ctypes
import importlib
import os
import re
import sys

import pytest
import torch

from helpers import ROOT, ref_test_kwargs


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "mask_bev_b200.h")).read()
    return sorted(set(re.findall(r"MBEV_API[^;(]*?\b(mbev_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mask_bev_b200 import _lib
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(syms)
    assert lib.mbev_abi_version() == _lib.ABI_VERSION
    assert b"sm_100a" in lib.mbev_build_info()
    assert lib.mbev_status_string(-2) == b"unsupported shape or configuration"


def test_struct_layout_matches_header():
    from mask_bev_b200 import _lib
    assert ctypes.sizeof(_lib.MbevGeometry) == 6 * 4 + 3 * 4 + 3 * 4 + 4 * 4
    p = _lib.MbevPfnParams
    assert p.in_dim.offset == 4 and p.units.offset == 20 and p.weight.offset == 40
    assert p.scale.offset == 72 and p.shift.offset == 104 and p.with_cluster_center.offset == 136
    assert ctypes.sizeof(p) == 136 + 5 * 4 + 6 * 4 + 4  # trailing pad to 8


def test_argument_validation_without_gpu():
    """Pure host-side checks of the ABI return codes (no kernel is launched)."""
    from mask_bev_b200 import _lib
    from mask_bev_b200 import functional as F_
    lib = _lib.load()
    geo = F_.make_geometry([0.16, 0.16, 40], [-40, -40, -20, 40, 40, 20], 32, 250000, 4, True)
    assert list(geo.grid) == [500, 500, 1]
    n = ctypes.c_size_t()
    assert lib.mbev_voxelize_workspace_bytes(ctypes.byref(geo), 2, 1000, ctypes.byref(n)) == 0 and n.value > 16000
    assert lib.mbev_voxelize_workspace_bytes(ctypes.byref(geo), 0, 1000, ctypes.byref(n)) == -1
    assert lib.mbev_voxelize_workspace_bytes(ctypes.byref(geo), 129, 1000, ctypes.byref(n)) == -1
    geo.num_feats = 2
    assert lib.mbev_voxelize_workspace_bytes(ctypes.byref(geo), 1, 10, ctypes.byref(n)) == -2
    with pytest.raises(_lib.MbevError):
        _lib.check(-3, "x")


def test_pillar_capacity_contract_ragged_batches_without_gpu():
    """Round-1 bug (GPUTEST_r01): the host side sized the pillar buffers as sum_f min(n_f, V, cells) while
    mbev_voxelize demanded min(total, B * min(V, cells)) — any ragged batch with one frame above min(V, cells) was
    refused with status -1. Both sides now use mbev_pillar_capacity; the check runs before the device is touched
    (fake non-null pointers, a zero-byte workspace: -3 = the capacity was accepted, -1 = refused)."""
    from mask_bev_b200 import _lib
    from mask_bev_b200 import functional as F_
    lib = _lib.load()
    fake = ctypes.c_void_p(4096)
    for grid_rng, vs, V, sizes in [((-8, 8, -6, 6), 0.5, 250000, [1500, 40, 900]),     # the failing round-1 case: 32 x 24 grid
                                   ((-8, 8, -6, 6), 0.5, 100, [1500, 0, 50, 99, 101]),  # max_voxels binds
                                   ((-40, 40, -40, 40), 0.16, 250000, [120000, 7, 0]),
                                   ((-8, 8, -6, 6), 0.5, 250000, [0])]:
        x0, x1, y0, y1 = grid_rng
        geo = F_.make_geometry([vs, vs, 4], [x0, y0, -2, x1, y1, 2], 8, V, 4, True)
        cells = geo.grid[0] * geo.grid[1] * geo.grid[2]
        off, total = F_._offsets(sizes)
        want = sum(min(s, V, cells) for s in sizes)
        assert lib.mbev_pillar_capacity(off, len(sizes), ctypes.byref(geo)) == want
        assert F_.pillar_capacity(geo, sizes) == max(1, want)
        args = lambda cap: (fake, off, len(sizes), ctypes.byref(geo), fake, fake, fake, fake, fake, cap, fake, 0, None)  # noqa: E731
        assert lib.mbev_voxelize(*args(want)) == -3, sizes       # capacity accepted, then "workspace too small"
        if want > 0:
            assert lib.mbev_voxelize(*args(want - 1)) == -1, sizes
    assert lib.mbev_pillar_capacity(None, 1, ctypes.byref(geo)) == -1


def test_capability_probes_without_gpu():
    """Shape / alignment probes of the fused entries are pure host code: they answer without a device."""
    from mask_bev_b200 import _lib
    from mask_bev_b200 import functional as F_
    import mask_bev_b200 as M
    from mask_bev_b200.synthetic import encoder_kwargs
    lib = _lib.load()
    null = ctypes.c_void_p(None)
    # K3 + LayerNorm: ny*nx must be a multiple of 4, batch within MBEV_MAX_BATCH
    assert lib.mbev_scatter_layernorm_supported(16, 128, 800, 800, null, null, null) == 1
    assert lib.mbev_scatter_layernorm_supported(2, 64, 25, 25, null, null, null) == 0
    assert lib.mbev_scatter_layernorm_supported(0, 64, 500, 500, null, null, null) == 0
    assert lib.mbev_scatter_layernorm_supported(129, 64, 500, 500, null, null, null) == 0
    n = ctypes.c_size_t()
    assert lib.mbev_scatter_layernorm_workspace_bytes(16, ctypes.byref(n)) == 0 and n.value >= 16 * 64 * 16
    assert lib.mbev_scatter_layernorm_workspace_bytes(0, ctypes.byref(n)) == -1
    # backward of K3 + LayerNorm: channels in groups of 4, plane a multiple of 4 cells
    assert lib.mbev_scatter_layernorm_backward_supported(16, 128, 800, 800) == 1
    assert lib.mbev_scatter_layernorm_backward_supported(2, 6, 16, 16) == 0
    assert lib.mbev_scatter_layernorm_backward_supported(2, 64, 25, 25) == 0
    assert lib.mbev_scatter_layernorm_backward_supported(129, 64, 16, 16) == 0
    assert lib.mbev_scatter_layernorm_backward_workspace_bytes(16, 128, 800, 800, ctypes.byref(n)) == 0
    assert n.value >= 16 * (5000 * 32 // 8) * 16   # one fp64 (S1, S2) pair per frame and CTA of the streaming pass
    assert lib.mbev_scatter_layernorm_backward_workspace_bytes(2, 6, 16, 16, ctypes.byref(n)) == -2
    st = lib.mbev_scatter_layernorm_backward(null, null, null, null, null, 0, 1, 4, 8, 8, null, null, null, null, null,
                                             null, 0, null)
    assert st == -1
    # K3 through the TMA engine: channels in groups of 4, plane a multiple of 4 cells, 16-byte aligned canvas
    assert lib.mbev_scatter_stream_supported(128, 800, 800, null) == 1
    assert lib.mbev_scatter_stream_supported(128, 25, 25, null) == 0
    assert lib.mbev_scatter_stream_supported(6, 16, 16, null) == 0
    assert lib.mbev_scatter_stream_supported(64, 16, 16, ctypes.c_void_p(4100)) == 0
    assert lib.mbev_scatter_forward_stream(null, null, 1, 64, 16, 16, null, 1, null) == -1
    assert lib.mbev_scatter_forward_stream(ctypes.c_void_p(4096), ctypes.c_void_p(4096), 1, 64, 16, 16,
                                           ctypes.c_void_p(4096), 0, null) == -1          # ctas_per_sm >= 1
    assert lib.mbev_scatter_forward_nhwc(ctypes.c_void_p(4096), ctypes.c_void_p(4096), 1, 6, 16, 16,
                                         ctypes.c_void_p(4096), null) == -2               # C % 4
    # K3 + LayerNorm: the walk must be one of the two schedules (checked before the device is touched)
    f = ctypes.c_void_p(4096)
    assert lib.mbev_scatter_layernorm_forward(f, f, f, 1, 4, 8, 8, f, f, 1e-3, 7, f, f, f, 1 << 20, null) == -1
    # the pipelined entry refuses missing or aliased streams / events before touching the device
    enc = M.MaskBevEncoder(**encoder_kwargs("kitti_b16"))
    cfg = enc._voxel_encoder._config()
    assert F_.pfn_path(cfg, 32) == "tcgen05" and F_.pfn_path(cfg, 100) == "tcgen05"
    geo = F_.make_geometry([0.1, 0.1, 40], [0, -40, -20, 80, 40, 20], 32, 250000, 4, True)
    off = (ctypes.c_int64 * 2)(0, 10)
    params = F_._pfn_struct(cfg, [None] * 3, None, None)
    st = lib.mbev_encode_batch_pipelined(null, null, off, 1, ctypes.byref(geo), ctypes.byref(params), null, null, null,
                                         null, null, 10, null, null, null, 0, null, 0, 1, null, null, null, null, null,
                                         null)
    assert st == -1
    s1, s2 = ctypes.c_void_p(8), ctypes.c_void_p(16)
    st = lib.mbev_encode_batch_pipelined(null, f, off, 1, ctypes.byref(geo), ctypes.byref(params), f, f, f, f, f, 10, f,
                                         f, f, 0, f, 0, 1, null, s1, s1, f, f, f)      # prep stream == pfn stream
    assert st == -1


def test_state_dict_keys_and_shapes_match_upstream():
    import mask_bev_b200 as M
    enc = M.MaskBevEncoder(**ref_test_kwargs(feat_channels=(128, 128, 128), T=32))
    sd = enc.state_dict()
    exp = {"_voxel_encoder.pfn_layers.0.linear.weight": (64, 11), "_voxel_encoder.pfn_layers.1.linear.weight": (64, 128),
           "_voxel_encoder.pfn_layers.2.linear.weight": (128, 128), "_voxel_encoder.pfn_layers.0.norm.weight": (64,),
           "_voxel_encoder.pfn_layers.2.norm.running_var": (128,), "_voxel_encoder.pfn_layers.1.norm.num_batches_tracked": (),
           "_layer_norm.weight": (128, 500, 500), "_layer_norm.bias": (128, 500, 500)}
    for k, shp in exp.items():
        assert k in sd and tuple(sd[k].shape) == shp, k
    assert len(sd) == 3 * 6 + 2
    bn = enc._voxel_encoder.pfn_layers[0].norm
    assert bn.eps == 1e-3 and bn.momentum == 0.01
    assert enc._voxel_layer.max_voxels == (250000, 250000) and enc._voxel_layer.deterministic
    assert enc._voxel_layer.grid_size.tolist() == [500, 500, 1]
    assert (enc._num_voxel_x, enc._num_voxel_y) == (500, 500)
    # oracle module has the same keys -> checkpoints are interchangeable
    from oracle import oracle as O
    pfn = O.make_pfn_oracle(in_channels=4, feat_channels=(128, 128, 128), with_distance=True)
    assert set(pfn.state_dict()) == {k[len("_voxel_encoder."):] for k in sd if k.startswith("_voxel_encoder.")}


def test_cpu_tensors_are_rejected_no_fallback():
    import mask_bev_b200 as M
    enc = M.MaskBevEncoder(**ref_test_kwargs())
    with pytest.raises(M.MbevError, match="no CPU path"):
        enc([torch.zeros(10, 4)])
    with pytest.raises(M.MbevError):
        enc._voxel_layer(torch.zeros(10, 4))
    with pytest.raises(M.MbevError):
        enc.middle_encode(torch.zeros(3, 64), torch.zeros(3, 4, dtype=torch.int32), 1)
    with pytest.raises(NotImplementedError):
        M.MaskBevEncoder(**{**ref_test_kwargs(), "encoding_type": "cosine"})


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mask_bev_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("the oracle in tests only", ""), f"{f} mentions the oracle"


@pytest.mark.skipif(not os.path.exists("/root/reference/mask_bev/models/encoders/mask_bev_encoders.py"),
                    reason="reference tree only exists in the build container")
def test_reference_encoder_file_imports_unchanged_against_shims():
    shim = os.path.join(ROOT, "mask_bev_b200", "shims")
    saved = list(sys.path)
    try:
        sys.path.insert(0, shim)
        sys.path.insert(0, "/root/reference")
        for m in [m for m in sys.modules if m.split(".")[0] in ("mmcv", "mmdet3d", "mask_bev")]:
            del sys.modules[m]
        mod = importlib.import_module("mask_bev.models.encoders.mask_bev_encoders")
        import mask_bev_b200 as M
        assert mod.Voxelization is M.Voxelization and mod.PillarFeatureNet is M.PillarFeatureNet
        enc = mod.MaskBevEncoder([128, 128, 128], (-40, 40), (-40, 40), (-20, 20), 0.16, 0.16, 40, 32, 'vanilla', 1,
                                 encoder_params=dict(with_distance=True), pc_point_dim=4)
        mine = M.MaskBevEncoder([128, 128, 128], (-40, 40), (-40, 40), (-20, 20), 0.16, 0.16, 40, 32, 'vanilla', 1,
                                encoder_params=dict(with_distance=True), pc_point_dim=4)
        assert {k: tuple(v.shape) for k, v in enc.state_dict().items()} == \
               {k: tuple(v.shape) for k, v in mine.state_dict().items()}
    finally:
        sys.path[:] = saved
        for m in [m for m in sys.modules if m.split(".")[0] in ("mmcv", "mmdet3d", "mask_bev")]:
            del sys.modules[m]


def test_reference_checkpoint_of_the_golden_run_loads_strictly():
    """The state dict saved from the reference's own MaskBevEncoder (tests/golden/encoder_reference.npz) loads into the
    product encoder with strict=True: same keys, same shapes (mask_bev_module.py:113-126 loads checkpoints by name)."""
    import mask_bev_b200 as M
    from helpers import encoder_reference
    _, weights, out, kw = encoder_reference()
    enc = M.MaskBevEncoder(**kw)
    res = enc.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()}, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert [enc._num_voxel_y, enc._num_voxel_x] == [int(v) for v in out["canvas_shape"]]


def test_header_is_plain_c_and_a_c_program_links_against_the_library(tmp_path):
    """The boundary is a C ABI: the header compiles as strict C99 (and as C++17), and a C program that takes the address
    of every declared entry point links against libmask_bev_b200.so and runs its host-only calls."""
    import shutil
    import subprocess
    from mask_bev_b200 import _lib
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    hdr = os.path.join(ROOT, "include", "mask_bev_b200.h")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr], check=True)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr], check=True)
    _lib.load()
    syms = _header_symbols()
    src = tmp_path / "link_all.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "mask_bev_b200.h"\n'
        "int main(void) {\n"
        "  const void *fns[] = {" + ", ".join(f"(const void *)&{s}" for s in syms) + "};\n"
        "  size_t n = sizeof fns / sizeof fns[0], i, bytes = 0;\n"
        "  for (i = 0; i < n; ++i) if (!fns[i]) return 2;\n"
        "  if (mbev_abi_version() != MBEV_ABI_VERSION) return 3;\n"
        "  if (!strstr(mbev_build_info(), \"sm_100a\")) return 4;\n"
        "  if (mbev_scatter_layernorm_workspace_bytes(16, &bytes) != 0 || bytes == 0) return 5;\n"
        "  if (mbev_scatter_layernorm_backward_supported(16, 128, 800, 800) != 1) return 6;\n"
        "  printf(\"%u entry points, ABI %d\\n\", (unsigned)n, mbev_abi_version());\n"
        "  return 0;\n}\n")
    exe = tmp_path / "link_all"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-l:" + os.path.basename(_lib.LIB_PATH), "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stderr)
    assert f"{len(syms)} entry points, ABI {_lib.ABI_VERSION}" in r.stdout


def _training_configs():
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", "training_configs.json")))


def _encoder_kwargs_of(cfg):
    """As MaskBevModule.__init__ derives them (mask_bev_module.py:39-43 defaults, :62, :72-75)."""
    z = cfg["z_range"]
    return dict(feat_channels=list(cfg["encoder_feat_channels"]), x_range=cfg["x_range"], y_range=cfg["y_range"],
                z_range=z, voxel_size_x=cfg["voxel_size"], voxel_size_y=cfg["voxel_size"], voxel_size_z=z[1] - z[0],
                max_num_points=cfg["max_num_points"], encoding_type=cfg.get("encoder_encoding_type", "vanilla"),
                fourier_enc_group=cfg.get("encoder_fourier_enc_group", 1), encoder_params=dict(with_distance=True),
                pc_point_dim=cfg.get("pc_point_dim", 4))


def test_every_reference_training_config_builds_and_fits_the_kernels():
    """SURVEY.md §8a parameter envelope: every file under the reference's configs/training (collected into
    tests/golden/training_configs.json by make_training_configs.py) builds the product encoder with the canvas the
    reference derives, and its shapes are inside what the kernels and their fused forms accept (host-side probes)."""
    import mask_bev_b200 as M
    from mask_bev_b200 import functional as F_
    cfgs = _training_configs()
    assert len(cfgs) == 11
    seen = set()
    for name, cfg in cfgs.items():
        kw = _encoder_kwargs_of(cfg)
        enc = M.MaskBevEncoder(**kw)
        nx = int((cfg["x_range"][1] - cfg["x_range"][0]) / cfg["voxel_size"])     # mask_bev_module.py:67-68
        ny = int((cfg["y_range"][1] - cfg["y_range"][0]) / cfg["voxel_size"])
        assert (enc._num_voxel_x, enc._num_voxel_y) == (nx, ny), name
        assert tuple(enc._layer_norm.weight.shape) == (kw["feat_channels"][-1], ny, nx), name
        C, B, T = kw["feat_channels"][-1], cfg.get("batch_size", 1), kw["max_num_points"]
        pcfg = enc._voxel_encoder._config()
        assert pcfg.in_dims[0] == kw["pc_point_dim"] + 3 + 3 + 1, name           # decoration width D
        assert F_.pfn_path(pcfg, T) in ("tcgen05", "fma"), name
        assert F_.scatter_layernorm_backward_supported(B, C, ny, nx), name       # fused LayerNorm pair usable
        assert _lib_probe_ln(B, C, ny, nx), name
        seen.add((tuple(kw["feat_channels"]), kw["pc_point_dim"], nx, ny))
    assert ((256, 128, 128), 4, 500, 500) in seen and ((128, 64, 128), 4, 500, 500) in seen
    assert any(s[1] == 3 for s in seen) and any(s[2] == 800 for s in seen)


def _lib_probe_ln(B, C, ny, nx):
    from mask_bev_b200 import _lib
    null = ctypes.c_void_p(None)
    return _lib.load().mbev_scatter_layernorm_supported(B, C, ny, nx, null, null, null) == 1


def test_training_configs_fixture_matches_the_reference_when_present():
    if not os.path.isdir("/root/reference/configs/training"):
        pytest.skip("reference tree not mounted (GPU box)")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_training_configs",
                                                  os.path.join(ROOT, "tests", "golden", "make_training_configs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.collect() == _training_configs()


def test_running_stats_update_matches_the_batchnorm_formula():
    """PillarFeatureNet._update_running_stats (multi-tensor form): running = (1 - m) * running + m * new with the unbiased
    variance over M = P*T slots, num_batches_tracked += 1; a step without pillars leaves the buffers untouched. Pure
    torch, so it runs on the CPU (the kernels only produce `batch_stats`)."""
    import copy

    import torch
    from mask_bev_b200.pillar_encoder import PillarFeatureNet
    torch.manual_seed(0)
    net = PillarFeatureNet(4, (128, 128, 128), with_distance=True, norm_cfg=dict(type='BN1d', eps=1e-3, momentum=0.01))
    ref = copy.deepcopy(net)
    bs = torch.rand(3, 2, 128)
    P, T = 1234, 32
    net._update_running_stats(bs, torch.tensor([P], dtype=torch.int32), T)   # a 1-element tensor, as the kernels hand it over
    M = P * T
    for l, (mine, layer) in enumerate(zip(net.pfn_layers, ref.pfn_layers)):
        bn, U = layer.norm, layer.units
        bn.running_mean.mul_(1 - 0.01).add_(bs[l, 0, :U] * 0.01)
        bn.running_var.mul_(1 - 0.01).add_(bs[l, 1, :U] * (M / (M - 1)) * 0.01)
        assert int(mine.norm.num_batches_tracked) == 1
        assert torch.allclose(mine.norm.running_mean, bn.running_mean, rtol=1e-6, atol=0)
        assert torch.allclose(mine.norm.running_var, bn.running_var, rtol=1e-6, atol=0)
    before = [l.norm.running_var.clone() for l in net.pfn_layers]
    net._update_running_stats(bs, torch.tensor([0], dtype=torch.int32), T)
    for b, l in zip(before, net.pfn_layers):
        assert torch.equal(b, l.norm.running_var) and int(l.norm.num_batches_tracked) == 1


def test_training_step_pair_validates_its_arguments_without_gpu():
    """mbev_pfn_forward_train_rows / mbev_pfn_backward_rows (ABI v11): host-side argument checks return before any launch
    — null pointers and a workspace smaller than mbev_pfn_backward_workspace_bytes are refused, an empty step is a no-op."""
    import torch
    from mask_bev_b200 import _lib
    from mask_bev_b200 import functional as F_
    lib = _lib.load()
    assert lib.mbev_abi_version() == 11
    cfg = F_.PfnConfig(in_channels=4, units=[64, 64, 128], in_dims=[11, 128, 128], with_cluster_center=True,
                       with_voxel_center=True, with_distance=True, legacy=True, voxel_center_dims=3, vx=0.16, vy=0.16,
                       vz=40.0, x_offset=-39.92, y_offset=-39.92, z_offset=0.0, eps=1e-3, gemm_path=0)
    ws = [torch.zeros(u, k) for u, k in zip(cfg.units, cfg.in_dims)]
    params = F_._pfn_struct(cfg, ws, None, None)
    n = ctypes.c_size_t()
    assert lib.mbev_pfn_backward_workspace_bytes(ctypes.byref(params), 32, 1000, 5000, ctypes.byref(n)) == 0
    rows = 5000 * (11 + 128 + 128 + 64 + 64 + 128) * 4   # X_l and Y_l of every layer are kept for the backward
    assert n.value > rows
    fake = ctypes.c_void_p(256)   # never dereferenced: every call below returns from the host-side checks
    three = (ctypes.c_void_p * F_.MAX_LAYERS)(256, 256, 256)
    none = ctypes.c_void_p(None)
    # forward: null workspace / null gamma array
    assert lib.mbev_pfn_forward_train_rows(fake, 4, fake, fake, fake, fake, 1000, 32, 5000, ctypes.byref(params), three,
                                           three, 1e-3, fake, fake, fake, none, n.value, none) == -1
    assert lib.mbev_pfn_forward_train_rows(fake, 4, fake, fake, fake, fake, 1000, 32, 5000, ctypes.byref(params), None,
                                           three, 1e-3, fake, fake, fake, fake, n.value, none) == -1
    # workspace too small -> MBEV_ERR_WORKSPACE (-3); empty step (capacity 0) -> OK without touching anything
    assert lib.mbev_pfn_forward_train_rows(fake, 4, fake, fake, fake, fake, 1000, 32, 5000, ctypes.byref(params), three,
                                           three, 1e-3, fake, fake, fake, fake, 1024, none) == -3
    assert lib.mbev_pfn_forward_train_rows(fake, 4, fake, fake, fake, fake, 0, 32, 0, ctypes.byref(params), three,
                                           three, 1e-3, fake, fake, fake, fake, 1024, none) == 0
    # backward: null dfeats, short workspace, inconsistent layer chain
    assert lib.mbev_pfn_backward_rows(fake, 1000, 4, 32, 5000, ctypes.byref(params), 1e-3, none, three, three, three,
                                      fake, n.value, none) == -1
    assert lib.mbev_pfn_backward_rows(fake, 1000, 4, 32, 5000, ctypes.byref(params), 1e-3, fake, three, three, three,
                                      fake, 1024, none) == -3
    params.in_dim[1] = 64
    assert lib.mbev_pfn_backward_rows(fake, 1000, 4, 32, 5000, ctypes.byref(params), 1e-3, fake, three, three, three,
                                      fake, n.value, none) == -1

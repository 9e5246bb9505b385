"""CPU: the packed collate (SURVEY.md §8 f3 host side) returns what the reference's list collate returns
(semantic_kitti_transforms.py:95-118) with the point clouds packed into one buffer, frame views in batch order."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _batch(with_meta=True, sizes=(7, 0, 3)):
    g = torch.Generator().manual_seed(0)
    items = []
    for i, n in enumerate(sizes):
        x = torch.randn((n, 4), generator=g)
        y = (torch.full((5,), i), torch.randn((5, 6, 6), generator=g))
        items.append((x, y, {"frame": i}) if with_meta else (x, y))
    return items


def test_packed_collate_mirrors_the_reference_list_collate():
    from mask_bev_b200.collate import PackedFrames, PackedListCollate
    batch = _batch()
    packed, (labels, masks), meta = PackedListCollate()(batch)
    # the reference's MaskListCollateHeight on the same batch
    ref_pc = [b[0] for b in batch]
    assert isinstance(packed, PackedFrames) and len(packed) == 3 and packed.sizes == [7, 0, 3]
    assert packed.offsets == [0, 7, 7, 10] and packed.points.shape == (10, 4) and packed.points.is_contiguous()
    for a, b in zip(packed.frames(), ref_pc):
        assert torch.equal(a, b)
    assert torch.equal(labels, torch.stack([b[1][0] for b in batch]))
    assert torch.equal(masks, torch.stack([b[1][1] for b in batch]))
    assert meta == [{"frame": 0}, {"frame": 1}, {"frame": 2}]
    # frames are VIEWS of the one buffer (one H2D copy moves them all)
    assert packed.frames()[2].data_ptr() == packed.points.data_ptr() + 7 * 4 * 4
    packed2, (l2, m2) = PackedListCollate()(_batch(with_meta=False))
    assert torch.equal(packed2.points, packed.points) and torch.equal(l2, labels)
    moved = packed.to("cpu")
    assert torch.equal(moved.points, packed.points) and moved.sizes == packed.sizes


def test_pack_rejects_ragged_feature_counts_and_empty_batches():
    from mask_bev_b200.collate import PackedFrames, pack_point_clouds
    with pytest.raises(ValueError):
        pack_point_clouds([torch.zeros(3, 4), torch.zeros(2, 3)])
    with pytest.raises(ValueError):
        pack_point_clouds([])
    with pytest.raises(ValueError):
        PackedFrames(torch.zeros(5, 4), [2, 2])
    p = pack_point_clouds([torch.zeros(2, 4, dtype=torch.float64)])
    assert p.points.dtype == torch.float32


def test_dataloader_with_packed_collate():
    from torch.utils.data import DataLoader
    from mask_bev_b200.collate import PackedListCollate
    items = _batch(sizes=(4, 2, 9, 1))
    loader = DataLoader(items, batch_size=2, collate_fn=PackedListCollate(), num_workers=0)
    got = list(loader)
    assert [b[0].sizes for b in got] == [[4, 2], [9, 1]]
    assert torch.equal(got[1][0].frames()[0], items[2][0])


def test_encoder_accepts_a_list_or_a_packed_batch_identically():
    """MaskBevEncoder's entry normalises both input forms to (concatenated points, frame sizes); a single frame and a
    packed batch are passed through without a copy."""
    from mask_bev_b200.collate import pack_point_clouds
    from mask_bev_b200.encoder import _as_points
    frames = [torch.randn(5, 4), torch.randn(0, 4), torch.randn(3, 4)]
    pts, sizes = _as_points(frames)
    assert sizes == [5, 0, 3] and torch.equal(pts, torch.cat(frames)) and pts.is_contiguous()
    packed = pack_point_clouds(frames)
    pts2, sizes2 = _as_points(packed)
    assert sizes2 == sizes and pts2.data_ptr() == packed.points.data_ptr() and torch.equal(pts2, pts)
    one, s1 = _as_points([frames[0]])
    assert one.data_ptr() == frames[0].data_ptr() and s1 == [5]

"""GPU: the data-parallel wrapper on a single rank (no process group) is the plain training step — same canvas, same
gradients bit for bit (the kernels' reductions are fixed-order) — and it reports the step's allreduce volume
(25 792 PFN floats + 2*C*ny*nx LayerNorm floats). The two-rank exchange itself is covered on CPU with
gloo (tests/test_data_parallel_gloo.py); under torchrun the same code runs over NCCL."""
import pytest
import torch

from helpers import encoder_pair, ref_test_kwargs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_single_rank_wrapper_equals_plain_step():
    from mask_bev_b200.data_parallel import FrontEndDataParallel, gradient_bytes
    from mask_bev_b200.synthetic import gen_frame
    kw = ref_test_kwargs(feat_channels=(128, 128, 128), T=32, vs=0.8)   # 100 x 100
    enc, _ = encoder_pair(kw, seed=4)
    enc = enc.to(DEV).train()
    frames = [torch.from_numpy(gen_frame(8000, 4, s)).to(DEV) for s in (1, 2, 3)]
    g = torch.Generator().manual_seed(5)
    R = torch.randn((3, 128, 100, 100), generator=g).to(DEV)
    state = {k: v.clone() for k, v in enc.state_dict().items()}
    assert gradient_bytes(enc) == (25792 + 2 * 128 * 100 * 100) * 4

    enc.zero_grad(set_to_none=True)
    out_a = enc(frames)
    (out_a * R).sum().backward()
    grads_a = {n: p.grad.clone() for n, p in enc.named_parameters()}

    enc.load_state_dict(state)
    enc.zero_grad(set_to_none=True)
    dp = FrontEndDataParallel(enc)
    out_b, owned = dp(frames)
    assert owned == [0, 1, 2]
    (out_b * R).sum().backward()
    rep = dp.reduce_gradients()
    assert rep.world == 1 and rep.collectives == 0
    assert torch.equal(out_a, out_b)
    for n, p in enc.named_parameters():
        assert torch.equal(p.grad, grads_a[n]), n

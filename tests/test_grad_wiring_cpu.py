"""CPU: the WIRING of the differentiable generator path (e4s2024_b200/stylegan2/grad.py: which style rows feed which layer, table algebra,
noise / bias / activation / ToRGB / skip order, region handling) with its CUDA-backed primitives replaced by plain torch stand-ins of
their documented semantics, against autograd through the oracle.  In fp64 both must agree to rounding: any difference is a wiring bug.
(The primitives themselves are checked on the GPU, tests/test_backward_gpu.py.)"""
import pytest
import torch
import torch.nn.functional as F

from e4s2024_b200 import synth
from oracle import e4s_oracle as orc


class _ModConv:
    """ModConvFn's contract: y = d[b, r(p)] * conv(x * s[b, r(p)]; scale W), r(p) by torch's legacy-nearest lookup of the label map."""

    @staticmethod
    def apply(x, weight, s, d, labels, mc, regions):
        B, _, H, W = x.shape
        up, k, cout = mc.upsample, mc.kernel_size, mc.out_channel
        ho, wo = (2 * H, 2 * W) if up else (H, W)
        if labels is not None:
            lh, lw = labels.shape[1:]
            ys = torch.floor(torch.arange(ho, dtype=torch.float32) * (lh / ho)).long().clamp(max=lh - 1)
            xs = torch.floor(torch.arange(wo, dtype=torch.float32) * (lw / wo)).long().clamp(max=lw - 1)
            reg = labels.long()[:, ys][:, :, xs]
        else:
            reg = torch.zeros(B, ho, wo, dtype=torch.long)
        wsc = mc.scale * weight[0]
        out = 0
        for r in range(regions):
            xm = x * s[:, r][:, :, None, None]
            if up:
                yt = F.conv_transpose2d(xm, wsc.transpose(0, 1), stride=2)
                yr = orc.upfirdn2d(yt, mc.blur.kernel.to(x.dtype), pad=mc.blur.pad)
            else:
                yr = F.conv2d(xm, wsc, padding=k // 2)
            if d is not None:
                yr = yr * d[:, r][:, :, None, None]
            out = out + yr * (reg == r)[:, None].to(x.dtype)
        return out


class _Linear:
    @staticmethod
    def apply(x, w, scale):
        return x @ (scale * w).t()


class _Ctx:
    def __init__(self, mask, lazy=False):
        self.onehot = True
        self.labels = mask.argmax(1).to(torch.uint8)


@pytest.mark.parametrize("size,rl,split", [(32, 18, 7), (64, 5, 5)])
def test_training_forward_wiring_matches_oracle_autograd(size, rl, split, monkeypatch):
    from e4s2024_b200.stylegan2 import grad as GR
    from e4s2024_b200.stylegan2.model import Generator
    monkeypatch.setattr(GR, "ModConvFn", _ModConv)
    monkeypatch.setattr(GR, "LinearFn", _Linear)
    monkeypatch.setattr(GR, "fused_leaky_relu", lambda x, b, ns=0.2, sc=2 ** 0.5: orc.fused_leaky_relu(x, b, ns, sc))
    monkeypatch.setattr(GR, "upfirdn2d", lambda x, k, up=1, down=1, pad=(0, 0): orc.upfirdn2d(x, k.to(x.dtype), up=up, down=down, pad=pad))
    monkeypatch.setattr(GR.E, "RegionCtx", _Ctx)
    K, B = 12, 2
    G = Generator(size, 512, 2, split_layer_idx=split, remaining_layer_idx=rl)
    synth.synth_module_weights(G, seed=4)
    G = G.double()
    mask = synth.onehot(synth.blocky_labels(B, K, 512, cells=16, seed=3), K).double()
    latent0 = synth.randn("wiring.latent", (B, K, G.n_latent, 512), 31).double()
    R = synth.randn("wiring.R", (B, 3, size, size), 31).double()
    latent = latent0.clone().requires_grad_(True)
    noise = [getattr(G.noises, f"noise_{i}") for i in range(G.num_layers)]
    img, feats = GR.generator_forward(G, latent, mask, noise)
    (img * R).sum().backward()
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in G.state_dict().items()}
    lo = latent0.clone().requires_grad_(True)
    img_o, feats_o = orc.generator_forward(sd, size, lo, mask, split_layer_idx=split, remaining_layer_idx=rl)
    (img_o * R).sum().backward()
    assert float((img - img_o).abs().max()) < 1e-10
    assert (feats is None) == (feats_o is None) and (feats is None or float((feats - feats_o).abs().max()) < 1e-10)
    assert float((latent.grad - lo.grad).abs().max()) < 1e-9 * max(1.0, float(lo.grad.abs().max()))
    checked = 0
    for name, p in G.named_parameters():
        want = sd[name].grad
        if want is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        assert p.grad is not None, name
        assert float((p.grad - want).abs().max()) <= 1e-9 * max(1.0, float(want.abs().max())), name
        checked += 1
    assert checked > 20

"""Net3 = regional style encoder + 12 LocalMLPs + mask-guided StyleGAN2 -- drop-in for
`models/networks.py` (LocalMLP :23-49, Net3 :51-277): same constructor (`Net3(opts)`), methods
(forward / get_style / get_style_vectors / cal_style_codes / gen_img), return tuples, state_dict
keys and the externally assigned attribute `latent_avg` [18,512]."""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L
from . import engine as E
from .encoders.psp_encoders import FSEncoder_PSP, RGB_PAD
from .engine import View
from .stylegan2.model import EqualLinear, Generator, grad_anchor, inference_only


class LocalMLP(nn.Module):
    """networks.py:23-49: EqualLinear -> nn.LeakyReLU() (slope 0.01) -> EqualLinear."""

    def __init__(self, dim_component=512, dim_style=512, num_w_layers=18, latent_squeeze_ratio=1):
        super().__init__()
        self.dim_component = dim_component
        self.dim_style = dim_style
        self.num_w_layers = num_w_layers
        self.mlp = nn.Sequential(EqualLinear(dim_component, dim_style // latent_squeeze_ratio, lr_mul=1), nn.LeakyReLU(),
                                 EqualLinear(dim_style // latent_squeeze_ratio, dim_style * num_w_layers, lr_mul=1))

    def rows_deferred(self, src, rows, row_stride, offset, out, extra_bias=None):
        """The two GEMMs of this MLP as un-launched problems (Net3 batches all 12 MLPs into two launches)."""
        pw0, b0 = self.mlp[0].packed()
        h, p0 = E.linear_rows(src, rows, row_stride, offset, pw0, bias=b0, act=L.ACT_LRELU, slope=self.mlp[1].negative_slope,
                              gain=1.0, launch=False)
        pw1, b1 = self.mlp[2].packed()
        _, p1 = E.linear_rows(h, rows, h.shape[1], 0, pw1, bias=b1 if extra_bias is None else extra_bias, out=out, launch=False)
        return p0, p1, h

    def rows(self, src, rows, row_stride, offset, out=None, extra_bias=None):
        pw0, b0 = self.mlp[0].packed()
        h = E.linear_rows(src, rows, row_stride, offset, pw0, bias=b0, act=L.ACT_LRELU, slope=self.mlp[1].negative_slope,
                          gain=1.0)
        pw1, b1 = self.mlp[2].packed()
        return E.linear_rows(h, rows, h.shape[1], 0, pw1, bias=b1 if extra_bias is None else extra_bias, out=out)

    def forward(self, x):
        x = x.contiguous().float()
        out = self.rows(x, x.shape[0], x.shape[1], 0)
        return out.view(-1, self.num_w_layers, self.dim_style)


class Net3(nn.Module):
    """FSEncoder + StyleGAN2 (networks.py:51-277)."""

    def __init__(self, opts):
        super().__init__()
        self.opts = opts
        assert self.opts.fsencoder_type in ["psp", "sean"]
        if self.opts.fsencoder_type != "psp":
            raise NotImplementedError("fsencoder_type='sean' is not used by the released pipelines (options default 'psp')")
        self.encoder = FSEncoder_PSP(mode="ir_se", opts=self.opts)
        dim_s_code = 256 + 512 + 512
        self.split_layer_idx = 5
        self.remaining_layer_idx = self.opts.remaining_layer_idx
        self.MLPs = nn.ModuleList()
        for _ in range(self.opts.num_seg_cls):
            self.MLPs.append(LocalMLP(dim_component=dim_s_code, dim_style=512,
                                      num_w_layers=self.remaining_layer_idx if self.remaining_layer_idx != 17 else 18))
        self.G = Generator(size=self.opts.out_size, style_dim=512, n_mlp=8, split_layer_idx=self.split_layer_idx,
                           remaining_layer_idx=self.remaining_layer_idx)
        for p in self.encoder.parameters():  # the encoder has no backward here (PTI tunes the MLPs and the generator from stored style vectors)
            p.requires_grad = False
        self._bias_cache = None
        E.install_pack_invalidation(self)
        self.eval()          # like Generator: train() is the explicit switch to the differentiable (PTI) path

    # ---- encoder --------------------------------------------------------------------------------
    def _encode(self, img, mask):
        """F.interpolate(img,(256,256),'bilinear') + encoder (networks.py:113-116), resize fused with the layout change."""
        x = L.resize_bilinear_nchw_to_nhwc(img.contiguous().float(), 256, 256, RGB_PAD, align_corners=False)
        return self.encoder.run(View(x), mask)

    # ---- style codes ----------------------------------------------------------------------------
    def _codes(self, style_vectors):
        """networks.py:223-253; the `+ latent_avg[:rl]` is folded into the second EqualLinear's bias."""
        sv = style_vectors.contiguous().float()
        b, k, d = sv.shape
        rl = self.remaining_layer_idx
        n_mlp_layers = self.MLPs[0].num_w_layers
        add_avg = bool(self.opts.start_from_latent_avg)
        if add_avg and getattr(self.opts, "learn_in_w", False):
            raise NotImplementedError("learn_in_w=True is not used by the released pipelines")
        total = 18 if (add_avg and rl != 17) else n_mlp_layers
        codes = torch.empty(b, k, total, 512, device=sv.device, dtype=torch.float32)
        la = self.latent_avg.to(sv.device).float() if add_avg else None
        biases = self._mlp_biases(la, n_mlp_layers) if add_avg else [None] * k
        # 24 tiny GEMMs (M = batch rows) -> two batched launches (first layers, then second layers)
        first, second, keep = [], [], []
        for i in range(k):
            p0, p1, h = self.MLPs[i].rows_deferred(sv, b, k * d, i * d, _RowsOut(codes, i, n_mlp_layers), extra_bias=biases[i])
            first.append(p0)
            second.append(p1)
            keep.append(h)
        L.conv_batched(first)
        L.conv_batched(second)
        if add_avg and rl != 17:
            codes[:, :, rl:] = la[rl:]
        return codes

    def _mlp_biases(self, la, n_layers):
        key = (la.data_ptr(), la._version) + tuple((m.mlp[2].bias.data_ptr(), m.mlp[2].bias._version) for m in self.MLPs)
        if self._bias_cache is None or self._bias_cache[0] != key:
            flat = la[:n_layers].reshape(-1)
            self._bias_cache = (key, [(m.mlp[2].bias.detach() * m.mlp[2].lr_mul + flat).contiguous() for m in self.MLPs])
        return self._bias_cache[1]

    # ---- public API (same signatures as the reference) ---------------------------------------------
    def forward(self, img, mask, resize=False, randomize_noise=True, return_latents=False):
        anchor = grad_anchor(self, (img,))
        with torch.no_grad():
            out = self._forward_impl(img, mask, resize, randomize_noise, return_latents)
        return inference_only(out, anchor)

    def _forward_impl(self, img, mask, resize=False, randomize_noise=True, return_latents=False):
        codes_vector, structure_feats = self._encode(img, mask)
        codes = self._codes(codes_vector)
        images1, result_latent, structure_feats_GT = self.G([codes], structure_feats, mask, input_is_latent=True,
                                                            randomize_noise=randomize_noise, return_latents=return_latents,
                                                            use_structure_code=False)
        if return_latents:
            return images1, structure_feats_GT, result_latent
        return images1, structure_feats_GT

    def get_style(self, img, mask):
        anchor = grad_anchor(self, (img,))
        with torch.no_grad():
            out = self._get_style_impl(img, mask)
        return inference_only(out, anchor)

    def _get_style_impl(self, img, mask):
        codes_vector, structure_feats = self._encode(img, mask)
        return structure_feats, self._codes(codes_vector)

    def get_style_vectors(self, img, mask):
        anchor = grad_anchor(self, (img,))
        with torch.no_grad():
            out = self._get_style_vectors_impl(img, mask)
        return inference_only(out, anchor)

    def _get_style_vectors_impl(self, img, mask):
        return self._encode(img, mask)

    def _codes_train(self, style_vectors):
        """_codes through the differentiable EqualLinear (stylegan2/grad.py): what the PTI coach back-propagates into the LocalMLPs."""
        from .stylegan2 import grad as GR
        sv = style_vectors.float()
        b, k, d = sv.shape
        rl = self.remaining_layer_idx
        add_avg = bool(self.opts.start_from_latent_avg)
        if add_avg and getattr(self.opts, "learn_in_w", False):
            raise NotImplementedError("learn_in_w=True is not used by the released pipelines")
        outs = []
        for i in range(k):
            m = self.MLPs[i]
            h = GR.equal_linear(m.mlp[0], sv[:, i].contiguous())
            h = torch.nn.functional.leaky_relu(h, m.mlp[1].negative_slope)
            outs.append(GR.equal_linear(m.mlp[2], h).view(b, m.num_w_layers, m.dim_style))
        codes = torch.stack(outs, dim=1)                                  # [B, K, L, 512]
        if add_avg:
            la = self.latent_avg.to(sv.device).float()
            codes = codes + la[:codes.shape[2]]
            if rl != 17:
                codes = torch.cat([codes, la[rl:].expand(b, k, 18 - rl, 512)], dim=2)
        return codes

    def _train_path(self, anchor):
        return anchor is not None and self.training and anchor.is_cuda

    def cal_style_codes(self, style_vectors):
        anchor = grad_anchor(self, (style_vectors,))
        if self._train_path(anchor):
            return self._codes_train(style_vectors)
        with torch.no_grad():
            out = self._cal_style_codes_impl(style_vectors)
        return inference_only(out, anchor)

    def _cal_style_codes_impl(self, style_vectors):
        return self._codes(style_vectors)

    def gen_img(self, struc_codes, style_codes, mask, randomize_noise=True, noise=None, return_latents=False):
        anchor = grad_anchor(self, (style_codes,))
        if self._train_path(anchor):
            return self._gen_img_impl(struc_codes, style_codes, mask, randomize_noise, noise, return_latents)      # G.forward takes its train() path
        with torch.no_grad():
            out = self._gen_img_impl(struc_codes, style_codes, mask, randomize_noise, noise, return_latents)
        return inference_only(out, anchor)

    def _gen_img_impl(self, struc_codes, style_codes, mask, randomize_noise=True, noise=None, return_latents=False):
        images, result_latent, structure_feats = self.G([style_codes], struc_codes, mask, input_is_latent=True,
                                                        randomize_noise=randomize_noise, noise=noise,
                                                        return_latents=return_latents, use_structure_code=False)
        if return_latents:
            return images, result_latent, structure_feats
        return images, -1, structure_feats


class _RowsOut:
    """Output window of LocalMLP i inside codes [B,K,L,512]: row b starts at codes[b,i,0,0]."""

    def __init__(self, codes, i, n_layers):
        self.t = codes[:, i]                     # [B, L, 512] strided view
        self._n = n_layers

    def data_ptr(self):
        return self.t.data_ptr()

    def stride(self, dim):
        assert dim == 0
        return self.t.stride(0)


def net3_state_shapes(opts):
    with torch.device("meta"):
        m = Net3(opts)
    return {k: tuple(v.shape) for k, v in m.state_dict().items()}

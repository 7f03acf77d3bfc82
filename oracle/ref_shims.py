"""Import the reference's OWN hot-path modules from /root/reference on a CPU-only box.

TEST INFRASTRUCTURE ONLY (used by oracle/make_golden.py and, when /root/reference is
present, by tests that cross-check the oracle live).  Nothing is copied: modules are
imported in place.  Shims:

 1. models.stylegan2.op.{fused_act,upfirdn2d} JIT-build CUDA extensions at import and have
    no CPU branch.  We pre-register stand-in modules: `upfirdn2d` executes the reference's
    own (dead, F never imported) `upfirdn2d_native` text from models/stylegan2/op/upfirdn2d.py
    with `F` injected; `fused_leaky_relu` is the formula of fused_bias_act_kernel.cu:26-47.
 2. swap_face_fine/face_parsing/model.py calls .cuda() at import -> identity on CPU.
 3. Resnet18.init_weight downloads weights -> model_zoo.load_url returns {}.
 4. BicubicDownSample(cuda=True) default -> callers set `.cuda = ''`.
"""
from __future__ import annotations

import ast
import os
import sys
import types

import torch
import torch.nn.functional as F
from torch import nn

REF_ROOT = os.environ.get("E4S_REF", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "models", "stylegan2"))


def _load_native_upfirdn2d():
    src_path = os.path.join(REF_ROOT, "models", "stylegan2", "op", "upfirdn2d.py")
    tree = ast.parse(open(src_path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "upfirdn2d_native"][0]
    ns = {"F": F, "torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), src_path, "exec"), ns)
    return ns["upfirdn2d_native"]


_installed = False


def install():
    """Idempotently put the reference on sys.path with the CPU shims in place."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    native = _load_native_upfirdn2d()

    def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
        # same argument mapping as models/stylegan2/op/upfirdn2d.py:87-121,142-147
        b, c, h, w = input.shape
        out = native(input.reshape(-1, h, w, 1), kernel, up, up, down, down, pad[0], pad[1], pad[0], pad[1])
        return out.reshape(b, c, out.shape[1], out.shape[2])

    def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
        shape = [1, -1] + [1] * (input.ndim - 2)
        return F.leaky_relu(input + bias.view(*shape), negative_slope) * scale

    class FusedLeakyReLU(nn.Module):
        def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
            super().__init__()
            self.bias = nn.Parameter(torch.zeros(channel))
            self.negative_slope, self.scale = negative_slope, scale

        def forward(self, input):
            return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)

    m_act = types.ModuleType("models.stylegan2.op.fused_act")
    m_act.FusedLeakyReLU, m_act.fused_leaky_relu = FusedLeakyReLU, fused_leaky_relu
    m_up = types.ModuleType("models.stylegan2.op.upfirdn2d")
    m_up.upfirdn2d = upfirdn2d
    sys.modules["models.stylegan2.op.fused_act"] = m_act
    sys.modules["models.stylegan2.op.upfirdn2d"] = m_up

    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    import torch.utils.model_zoo as model_zoo
    model_zoo.load_url = lambda *a, **k: {}
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _installed = True


def generator_cls():
    install()
    from models.stylegan2.model import Generator
    return Generator


def stylegan_module():
    install()
    import models.stylegan2.model as m
    return m


def net3_cls():
    install()
    from models.networks import Net3
    return Net3


def encoder_cls():
    install()
    from models.encoders.psp_encoders import FSEncoder_PSP
    return FSEncoder_PSP


def bisenet_module():
    install()
    import swap_face_fine.face_parsing.model as m
    return m


def bicubic_cls():
    """BicubicDownSample lives in face_parsing_demo.py, which imports cv2/datasets; exec only the class."""
    install()
    path = os.path.join(REF_ROOT, "swap_face_fine", "face_parsing", "face_parsing_demo.py")
    tree = ast.parse(open(path).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "BicubicDownSample"][0]
    ns = {"torch": torch, "nn": nn, "F": F}
    exec(compile(ast.Module(body=[cls], type_ignores=[]), path, "exec"), ns)
    return ns["BicubicDownSample"]


def seg19_to_seg12_fn():
    install()
    path = os.path.join(REF_ROOT, "datasets", "dataset.py")
    tree = ast.parse(open(path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef)
          and n.name == "__ffhq_masks_to_faceParser_mask_detailed"][0]
    import numpy as np
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns["__ffhq_masks_to_faceParser_mask_detailed"]


def net3_opts(out_size=1024, remaining_layer_idx=13, num_seg_cls=12):
    """Fields Net3 reads (options/our_swap_face_pipeline_options.py:12-18,50)."""
    return types.SimpleNamespace(fsencoder_type="psp", remaining_layer_idx=remaining_layer_idx,
                                 num_seg_cls=num_seg_cls, out_size=out_size, train_G=False,
                                 start_from_latent_avg=True, learn_in_w=False)

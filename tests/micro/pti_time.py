"""Timing aid: one PTI step of the drop-in at 1024^2, batch 1 (reference training/video_swap_ft_coach.py:262-299): stored style vectors ->
cal_style_codes -> gen_img -> L2 loss -> backward -> Adam step, CUDA events; and the same forward in eval() mode for scale."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from e4s2024_b200 import synth
from e4s2024_b200.networks import Net3
from e4s2024_b200.stylegan2 import grad as GR
from oracle.ref_shims import net3_opts

net = Net3(net3_opts(out_size=1024, remaining_layer_idx=13))
synth.synth_module_weights(net, seed=9)
net = net.cuda()
net.latent_avg = synth.randn("net3.latent_avg", (18, 512), 9, 0.1).cuda()
mask = synth.onehot(synth.make_labels("face", 1, 12, 512, seed=1), 12).cuda()
sv = synth.randn("pti.sv", (1, 12, 1280), 3).cuda()
target = synth.smooth_image("pti.target", 1, 1024, 5).cuda()
res = {}
for engine in ("tc", "f32"):
    GR.set_train_engine(engine)
    net.train()
    opt = torch.optim.Adam([p for p in net.parameters() if p.requires_grad], lr=1e-6)

    def step():
        img = net.gen_img(None, net.cal_style_codes(sv), mask, randomize_noise=False)[0]
        loss = torch.nn.functional.mse_loss(img, target)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    step(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record(); torch.cuda.synchronize()
    res[f"pti_step_ms_{engine}"] = e0.elapsed_time(e1) / 3
    res[f"peak_mem_gb_{engine}"] = torch.cuda.max_memory_allocated() / 2 ** 30
GR.set_train_engine("tc")
net.eval()
with torch.no_grad():
    net.gen_img(None, net.cal_style_codes(sv), mask, randomize_noise=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        net.gen_img(None, net.cal_style_codes(sv), mask, randomize_noise=False)
    e1.record(); torch.cuda.synchronize()
res["eval_forward_ms"] = e0.elapsed_time(e1) / 5
print(json.dumps(res))

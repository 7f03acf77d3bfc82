"""Timing aid: the full swap hot path (BASELINE config 5 per GPU): parse -> one-hot -> Net3 (encoder + MLPs + generator)
at B faces, stage by stage, CUDA events, synthetic weights."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from e4s2024_b200 import synth, _lib as L, engine as E
from e4s2024_b200.networks import Net3
from e4s2024_b200.face_parsing.face_parsing_demo import FaceParser
from oracle.ref_shims import net3_opts

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
net = Net3(net3_opts(out_size=1024, remaining_layer_idx=13))
synth.synth_module_weights(net, seed=9)
net = net.cuda()
net.latent_avg = synth.randn("net3.latent_avg", (18, 512), 9, 0.1).cuda()
parser = FaceParser(seg_ckpt=None, size=1024, device="cuda")
synth.synth_module_weights(parser.seg, seed=10)
parser.seg.cuda()
img = synth.smooth_image("swap.img", B, 1024, 13).cuda()          # [-1, 1]
img01 = (img + 1) / 2


def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r


ms_parse, lab = t(lambda: parser.parse_batch(img01))
ms_onehot, mask = t(lambda: L.labels_to_onehot(lab, 12))
ms_enc, (vec, _) = t(lambda: net.get_style_vectors(img, mask))
ms_codes, codes = t(lambda: net.cal_style_codes(vec))
ms_gen, _ = t(lambda: net.gen_img(None, codes, mask, randomize_noise=False))
ms_all, _ = t(lambda: net(img, L.labels_to_onehot(parser.parse_batch(img01), 12), randomize_noise=False))
print(json.dumps({"batch": B, "ms": {"parse": ms_parse, "onehot": ms_onehot, "encoder": ms_enc, "mlps": ms_codes, "generator": ms_gen, "all": ms_all},
                  "faces_per_s": B / ms_all * 1e3}))
if os.environ.get("E4S_NCU"):          # ncu --profile-from-start off: one clean full-path pass
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    net(img, L.labels_to_onehot(parser.parse_batch(img01), 12), randomize_noise=False)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()

"""Experiment: FaceParser.parse_batch (B=16, 1024^2 in) in the three BiSeNet modes: time, and label agreement with the exact-fp32 mode."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from e4s2024_b200 import synth
from e4s2024_b200.face_parsing import resnet as rn
from e4s2024_b200.face_parsing.face_parsing_demo import FaceParser
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
parser = FaceParser(seg_ckpt=None, size=1024, device="cuda")
synth.synth_module_weights(parser.seg, seed=10)
parser.seg.cuda()
img01 = ((synth.smooth_image("swap.img", B, 1024, 13) + 1) / 2).cuda()
res, labs = {}, {}
for mode in ("f32", "tc", "tc16"):
    rn.set_bisenet_engine(mode)
    parser.parse_batch(img01); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lab = parser.parse_batch(img01)
    e1.record(); torch.cuda.synchronize()
    labs[mode] = lab
    res[mode] = {"ms": e0.elapsed_time(e1) / 5}
for mode in ("tc", "tc16"):
    res[mode]["labels_differing_from_f32"] = int((labs[mode] != labs["f32"]).sum())
res["labels_total"] = int(labs["f32"].numel())
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/parser_modes.json", "w"))

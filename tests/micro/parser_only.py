"""Profiling aid: one batched FaceParser.parse_batch (BiSeNet on the exact-fp32 engine), B=16."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from e4s2024_b200 import synth
from e4s2024_b200.face_parsing.face_parsing_demo import FaceParser
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
parser = FaceParser(seg_ckpt=None, size=1024, device="cuda")
synth.synth_module_weights(parser.seg, seed=10)
parser.seg.cuda()
img01 = ((synth.smooth_image("swap.img", B, 1024, 13) + 1) / 2).cuda()
parser.parse_batch(img01); torch.cuda.synchronize()
torch.cuda.profiler.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); parser.parse_batch(img01); e1.record(); torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("parse ms", e0.elapsed_time(e1))

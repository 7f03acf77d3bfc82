"""Where does the end-to-end step lose time against the resident step?  HostPipeline variants at the bench workload."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench as B
from e4s2024_b200.serving import HostPipeline
from e4s2024_b200.utils.torch_utils import labelMap2OneHot
dev = torch.device("cuda", 0)
G = B.build_generator(dev)
latent_h, mask_h = B.make_generator_inputs(B.BATCH, seed=1)
latent_h = latent_h.pin_memory()
labels_h = mask_h.argmax(1, keepdim=True).to(torch.uint8).pin_memory()
out_h = torch.empty(B.BATCH, 3, B.SIZE, B.SIZE).pin_memory()
mask_d = mask_h.to(dev); latent_d = latent_h.to(dev); labels_d = labels_h.to(dev)
STEPS = 20

def timed(fn):
    fn(3); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); fn(STEPS); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / STEPS, (time.perf_counter() - t0) * 1e3 / STEPS

def resident(n):
    for _ in range(n):
        G([latent_d], None, mask_d, input_is_latent=True, randomize_noise=False)

def resident_onehot(n):
    for _ in range(n):
        G([latent_d], None, labelMap2OneHot(labels_d, B.K), input_is_latent=True, randomize_noise=False)

def make(pipe_out, h2d=True):
    fn = lambda lat, lab: G([lat], None, labelMap2OneHot(lab, B.K), input_is_latent=True, randomize_noise=False)[0]
    pipe = HostPipeline(fn, dev)
    def run(n):
        for _ in range(n):
            pipe.submit((latent_h, labels_h), out_h if pipe_out else None)
        pipe.drain()
    return run

def d2h_only(n):
    img = torch.empty(B.BATCH, 3, B.SIZE, B.SIZE, device=dev)
    for _ in range(n):
        out_h.copy_(img, non_blocking=True)

for name, fn in (("resident", resident), ("resident + on-device one-hot", resident_onehot), ("pipeline, no D2H", make(False)),
                 ("pipeline, full", make(True)), ("D2H copy alone (201 MB)", d2h_only)):
    ms, wall = timed(fn)
    print(f"{name:32s} {ms:7.2f} ms/step (wall {wall:7.2f})")

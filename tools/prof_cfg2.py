"""Dev tool: a few device-resident launches of one BASELINE workload (cfg1..cfg4) for timing and ncu:
    python tools/prof_cfg2.py cfg2 16 [fast|exact|f16|fp32]    # prints the kernel(s) used and us/image"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rust-image-transform_b200"))
import torch
import imagekit_cuda as ik
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
shapes = {"cfg2": (3840, 2160, 4, 1920, 1080, 4), "cfg3": (4032, 3024, 3, 400, 300, 4), "cfg1": (1920, 1080, 3, 400, 225, 4),
          "cfg4": (1920, 1080, 3, 3840, 2160, 2)}
sw, sh, ch, dw, dh, filt = shapes[wl]
ctx = ik.Context([0])
if len(sys.argv) > 3:
    ctx.set_mode({"fast": ik.MODE_FAST, "exact": ik.MODE_EXACT, "f16": ik.MODE_FAST_F16, "fp32": ik.MODE_FAST_FP32}[sys.argv[3]])
src = torch.randint(0, 256, (batch, sh, sw, ch), dtype=torch.uint8, device="cuda")
dst = torch.zeros((batch, dh, dw, ch), dtype=torch.uint8, device="cuda")
jobs = [(src[i].data_ptr(), sw, sh, sw * ch, dst[i].data_ptr(), dw, dh, dw * ch, ch, filt) for i in range(batch)]
b = ctx.prepare_batch(0, jobs)
s = torch.cuda.Stream()
for _ in range(4):
    b.launch(s.cuda_stream)
s.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
for _ in range(10):
    b.launch(s.cuda_stream)
e1.record(s)
s.synchronize()
print(b.describe()); print(wl, "batch", batch, "ms/launch", e0.elapsed_time(e1) / 10, "us/image", e0.elapsed_time(e1) / 10 / batch * 1e3)

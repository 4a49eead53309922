import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rust-image-transform_b200"))
import torch, imagekit_cuda as ik
wl = sys.argv[1]; batch = int(sys.argv[2])
shapes = {"cfg2": (3840, 2160, 4, 1920, 1080, 4), "cfg3": (4032, 3024, 3, 400, 300, 4)}
sw, sh, ch, dw, dh, filt = shapes[wl]
ctx = ik.Context([0])
src = torch.randint(0, 256, (batch, sh, sw, ch), dtype=torch.uint8, device="cuda")
dst = torch.zeros((batch, dh, dw, ch), dtype=torch.uint8, device="cuda")
jobs = [(src[i].data_ptr(), sw, sh, sw * ch, dst[i].data_ptr(), dw, dh, dw * ch, ch, filt) for i in range(batch)]
b = ctx.prepare_batch(0, jobs)
s = torch.cuda.Stream()
for _ in range(3): b.launch(s.cuda_stream)
s.synchronize()
L = ik._lib.load()
out = (C.c_ulonglong * 8)()
L.ikc_debug_timing(out, 1)
b.launch(s.cuda_stream); s.synchronize()
L.ikc_debug_timing(out, 0)
v = list(out); n = v[7]
names = ["V", "bar V->H", "H (all)", "H fast loop", "bar H->V", "-", "total"]
print(wl, "CTAs", n, {k: round(x / n) for k, x in zip(names, v[:7])}, "cycles per CTA (thread 0)")

"""Writes the golden inputs as raw rasters + a manifest for oracle/ref_harness (the Rust project that runs the real
`image` 0.25.8 over them).  Output: oracle/ref_harness/inputs/{cases.txt, <name>.src.bin}.  Not committed: regenerate."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, os.path.join(HERE, ".."))
from conftest import checker, photo_like, splitmix_noise  # noqa: E402

OUT = os.path.join(ROOT, "oracle", "ref_harness", "inputs")

# the self-generated regression cases (tests/golden/resize_golden.npz) ...
g = np.load(os.path.join(HERE, "resize_golden.npz"))
names = sorted({k.split("/")[0] for k in g.files})
cases = [(n, g[n + "/src"], int(g[n + "/meta"][0]), int(g[n + "/meta"][1]), int(g[n + "/meta"][2]), "exact") for n in names]
# ... plus shapes of the kind the kernels specialise for (2:1 Rgba, thumbnails, 2x upscale, 16-row seams) and the
# fit-within semantics of DynamicImage::resize that the reference's resize_image calls
extra = [
    ("rgba_2to1_l3", splitmix_noise((96, 128, 4), image_id=201), 64, 48, 4, "exact"),
    ("rgba_2to1_edges_l3", checker((100, 200, 4)), 100, 50, 4, "exact"),
    ("rgb_thumb_l3", photo_like((151, 201, 3)), 40, 30, 4, "exact"),
    ("rgb_up2_catmull", splitmix_noise((40, 60, 3), image_id=202), 120, 80, 2, "exact"),
    ("rgb_odd_l3", splitmix_noise((97, 131, 3), image_id=203), 53, 41, 4, "exact"),
    ("rgb_fit_w400", photo_like((108, 192, 3)), 40, 1000, 4, "fit"),
    ("rgba_fit_h", splitmix_noise((90, 160, 4), image_id=204), 1000, 45, 4, "fit"),
]
cases += extra
os.makedirs(OUT, exist_ok=True)
with open(os.path.join(OUT, "cases.txt"), "w") as f:
    f.write("# name h w channels dw dh filter mode\n")
    for name, src, dw, dh, filt, mode in cases:
        s = src if src.ndim == 3 else src[:, :, None]
        h, w, c = s.shape
        np.ascontiguousarray(s, np.uint8).tofile(os.path.join(OUT, name + ".src.bin"))
        f.write(f"{name} {h} {w} {c} {dw} {dh} {filt} {mode}\n")
print("wrote", len(cases), "inputs to", os.path.normpath(OUT))

"""Packs the outputs of oracle/ref_harness (the real `image` 0.25.8) into tests/golden/reference_image_0_25_8.npz:
per case <name>/src, <name>/meta = [dw, dh, filter, fit?], <name>/ref.  Commit the .npz; the pinned-parity tests then run."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
IN = os.path.join(HERE, "..", "..", "oracle", "ref_harness", "inputs")
out = {}
for line in open(os.path.join(IN, "cases.txt")):
    t = line.split()
    if len(t) != 8 or t[0].startswith("#"):
        continue
    name, h, w, c, dw, dh, filt = t[0], *map(int, t[1:7])
    fit = t[7] == "fit"
    ref_path = os.path.join(IN, name + ".ref.bin")
    if not os.path.exists(ref_path):
        raise SystemExit(f"{ref_path} missing: run `cargo run --release -- inputs` in oracle/ref_harness first")
    ow, oh = (dw, dh) if not fit else map(int, open(os.path.join(IN, name + ".ref.dims")).read().split())
    out[name + "/src"] = np.fromfile(os.path.join(IN, name + ".src.bin"), np.uint8).reshape(h, w, c)
    out[name + "/meta"] = np.array([dw, dh, filt, int(fit)], np.int32)
    out[name + "/ref"] = np.fromfile(ref_path, np.uint8).reshape(oh, ow, c)
np.savez_compressed(os.path.join(HERE, "reference_image_0_25_8.npz"), **out)
print("packed", len(out) // 3, "reference cases")

"""C-ABI checks that need no GPU: the library loads, exports every symbol include/imagekit_cuda.h
declares, its host-built weight tables equal the oracle's bit for bit, and it fails loudly (no CPU
fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "imagekit_cuda.h")


def _declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"IKC_API[^;(]*?\b(ikc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(ik):
    L = ik._lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert sorted(ik._lib.EXPORTS) == declared
    assert L.ikc_version() == 1


def test_job_struct_layout_matches_header(ik):
    # struct ikc_job: 2 pointers, 4 u32, 2 size_t, 4 i32 = 64 bytes on LP64
    assert ctypes.sizeof(ik._lib.Job) == 64
    assert ik._lib.Job.sw.offset == 16 and ik._lib.Job.src_pitch.offset == 32 and ik._lib.Job.channels.offset == 48


@pytest.mark.parametrize("filt,n_in,n_out", [(4, 2160, 1080), (4, 3024, 300), (4, 1080, 225), (2, 1080, 2160),
                                             (3, 777, 123), (1, 50, 500), (0, 100, 37), (4, 600, 1), (4, 2, 200),
                                             (4, 4032, 400), (2, 640, 640), (4, 1, 1)])
def test_weight_tables_bit_identical_to_oracle(ik, oracle, filt, n_in, n_out):
    from imagekit_cuda import engine
    l1, c1, w1 = engine.pass_table(filt, n_in, n_out)
    l2, c2, w2 = oracle.pass_table(filt, n_in, n_out)
    m = min(w1.shape[1], w2.shape[1])
    assert np.array_equal(l1, l2) and np.array_equal(c1, c2)
    assert c1.max() <= m
    assert np.array_equal(w1[:, :m].view(np.uint32), w2[:, :m].view(np.uint32))


def test_no_cpu_fallback_without_gpu(ik):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ik.ImageKitError) as e:
        ik.Context()
    assert "no CPU fallback" in str(e.value)
    img = ik.DynamicImage.new_rgb8(80, 60)
    assert ik.resize_image(img, None, None) is img          # transform.rs:67-69 needs no device
    with pytest.raises(ik.ImageKitError):
        ik.resize_image(img, 40, None)


def test_error_paths_without_context(ik):
    L = ik._lib.load()
    assert L.ikc_resize_u8(None, None, 1, 1, 1, 1, None, 1, 1, 1, 4) == ik._lib.ERR_INVALID_ARG
    assert L.ikc_device_count(None) == 0
    assert L.ikc_batch_launch(None, None) == ik._lib.ERR_INVALID_ARG
    assert b"null" in L.ikc_last_error()


# ---- what the planner derives for the specialised kernels (host logic; checked against the oracle's tables) ----
@pytest.mark.parametrize("filt,n_in,n_out", [(4, 2160, 1080), (4, 3840, 1920), (4, 1200, 300), (4, 3024, 300),
                                             (4, 1080, 225), (2, 500, 250), (4, 777, 388), (4, 64, 32)])
def test_uniform_stretch_of_downscale_passes(ik, oracle, filt, n_in, n_out):
    from imagekit_cuda import engine
    info = engine.pass_info(filt, n_in, n_out)
    left, count, w = oracle.pass_table(filt, n_in, n_out)
    right = left.astype(np.int64) + count
    # ring size = most windows over any source index
    cover = np.zeros(n_in + 1, np.int64)
    np.add.at(cover, left, 1)
    np.add.at(cover, right, -1)
    assert info["ring_k"] == int(np.cumsum(cover)[:-1].max())
    lo, hi, step = info["uni_lo"], info["uni_hi"], info["uni_step"]
    if n_in % n_out == 0 and n_out >= 8 * info["ring_k"]:
        assert step == n_in // n_out and hi - lo >= n_out - 8, (lo, hi, step)   # all but the clamped borders
    if step:
        k = info["ring_k"]
        assert 1 <= lo < hi <= n_out
        for o in range(lo, hi):
            assert right[o] - right[o - 1] == step and count[o] == k * step
            if o >= k:
                assert left[o] == right[o - k]
            assert np.array_equal(w[o, :count[o]].view(np.uint32), w[lo, :count[lo]].view(np.uint32))
        # maximal: the outputs just outside break one of the conditions
        for o in (lo - 1, hi):
            if 1 <= o < n_out:
                same = (right[o] - right[o - 1] == step and count[o] == k * step and (o < k or left[o] == right[o - k]) and
                        np.array_equal(w[o, :count[o]].view(np.uint32), w[lo, :count[lo]].view(np.uint32)))
                assert not same
    else:
        assert n_in % n_out != 0 or n_out < 8 * info["ring_k"]


@pytest.mark.parametrize("filt,n_in", [(2, 1080), (2, 1920), (4, 400), (1, 77), (0, 64), (3, 200), (2, 1), (4, 3)])
def test_exact_2x_upscale_frame(ik, oracle, filt, n_in):
    from imagekit_cuda import engine
    info = engine.pass_info(filt, n_in, 2 * n_in)
    left, count, _ = oracle.pass_table(filt, n_in, 2 * n_in)
    base = np.arange(2 * n_in) // 2 + info["up2_off"]
    assert info["up2_taps"] >= 1
    assert np.all(left >= base) and np.all(left + count <= base + info["up2_taps"])      # every window fits its frame
    assert np.any(left == base) and np.any(left + count == base + info["up2_taps"])      # and the frame is tight
    assert 0 <= info["up2_uni_lo"] < info["up2_uni_hi"] <= n_in
    assert engine.pass_info(filt, n_in, 2 * n_in + 1)["up2_taps"] == 0                   # not an exact 2x upscale


def test_header_is_plain_c_and_links(tmp_path):
    """The boundary is a C ABI: include/imagekit_cuda.h must compile as C99 and a C program must link the library."""
    import os, shutil, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "rust-image-transform_b200")
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text('#include "imagekit_cuda.h"\n'
                   "int main(void) {\n"
                   "    ikc_pass_info_t info; ikc_job job; uint32_t tw = 0, th = 0;\n"
                   "    (void)job;\n"
                   "    if (ikc_version() <= 0) return 1;\n"
                   "    if (ikc_target_dims(1920, 1080, 1, 400, 0, 0, &tw, &th) != 0 || tw != 400 || th != 225) return 2;\n"
                   "    if (ikc_pass_info(IKC_FILTER_LANCZOS3, 2160, 1080, &info) != 0 || info.ring_k != 6 || info.uni_step != 2) return 3;\n"
                   "    return IKC_CHANNELS(4, 3) == (4 | (3 << 8)) ? 0 : 4;\n"
                   "}\n")
    exe = tmp_path / "abi"
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                           str(src), "-o", str(exe), "-L", lib_dir, "-limagekit_cuda", f"-Wl,-rpath,{lib_dir}"])
    assert subprocess.call([str(exe)]) == 0


def test_struct_layouts_match_the_header(ik, tmp_path):
    """The ctypes mirrors (and, field for field, the Rust crate's ffi.rs) must lay out ikc_stats_t, ikc_job and
    ikc_pass_info_t exactly as the C header does: sizes and the offsets of the last fields, printed by a C program."""
    import os, re, shutil, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "layout.c"
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "imagekit_cuda.h"\n'
                   "int main(void) {\n"
                   '    printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(ikc_stats_t), offsetof(ikc_stats_t, staging_trims),\n'
                   "           offsetof(ikc_stats_t, submit_batches), sizeof(ikc_job), offsetof(ikc_job, device), sizeof(ikc_pass_info_t));\n"
                   "    return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.check_call([gcc, "-std=c99", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)])
    c_stats, c_trims, c_submit, c_job, c_job_dev, c_info = (int(x) for x in subprocess.check_output([str(exe)]).split())
    S, J, P = ik._lib.Stats, ik._lib.Job, ik._lib.PassInfo
    assert (ctypes.sizeof(S), S.staging_trims.offset, S.submit_batches.offset) == (c_stats, c_trims, c_submit)
    assert (ctypes.sizeof(J), J.device.offset) == (c_job, c_job_dev)
    assert ctypes.sizeof(P) == c_info
    # the Rust mirror: same field names in the same order as the header's struct (all u64)
    hdr = open(os.path.join(root, "include", "imagekit_cuda.h")).read()
    body = hdr[hdr.index("typedef struct ikc_stats_t {"):hdr.index("} ikc_stats_t;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    c_fields = [f.strip() for decl in re.findall(r"uint64_t([^;]*);", body) for f in decl.split(",")]
    rs = open(os.path.join(root, "rust-image-transform_b200", "crate", "src", "ffi.rs")).read()
    rs_body = rs[rs.index("pub struct ikc_stats_t {"):]
    rs_body = rs_body[:rs_body.index("}")]
    rs_fields = re.findall(r"pub (\w+): u64", rs_body)
    assert c_fields == rs_fields == [n for n, _ in S._fields_]


def test_rust_bindings_name_only_exported_entry_points(ik):
    """Every extern "C" function the Rust crate binds (crate/src/ffi.rs; it cannot be compiled here) is declared in the
    header and exported by the library, with the same number of parameters."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "imagekit_cuda.h")).read(), flags=re.S)
    c_decl = {m.group(1): m.group(2) for m in re.finditer(r"IKC_API[^;(]*?\b(ikc_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S)}
    rs = open(os.path.join(root, "rust-image-transform_b200", "crate", "src", "ffi.rs")).read()
    rs = re.sub(r"//[^\n]*", "", rs)
    rs_decl = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (ikc_\w+)\s*\(([^)]*)\)", rs, flags=re.S)}
    assert rs_decl, "no bindings found in ffi.rs"
    lib = ik._lib.load()

    def n_params(text):
        text = text.strip()
        return 0 if text in ("", "void") else len([p for p in text.split(",") if p.strip()])
    for name, params in rs_decl.items():
        assert name in c_decl, f"{name} is bound in ffi.rs but not declared in imagekit_cuda.h"
        assert hasattr(lib, name), f"{name} is not exported by libimagekit_cuda.so"
        assert n_params(params) == n_params(c_decl[name]), (name, params, c_decl[name])


# ---- band form of a downscale pass: the weight tiles of the tensor-core vertical pass (host logic) ----
@pytest.mark.parametrize("filt,n_in,n_out", [(4, 2160, 1080), (4, 3024, 300), (4, 1080, 225), (4, 1080, 1080), (4, 1000, 999),
                                             (4, 777, 388), (2, 500, 250), (1, 640, 123), (3, 333, 100), (0, 100, 37),
                                             (4, 16, 16), (4, 17, 1), (4, 4032, 400)])
def test_band_tiles_reproduce_the_pass(ik, oracle, filt, n_in, n_out):
    from imagekit_cuda import engine
    band = engine.pass_band(filt, n_in, n_out)
    assert band is not None
    n, gbase, hi, lo = band
    assert n in (32, 48)
    chunks = hi.shape[0]
    assert chunks == (n_in + 15) // 16 and gbase[-1] == (n_out + 15) // 16
    assert np.all(np.diff(gbase[:-1]) >= 0)
    left, count, w = oracle.pass_table(filt, n_in, n_out)
    # dense matrix rebuilt from the tiles: (hi + lo) * 2^-14, accumulated where the tiles put it
    dense = np.zeros((n_out + 64, chunks * 16), np.float64)
    for c in range(chunks):
        rows = slice(16 * gbase[c], 16 * gbase[c] + n)
        dense[rows, 16 * c:16 * c + 16] += (hi[c].astype(np.float64) + lo[c].astype(np.float64)) / 16384.0
    want = np.zeros_like(dense)
    for o in range(n_out):
        want[o, left[o]:left[o] + count[o]] = w[o, :count[o]]
    assert np.all(dense[n_out:] == 0) and np.all(dense[:, n_in:] == 0)
    # hi + lo carries the weight to ~2^-22 of its own size (f16 hi: 11 bits, lo: 11 more)
    err = np.abs(dense - want)
    assert np.all(err <= np.abs(want) * 2.0 ** -21 + 2.0 ** -38), float(err.max())
    # a chunk's window never reaches a group that earlier chunks had already finished
    right = left.astype(np.int64) + count
    for c in range(chunks):
        done_before = [g for g in range(gbase[-1]) if right[min(16 * g + 15, n_out - 1)] <= 16 * c]
        assert all(g < gbase[c] for g in done_before)


def test_upscale_pass_has_no_band_form(ik):
    from imagekit_cuda import engine
    assert engine.pass_band(4, 100, 200) is None
    assert engine.pass_band(4, 8, 4) is None      # fewer source indices than one chunk


# ---- 8-bit band form: the integer weight tiles of the i8 tensor-core vertical pass (host logic) ----
@pytest.mark.parametrize("filt,n_in,n_out", [(4, 2160, 1080), (4, 3024, 300), (4, 1080, 225), (4, 777, 388), (2, 500, 250),
                                             (1, 640, 123), (3, 333, 100), (0, 100, 37), (4, 4032, 400), (4, 64, 32),
                                             (4, 33, 1)])
def test_band8_tiles_are_the_quantised_pass(ik, oracle, filt, n_in, n_out):
    from imagekit_cuda import engine
    band = engine.pass_band8(filt, n_in, n_out)
    assert band is not None
    limbs, shift, gbase, dig = band
    chunks = dig.shape[0]
    assert limbs == 2 and chunks == (n_in + 31) // 32 and gbase[-1] == (n_out + 7) // 8
    assert np.all(np.diff(gbase[:-1]) >= 0)
    assert dig[:, 1:].min() >= -128 and dig[:, 1:].max() <= 127        # low digits (base 256)
    left, count, w = oracle.pass_table(filt, n_in, n_out)
    # integer weights rebuilt from the digits, accumulated where the tiles put them (window position p = output 8 * gbase + p)
    W = np.zeros((n_out + 64, chunks * 32), np.int64)
    for c in range(chunks):
        val = np.zeros((32, 32), np.int64)
        for d in range(limbs):
            val = val * 256 + dig[c, d].astype(np.int64)
        for pos in range(32):
            if not val[pos].any():
                continue
            o = 8 * gbase[c] + pos
            W[o, 32 * c:32 * c + 32] += val[pos]
    assert not W[n_out:].any() and not W[:, n_in:].any()
    scale = float(2 ** shift)
    for o in range(n_out):
        row = W[o]
        nz = np.flatnonzero(row)
        assert nz.size == 0 or (nz[0] >= left[o] and nz[-1] < left[o] + count[o])
        assert row.sum() == 2 ** shift                                  # a flat area stays exactly flat
        err = np.abs(row[left[o]:left[o] + count[o]] / scale - w[o, :count[o]].astype(np.float64))
        assert err.max() <= (count[o] / 2 + 1) / scale                  # half an LSB, plus the sum correction on one tap
    assert np.abs(W).max() <= 127 * 256 + 127


@pytest.mark.parametrize("filt,n_in,n_out", [(4, 2160, 1080), (4, 777, 388), (2, 500, 250), (4, 64, 32), (4, 700, 400), (4, 1300, 600)])
def test_band8t_tiles_hold_the_same_integers_as_band8(ik, filt, n_in, n_out):
    """The row-band tiles (A operand of banded8t.cu) and the chunk-window tiles (B operand of banded8.cu) are two layouts of
    one integer weight matrix."""
    from imagekit_cuda import engine
    limbs, shift, gbase, dig = engine.pass_band8(filt, n_in, n_out)
    band_t = engine.pass_band8t(filt, n_in, n_out)
    assert band_t is not None
    nc, k_lo, t, rows = band_t
    bands = t.shape[0]
    assert 96 <= rows <= 128 and bands == (n_out + rows - 1) // rows and 1 <= nc <= 10
    if (n_in, n_out) == (2160, 1080):
        assert (rows, nc, bands) == (120, 8, 9)          # exactly 2:1: 120-row bands span 8 chunks, 128-row bands 9
    W8 = np.zeros((n_out + 64, dig.shape[0] * 32 + 512), np.int64)
    for c in range(dig.shape[0]):
        val = dig[c, 0].astype(np.int64) * 256 + dig[c, 1].astype(np.int64)
        for pos in np.flatnonzero(val.any(axis=1)):
            W8[8 * gbase[c] + pos, 32 * c:32 * c + 32] += val[pos]
    WT = np.zeros_like(W8)
    for r in range(bands):
        for c in range(nc):
            val = t[r, c, 0].astype(np.int64) * 256 + t[r, c, 1].astype(np.int64)
            y0 = k_lo[r] + 32 * c
            live = min(rows, WT.shape[0] - rows * r)
            assert not val[live:].any()
            WT[rows * r:rows * r + live, y0:y0 + 32] += val[:live]
    assert np.array_equal(W8[:n_out], WT[:n_out]) and not WT[n_out:].any()


def test_band8t_needs_a_ratio_near_two(ik):
    from imagekit_cuda import engine
    assert engine.pass_band8t(4, 3024, 300) is None     # a band of 128 outputs spans far more than 10 chunks
    assert engine.pass_band8t(4, 100, 150) is None      # an upscale that is not 2x


@pytest.mark.parametrize("filt,n_in", [(2, 1080), (4, 100), (1, 77), (3, 640)])
def test_band8t_of_a_2x_upscale_is_the_quantised_pass(ik, oracle, filt, n_in):
    """Exact 2x upscales get the row-band tiles too (banded8u.cu): three chunks per band, rows sum to a power of two, every
    integer weight within half a unit (plus the sum correction) of w * 2^shift."""
    from imagekit_cuda import engine
    n_out = 2 * n_in
    nc, k_lo, t, rows = engine.pass_band8t(filt, n_in, n_out)
    assert rows == 128 and 1 <= nc <= 4 and t.shape[0] == (n_out + 127) // 128
    left, count, w = oracle.pass_table(filt, n_in, n_out)
    W = np.zeros((t.shape[0] * 128, n_in + 256), np.int64)
    for r in range(t.shape[0]):
        for c in range(nc):
            val = t[r, c, 0].astype(np.int64) * 256 + t[r, c, 1].astype(np.int64)
            W[128 * r:128 * r + 128, k_lo[r] + 32 * c:k_lo[r] + 32 * c + 32] += val
    assert not W[n_out:].any() and not W[:, n_in:].any()
    total = int(W[0].sum())
    shift = total.bit_length() - 1
    assert total == 1 << shift and 10 <= shift <= 21
    for o in range(n_out):
        row = W[o]
        assert row.sum() == total
        nz = np.flatnonzero(row)
        assert nz[0] >= left[o] and nz[-1] < left[o] + count[o]
        err = np.abs(row[left[o]:left[o] + count[o]] / float(total) - w[o, :count[o]].astype(np.float64))
        assert err.max() <= (count[o] / 2 + 1) / float(total)


def test_band8_needs_a_narrow_chunk_window(ik):
    from imagekit_cuda import engine
    assert engine.pass_band8(4, 100, 200) is None       # upscale
    assert engine.pass_band8(4, 1080, 1080) is None     # 32 source rows touch more than 32 outputs
    assert engine.pass_band8(4, 1000, 700) is None

"""Algorithm-level simulation (numpy, float64 accumulate) of the fused kernel's ring schedule:
vertical ring with chunk pre-roll, horizontal ring with x-segments + head/tail carries.
Dev tool only: checks the index logic of the CUDA kernel against the oracle tables."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from oracle import oracle as o


def ring_tables(filt, n_in, n_out):
    left, cnt, w = o.pass_table(filt, n_in, n_out)
    left = left.astype(int); cnt = cnt.astype(int); right = left + cnt
    cover = np.zeros(n_in + 1, int)
    for i in range(n_out):
        cover[left[i]] += 1; cover[right[i]] -= 1
    K = int(np.cumsum(cover)[:-1].max())
    ring = np.zeros((n_in, K))
    for i in range(n_out):
        for t in range(cnt[i]):
            assert ring[left[i] + t, i % K] == 0
            ring[left[i] + t, i % K] = w[i, t]
    return left, right, ring, K


def v_pass_chunk(src, left, right, ring, K, oy0, oy1, GV):
    """emulate one CTA's vertical march for outputs [oy0, oy1); returns tmp rows + rows consumed"""
    assert oy0 % K == 0 and GV % K == 0
    n_out = len(left)
    acc = np.zeros((K,) + src.shape[1:])
    y = left[oy0]; y_first = y
    out = {}
    oo = oy0 - K
    def body(oo, y):
        for c in range(K):
            ov = oo + c
            yend = right[ov] if (0 <= ov < oy1) else y
            while y < yend:
                for j in range(K):
                    acc[j] += ring[y, j] * src[y]
                y += 1
            if oy0 <= ov < oy1:
                out[ov] = acc[c].copy()
            acc[c] = 0
        return y
    y = body(oo, y); oo += K
    g0 = oy0
    while g0 < oy1:
        for m in range(GV // K):
            y = body(oo, y); oo += K
        g0 += GV
    assert y == right[oy1 - 1], (y, right[oy1 - 1])
    return out, (y_first, y)


def h_pass_strip(tmp_row, left, right, ring, K, ox0, ox1, n_seg_max, max_count):
    """emulate horizontal ring for outputs [ox0,ox1) of one row with segments + carries"""
    xl, xr = left[ox0], right[ox1 - 1]
    n_seg = max(1, min(n_seg_max, (xr - xl) // (max_count + 2 * K + 2)))
    seglen = -(-(xr - xl) // n_seg)
    S = [min(xl + s * seglen, xr) for s in range(n_seg)] + [xr]
    final = {}
    heads = {}; tails = {}; ofirst = {}
    for s in range(n_seg):
        Ss, Se = S[s], S[s + 1]
        # first output ending after Ss
        of = max(0, ox0 - K + 1)
        while of < ox1 and right[of] <= Ss:
            of += 1
        ofirst[s] = of
        oo = (of // K) * K          # absolute residues: slot = output index mod K
        acc = np.zeros((K,) + tmp_row.shape[1:])
        x = Ss; heads[s] = []
        done = False
        while not done:
            for c in range(K):
                ov = oo + c
                if ov >= ox1:
                    done = True; break
                r_end = right[ov]
                xend = min(r_end, Se)
                while x < xend:
                    for j in range(K):
                        acc[j] += ring[x, j] * tmp_row[x]
                    x += 1
                if r_end > Se:
                    done = True; break
                if r_end > Ss and ov >= ox0:
                    if left[ov] >= Ss:
                        final[ov] = acc[c].copy()
                    else:
                        # head i must land on an already consumed pixel slot
                        assert x > Ss + len(heads[s]), "head slot not consumed yet"
                        heads[s].append((ov, acc[c].copy()))
                acc[c] = 0
            oo += K
        tails[s] = acc.copy()
        assert len(heads[s]) <= K
        assert Se - Ss >= 2 * K or n_seg == 1
    for s in range(1, n_seg):
        for (ov, part) in heads[s]:
            tot = part.copy()
            sg = s - 1
            while sg >= 0 and S[sg + 1] > left[ov]:
                tot += tails[sg][ov % K]
                sg -= 1
            final[ov] = tot
    return final, n_seg


def run(sw, sh, C, dw, dh, filt, rng, GV_rows=32, n_seg_max=8, strip_out=None, chunk=None):
    src = rng.integers(0, 256, (sh, sw, C)).astype(np.float64)
    lv, rv, ringv, KV = ring_tables(filt, sh, dh)
    lh, rh, ringh, KH = ring_tables(filt, sw, dw)
    GV = (GV_rows // KV) * KV
    # reference (float64) via tables
    _, cntv, wv = o.pass_table(filt, sh, dh); _, cnth, wh = o.pass_table(filt, sw, dw)
    tmp_ref = np.stack([sum(wv[i, t] * src[lv[i] + t] for t in range(cntv[i])) for i in range(dh)])
    out_ref = np.stack([sum(wh[i, t] * tmp_ref[:, lh[i] + t] for t in range(cnth[i])) for i in range(dw)], 1)
    chunk = chunk or dh
    chunk = max(GV, (chunk // GV) * GV)
    strip_out = strip_out or dw
    out = np.zeros((dh, dw, C))
    maxc_h = int((rh - lh).max())
    for oy0 in range(0, dh, chunk):
        oy1 = min(dh, oy0 + chunk)
        tmp, _ = v_pass_chunk(src, lv, rv, ringv, KV, oy0, oy1, GV)
        for oy in range(oy0, oy1):
            assert np.allclose(tmp[oy], tmp_ref[oy], atol=1e-9)
            for ox0 in range(0, dw, strip_out):
                ox1 = min(dw, ox0 + strip_out)
                fin, nseg = h_pass_strip(tmp[oy], lh, rh, ringh, KH, ox0, ox1, n_seg_max, maxc_h)
                for ox in range(ox0, ox1):
                    out[oy, ox] = fin[ox]
    err = np.abs(out - out_ref).max()
    print(f"{sw}x{sh}x{C}->{dw}x{dh} f={filt} KV={KV} KH={KH} GV={GV} chunk={chunk} strip={strip_out} err={err:.2e}")
    assert err < 1e-8


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    run(64, 48, 1, 32, 24, 4, rng, strip_out=16, chunk=12)
    run(200, 90, 1, 97, 41, 4, rng, strip_out=40, chunk=18)
    run(403, 77, 1, 40, 8, 4, rng, strip_out=13)
    run(300, 60, 2, 150, 30, 2, rng, strip_out=150, chunk=8)
    run(500, 70, 1, 250, 35, 3, rng, strip_out=123, chunk=14)
    run(640, 100, 1, 320, 33, 1, rng, strip_out=100)
    run(1000, 40, 1, 333, 13, 4, rng, strip_out=333)
    run(1280, 37, 1, 640, 19, 4, rng, strip_out=123)
    print("ring schedule OK")

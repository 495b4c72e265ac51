"""Index-level simulation of kernels_block2.cu (strips, windows, row pipeline) in float64 NumPy.

Not a test of the CUDA code: it replays the kernel's tile geometry (window offsets, the st.shared scatter of the
layer-1 epilogue into the layer-2 tile, row-block halos, ownership masks, residual ring addressing) and checks that
every output element equals a direct evaluation of conv+clip+pool -> conv+clip+pool -> join.  Run on the CPU.
"""
import sys
import numpy as np
import torch
import torch.nn.functional as F

STRIP, RESPX, RESROWS = 103, 112, 8


def direct(R2, w2, w3, A, B, C, out_side):
    def layer(x, w):
        y = F.conv2d(torch.from_numpy(x.transpose(2, 0, 1)[None]), torch.from_numpy(w.transpose(3, 2, 0, 1)))
        y = y.clamp(0, 1)
        y = F.avg_pool2d(y, 4, 1) * 16
        return y[0].numpy().transpose(1, 2, 0)
    P2 = layer(R2, w2)
    P3 = layer(P2, w3)
    S, SS = out_side, R2.shape[0]
    scale = np.float32(SS) / np.float32(S)
    out = np.zeros_like(P3)
    for y in range(S):
        fy = np.float32(y) * scale
        y0 = int(fy); y1 = min(y0 + 1, SS - 1); ty = float(fy - np.float32(y0))
        for x in range(S):
            fx = np.float32(x) * scale
            x0 = int(fx); x1 = min(x0 + 1, SS - 1); tx = float(fx - np.float32(x0))
            top = R2[y0, x0] + (R2[y0, x1] - R2[y0, x0]) * tx
            bot = R2[y1, x0] + (R2[y1, x1] - R2[y1, x0]) * tx
            out[y, x] = A * P3[y, x] + B * (top + (bot - top) * ty) + C
    return out


def tile_conv(stage_rows, w):
    """stage_rows: [3][128+2][32] (three input rows as dense 128-lane planes + 2 over-read lanes) -> conv row [128][32]"""
    acc = np.zeros((128, w.shape[3]))
    for dy in range(3):
        for dx in range(3):
            acc += stage_rows[dy][dx:dx + 128] @ w[dy, dx]
    return acc


def hpool(v):  # [128][C] per 32-lane window: lane l sums lanes l..l+3 of the same window (shuffle semantics)
    out = np.zeros_like(v)
    for q in range(4):
        w = v[32 * q:32 * q + 32]
        pad = np.concatenate([w, w[-1:].repeat(3, 0)])  # shfl_down beyond the warp returns the own value: garbage
        out[32 * q:32 * q + 32] = pad[0:32] + pad[1:33] + pad[2:34] + pad[3:35]
    return out


def sim(R2, w2, w3, A, B, C, out_side, rows_per_item):
    in_side = R2.shape[0]
    scale = np.float32(in_side) / np.float32(out_side)
    n_strips = -(-out_side // STRIP)
    out = np.full((out_side, out_side, 32), np.nan)
    writes = np.zeros((out_side, out_side), int)
    garbage = 1e3  # finite garbage for over-reads
    for strip in range(n_strips):
        x0 = min(strip * STRIP, out_side - STRIP)
        own_lo, own_hi = strip * STRIP, min((strip + 1) * STRIP, out_side)
        jb = int(np.float32(x0) * scale)
        for po0 in range(0, out_side, rows_per_item):
            npo = min(rows_per_item, out_side - po0)
            nconv3 = (npo + 3 + 1) & ~1
            nin2 = nconv3 + 2
            nconv2 = nin2 + 4
            nin1 = nconv2 + 2
            # ---- layer 1 input rows as windowed tiles
            def l1_row(r):
                t = np.full((130, 32), garbage)
                y = po0 + r
                for q in range(4):
                    for l in range(32):
                        x = x0 + 27 * q + l
                        t[32 * q + l] = R2[y, x] if (y < in_side and x < in_side) else garbage
                return t
            rows1 = [l1_row(r) for r in range(nin1)]
            conv2 = [np.clip(tile_conv(rows1[c:c + 3], w2), 0, 1) for c in range(nconv2)]
            # ---- layer-1 epilogue: P2 rows into layer-2 tiles via the sts scatter
            p2_tiles = [np.full((130, 32), np.nan) for _ in range(nin2)]
            for t in range(nconv2 // 2):
                for k, row in ((0, 2 * t - 3), (1, 2 * t - 2)):
                    if 0 <= row < nin2:
                        v = conv2[row] + conv2[row + 1] + conv2[row + 2] + conv2[row + 3]
                        hp = hpool(v)
                        for q in range(4):
                            for l in range(27):
                                ia = 101 + l if q == 3 else 32 * q + l
                                ib = -1
                                if l < 5:
                                    ib = {1: 27 + l, 2: 59 + l, 3: 91 + l}.get(q, -1)
                                elif q == 2 and l >= 22:
                                    ib = 74 + l
                                p2_tiles[row][ia] = hp[32 * q + l]
                                if ib >= 0:
                                    p2_tiles[row][ib] = hp[32 * q + l]
            for tl in p2_tiles:
                assert not np.isnan(tl[:128]).any(), "layer-2 tile lane never written"
                tl[128:] = garbage
            conv3 = [np.clip(tile_conv(p2_tiles[c:c + 3], w3), 0, 1) for c in range(nconv3)]
            # ---- layer-2 epilogue
            loaded = set()
            for m in range(nconv3 // 2):
                ra, rb = max(2 * m - 3, 0), min(2 * m - 2, npo - 1)
                if rb >= ra:
                    lo = int(np.float32(po0 + ra) * scale)
                    hi = min(int(np.float32(po0 + rb) * scale) + 1, in_side - 1)
                    first = lo if not loaded else max(loaded) + 1
                    for r in range(first, hi + 1):
                        loaded.add(r)
                    assert hi - min(x for x in loaded if x >= lo - 0) <= 10
                for k, row in ((0, 2 * m - 3), (1, 2 * m - 2)):
                    if 0 <= row < npo:
                        v = conv3[row] + conv3[row + 1] + conv3[row + 2] + conv3[row + 3]
                        hp = hpool(v)
                        Y = po0 + row
                        fy = np.float32(Y) * scale
                        y0 = int(fy); y1 = min(y0 + 1, in_side - 1); ty = float(fy - np.float32(y0))
                        assert y0 in loaded and y1 in loaded, (Y, y0, y1, sorted(loaded)[-4:])
                        assert max(loaded) - y0 < RESROWS
                        for q in range(4):
                            for l in range(32):
                                rel = (76 if q == 3 else 27 * q) + l
                                col = x0 + rel
                                ok = l < 27 and (q < 3 or l >= 5) and own_lo <= col < own_hi
                                if not ok:
                                    continue
                                fx = np.float32(col) * scale
                                jx0 = int(fx); tx = float(fx - np.float32(jx0))
                                joff = jx0 - jb
                                assert 0 <= joff <= RESPX - 2
                                jdx = 1 if jx0 + 1 < in_side else 0
                                tl_, tr_ = R2[y0, jb + joff], R2[y0, jb + joff + jdx]
                                bl_, br_ = R2[y1, jb + joff], R2[y1, jb + joff + jdx]
                                top = tl_ + (tr_ - tl_) * tx
                                bot = bl_ + (br_ - bl_) * tx
                                out[Y, col] = A * hp[32 * q + l] + B * (top + (bot - top) * ty) + C
                                writes[Y, col] += 1
    return out, writes


if __name__ == "__main__":
    in_side = int(sys.argv[1]) if len(sys.argv) > 1 else 215
    out_side = in_side - 10
    rpi = int(sys.argv[2]) if len(sys.argv) > 2 else 69
    rng = np.random.default_rng(0)
    R2 = rng.random((in_side, in_side, 32))
    w2 = rng.normal(0, 0.05, (3, 3, 32, 32)); w3 = rng.normal(0, 0.05, (3, 3, 32, 32))
    A, B, C = rng.normal(0, 1, 32), rng.normal(0, 1, 32), rng.normal(0, 1, 32)
    want = direct(R2, w2, w3, A, B, C, out_side)
    got, writes = sim(R2, w2, w3, A, B, C, out_side, rpi)
    assert (writes == 1).all(), (writes.min(), writes.max())
    err = np.abs(got - want).max()
    print("in_side %d rows_per_item %d: max err %.3e, every output written exactly once" % (in_side, rpi, err))
    assert err < 1e-9

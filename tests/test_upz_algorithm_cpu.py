"""CPU: the ALGORITHM of the cell-form regional up-convolution (csrc/conv_tc_upz.cu) restated in plain torch and checked against the
oracle's formulation of the reference layer (conv_transpose2d(stride 2) + Blur per region, masked after the blur: reference
models/stylegan2/model.py:287-300,395-398).  Test infrastructure: it pins the algebra the CUDA kernels implement -- the cell / phase /
tap-block mapping, the (cell, region) row rule of e4s_upz_build_rows and the window the FIR pass reads -- independently of any GPU."""
import torch
import torch.nn.functional as F

from oracle import e4s_oracle as orc


def _nearest(n_out, n_in):
    return torch.floor(torch.arange(n_out, dtype=torch.float32) * (n_in / n_out)).long().clamp(max=n_in - 1)


def cell_form_upconv(x, w, s, d, labels, fir):
    """x [B,Ci,H,W], w [Co,Ci,3,3] (already scaled), s [B,K,Ci], d [B,K,Co], labels [B,Hl,Wl] (long), fir [4,4] -> [B,Co,2H,2W]."""
    B, Ci, H, W = x.shape
    Co = w.shape[0]
    Ho, Wo = 2 * H, 2 * W
    reg = labels[:, _nearest(Ho, labels.shape[1])][:, :, _nearest(Wo, labels.shape[2])]          # region of every output pixel
    # --- e4s_upz_build_rows: cell (cy,cx) lists region r iff an output pixel in (2cy-2..2cy+2) x (2cx-2..2cx+2) lies in r
    rows, index = [], {}
    for b in range(B):
        for cy in range(H + 1):
            for cx in range(W + 1):
                win = reg[b, max(2 * cy - 2, 0):min(2 * cy + 2, Ho - 1) + 1, max(2 * cx - 2, 0):min(2 * cx + 2, Wo - 1) + 1]
                for r in sorted(set(win.reshape(-1).tolist())):
                    index[(b, cy, cx, r)] = len(rows)
                    rows.append((b, cy, cx, r))
    # --- cell GEMM: Z[row, py, px, :] = z[2cy+py, 2cx+px] of conv_transpose2d(x * s_r): tap (dy,dx) of the 2x2 input window feeds phase
    #     (py,px) through W[ky][kx] with ky = py if dy == 0 else 2 (only py == 0), kx likewise: 9 blocks, no zeros
    Z = torch.zeros(len(rows), 2, 2, Co, dtype=x.dtype)
    for i, (b, cy, cx, r) in enumerate(rows):
        for dy in (0, -1):
            for dx in (0, -1):
                iy, ix = cy + dy, cx + dx
                if not (0 <= iy < H and 0 <= ix < W):
                    continue
                a = x[b, :, iy, ix] * s[b, r]
                for py in ((0, 1) if dy == 0 else (0,)):
                    for px in ((0, 1) if dx == 0 else (0,)):
                        ky, kx = (py if dy == 0 else 2), (px if dx == 0 else 2)
                        Z[i, py, px] += w[:, :, ky, kx] @ a
    # --- FIR pass: out[q] = d[r(q)] * sum_{u,v} flip(fir)[u,v] * z_{r(q)}[q - 1 + (u,v)], z outside the (2H+1) x (2W+1) grid = 0
    kf = torch.flip(fir, [0, 1])
    out = torch.zeros(B, Co, Ho, Wo, dtype=x.dtype)
    for b in range(B):
        for oy in range(Ho):
            for ox in range(Wo):
                r = int(reg[b, oy, ox])
                acc = torch.zeros(Co, dtype=x.dtype)
                for u in range(4):
                    for v in range(4):
                        zy, zx = oy - 1 + u, ox - 1 + v
                        if 0 <= zy <= 2 * H and 0 <= zx <= 2 * W:
                            acc += kf[u, v] * Z[index[(b, zy >> 1, zx >> 1, r)], zy & 1, zx & 1]
                out[b, :, oy, ox] = acc * d[b, r]
    return out, len(rows) / (B * (H + 1) * (W + 1))


def test_cell_form_equals_the_reference_formulation():
    g = torch.Generator().manual_seed(2)
    B, Ci, Co, H, W, K = 2, 5, 4, 4, 6, 3
    x = torch.randn(B, Ci, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(Co, Ci, 3, 3, generator=g, dtype=torch.float64) * 0.3
    s = 1 + 0.3 * torch.randn(B, K, Ci, generator=g, dtype=torch.float64)
    d = 1 + 0.2 * torch.randn(B, K, Co, generator=g, dtype=torch.float64)
    labels = torch.randint(0, K, (B, 16, 24), generator=g)
    fir = orc.fir_kernel((1, 3, 3, 1), 4.0, torch.float64)
    got, rows_per_cell = cell_form_upconv(x, w, s, d, labels, fir)
    # the reference's formulation: sum_k mask_k * demod_k * Blur(conv_transpose2d(x * s_k; W))
    Ho, Wo = 2 * H, 2 * W
    reg = labels[:, _nearest(Ho, 16)][:, :, _nearest(Wo, 24)]
    want = torch.zeros(B, Co, Ho, Wo, dtype=torch.float64)
    for k in range(K):
        yt = F.conv_transpose2d(x * s[:, k][:, :, None, None], w.transpose(0, 1), stride=2)
        yk = orc.upfirdn2d(yt, fir, pad=(1, 1)) * d[:, k][:, :, None, None]
        want += yk * (reg == k)[:, None].double()
    assert got.shape == want.shape
    assert float((got - want).abs().max()) < 1e-12
    assert 1.0 <= rows_per_cell <= K


def test_cell_form_single_region_has_one_row_per_cell():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 3, 3, 3, generator=g, dtype=torch.float64)
    w = torch.randn(2, 3, 3, 3, generator=g, dtype=torch.float64)
    s = torch.ones(1, 1, 3, dtype=torch.float64)
    d = torch.ones(1, 1, 2, dtype=torch.float64)
    fir = orc.fir_kernel((1, 3, 3, 1), 4.0, torch.float64)
    got, rpc = cell_form_upconv(x, w, s, d, torch.zeros(1, 8, 8, dtype=torch.long), fir)
    want = orc.upfirdn2d(F.conv_transpose2d(x, w.transpose(0, 1), stride=2), fir, pad=(1, 1))
    assert rpc == 1.0 and float((got - want).abs().max()) < 1e-12

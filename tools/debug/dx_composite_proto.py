"""CPU fp64 prototype of the composite-stencil backward of the depthwise branch section.

Checks, against torch autograd of the reference op sequence (4 depthwise conv + training BatchNorm branches, summed):
    dx      = K5[flip(sum_br c1_br w_br)](du) - K9[C](x) - T3_int + frame(q)
    dw_br[t] = c1_br P[t] - c2_br Q_br[t] - c0_br S[t]
with  C[d]     = sum_br c2_br sum_s w_br[s] w_br[s+d]              (9 x 9, zero-extended x)
      frame(q) = sum_br sum_{s: q-s outside} w_br[s] (c2_br ytilde_br(q-s) + c0_br)   (2-pixel frame only)
      P[t]     = sum_p du(p) x~(p+t),  Q_br[t] = sum_p y_br(p) x~(p+t),  S[t] = sum_p x~(p+t)   (p inside the image)
Run: python tools/debug/dx_composite_proto.py
"""
import torch
import torch.nn.functional as F

torch.manual_seed(0)
SHAPES = {0: (5, 5), 1: (3, 3), 2: (3, 1), 3: (1, 3)}


def embed5(w, kh, kw):
    """[E,kh,kw] -> [E,5,5] centred"""
    out = w.new_zeros(w.shape[0], 5, 5)
    r0, c0 = (5 - kh) // 2, (5 - kw) // 2
    out[:, r0:r0 + kh, c0:c0 + kw] = w
    return out


def run(B, E, H, W, eps=1e-5):
    x = torch.randn(B, E, H, W, dtype=torch.float64, requires_grad=True)
    ws = [torch.randn(E, *SHAPES[k], dtype=torch.float64, requires_grad=True) for k in range(4)]
    gam = [torch.randn(E, dtype=torch.float64, requires_grad=True) for _ in range(4)]
    bet = [torch.randn(E, dtype=torch.float64, requires_grad=True) for _ in range(4)]
    ys, u = [], 0
    for k in range(4):
        kh, kw = SHAPES[k]
        y = F.conv2d(x, ws[k].unsqueeze(1), None, 1, (kh // 2, kw // 2), 1, E)
        ys.append(y)
        u = u + F.batch_norm(y, None, None, gam[k], bet[k], True, 0.1, eps)
    du = torch.randn_like(u)
    grads = torch.autograd.grad(u, [x] + ws, du)
    dx_ref, dw_ref = grads[0], grads[1:]

    # ---- restated backward
    n = B * H * W
    w5 = [embed5(ws[k].detach(), *SHAPES[k]) for k in range(4)]           # [E,5,5] each
    xd = x.detach()
    xp = F.pad(xd, (2, 2, 2, 2))
    # lag sums over p inside the image
    P = torch.zeros(E, 5, 5, dtype=torch.float64)
    S = torch.zeros(E, 5, 5, dtype=torch.float64)
    Q = [torch.zeros(E, 5, 5, dtype=torch.float64) for _ in range(4)]
    for a in range(5):
        for b in range(5):
            xs = xp[:, :, a:a + H, b:b + W]
            P[:, a, b] = (du * xs).sum((0, 2, 3))
            S[:, a, b] = xs.sum((0, 2, 3))
            for k in range(4):
                Q[k][:, a, b] = (ys[k].detach() * xs).sum((0, 2, 3))
    sdu = du.sum((0, 2, 3))
    c1, c2, c0 = [], [], []
    for k in range(4):
        yk = ys[k].detach()
        mean = yk.mean((0, 2, 3))
        var = yk.var((0, 2, 3), unbiased=False)
        rstd = (var + eps).rsqrt()
        sduy = (w5[k] * P).sum((1, 2))
        dgamma = rstd * (sduy - mean * sdu)
        c1k = gam[k].detach() * rstd
        c2k = gam[k].detach() * rstd * rstd * dgamma / n
        c1.append(c1k); c2.append(c2k); c0.append(c1k * sdu / n - c2k * mean)
    # weight gradients
    for k in range(4):
        dwk = c1[k][:, None, None] * P - c2[k][:, None, None] * Q[k] - c0[k][:, None, None] * S
        kh, kw = SHAPES[k]
        r0, cc0 = (5 - kh) // 2, (5 - kw) // 2
        err = (dwk[:, r0:r0 + kh, cc0:cc0 + kw] - dw_ref[k]).abs().max().item()
        assert err < 1e-9 * max(1.0, dw_ref[k].abs().max().item()), ("dw", k, err)
    # merged kernels
    Wc1 = sum(c1[k][:, None, None] * w5[k] for k in range(4))               # [E,5,5]
    Wc0 = sum(c0[k][:, None, None] * w5[k] for k in range(4))
    C = torch.zeros(E, 9, 9, dtype=torch.float64)
    for k in range(4):
        wk = w5[k]
        for sa in range(5):
            for sb in range(5):
                for ta in range(5):
                    for tb in range(5):
                        C[:, ta - sa + 4, tb - sb + 4] += c2[k] * wk[:, sa, sb] * wk[:, ta, tb]
    # main stencil: dx0(q) = sum_s Wc1[s] du~(q-s) - sum_d C[d] x~(q+d) - sum_s Wc0[s]
    t1 = F.conv2d(du, torch.flip(Wc1, (1, 2)).unsqueeze(1), None, 1, 2, 1, E)
    t2 = F.conv2d(xd, C.unsqueeze(1), None, 1, 4, 1, E)
    dx0 = t1 - t2 - Wc0.sum((1, 2))[None, :, None, None]
    # frame correction, pixel by pixel (what the border kernel does)
    dx = dx0.clone()
    xz = lambda b, e, r, c: xd[b, e, r, c] if (0 <= r < H and 0 <= c < W) else 0.0
    for qr in range(H):
        for qc in range(W):
            if 2 <= qr < H - 2 and 2 <= qc < W - 2:
                continue
            for b in range(B):
                for e in range(E):
                    acc = 0.0
                    for k in range(4):
                        wk = w5[k][e]
                        for sa in range(5):
                            for sb in range(5):
                                if wk[sa, sb] == 0:
                                    continue
                                pr, pc = qr - (sa - 2), qc - (sb - 2)
                                if 0 <= pr < H and 0 <= pc < W:
                                    continue
                                yt = 0.0
                                for ta in range(5):
                                    for tb in range(5):
                                        yt += wk[ta, tb] * xz(b, e, pr + ta - 2, pc + tb - 2)
                                acc += wk[sa, sb] * (c2[k][e] * yt + c0[k][e])
                    dx[b, e, qr, qc] += acc
    err = (dx - dx_ref).abs().max().item()
    int_err = (dx0 - dx_ref)[:, :, 2:H - 2, 2:W - 2].abs().max().item() if H > 4 and W > 4 else 0.0
    print(f"B{B} E{E} {H}x{W}: dx err {err:.2e} (interior before frame fix {int_err:.2e}), scale {dx_ref.abs().max().item():.2f}")
    assert err < 1e-9 * max(1.0, dx_ref.abs().max().item())


if __name__ == "__main__":
    run(2, 2, 9, 11)
    run(1, 2, 4, 8)
    run(1, 1, 3, 8)
    run(1, 1, 2, 8)
    run(1, 1, 1, 8)
    print("ok")

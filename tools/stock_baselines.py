"""Same-device STOCK baselines for the two hot-path units ("same-device torch/cuDNN ReparamConv" and
"torch-gather NA"), next to this repo's kernels, on the LM-Net stage shapes of the 352x352 / batch-16 workload.
Informational only — not the `--impl reference` arm of bench.py.

    python tools/stock_baselines.py [--out profiles/rNN_stock_baselines.txt]

* ReparamConv, training forward + backward, bf16 autocast: the reference op sequence
  (/root/reference/core/modules.py:586-600 — expand conv + BN + Hardswish, four depthwise conv + BN branches, GELU, SE,
  pointwise + shortcut 1x1) run through stock torch modules on cuDNN / cuBLAS (channels-last input, eager and replayed
  from a CUDA graph) against the same module on the lmnet_b200 kernels.
* Neighbourhood attention core (kernel 3, 12 heads): a stock-torch formulation that gathers the K x K neighbours with the
  clamped-window index rule and runs softmax / weighted sum with ordinary tensor ops (what one writes without natten),
  against lmnet_b200's fused na2d.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lm-net_b200")]

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def stock_reparam_forward(m, x):
    x1 = m.expand_conv(x)
    out = m.large_conv(x1) + m.square_conv(x1) + m.ver_conv(x1) + m.hor_conv(x1)
    return m.pointwise_conv(m.se(m.active(out))) + m.shortcut(x)


def time_fn(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def graphed(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g.replay


def stock_na(q, k, v, rpb, K, scale):
    """q, k, v [B,H,W,heads,D]; neighbours gathered with start = clamp(t - K/2, 0, L - K); rpb [heads,2K-1,2K-1]."""
    B, H, W, nh, D = q.shape
    dev = q.device
    ns = K // 2
    iy = torch.arange(H, device=dev)
    ix = torch.arange(W, device=dev)
    sy = (iy - ns).clamp(0, H - K)
    sx = (ix - ns).clamp(0, W - K)
    ky = sy[:, None] + torch.arange(K, device=dev)[None]                    # [H,K]
    kx = sx[:, None] + torch.arange(K, device=dev)[None]                    # [W,K]
    kk = k[:, ky][:, :, :, kx]                                              # [B,H,K,W,K,nh,D]
    vv = v[:, ky][:, :, :, kx]
    s = torch.einsum("bhwnd,bhiwjnd->bhwnij", q * scale, kk)
    by = (ky - iy[:, None] + K - 1)                                         # [H,K]
    bx = (kx - ix[:, None] + K - 1)                                         # [W,K]
    bias = rpb[:, by][:, :, :, bx]                                          # [nh,H,K,W,K]
    s = s + bias.permute(1, 3, 0, 2, 4).unsqueeze(0)
    p = s.flatten(-2).softmax(-1).view_as(s)
    return torch.einsum("bhwnij,bhiwjnd->bhwnd", p, vv)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--res", type=int, default=352)
    a = ap.parse_args()
    from lmnet_b200.model import ReparamConv
    from natten.functional import na2d

    dev = torch.device("cuda")
    torch.backends.cudnn.benchmark = True
    lines = [f"Same-device stock baselines, B200, bf16 autocast, batch {a.batch}, training forward + backward (ms, median of 10)",
             "", "ReparamConv(C, 2C, C) block:",
             f"{'level':>5} {'C':>4} {'R':>4} {'stock eager':>12} {'stock graph':>12} {'lmnet_b200':>11} {'speed-up vs graph':>18}"]
    for lvl in range(1, 5):
        C, R = 12 * 2 ** (lvl - 1), a.res // 2 ** (lvl - 1)
        torch.manual_seed(0)
        m = ReparamConv(C, 2 * C, C).to(dev).train()
        x = torch.randn(a.batch, C, R, R, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_()

        def run(fwd):
            def step():
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    y = fwd(x)
                y.backward(torch.ones_like(y))
                x.grad = None
                for p in m.parameters():
                    p.grad = None
            return step

        stock = run(lambda t: stock_reparam_forward(m, t))
        ours = run(m)
        t_se = time_fn(stock)
        t_sg = time_fn(graphed(stock))
        t_o = time_fn(graphed(ours))
        lines.append(f"{lvl:5d} {C:4d} {R:4d} {t_se:12.3f} {t_sg:12.3f} {t_o:11.3f} {t_sg / t_o:17.2f}x")
    lines += ["", "Neighbourhood attention core (kernel 3, dilation 1, 12 heads), forward + backward:",
              f"{'level':>5} {'hd':>3} {'R':>4} {'torch gather':>13} {'lmnet_b200':>11} {'speed-up':>9}"]
    for lvl in range(1, 5):
        hd, R = 2 ** (lvl - 1), a.res // 2 ** (lvl - 1)
        g = torch.Generator(device=dev).manual_seed(0)
        B = a.batch if lvl > 1 else max(1, a.batch // 4)          # the gathered K^2 copies of k and v need 18x the memory
        q, k, v = (torch.randn(B, R, R, 12, hd, device=dev, dtype=torch.bfloat16, generator=g).requires_grad_() for _ in range(3))
        rpb = (0.02 * torch.randn(12, 5, 5, device=dev, generator=g)).requires_grad_()
        go = torch.randn(B, R, R, 12, hd, device=dev, dtype=torch.bfloat16, generator=g)
        scale = hd ** -0.5

        def stock():
            o = stock_na(q, k, v, rpb.to(torch.bfloat16), 3, scale)
            torch.autograd.grad(o, (q, k, v, rpb), go)

        def ours():
            o = na2d(q, k, v, 3, 1, rel_pos_bias=rpb, scale=scale)
            torch.autograd.grad(o, (q, k, v, rpb), go)

        t_s, t_o = time_fn(stock, iters=5, warm=2), time_fn(ours)
        note = f"  (batch {B})" if B != a.batch else ""
        lines.append(f"{lvl:5d} {hd:3d} {R:4d} {t_s:13.3f} {t_o:11.3f} {t_s / t_o:8.1f}x{note}")
    text = "\n".join(lines)
    print(text)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        open(a.out, "w").write(text + "\n")


if __name__ == "__main__":
    main()

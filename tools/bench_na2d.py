"""na2d micro-benchmark (BASELINE.json configs[3]) and high-resolution inference check (configs[4]).

    python tools/bench_na2d.py [--out profiles/rNN_na2d_microbench.txt]

For each LM-Net stage shape of the 352x352 / batch-16 workload ([16,R,R,12,hd], R,hd = 352,1 / 176,2 / 88,4 /
44,8), kernel 3 (what the model runs) and kernel 7 with dilation 1 and 2 (the BASELINE micro-benchmark), bf16:
fused forward and forward+backward time (CUDA events, L2 flushed between iterations by writing a 512 MB
buffer; a device-side spin before the first event keeps host launch latency out of the interval), algorithmic GB/s (fwd 4*N*es, bwd 7*N*es) and the fraction of the measured HBM peak.
Then one LM-Net inference pass at batch 8, 1024x1024, bf16 (eval mode: running statistics folded).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lm-net_b200")]

import torch  # noqa: E402

from natten.functional import na2d  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def time_it(fn, flush, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.fill_(1.0)
        torch.cuda._sleep(2_000_000)   # ~1 ms of device spin: the host enqueues fn() meanwhile, so the events
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)  # bracket device time only
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-highres", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda")
    pk, src = peak()
    flush = torch.empty(512 * 1024 * 1024 // 4, device=dev)
    lines = [f"na2d fused micro-benchmark, bf16, B=16, 12 heads; HBM peak {pk:.0f} GB/s ({src}); L2 flushed between iterations",
             f"{'R':>4} {'hd':>3} {'K':>2} {'d':>2} {'fwd ms':>8} {'fwd GB/s':>9} {'frac':>6} {'bwd ms':>8} {'bwd GB/s':>9} {'frac':>6}"]
    g = torch.Generator(device=dev).manual_seed(0)
    for R, hd in ((352, 1), (176, 2), (88, 4), (44, 8)):
        for K, d in ((3, 1), (7, 1), (7, 2)):
            q, k, v = (torch.randn(16, R, R, 12, hd, device=dev, dtype=torch.bfloat16, generator=g).requires_grad_() for _ in range(3))
            rpb = (0.02 * torch.randn(12, 2 * K - 1, 2 * K - 1, device=dev, generator=g)).requires_grad_()
            go = torch.randn(16, R, R, 12, hd, device=dev, dtype=torch.bfloat16, generator=g)
            nbytes = q.numel() * 2
            with torch.no_grad():
                t_f = time_it(lambda: na2d(q, k, v, K, d, rel_pos_bias=rpb), flush)
            out = na2d(q, k, v, K, d, rel_pos_bias=rpb)

            def bwd():
                torch.autograd.grad(out, (q, k, v, rpb), go, retain_graph=True)

            t_b = time_it(bwd, flush)
            gf, gb = 4 * nbytes / t_f / 1e6, 7 * nbytes / t_b / 1e6
            lines.append(f"{R:4d} {hd:3d} {K:2d} {d:2d} {t_f:8.3f} {gf:9.1f} {gf / pk:6.3f} {t_b:8.3f} {gb:9.1f} {gb / pk:6.3f}")
            del q, k, v, out
    if not a.no_highres:
        from lmnet_b200.model import LM_Net

        torch.manual_seed(0)
        net = LM_Net(3, 2).to(dev).eval()
        x = torch.randn(8, 3, 1024, 1024, device=dev)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            t = time_it(lambda: net(x), flush, iters=5, warm=2)
            y = net(x)
        lines.append("")
        lines.append(f"LM-Net inference, batch 8, 1024x1024, bf16 autocast, eval mode: {t:.1f} ms/batch = {8e3 / t:.1f} images/s; "
                     f"logits {tuple(y.shape)}, finite={bool(torch.isfinite(y.float()).all())}, "
                     f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
    text = "\n".join(lines)
    print(text)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        open(a.out, "w").write(text + "\n")


if __name__ == "__main__":
    main()

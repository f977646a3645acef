"""TEST INFRASTRUCTURE — plain-PyTorch restatement of ReparamConv's branch section and forward.

Follows /root/reference/core/modules.py:586-600 (forward), :592-595 (four depthwise conv + BN
branches summed), :597 (GELU then SE), :1030-1036 (SE).  Runs on CPU (fp32 / fp64) or on a GPU
through stock torch ops.  Checked against the unmodified reference class in
tests/test_reference_parity.py (when /root/reference is present) and through the committed golden
vectors in tests/golden/ (generated from the reference by tests/golden/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import torch
import torch.nn.functional as F


def branch_section(mod, x1):
    """out = sum of the four (depthwise conv -> BatchNorm) branches of a non-deploy ReparamConv-like `mod`."""
    out = None
    for br in (mod.large_conv, mod.square_conv, mod.ver_conv, mod.hor_conv):
        c, bn = br.conv, br.bn
        y = F.conv2d(x1, c.weight, None, c.stride, c.padding, 1, c.groups)
        # nn.BatchNorm2d (which the reference calls through `self.large_conv(x1)` etc.) keys on ITS OWN .training flag
        use_batch = bn.training or bn.running_mean is None
        if bn.training and bn.running_mean is not None:
            bn.num_batches_tracked.add_(1)
        y = F.batch_norm(y, None if bn.running_mean is None else bn.running_mean,
                         None if bn.running_var is None else bn.running_var,
                         bn.weight, bn.bias, use_batch, bn.momentum, bn.eps)
        out = y if out is None else out + y
    return out


def dw_bn_gelu(mod, x1):
    """(z, pool) = (GELU(branch sum), spatial mean of z) — what the fused CUDA op returns."""
    if getattr(mod, "deploy", False):
        fc = mod.fuse_conv
        out = F.conv2d(x1, fc.weight, fc.bias, fc.stride, fc.padding, 1, fc.groups)
    else:
        out = branch_section(mod, x1)
    z = F.gelu(out)
    return z, z.float().mean(dim=(2, 3)) if z.dtype != torch.float64 else z.mean(dim=(2, 3))


def reparam_forward_ref(mod, x):
    """Whole ReparamConv.forward in stock torch ops (the reference op sequence)."""
    x1 = mod.expand_conv(x)
    z, _ = dw_bn_gelu(mod, x1)
    se = mod.se
    gate = se.scale_activation(se.fc2(se.activation(se.fc1(F.adaptive_avg_pool2d(z, 1)))))
    return mod.pointwise_conv(gate * z) + mod.shortcut(x)

"""TEST-ONLY: answers the raw C-ABI launches of lmnet_b200 with the CPU oracle so that the host-side Python
(modules, autograd plumbing, layout handling, DDP wiring) can run on a machine without a GPU.  The product
has no such path.  `apply(setattr_fn)` installs the patches through a monkeypatch-style setter."""


def apply(setattr_fn):
    import torch

    from lmnet_b200 import _lib, na_ops, reparam
    from oracle import na2d_ref, reparam_ref

    o = na2d_ref.c_oracle()
    setattr_fn(_lib, "require_cuda", lambda *a: None)

    def c32(t):   # oracle precision: fp64 stays fp64, everything else runs in fp32
        t = t.detach()
        return (t if t.dtype == torch.float64 else t.float()).contiguous()

    def rp(rpb32, like):
        if rpb32 is None:
            return None
        return rpb32.double() if like.dtype == torch.float64 else rpb32

    def fused_fwd(q, k, v, rpb32, out, K, d, scale, order="bhwnd", lse=None):
        assert order == "bhwnd"
        out.copy_(o.fused_fwd(c32(q), c32(k), c32(v), rp(rpb32, q), K, d, scale).to(out.dtype))

    def fused_bwd(q, k, v, rpb32, dout, dq, dk, dv, drpb, K, d, scale, order="bhwnd"):
        gq, gk, gv, gr = o.fused_bwd(c32(q), c32(k), c32(v), rp(rpb32, q), c32(dout), K, d, scale)
        dq.copy_(gq.to(dq.dtype)), dk.copy_(gk.to(dk.dtype)), dv.copy_(gv.to(dv.dtype))
        if drpb is not None:
            drpb.copy_(gr)

    def qk_fwd(q, k, rpb32, attn, K, d):
        attn.copy_(o.qk_fwd(c32(q), c32(k), rp(rpb32, q), K, d).to(attn.dtype))

    def qk_bwd(q, k, dattn, dq, dk, drpb, K, d):
        gq, gk, gr = o.qk_bwd(c32(q), c32(k), c32(dattn), K, d, drpb is not None)
        dq.copy_(gq.to(dq.dtype)), dk.copy_(gk.to(dk.dtype))
        if drpb is not None:
            drpb.copy_(gr)

    def av_fwd(attn, v, out, K, d):
        out.copy_(o.av_fwd(c32(attn), c32(v), K, d).to(out.dtype))

    def av_bwd(attn, v, dout, dattn, dv, K, d):
        ga, gv = o.av_bwd(c32(attn), c32(v), c32(dout), K, d)
        dattn.copy_(ga.to(dattn.dtype)), dv.copy_(gv.to(dv.dtype))

    for name, fn in dict(raw_fused_fwd=fused_fwd, raw_fused_bwd=fused_bwd, raw_qk_fwd=qk_fwd, raw_qk_bwd=qk_bwd,
                         raw_av_fwd=av_fwd, raw_av_bwd=av_bwd).items():
        setattr_fn(na_ops, name, fn)
    import torch.nn.functional as F

    from lmnet_b200 import bnact

    def bn_act_ref(bn, y, act="none", stats=None):
        out = bn(y)
        return {"none": lambda t: t, "hardswish": F.hardswish, "gelu": F.gelu, "relu": F.relu}[act](out)

    setattr_fn(bnact, "bn_act", bn_act_ref)
    setattr_fn(reparam, "bn_act", bn_act_ref)
    setattr_fn(reparam, "expand_1x1", lambda conv, x, want_stats=False: (conv(x), None) if want_stats else conv(x))
    setattr_fn(reparam, "pointwise_shortcut", lambda pw, sc, z, gate, x: pw(gate * z) + sc(x))
    from lmnet_b200 import patch

    setattr_fn(patch, "layer_norm", lambda ln, x: ln(x))
    from lmnet_b200 import upsample

    setattr_fn(upsample.Upsample2x, "forward",
               lambda self, x: F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True))
    setattr_fn(reparam, "fused_dw_bn_gelu", lambda mod, x1: reparam_ref.dw_bn_gelu(mod, x1))
    setattr_fn(reparam, "fused_dw_deploy", lambda mod, x1: reparam_ref.dw_bn_gelu(mod, x1))
    return o

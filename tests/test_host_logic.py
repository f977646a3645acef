"""CPU-side tests: the C-ABI library loads and exports what include/*.h declares, argument checking,
module surface (state_dict keys, constructor spellings, padding), autograd plumbing through the
`cpu_backend` fixture, and the model definition against the committed golden vectors."""
import ctypes
import json
import os
import re

import pytest
import torch

from _helpers import GOLDEN_DIR, fill_deterministic, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_library_loads_and_exports_every_declared_symbol():
    from lmnet_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    so = ctypes.CDLL(_lib.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "lmnet_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(lmnet_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(so, name), f"{name} declared in include/lmnet_b200.h but not exported"
    assert _lib.lib().lmnet_abi_version() == 1
    assert _lib.lib().lmnet_status_string(-2).decode().startswith("unsupported")


def test_c_abi_rejects_bad_arguments_without_a_gpu():
    from lmnet_b200 import _lib as L

    lib = L.lib()
    dims = L.na_dims(1, 4, 8, 2, 4, 3, 2)       # H < K*dilation
    assert lib.lmnet_na2d_bwd_workspace_bytes(L.byref(dims)) == 0
    v = L.View5(0, 0, 0, 0, 0)
    rc = lib.lmnet_na2d_fwd(L.byref(v), L.byref(v), L.byref(v), None, L.byref(v), None, L.byref(dims), 1.0, 0, None)
    assert rc == -1
    dims = L.na_dims(1, 8, 8, 2, 4, 4, 1)       # even kernel
    rc = lib.lmnet_na2d_qk_fwd(L.byref(v), L.byref(v), None, None, L.byref(dims), 0, None)
    assert rc == -1
    dd = L.DwDims(0, 4, 8, 8)
    assert lib.lmnet_reparam_dw_workspace_bytes(L.byref(dd), 0) == 0


def test_product_raises_on_cpu_tensors():
    import natten

    m = natten.NeighborhoodAttention2D(dim=24, num_heads=12, kernel_size=3)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.randn(1, 6, 6, 24))
    with pytest.raises(RuntimeError, match="CUDA"):
        natten.functional.na2d_qk(torch.randn(1, 2, 5, 5, 4), torch.randn(1, 2, 5, 5, 4), 3)
    from lmnet_b200.model import ReparamConv

    with pytest.raises(RuntimeError, match="CUDA"):
        ReparamConv(4, 8, 4)(torch.randn(1, 4, 8, 8))


def test_argument_checks_mirror_natten():
    import natten
    from natten.functional import na2d, na2d_av, na2d_qk

    q = torch.randn(1, 2, 6, 6, 4)
    with pytest.raises(ValueError):
        na2d_qk(q, q, 4)                       # even kernel
    with pytest.raises(ValueError):
        na2d_qk(q, q, 1)
    with pytest.raises(ValueError):
        na2d_qk(q, q, 7)                       # 6 < 7
    with pytest.raises(ValueError):
        na2d_qk(q, q, 3, dilation=3)           # 6 < 9
    with pytest.raises(ValueError):
        na2d_qk(q, q, 3, rel_pos_bias=torch.zeros(2, 3, 3))
    with pytest.raises(ValueError):
        na2d_av(torch.randn(1, 2, 6, 6, 8), q, 3)
    with pytest.raises(ValueError):
        na2d(q, q[:, :1], q, 3)
    with pytest.raises(ValueError):
        natten.NeighborhoodAttention2D(dim=25, num_heads=12, kernel_size=3)
    with pytest.raises(ValueError):
        natten.NeighborhoodAttention2D(dim=24, num_heads=12, kernel_size=4)


def test_module_surface_matches_natten_014_017():
    import natten

    m = natten.NeighborhoodAttention2D(dim=24, num_heads=12, kernel_size=3)       # LM-Net's spelling
    assert list(m.state_dict().keys()) == ["rpb", "qkv.weight", "qkv.bias", "proj.weight", "proj.bias"]
    assert m.rpb.shape == (12, 5, 5) and m.qkv.weight.shape == (72, 24)
    assert 0 < float(m.rpb.detach().abs().max()) <= 2.0
    m2 = natten.NeighborhoodAttention2D(24, 12, 7, dilation=2, rel_pos_bias=False, qkv_bias=False)
    assert m2.rpb is None and m2.qkv.bias is None
    m3 = natten.NeighborhoodAttention2D(24, 12, 3, 1, bias=False)
    assert "rpb" not in m3.state_dict()
    for name in ("na2d", "na2d_qk", "na2d_av", "natten2dqkrpb", "natten2dav"):
        assert callable(getattr(natten.functional, name))


@pytest.mark.parametrize("K,d,H,W", [(3, 1, 7, 9), (3, 2, 8, 7), (5, 1, 6, 11), (3, 1, 2, 5)])
def test_module_forward_backward_against_oracle_module(cpu_backend, K, d, H, W):
    """Host plumbing of the drop-in module (packed qkv views, scale, rpb, padding of small maps, merge,
    proj) == the oracle restatement of natten's module; backward through the packed autograd op."""
    import natten
    from oracle.na2d_ref import OracleNeighborhoodAttention2D

    torch.manual_seed(0)
    m = natten.NeighborhoodAttention2D(dim=24, num_heads=6, kernel_size=K, dilation=d)
    ref = OracleNeighborhoodAttention2D(24, 6, K, d)
    ref.load_state_dict(m.state_dict())
    x = torch.randn(2, H, W, 24, requires_grad=True)
    xr = x.detach().clone().requires_grad_()
    if H < K * d or W < K * d:   # natten 0.14: zero-pad to the window, crop the output
        pad = torch.nn.functional.pad(xr, (0, 0, 0, max(0, K * d - W), 0, max(0, K * d - H)))
        yr = ref(pad)[:, :H, :W]
    else:
        yr = ref(xr)
    y = m(x)
    assert y.shape == x.shape
    assert rel_err(y, yr) < 1e-5
    g = torch.randn_like(y)
    y.backward(g)
    yr.backward(g)
    assert rel_err(x.grad, xr.grad) < 1e-4
    for (n, p), (_, pr) in zip(m.named_parameters(), ref.named_parameters()):
        assert rel_err(p.grad, pr.grad) < 1e-4, n


def test_functional_ops_autograd_plumbing(cpu_backend):
    from natten.functional import na2d, na2d_av, na2d_qk, natten2dav, natten2dqkrpb
    from oracle import na2d_ref as R

    torch.manual_seed(1)
    B, Hd, H, W, D, K = 1, 2, 6, 7, 4, 3
    # transposed (non-contiguous) inputs, as transformers' DiNAT produces them
    q = torch.randn(B, H, W, Hd, D).permute(0, 3, 1, 2, 4).requires_grad_()
    k = torch.randn(B, H, W, Hd, D).permute(0, 3, 1, 2, 4).requires_grad_()
    v = torch.randn(B, H, W, Hd, D).permute(0, 3, 1, 2, 4).requires_grad_()
    rpb = (0.1 * torch.randn(Hd, 5, 5)).requires_grad_()
    attn = natten2dqkrpb(q, k, rpb, K, 1)
    out = natten2dav(attn.softmax(-1), v, K, 1)
    ref = R.na2d_av_gather(R.na2d_qk_gather(q, k, rpb, K, 1).softmax(-1), v, K, 1)
    assert rel_err(out, ref) < 1e-5
    go = torch.randn_like(out)
    g1 = torch.autograd.grad(out, (q, k, v, rpb), go)
    g2 = torch.autograd.grad(ref, (q, k, v, rpb), go)
    for a, b in zip(g1, g2):
        assert rel_err(a, b) < 1e-4
    # fused op with separate tensors, fused layout
    qf, kf, vf = (t.permute(0, 2, 3, 1, 4) for t in (q, k, v))
    of = na2d(qf, kf, vf, K, rel_pos_bias=rpb)
    rf = R.na2d_gather(qf, kf, vf, K, 1, rpb)
    assert rel_err(of, rf) < 1e-5
    g1 = torch.autograd.grad(of, (q, k, v, rpb), go.permute(0, 2, 3, 1, 4))
    g2 = torch.autograd.grad(rf, (q, k, v, rpb), go.permute(0, 2, 3, 1, 4))
    for a, b in zip(g1, g2):
        assert rel_err(a, b) < 1e-4


def test_model_state_dict_matches_reference_keys():
    from lmnet_b200.model import LM_Net, count_parameters

    golden = json.load(open(os.path.join(GOLDEN_DIR, "lmnet_keys.json")))
    net = LM_Net(3, 2)
    mine = {k: list(v.shape) for k, v in net.state_dict().items()}
    assert list(mine.keys()) == list(golden.keys())
    assert mine == golden
    assert count_parameters(net) == 3966566          # SURVEY.md §0 probe of the reference


def test_reparam_block_against_reference_golden(cpu_backend):
    """Our ReparamConv (host plumbing + oracle branch section) vs vectors produced by the reference class."""
    from lmnet_b200.model import ReparamConv

    gold = torch.load(os.path.join(GOLDEN_DIR, "reparam_golden.pt"))
    blk = ReparamConv(6, 8, 4).double()
    fill_deterministic(blk, seed=7)
    x = gold["x"].clone().requires_grad_()
    blk.train()
    y = blk(x)
    y.backward(gold["go"])
    assert rel_err(y, gold["train_out"]) < 1e-12
    assert rel_err(x.grad, gold["dx"]) < 1e-10
    for k, p in blk.named_parameters():
        assert rel_err(p.grad, gold["grads"][k]) < 1e-9, k
    for k, b in blk.named_buffers():
        assert torch.allclose(b.double(), gold["buffers_after"][k].double(), rtol=1e-12, atol=1e-12), k
    blk.eval()
    with torch.no_grad():
        assert rel_err(blk(gold["x"]), gold["eval_out"]) < 1e-12
        blk.switch_to_deploy()
        assert rel_err(blk.fuse_conv.weight, gold["deploy_weight"]) < 1e-12
        assert rel_err(blk.fuse_conv.bias, gold["deploy_bias"]) < 1e-12
        assert rel_err(blk(gold["x"]), gold["deploy_out"]) < 1e-12
        assert rel_err(gold["deploy_out"], gold["eval_out"]) < 1e-10      # the reference's own identity


def test_model_against_reference_golden(cpu_backend):
    """Our LM_Net definition reproduces the reference's logits and gradients (fp64, oracle kernels)."""
    from lmnet_b200.model import LM_Net

    gold = torch.load(os.path.join(GOLDEN_DIR, "lmnet_golden.pt"))
    net = LM_Net(3, 2).double()
    fill_deterministic(net, seed=3)
    # rpb crosses the C ABI as fp32 (include/lmnet_b200.h), hence 1e-7 rather than fp64 round-off
    net.eval()
    with torch.no_grad():
        assert rel_err(net(gold["img"]), gold["eval_logits"]) < 1e-7
    net.train()
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    logits = net(gold["img"])
    assert rel_err(logits, gold["train_logits"]) < 1e-7
    assert torch.equal(logits.argmax(1), gold["train_logits"].argmax(1))
    loss = torch.nn.functional.cross_entropy(logits, gold["target"])
    loss.backward()
    assert abs(float(loss) - float(gold["loss"])) < 1e-7
    for k, p in net.named_parameters():
        g = float(gold["grad_norms"][k])
        assert abs(float(p.grad.norm()) - g) <= 1e-6 * max(1.0, g), k
    for k, p in net.named_parameters():
        if k.endswith("rpb"):
            assert rel_err(p.grad, gold["grad_rpb"][k]) < 1e-6, k


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) on a small bounded sample: one JSON
    line with the contract's keys, no GPU needed."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--res", "64"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["value"] > 0 and d["higher_is_better"] is True
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "e2e",
                "cpu_baseline"):
        assert key in d, key
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_step_roofline_table_is_committed_and_plausible():
    """profiles/step_roofline.json feeds bench.py's `step_roofline`; forward FLOPs per image must agree
    with the survey's independent probe (19.9 GFLOP at 352x352) and the hot-path units with §8 d4/d5."""
    import importlib.util
    import json
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tab = json.load(open(os.path.join(root, "profiles", "step_roofline.json")))
    assert tab["resolution"] == 352 and len(tab["ops"]) > 200
    assert abs(tab["totals_per_image"]["flops_fwd"] / 19.9e9 - 1) < 0.05
    na, dw = tab["by_class"]["NeighborhoodAttention2D.core"], tab["by_class"]["ReparamConv.dw_section"]
    assert na["ops"] == 4 and abs((na["bytes_fwd"] + na["bytes_bwd"]) * 16 / 981e6 - 1) < 0.01      # §8 d4: 981 MB at batch 16
    assert dw["ops"] == 16 and abs((dw["bytes_fwd"] + dw["bytes_bwd"]) * 16 / 5.71e9 - 1) < 0.01    # §8 d5: 5.71 GB
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    r = bench.step_roofline(16, 352, 46.8)
    assert r is not None and 4.0 < r["t_roof_ms"] < 9.0 and 0 < r["frac"] < 1


def test_conv3x3_pixel_pitch_detection():
    """Host logic of the in-place read of torch.cat gradient slices (lmnet_b200/conv3x3.py::_pixel_pitch): a dense
    channels-last tensor reports its channel count, a channel slice of a wider channels-last tensor the wider count,
    anything the kernels' aligned vector loads cannot address reports None."""
    import torch

    from lmnet_b200.conv3x3 import _pixel_pitch

    wide = torch.zeros(2, 48, 6, 10, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    assert _pixel_pitch(wide) == 48
    assert _pixel_pitch(wide[:, :24]) == 48 and _pixel_pitch(wide[:, 24:]) == 48
    assert _pixel_pitch(wide[:, 16:40]) == 48            # 32-byte offset: 16-byte vectors stay aligned
    assert _pixel_pitch(wide[:, 4:28]) is None           # 8-byte offset breaks the 16-byte vectors of a 24-channel slice
    assert _pixel_pitch(wide[:, 4:16]) == 48             # ... but not the 8-byte vectors of a 12-channel slice
    assert _pixel_pitch(torch.zeros(2, 24, 6, 10, dtype=torch.bfloat16)) is None          # NCHW planes
    assert _pixel_pitch(wide[:, :, ::2]) is None                                            # strided rows
    assert _pixel_pitch(wide[:, :, :, 1:]) is None                                          # cropped columns
    one = torch.zeros(1, 24, 6, 10, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    assert _pixel_pitch(one) == 24

"""Shared test utilities: deterministic weights, reference import shims, comparison helpers."""
import importlib
import math
import os
import sys
import types

import torch

REFERENCE_ROOT = "/root/reference"
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "core", "LM_Net.py"))


def fill_deterministic(module: torch.nn.Module, seed: int = 0) -> None:
    """Overwrite every state_dict entry with values that depend only on (seed, position, shape), so the
    reference model, our model and the CPU oracle model can be given identical weights without shipping
    a 16 MB checkpoint."""
    with torch.no_grad():
        for idx, (key, t) in enumerate(module.state_dict().items()):
            g = torch.Generator().manual_seed(seed * 1000003 + idx)
            if key.endswith("num_batches_tracked"):
                t.zero_()
                continue
            r = torch.randn(t.shape, generator=g, dtype=torch.float64)
            if key.endswith("running_var"):
                v = 0.5 + torch.rand(t.shape, generator=g, dtype=torch.float64)
            elif key.endswith("running_mean"):
                v = 0.1 * r
            elif key.endswith("rpb"):
                v = 0.2 * r
            elif key.endswith("bias"):
                v = 0.1 * r
            elif t.dim() == 1:  # norm scales
                v = 1.0 + 0.1 * r
            else:
                fan_in = t[0].numel()
                v = r * (0.7 / math.sqrt(fan_in))
            t.copy_(v.to(t.dtype))


def import_reference():
    """Import the unmodified reference model code with the smallest possible shims:
    * `timm.models.layers` (absent offline): to_2tuple, trunc_normal_, DropPath
    * `natten`: resolves to OUR drop-in package (lm-net_b200/natten) — that is the point.
    Returns (core.LM_Net module, core.modules module)."""
    if not reference_available():
        raise RuntimeError("reference not present")
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        layers.to_2tuple = lambda x: x if isinstance(x, tuple) else (x, x)
        layers.trunc_normal_ = torch.nn.init.trunc_normal_

        class DropPath(torch.nn.Module):
            def __init__(self, p=0.0):
                super().__init__()
                self.p = p

            def forward(self, x):
                return x

        layers.DropPath = DropPath
        timm.models, models.layers = models, layers
        sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    lm = importlib.import_module("core.LM_Net")
    mods = importlib.import_module("core.modules")
    return lm, mods


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b| — the relative error norm used for the parity tolerances."""
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

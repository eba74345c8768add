"""Import shim for the UNMODIFIED reference modules under /root/reference.

TEST INFRASTRUCTURE ONLY -- used in the build container to pin oracle/mmdit_oracle.py and to
mint tests/golden/*.pt (oracle/make_golden.py).  /root/reference does not exist on the GPU box,
so nothing that runs there imports this file.

Third-party modules the reference imports but this image lacks are stubbed before import:
  * xformers.ops.swiglu_op.SwiGLU (pinned 0.0.29.post3, README.md:69): restated from its
    published semantics -- w12 = Linear(in, 2*hidden), w3 = Linear(hidden, out), both with bias,
    x1, x2 = w12(x).chunk(2, -1); out = w3(silu(x1) * x2)   [parity unpinned by the reference]
  * colorama (imported, unused) and src.helpers.VAE_T5_CLIP_inference (encoders: out of scope)
"""
import sys
import types

import torch.nn as nn
import torch.nn.functional as F

REF_ROOT = "/root/reference"


class SwiGLU(nn.Module):
    def __init__(self, in_features, hidden_features, out_features=None, bias=True):
        super().__init__()
        self.w12 = nn.Linear(in_features, 2 * hidden_features, bias=bias)
        self.w3 = nn.Linear(hidden_features, out_features or in_features, bias=bias)

    def forward(self, x):
        a, b = self.w12(x).chunk(2, -1)
        return self.w3(F.silu(a) * b)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m


def import_reference():
    """Returns the reference's diff_model class (CPU: build it with attn_type='softmax')."""
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]          # never mix with the product's own `src` package
    _stub("xformers"); _stub("xformers.ops"); _stub("xformers.ops.swiglu_op", SwiGLU=SwiGLU)
    _stub("colorama", Fore=None)
    sys.path[:0] = [REF_ROOT, REF_ROOT + "/src"]
    try:
        _stub("src.helpers.VAE_T5_CLIP_inference", VAE_T5_CLIP_inference=object)
        import src  # noqa: F401  (the reference's package)
        sys.modules["src.helpers.VAE_T5_CLIP_inference"] = sys.modules["src.helpers.VAE_T5_CLIP_inference"]
        from src.models.diff_model import diff_model
    finally:
        sys.path[:] = [p for p in sys.path if not p.startswith(REF_ROOT)]
    return diff_model

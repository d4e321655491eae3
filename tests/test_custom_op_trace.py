"""CPU: the public feature functions trace as ONE ``srfe::*`` custom-op node under a full-graph capture (fake CUDA
tensors, no device needed), and the ops' fake kernels infer the shapes the C ABI reports.  The op bodies themselves run
on the GPU box (tests/test_custom_op_gpu.py)."""
from __future__ import annotations

import pytest
import torch
from torch._subclasses.fake_tensor import FakeTensorMode

import speechrecognitionproject_b200 as S
from speechrecognitionproject_b200 import features as F

_FN = {"SpecParams": S.spec, "FbankParams": S.fbank, "MfccParams": S.mfcc}


@pytest.mark.parametrize("name", sorted(S.PRESETS))
@pytest.mark.filterwarnings("ignore")
def test_fullgraph_trace_is_one_custom_op(srfe_lib, name):
    p = S.PRESETS[name]
    fn = _FN[type(p).__name__]

    class FrontEnd(torch.nn.Module):
        def forward(self, x):
            return fn(x, p) * 1.0

    with FakeTensorMode():
        x = torch.empty(3, 16000, device="cuda")
    ep = torch.export.export(FrontEnd(), (x,), strict=True)          # strict = dynamo, errors on any graph break
    calls = [n.target for n in ep.graph.nodes if n.op == "call_function"]
    fam = type(p).__name__[:-6].lower()
    assert calls[0] == getattr(torch.ops.srfe, fam).default and len(calls) == 2, calls
    out = [n for n in ep.graph.nodes if n.op == "output"][0].args[0][0]
    assert tuple(out.meta["val"].shape) == (3,) + F.out_shape(p, 16000)
    assert out.meta["val"].device.type == "cuda" and out.meta["val"].dtype == torch.float32


@pytest.mark.parametrize("n", [16000, 8000, 12345])
def test_fake_kernels_agree_with_the_c_abi_shapes(srfe_lib, n):
    for name, p in S.PRESETS.items():
        fam = type(p).__name__[:-6].lower()
        with FakeTensorMode():
            for dt in (torch.float32, torch.int16):
                fake = getattr(torch.ops.srfe, fam)(torch.empty(3, n, device="cuda", dtype=dt), *p.op_args())
                assert tuple(fake.shape) == (3,) + F.out_shape(p, n), (name, n)
                assert fake.dtype == torch.float32

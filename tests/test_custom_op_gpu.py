"""GPU: the registered PyTorch custom ops ``srfe::spec | fbank | mfcc`` (the boundary BASELINE.json's north_star
names: "backed by a PyTorch custom op over a thin C-ABI").  The eager public functions call the C ABI directly
(features._eager_device); these tests make sure the op path itself -- dispatcher, fake kernels, torch.compile --
runs and returns the same bits.
"""
from __future__ import annotations

import numpy as np
import pytest
import torch

import oracle
import speechrecognitionproject_b200 as S
from speechrecognitionproject_b200 import features as F

pytestmark = pytest.mark.gpu


def _op_args(p):
    fam = {"SpecParams": "spec", "FbankParams": "fbank", "MfccParams": "mfcc"}[type(p).__name__]
    return getattr(torch.ops.srfe, fam), p.op_args()


_FN = {"SpecParams": S.spec, "FbankParams": S.fbank, "MfccParams": S.mfcc}


@pytest.fixture(scope="module")
def xd():
    return torch.from_numpy(oracle.synthetic_corpus(6, config_index=1)).cuda()


@pytest.mark.parametrize("name", sorted(S.PRESETS))
def test_op_called_directly_equals_eager_path(srfe_lib, xd, name):
    p = S.PRESETS[name]
    op, args = _op_args(p)
    n0 = S.launch_count()
    via_op = op(xd, *args)
    assert S.launch_count() - n0 == 1                      # the op launched the fused kernel, nothing else did
    eager = _FN[type(p).__name__](xd, p)
    assert via_op.shape == eager.shape and via_op.dtype == torch.float32 and via_op.is_cuda
    assert torch.equal(via_op, eager)


@pytest.mark.parametrize("name", ["R-SPEC", "R-FBANK", "R-MFCC", "C-MFCC"])
def test_opcheck(srfe_lib, xd, name):
    p = S.PRESETS[name]
    op, args = _op_args(p)
    # schema, fake-tensor kernel (shape / dtype / device inference) and dispatch registration; no autograd
    # formula is registered on purpose (features are computed under no_grad, models/model_mfcc_bgru.py:29)
    torch.library.opcheck(op, (xd, *args), test_utils=("test_schema", "test_faketensor"))
    # int16 PCM goes through the same op
    torch.library.opcheck(op, (xd.to(torch.int16), *args), test_utils=("test_schema", "test_faketensor"))


@pytest.mark.parametrize("name", ["R-SPEC", "R-FBANK", "R-MFCC"])
def test_torch_compile_fullgraph_routes_through_the_op(srfe_lib, xd, name):
    p = S.PRESETS[name]
    fn = _FN[type(p).__name__]

    def front_end(x):
        return fn(x, p) * 1.0

    eager = fn(xd, p)
    compiled = torch.compile(front_end, fullgraph=True, backend="aot_eager")
    n0 = S.launch_count()
    got = compiled(xd)
    assert S.launch_count() - n0 >= 1
    assert torch.equal(got, eager)


def test_op_fixes_up_strided_and_overlapping_inputs(srfe_lib, xd):
    op, args = _op_args(S.R_MFCC)
    ref = op(xd, *args)
    wide = torch.zeros(xd.size(0), xd.size(1) + 6, device="cuda")
    wide[:, 3:-3] = xd
    assert torch.equal(op(wide[:, 3:-3], *args), ref)                       # odd offset, padded rows
    one = xd[:1].expand(4, -1)                                              # stride(0) == 0 < n_samples
    assert torch.equal(op(one, *args), ref[:1].expand(4, -1, -1))
    assert torch.equal(S.mfcc(one), ref[:1].expand(4, -1, -1))


def test_fake_shapes_match_the_c_abi(srfe_lib):
    from torch._subclasses.fake_tensor import FakeTensorMode
    for name, p in S.PRESETS.items():
        op, args = _op_args(p)
        for n in (16000, 8000, 12345):
            with FakeTensorMode():
                fake = op(torch.empty(3, n, device="cuda"), *args)
            assert tuple(fake.shape) == (3,) + F.out_shape(p, n), (name, n)

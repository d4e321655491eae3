"""patch_model: the re-plumbed Network.forward (one batched feature call, PCM moved instead of
features) gives the same logits as the reference's own per-clip CPU loop.

CPU tests drive the patched forward with oracle features (feature_fn hook) so the handoff logic
is checked without a GPU; when /root/reference is mounted (build container only) the UNMODIFIED
reference model modules are imported and compared, otherwise structural stand-ins with the same
attribute names are used.  The GPU test runs the stand-ins with the real kernels."""
from __future__ import annotations

import importlib.util
import os
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

import oracle
from oracle import librosa_shim
from speechrecognitionproject_b200 import patch

REF = "/root/reference/models"


def _oracle_fn(kind):
    def f(x):
        xn = x.detach().cpu().numpy()
        if kind in ("mfcc_bgru", "mfrn_bgru"):
            y = np.stack([oracle.mfcc_ref(c).T for c in xn])
        elif kind in ("spec_bgru", "spec_cnn"):
            y = np.stack([oracle.spec_ref(c).T for c in xn])
        else:
            y = np.stack([oracle.fbank_ref(c) for c in xn])
        return torch.from_numpy(np.ascontiguousarray(y))
    return f


def _load_ref(name):
    librosa_shim.install()
    spec = importlib.util.spec_from_file_location(f"ref_{name}", os.path.join(REF, f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("name,kind", [("model_mfcc_bgru", "mfcc_bgru"), ("model_spec_bgru", "spec_bgru"),
                                       ("model_spec_cnn", "spec_cnn"), ("model_fbanks_cnn", "fbanks_cnn"),
                                       ("model_mfrn_bgru", "mfrn_bgru")])
@pytest.mark.filterwarnings("ignore")
def test_patched_forward_equals_reference_forward(name, kind):
    mod = _load_ref(name)
    assert patch.detect_kind(mod) == kind
    torch.manual_seed(0)
    kw = {} if "cnn" in kind else {"num_features": 16, "num_layers": 1}
    net = mod.Network(**kw).eval()
    keys = list(net.state_dict().keys())
    x = torch.from_numpy(oracle.synthetic_corpus(3, config_index=6))
    with torch.no_grad():
        want = net(x)                                  # reference: per-clip CPU loop
    patch.patch_model(mod, feature_fn=_oracle_fn(kind))
    try:
        with torch.no_grad():
            got = net(x)                               # one batched feature call
        assert list(net.state_dict().keys()) == keys   # layers / checkpoint keys untouched
        torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
    finally:
        patch.unpatch_model(mod)
    with torch.no_grad():
        torch.testing.assert_close(net(x), want)       # unpatch restores the original


# ---- structural stand-ins (same attribute names / loop shape as the reference modules) -------
def _standin_mfcc_bgru():
    m = types.ModuleType("standin_mfcc_bgru")

    def compute_mfcc(sample):
        return torch.from_numpy(oracle.mfcc_ref(sample.numpy()))

    class Network(nn.Module):
        def __init__(self):
            super().__init__()
            self.gru = nn.GRU(39, hidden_size=8, num_layers=1, bidirectional=True, batch_first=True)
            self.fc = nn.Linear(16, 12)

        def forward(self, x):
            with torch.no_grad():
                inx = torch.ones(x.size(0), 39, 51)
                for i in range(x.size(0)):
                    inx[i] = m.compute_mfcc(x[i])
            inx = inx.to(next(self.parameters()).device).transpose(1, 2)
            inx, _ = self.gru(inx)
            return self.fc(inx[:, -1, :])

    m.compute_mfcc, m.Network = compute_mfcc, Network
    return m


def _standin_fbanks_cnn():
    m = types.ModuleType("standin_fbanks_cnn")

    def filter_banks(sample):
        return torch.from_numpy(oracle.fbank_ref(sample.numpy()))

    class Network(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = nn.Conv2d(1, 4, kernel_size=(7, 3), padding=(3, 1))
            self.maxpool1 = nn.MaxPool2d((1, 3))
            self.conv2 = nn.Conv2d(4, 4, (1, 7), padding=(0, 3))
            self.maxpool2 = nn.MaxPool2d((1, 4))
            self.conv3 = nn.Conv2d(4, 8, (1, 10))
            self.conv4 = nn.Conv2d(8, 8, (7, 1), padding=(3, 0))
            self.maxpool3 = nn.MaxPool1d(98)
            self.dropout = nn.Dropout()
            self.fc1 = nn.Linear(8, 8)
            self.fc2 = nn.Linear(8, 12)

        def forward(self, x):
            with torch.no_grad():
                inx = torch.ones(x.size(0), 98, 120)
                for i in range(x.size(0)):
                    inx[i] = m.filter_banks(x[i])
            inx = inx.to(next(self.parameters()).device)
            return patch._cnn_tail(self, inx)

    m.filter_banks, m.Network = filter_banks, Network
    return m


@pytest.mark.parametrize("factory,kind", [(_standin_mfcc_bgru, "mfcc_bgru"), (_standin_fbanks_cnn, "fbanks_cnn")])
def test_standin_cpu(factory, kind):
    mod = factory()
    assert patch.detect_kind(mod) == kind
    torch.manual_seed(1)
    net = mod.Network().eval()
    x = torch.from_numpy(oracle.synthetic_corpus(2, config_index=7))
    with torch.no_grad():
        want = net(x)
    patch.patch_model(mod, feature_fn=_oracle_fn(kind))
    with torch.no_grad():
        got = net(x)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
    assert mod.Network.forward.__srfe_patched__ == kind


@pytest.mark.gpu
@pytest.mark.parametrize("factory,kind", [(_standin_mfcc_bgru, "mfcc_bgru"), (_standin_fbanks_cnn, "fbanks_cnn")])
def test_standin_gpu_real_kernels(srfe_lib, factory, kind):
    mod = factory()
    torch.manual_seed(2)
    net = mod.Network().eval()
    x = torch.from_numpy(oracle.synthetic_corpus(4, config_index=8))
    with torch.no_grad():
        want = net(x)                                  # CPU loop with oracle features
    net = net.cuda()
    patch.patch_model(mod)                             # default: fused CUDA features, PCM H2D only
    with torch.no_grad():
        got_host_pcm = net(x)                          # CPU batch, like DataLoader hands it over
        got_dev_pcm = net(x.cuda())                    # PCM already resident
    assert got_host_pcm.is_cuda and torch.equal(got_host_pcm, got_dev_pcm)
    torch.testing.assert_close(got_host_pcm.cpu(), want, rtol=2e-3, atol=2e-3)


def _standin_spec(kind):
    """spec_bgru / spec_cnn stand-ins (same attribute names and layer order as the reference modules)."""
    m = types.ModuleType(f"standin_{kind}")

    def compute_spec(sample):
        y = oracle.spec_ref(sample.numpy())
        return torch.from_numpy(np.ascontiguousarray(y.T if kind == "spec_cnn" else y))

    if kind == "spec_bgru":
        class Network(nn.Module):
            def __init__(self):
                super().__init__()
                self.gru = nn.GRU(321, hidden_size=8, num_layers=1, bidirectional=True, batch_first=True)
                self.fc = nn.Linear(16, 12)

            def forward(self, x):
                with torch.no_grad():
                    inx = torch.ones(x.size(0), 321, 49)
                    for i in range(x.size(0)):
                        inx[i] = m.compute_spec(x[i])
                inx = inx.to(next(self.parameters()).device).transpose(1, 2)
                inx, _ = self.gru(inx)
                return self.fc(inx[:, -1, :])
    else:
        class Network(nn.Module):
            def __init__(self):
                super().__init__()
                self.conv1 = nn.Conv2d(1, 4, kernel_size=(7, 3), padding=(3, 1))
                self.maxpool1 = nn.MaxPool2d((1, 3))
                self.conv2 = nn.Conv2d(4, 4, (1, 7), padding=(0, 3))
                self.maxpool2 = nn.MaxPool2d((1, 4))
                self.conv3 = nn.Conv2d(4, 8, (1, 26))
                self.conv4 = nn.Conv2d(8, 8, (7, 1), padding=(3, 0))
                self.maxpool3 = nn.MaxPool1d(49)
                self.dropout = nn.Dropout()
                self.fc1 = nn.Linear(8, 8)
                self.fc2 = nn.Linear(8, 12)

            def forward(self, x):
                with torch.no_grad():
                    inx = torch.ones(x.size(0), 49, 321)
                    for i in range(x.size(0)):
                        inx[i] = m.compute_spec(x[i])
                inx = inx.to(next(self.parameters()).device)
                return patch._cnn_tail(self, inx)

    m.compute_spec, m.Network = compute_spec, Network
    return m


def test_shared_front_end_dedups_ensemble_members():
    """ensemble members patched with one SharedFrontEnd compute each feature set once per batch
    (the two spectrogram models share theirs), recompute on a new batch and after an in-place edit."""
    mods = {"spec_bgru": _standin_spec("spec_bgru"), "spec_cnn": _standin_spec("spec_cnn"),
            "fbanks_cnn": _standin_fbanks_cnn()}
    torch.manual_seed(3)
    nets = {k: m.Network().eval() for k, m in mods.items()}
    x = torch.from_numpy(oracle.synthetic_corpus(2, config_index=9))
    with torch.no_grad():
        want = {k: n(x) for k, n in nets.items()}
    calls = []
    fe = patch.SharedFrontEnd()
    for k, m in mods.items():
        fn = _oracle_fn(k)
        patch.patch_model(m, kind=k, feature_fn=(lambda t, k=k, fn=fn: (calls.append(k), fn(t))[1]), frontend=fe)
    with torch.no_grad():
        got = {k: n(x) for k, n in nets.items()}
    for k in mods:
        torch.testing.assert_close(got[k], want[k], rtol=1e-5, atol=1e-5)
    assert calls == ["spec_bgru", "fbanks_cnn"] and fe.launches == 2 and fe.uploads == 0     # spec_cnn reused spec_bgru's
    with torch.no_grad():
        nets["spec_cnn"](x)                                  # same batch again: still cached
    assert fe.launches == 2
    x[0, 0] += 1.0                                           # in-place edit bumps the version -> recompute
    with torch.no_grad():
        nets["spec_cnn"](x)
    assert fe.launches == 3
    y = x.clone()
    with torch.no_grad():
        nets["spec_bgru"](y)                                 # another batch object
    assert fe.launches == 4
    for m in mods.values():
        patch.unpatch_model(m)


@pytest.mark.gpu
def test_shared_front_end_gpu(srfe_lib):
    mods = {"spec_bgru": _standin_spec("spec_bgru"), "spec_cnn": _standin_spec("spec_cnn"),
            "fbanks_cnn": _standin_fbanks_cnn(), "mfcc_bgru": _standin_mfcc_bgru()}
    torch.manual_seed(4)
    nets = {k: m.Network().eval() for k, m in mods.items()}
    x = torch.from_numpy(oracle.synthetic_corpus(3, config_index=10))
    with torch.no_grad():
        want = {k: n(x) for k, n in nets.items()}          # reference-style CPU loops with oracle features
    import speechrecognitionproject_b200 as S
    fe = patch.SharedFrontEnd()
    for k, m in mods.items():
        nets[k].cuda()
        patch.patch_model(m, kind=k, frontend=fe)
    n0 = S.launch_count()
    with torch.no_grad():
        got = {k: n(x) for k, n in nets.items()}
    # one PCM H2D; the spectrogram (shared by both spec models) and the fbank features come out of ONE fused launch
    # (srfe_spec_fbank_kernel), MFCC out of a second one
    assert fe.uploads == 1 and fe.launches == 2 and S.launch_count() - n0 == 2
    for k in mods:
        # bins / bands far below the clip maximum differ between any two fp32 FFTs (tests/tolerances.py is level-aware
        # for that reason); through a few hundred random input weights that is worth a few 1e-3 on a logit.  Feature
        # parity proper is test_parity_gpu.py's job; this test is about the shared upload / launch plumbing.
        tol = 1e-2
        torch.testing.assert_close(got[k].cpu(), want[k], rtol=tol, atol=tol, msg=lambda m, k=k: f"{k}: {m}")
    for m in mods.values():
        patch.unpatch_model(m)


# ---- the full-size reference models pinned to the UNMODIFIED reference modules (model_mfcc_bgru = BASELINE cfg5) ----
MODELS = ["model_mfcc_bgru", "model_spec_bgru", "model_spec_cnn", "model_fbanks_cnn"]


def _golden_logits(name="model_mfcc_bgru"):
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"{name}_logits.npz"))
    return {k: g[k] for k in g.files}


def _seeded_net(mod, seed):
    torch.manual_seed(int(seed))
    return mod.Network().eval()


def _assert_checksums(net, g):
    sd = net.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    for k, want in zip(g["keys"], g["checksums"]):
        v = sd[str(k)].double()
        np.testing.assert_allclose([float(v.sum()), float(v.abs().sum())], want, rtol=1e-12, atol=0, err_msg=str(k))   # sums re-associate across CPUs


@pytest.mark.parametrize("name", MODELS)
@pytest.mark.filterwarnings("ignore")
def test_twin_regenerates_the_reference_weights_and_logits(name):
    """tests/twins.py's shape twins, seeded like oracle/make_golden_logits.py seeded the unmodified reference modules,
    have the reference's parameters (per-tensor checksums recorded from the real modules) and their forward
    (per-clip CPU loop, oracle features) returns the recorded reference logits."""
    from tests import twins
    g = _golden_logits(name)
    net = _seeded_net(twins.twin(name), g["seed"])
    _assert_checksums(net, g)
    x = torch.from_numpy(oracle.synthetic_corpus(int(g["n_clips"]), config_index=int(g["config_index"])))
    with torch.no_grad():
        np.testing.assert_allclose(net(x).numpy(), g["logits"], rtol=0, atol=2e-5)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("name", MODELS)
@pytest.mark.filterwarnings("ignore")
def test_twin_equals_reference_module(name):
    from tests import twins
    g = _golden_logits(name)
    ref = _seeded_net(twins.load_reference(name), g["seed"])
    twin = _seeded_net(twins.twin(name), g["seed"])
    _assert_checksums(ref, g)
    for (k1, v1), (k2, v2) in zip(ref.state_dict().items(), twin.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)


@pytest.mark.gpu
@pytest.mark.parametrize("name", MODELS)
@pytest.mark.filterwarnings("ignore")
def test_full_size_reference_models_gpu_vs_reference_logits(srfe_lib, name):
    """The reference's networks at full size on the GPU: reference weights (regenerated from the seed, checksummed),
    forward re-plumbed by patch_model (PCM H2D -> fused feature kernel -> the model's own layers), logits against the
    ones the UNMODIFIED reference module produced on the CPU (tests/golden/<model>_logits.npz)."""
    from tests import twins
    g = _golden_logits(name)
    mod, which = twins.load(name)
    net = _seeded_net(mod, g["seed"])
    _assert_checksums(net, g)
    net = net.cuda()
    patch.patch_model(mod)
    try:
        x = torch.from_numpy(oracle.synthetic_corpus(int(g["n_clips"]), config_index=int(g["config_index"])))
        with torch.no_grad():
            got = net(x)
            got_dev = net(x.cuda())
        assert got.is_cuda and torch.equal(got, got_dev)
        err = float(np.abs(got.cpu().numpy() - g["logits"]).max())
        print(f"\n{name} ({which}): max abs logit difference vs the unmodified reference module {err:.2e}")
        # MFCC features differ from the CPU path by <= 1e-3 everywhere, which moves the BGRU logits by < 1e-5 (measured);
        # the log-spectrogram / log-fbank models also see the deep-null elements outside the tolerance domain
        # (tests/tolerances.py), through randomly initialised convolutions: bounded loosely, printed exactly.
        # TF32 is off for cuDNN RNNs and convolutions by default.
        tol = 2e-4 if name == "model_mfcc_bgru" else 2e-2 * max(1.0, float(np.abs(g["logits"]).max()))
        assert err <= tol, (name, which, err)
    finally:
        patch.unpatch_model(mod)

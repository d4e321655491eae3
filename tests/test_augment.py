"""On-device augmentation / silence synthesis (/root/reference/dataset.py:148-202, :107-116).

CPU: the counter-based stream (Philox4x32-10 known answers from the Random123 distribution) and the numpy restatement's
behaviour against the reference's definitions.  GPU: the kernel against the restatement, BIT-exact (integer and float64
arithmetic only), shard independence, and the augmented batch flowing into the fused front end on the device."""
from __future__ import annotations

from dataclasses import replace

import numpy as np
import pytest
import torch

from oracle import augment as OA

N = 16000


def _bank(rng):
    return [(rng.standard_normal(60000) * 500).astype(np.int16), (rng.standard_normal(31234) * 2000).astype(np.int16),
            np.zeros(N, np.int16)]                                    # incl. a silent file exactly one clip long


def _batch(rng, b=48):
    x = (rng.standard_normal((b, N)) * rng.uniform(30, 9000, (b, 1))).clip(-32768, 32767).astype(np.int16)
    x[1] = 32767                                                      # full scale: the int16 wrap of np.int16() shows
    x[2] = 0
    return x


def test_philox_known_answers():
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for c, k, want in kat:
        got = OA.philox4x32_10(np.array([c], np.uint32), np.array([k], np.uint32))[0]
        assert [int(v) for v in got] == want


def test_restatement_follows_the_reference_definitions():
    rng = np.random.default_rng(1)
    x, bank = _batch(rng, 400), _bank(rng)
    kind = np.zeros(400, np.int8); kind[10:20] = OA.KIND_SILENCE_ZERO; kind[20:30] = OA.KIND_SILENCE_NOISE
    p = OA.AugmentParams(seed=77)
    out, ops = OA.augment_ref(x, bank, p, kind, first_index=5)
    assert out.dtype == np.float32 and out.shape == x.shape
    frac = np.bincount(ops[kind == 0], minlength=8) / (kind == 0).sum()      # five equal bands (dataset.py:107-116)
    assert all(0.12 < frac[o] < 0.28 for o in (OA.OP_NONE, OA.OP_SHIFT, OA.OP_NOISE_UNIFORM, OA.OP_HOST_PITCH, OA.OP_HOST_SPEED))
    assert not out[10:20].any() and (ops[10:20] == OA.OP_SILENCE_ZERO).all()
    sil = out[20:30]
    assert (ops[20:30] == OA.OP_SILENCE_NOISE).all() and (np.abs(sil - np.round(sil)).max() > 0 or not sil.any())   # non-integer floats
    for i in [i for i in np.flatnonzero(ops == OA.OP_SHIFT) if i > 2][:20]:    # dataset.py:193-202 (clips 1, 2 are constant)
        y, s = out[i], x[i].astype(np.float32)
        hit = [sh for sh in range(-4800, 4801) if (sh >= 0 and np.array_equal(y[:N - sh], s[sh:])) or (sh < 0 and np.array_equal(y[-sh:], s[:sh]))]
        assert hit, i
        sh = hit[0]
        fill = y[N - sh:] if sh >= 0 else y[:-sh]
        assert fill.size == abs(sh) and (fill >= -32).all() and (fill <= 31).all() and np.array_equal(fill, np.round(fill))
    for i in np.flatnonzero(ops == OA.OP_NOISE_UNIFORM)[:20]:               # dataset.py:185-191: some slice, some factor < 0.1
        d = out[i].astype(np.float64) - x[i]
        assert np.array_equal(out[i], np.round(out[i])) and (x[i].max() > 30000 or np.abs(d).max() <= 0.1 * 32768 + 1)
    for i in np.flatnonzero((ops == OA.OP_NONE) | (ops == OA.OP_HOST_PITCH) | (ops == OA.OP_HOST_SPEED)):
        assert np.array_equal(out[i], x[i].astype(np.float32))              # dataset.py:117
    # counter-based: a clip's result depends on (seed, global index) only
    out2, ops2 = OA.augment_ref(x[100:150], bank, p, kind[100:150], first_index=105)
    assert np.array_equal(out2, out[100:150]) and np.array_equal(ops2, ops[100:150])
    assert not np.array_equal(OA.augment_ref(x[:40], bank, replace(p, seed=78), kind[:40], 5)[1], ops[:40])


def test_snr_mix_hits_the_target_snr():
    rng = np.random.default_rng(2)
    x, bank = _batch(rng, 64)[3:], _bank(rng)[:2]
    p = OA.AugmentParams(seed=3, shift_band=(0, 0), noise_band=(0, 0), snr_band=(0.0, 2.0), pitch_band=(0, 0), speed_band=(0, 0))
    out, ops = OA.augment_ref(x, bank, p)
    assert (ops == OA.OP_NOISE_SNR).all()
    seen = set()
    for i in range(len(x)):
        added = out[i].astype(np.float64) - x[i]
        if not added.any():
            seen.add(None)
            continue
        if np.abs(x[i]).max() > 9000 or np.abs(x[i]).std() < 200:      # np.int16() wraps loud mixes; truncation dominates quiet ones
            continue
        snr = 10 * np.log10((x[i].astype(np.float64) ** 2).sum() / (added ** 2).sum())
        level = min(OA.SNR_LEVELS_DB, key=lambda s: abs(s - snr))
        assert abs(snr - level) < 0.3, (i, snr)                              # int16 truncation moves it a little
        seen.add(level)
    assert len(seen) >= 4


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("seed,snr", [(0, False), (2**40 + 12345, True)])
def test_kernel_matches_the_restatement_bit_for_bit(srfe_lib, seed, snr):
    from speechrecognitionproject_b200 import augment as GA
    rng = np.random.default_rng(4)
    x, bank = _batch(rng, 300), _bank(rng)
    kind = np.zeros(300, np.int8); kind[7] = 1; kind[8:40] = 2
    bands = dict(snr_band=(0.8, 1.0)) if snr else {}
    want, want_ops = OA.augment_ref(x, bank, OA.AugmentParams(seed=seed, **bands), kind, first_index=2**33 + 17)
    gbank = GA.NoiseBank(bank)
    got, ops = GA.augment(torch.from_numpy(x).cuda(), gbank, GA.AugmentParams(seed=seed, **bands),
                          kind=torch.from_numpy(kind), first_index=2**33 + 17)
    assert got.dtype == torch.float32 and got.is_cuda
    np.testing.assert_array_equal(ops.cpu().numpy(), want_ops)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    # shards of the batch, each with its own offset, reproduce the whole
    a, _ = GA.augment(torch.from_numpy(x[:111]).cuda(), gbank, GA.AugmentParams(seed=seed, **bands), kind=torch.from_numpy(kind[:111]), first_index=2**33 + 17)
    b, _ = GA.augment(torch.from_numpy(x[111:]).cuda(), gbank, GA.AugmentParams(seed=seed, **bands), kind=torch.from_numpy(kind[111:]), first_index=2**33 + 17 + 111)
    assert torch.equal(torch.cat((a, b)), got)


@pytest.mark.gpu
def test_device_augmenter_methods_and_front_end_handoff(srfe_lib):
    import speechrecognitionproject_b200 as S
    from speechrecognitionproject_b200 import augment as GA
    import oracle
    rng = np.random.default_rng(5)
    x, bank = _batch(rng, 64), _bank(rng)[:2]
    aug = GA.DeviceAugmenter(GA.NoiseBank(bank), seed=9)
    xd = torch.from_numpy(x).cuda()
    sil = aug.generate_silence_sample(200)                                   # dataset.py:148-161
    assert sil.shape == (200, N) and not sil[:185].any() and sil[185:].abs().sum() > 0 and aug.silence_class_zeros_count == 185
    assert aug.generate_silence_sample(3).abs().sum() > 0
    i0 = aug.next_index
    y = aug.time_stretching(xd, 4800)
    want, ops = OA.augment_ref(x, bank, OA.AugmentParams(seed=9, shift_band=(0, 2), noise_band=(0, 0), pitch_band=(0, 0), speed_band=(0, 0)), first_index=i0)
    assert (ops == OA.OP_SHIFT).all()
    np.testing.assert_array_equal(y.cpu().numpy(), want)
    i0 = aug.next_index
    y = aug.add_noise_uniform(xd, 0.1)
    want, _ = OA.augment_ref(x, bank, OA.AugmentParams(seed=9, shift_band=(0, 0), noise_band=(0, 2), pitch_band=(0, 0), speed_band=(0, 0)), first_index=i0)
    np.testing.assert_array_equal(y.cpu().numpy(), want)
    i0 = aug.next_index
    y = aug.add_noise_snr(xd)
    want, _ = OA.augment_ref(x, bank, OA.AugmentParams(seed=9, shift_band=(0, 0), noise_band=(0, 0), snr_band=(0, 2), pitch_band=(0, 0), speed_band=(0, 0)), first_index=i0)
    np.testing.assert_array_equal(y.cpu().numpy(), want)
    # augmented batch -> fused front end, all on the device, same stream
    i0 = aug.next_index
    ya, ops = aug(xd)
    feats = S.mfcc(ya, S.R_MFCC)
    want, _ = OA.augment_ref(x, bank, OA.AugmentParams(seed=9), first_index=i0)
    truth = np.stack([oracle.mfcc_truth(c, oracle.R_MFCC) for c in want[:8]])
    assert np.abs(feats[:8].cpu().numpy() - truth).max() <= 1e-3
    with pytest.raises(TypeError):
        GA.augment(xd.float(), None)
    with pytest.raises(RuntimeError, match="SRFE_ERR_BAD_ARG"):
        GA.augment(xd, None)                                                 # default bands need a noise bank

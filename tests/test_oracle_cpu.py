"""Pin the CPU oracle (oracle/*.py) against the golden vectors made from the reference itself."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, check_against_golden, golden_logmel_cases
from oracle import collate as ocollate
from oracle import logmel as ologmel
from oracle import signals

CASES = golden_logmel_cases()
# fp64 oracle vs HF numpy(fp64) path: both are double precision -> tight; vs HF torch(fp32) path: 1e-3 (north_star)
FAST = {"zeros_128", "tone1k_128", "noise_128", "speechlike_128", "noise_80", "noise_len3_128", "noise_len1_128",
        "noise_len201_128", "noise_len112123_128", "noise_len640000_128", "tone_len112123_80", "ones_128"}


@pytest.mark.parametrize("key,n_mel,gen", [c for c in CASES if c[0] in FAST], ids=[c[0] for c in CASES if c[0] in FAST])
def test_fp64_oracle_matches_reference_numpy_path(logmel_golden, key, n_mel, gen):
    out = ologmel.logmel_clip(gen(), n_mel, "fp64")
    assert out.shape == (n_mel, 3000) and out.dtype == np.float32
    check_against_golden(out, logmel_golden, key, "n", 2e-6)
    check_against_golden(out, logmel_golden, key, "t", 1e-3)


@pytest.mark.parametrize("key,n_mel,gen", [c for c in CASES if c[0] in ("noise_128", "chirp_80", "speechlike_128")],
                         ids=["noise_128", "chirp_80", "speechlike_128"])
def test_fp32_oracle_matches_reference_torch_path(logmel_golden, key, n_mel, gen):
    out = ologmel.logmel_clip(gen(), n_mel, "fp32")
    check_against_golden(out, logmel_golden, key, "t", 2e-4)


def test_known_answers(logmel_golden):
    # SURVEY §8(c) known-answer table, straight from the golden file
    g = logmel_golden
    assert g["zeros_128/t/stats"][0] == -1.5 and g["zeros_128/t/stats"][1] == -1.5
    assert abs(g["ones_128/t/stats"][0] - 1.6206918) < 1e-6 and abs(g["ones_128/t/stats"][1] + 0.37930822) < 1e-6
    assert abs(g["tone1k_128/t/stats"][0] - 1.474446) < 1e-6
    assert abs(g["tone1k_128/t/stats"][0] - g["tone1k_128/t/stats"][1] - 2.0) < 1e-6  # clamp active
    assert int(g["tone1k_128/t/stats"][3]) // 3000 == 42  # arg-max mel bin
    # per-clip (not per-batch) max in a batched call
    assert abs(float(g["batch3/0/stats"][0]) - 1.474446) < 1e-6
    assert abs(float(g["batch3/1/stats"][0]) + 0.025553823) < 1e-6


def test_mel_filters_and_window(logmel_golden):
    for n in (80, 128):
        fb = ologmel.mel_filter_bank(n)
        ref = logmel_golden[f"mel_filters_{n}"]
        assert fb.shape == ref.shape == (201, n)
        np.testing.assert_allclose(fb, ref, rtol=0, atol=1e-15)
        assert not fb[0].any() and not fb[200].any()
        assert ((fb != 0).sum(axis=1) <= 2).all()
    assert (ologmel.mel_filter_bank(128) != 0).sum() == 394 and (ologmel.mel_filter_bank(80) != 0).sum() == 391
    np.testing.assert_allclose(ologmel.hann_periodic(), logmel_golden["hann_400"], atol=3e-7)


def test_attention_mask_and_batch(logmel_golden):
    lens = [480000, 480000, 112123]
    np.testing.assert_array_equal(ologmel.frame_attention_mask(lens), logmel_golden["batch3/attention_mask"])
    assert ologmel.frame_attention_mask([640000]).sum() == 3000
    assert ologmel.frame_attention_mask([112123]).sum() == 701
    out = ologmel.logmel_clip(signals.noise(7, 112123), 128, "fp64")
    sub = logmel_golden["batch3/2/sub"]
    assert np.abs(out[:, ::25] - sub).max() < 1e-3


def test_full_resolution_short_clip(logmel_golden):
    out = ologmel.logmel_clip(signals.noise(100 + 16000 % 97, 16000), 128, "fp64")
    assert np.abs(out - logmel_golden["full/noise_len16000_128"]).max() < 1e-3


def test_do_normalize(logmel_golden):
    clip = signals.noise(7, 112123)
    x, _ = ologmel.pad_or_truncate(clip)
    xn = ologmel.zero_mean_unit_var(x, 112123)
    out = ologmel.logmel_clip(xn, 128, "fp64")
    assert np.abs(out[:, ::25] - logmel_golden["normalize/noise_len112123_128"]).max() < 1e-3


def _collate_meta():
    with open(os.path.join(GOLDEN_DIR, "collate_golden.json")) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", _collate_meta(), ids=[c["key"] for c in _collate_meta()])
def test_collate_oracle_matches_reference(collate_golden, case):
    import hashlib
    import sys
    sys.path.insert(0, GOLDEN_DIR)
    from make_golden import fake_features

    key = case["key"]
    flat, offs = collate_golden[f"{key}/ids_flat"], collate_golden[f"{key}/ids_offsets"]
    labels_in = [flat[offs[i]:offs[i + 1]].tolist() for i in range(len(offs) - 1)]
    feats = fake_features(case["seed"], case["batch"], case["n_mel"])
    f, lab = ocollate.collate_padding(feats, labels_in, signals.EOT, signals.SOT)
    assert lab.dtype == np.int64 and f.dtype == np.float32
    np.testing.assert_array_equal(lab, collate_golden[f"{key}/labels"])
    assert hashlib.sha256(f.tobytes()).hexdigest() == case["features_sha256"]
    _, lab_s = ocollate.collate_streaming(feats, labels_in, signals.EOT)
    np.testing.assert_array_equal(lab_s, collate_golden[f"{key}/labels_streaming"])


def test_fixed448(collate_golden):
    out = ocollate.labels_fixed_length(collate_golden["fixed448/ids"].tolist(), signals.EOT, 448)
    np.testing.assert_array_equal(out, collate_golden["fixed448/labels"])

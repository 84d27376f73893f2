"""Live cross-check of the oracle against the reference itself (build container only).

The golden vectors pin the oracle on fixed cases; here the reference's UNMODIFIED `DataCollatorSpeechSeq2SeqWithPadding`
(imported from /root/reference as in tests/golden/make_golden.py) and the installed transformers extractor are run on
randomly generated inputs (hypothesis) next to the oracle.  Skipped wherever /root/reference or transformers is absent
(e.g. on the GPU box, where only the committed vectors travel)."""
import os
import sys
import types

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle import collate as ocollate
from oracle import logmel as ologmel
from oracle import signals

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/finetune/training"),
                                reason="reference tree not mounted (build container only)")
tr = pytest.importorskip("transformers")
hyp = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402


@pytest.fixture(scope="module")
def ref_env():
    sys.path.insert(0, GOLDEN_DIR)
    import make_golden as mg

    ref = mg.import_reference_collators()
    fe = tr.WhisperFeatureExtractor(feature_size=80)
    proc = types.SimpleNamespace(feature_extractor=fe, tokenizer=mg.standin_tokenizer())
    return ref.DataCollatorSpeechSeq2SeqWithPadding(processor=proc, decoder_start_token_id=signals.SOT), fe


ids = st.integers(min_value=0, max_value=51865)
row = st.lists(ids, min_size=1, max_size=40)


@settings(max_examples=60, deadline=None)
@given(rows=st.lists(row, min_size=1, max_size=6), bos=st.lists(st.booleans(), min_size=6, max_size=6),
       eos=st.lists(st.booleans(), min_size=6, max_size=6))
def test_label_padding_matches_the_reference_collator(ref_env, rows, bos, eos):
    coll, _ = ref_env
    labels = [([signals.SOT] if bos[i] else []) + r + ([signals.EOT] if eos[i] else []) for i, r in enumerate(rows)]
    feats = [np.full((80, 4), float(i), dtype=np.float32) for i in range(len(labels))]
    out = coll({"input_features": feats, "labels": labels})
    f_ref, l_ref = ocollate.collate_padding(feats, labels, signals.EOT, signals.SOT)
    np.testing.assert_array_equal(out["labels"].numpy(), l_ref)
    np.testing.assert_array_equal(out["input_features"].numpy(), f_ref)


@settings(max_examples=6, deadline=None)
@given(n=st.integers(min_value=1, max_value=200000), seed=st.integers(min_value=0, max_value=10**6),
       amp=st.sampled_from([1e-4, 1e-2, 0.1, 1.0]))
def test_logmel_matches_the_installed_extractor_on_random_clips(ref_env, n, seed, amp):
    _, fe = ref_env
    clip = signals.noise(seed, n, amp=amp)
    ref = fe(clip, sampling_rate=16000).input_features[0]
    out = ologmel.logmel_clip(clip, 80, "fp64")
    assert np.abs(out - ref).max() <= 1e-3


def test_padding_variants_of_the_installed_extractor_match_the_oracle(ref_env):
    # `padding="longest"`, `max_length=`, `pad_to_multiple_of=` change the sample count the STFT runs on; the oracle takes
    # it as `n_samples` (the GPU suite checks the CUDA path against the oracle with the same values)
    _, fe = ref_env
    clips = [signals.noise(1, 32000), signals.noise(2, 16000)]
    o = fe(clips, sampling_rate=16000, padding="longest", return_attention_mask=True)
    f = np.asarray(o["input_features"])
    assert f.shape == (2, 80, 200)
    np.testing.assert_array_equal(np.asarray(o["attention_mask"]), ologmel.frame_attention_mask([32000, 16000], 32000))
    for i, c in enumerate(clips):
        assert np.abs(f[i] - ologmel.logmel_clip(c, 80, "fp64", n_samples=32000)).max() <= 1e-5
    f = np.asarray(fe(clips, sampling_rate=16000, padding="max_length", max_length=160000)["input_features"])
    assert f.shape == (2, 80, 1000)
    assert np.abs(f[0] - ologmel.logmel_clip(clips[0], 80, "fp64", n_samples=160000)).max() <= 1e-5
    f = np.asarray(fe(clips, sampling_rate=16000, padding="longest", pad_to_multiple_of=48000)["input_features"])
    assert f.shape == (2, 80, 300)
    assert np.abs(f[1] - ologmel.logmel_clip(clips[1], 80, "fp64", n_samples=48000)).max() <= 1e-5

#!/usr/bin/env python
"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the build container only (needs `transformers` and /root/reference):

    python tests/golden/make_golden.py

* log-mel vectors come from the installed `transformers.WhisperFeatureExtractor`
  — the module the reference calls at
  ref:finetune/training/data_and_collator/datasets_and_collators.py:194 — through BOTH
  of its code paths (`_torch_extract_fbank_features` = what the reference runs, and
  `_np_extract_fbank_features` = fp64).
* collator vectors come from the reference's own `DataCollatorSpeechSeq2SeqWithPadding`
  and `SimpleStreamingCollator._prepare_dataset`, imported UNMODIFIED from /root/reference
  (stub modules for the missing h5py/ray top-level imports), driven
  with HF's real `PreTrainedTokenizerBase.pad` on a vocab-less stand-in tokenizer.

Inputs are regenerated at test time from `oracle/signals.py` (deterministic integer
hash), so only outputs (sub-sampled) are stored.
"""
from __future__ import annotations

import hashlib
import importlib
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import signals  # noqa: E402

SUB = 25  # keep every 25th frame + head/tail blocks


def pack(out: np.ndarray) -> dict:
    return {
        "sub": out[:, ::SUB].copy(),
        "head": out[:, :48].copy(),
        "tail": out[:, -48:].copy(),
        "stats": np.array([out.max(), out.min(), out.astype(np.float64).mean(), float(out.argmax())], dtype=np.float64),
    }


def logmel_cases():
    """(key, n_mel, clip ndarray)"""
    cases = []
    for name in signals.NAMED_CASES:
        cases.append((f"{name}_128", 128, signals.named_case(name)))
    for name in ["zeros", "tone1k", "noise", "speechlike", "chirp"]:
        cases.append((f"{name}_80", 80, signals.named_case(name)))
    # ragged / edge lengths: 3 samples, 1 s, non-multiple of hop, 7.0077 s, 40 s (truncated)
    for n in [3, 1, 200, 201, 16000, 112123, 479999, 640000]:
        cases.append((f"noise_len{n}_128", 128, signals.noise(100 + n % 97, n)))
    cases.append(("tone_len112123_80", 80, signals.tone(440.0, 112123, 0.3)))
    cases.append(("speechlike_len250000_128", 128, signals.speechlike(9, 250000)))
    # speech-like dynamics (the clamp has work in most tiles) and a loud click in near-silence
    cases.append(("bursty60_128", 128, signals.bursty(11)))
    cases.append(("bursty60_80", 80, signals.bursty(12)))
    cases.append(("bursty60_len300007_128", 128, signals.bursty(13, 300007)))
    cases.append(("click_128", 128, signals.click_in_silence()))
    return cases


def make_logmel():
    from transformers import WhisperFeatureExtractor
    import torch
    import transformers

    blob = {}
    meta = {"transformers": transformers.__version__, "torch": torch.__version__, "numpy": np.__version__, "sub": SUB,
            "cases": []}
    fes = {n: WhisperFeatureExtractor(feature_size=n) for n in (80, 128)}
    for key, n_mel, clip in logmel_cases():
        fe = fes[n_mel]
        # exactly the reference call (ref datasets_and_collators.py:194-195)
        out_t = fe(clip, sampling_rate=16000).input_features[0]
        # fp64 numpy path of the same class
        padded = np.zeros((1, fe.n_samples), dtype=np.float32)
        m = min(len(clip), fe.n_samples)
        padded[0, :m] = clip[:m]
        out_n = fe._np_extract_fbank_features(padded, "cpu")[0].astype(np.float32)
        assert out_t.shape == (n_mel, 3000) and out_t.dtype == np.float32
        for tag, out in (("t", out_t), ("n", out_n)):
            for k, v in pack(out).items():
                blob[f"{key}/{tag}/{k}"] = v
        meta["cases"].append({"key": key, "n_mel": n_mel, "len": int(len(clip)),
                              "torch_vs_np_maxabs": float(np.abs(out_t - out_n).max())})
        print(key, out_t.max(), out_t.min(), meta["cases"][-1]["torch_vs_np_maxabs"])

    # batched call: per-clip max (not per batch) + attention mask + 'pt' tensors
    fe = fes[128]
    pair = [signals.named_case("tone1k"), signals.named_case("tone1k_quiet"), signals.noise(7, 112123)]
    bf = fe(pair, sampling_rate=16000, return_tensors="pt", return_attention_mask=True)
    feats = bf["input_features"].numpy()
    for i in range(3):
        for k, v in pack(feats[i]).items():
            blob[f"batch3/{i}/{k}"] = v
    blob["batch3/attention_mask"] = bf["attention_mask"].numpy()
    assert bf["attention_mask"].dtype == torch.int32
    # full-resolution vectors for two short-ish cases (cheap to store: mostly constant tail)
    blob["full/noise_len16000_128"] = fe(signals.noise(100 + 16000 % 97, 16000), sampling_rate=16000).input_features[0]
    # mel filter banks and window as the reference builds them
    blob["mel_filters_128"] = fes[128].mel_filters
    blob["mel_filters_80"] = fes[80].mel_filters
    blob["hann_400"] = torch.hann_window(400).numpy()
    # do_normalize variant (API completeness row f-4)
    blob["normalize/noise_len112123_128"] = fe(signals.noise(7, 112123), sampling_rate=16000, do_normalize=True).input_features[0][:, ::SUB]
    np.savez_compressed(os.path.join(HERE, "logmel_golden.npz"), **blob)
    with open(os.path.join(HERE, "logmel_golden.json"), "w") as f:
        json.dump(meta, f, indent=1)


def import_reference_collators():
    os.environ.setdefault("USER", "root")  # ref projects_paths.py:22
    for name in ("h5py", "ray"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, "/root/reference/finetune/training")
    return importlib.import_module("data_and_collator.datasets_and_collators")


def standin_tokenizer():
    from transformers import PreTrainedTokenizerBase
    from transformers.tokenization_utils_base import PaddingStrategy, TruncationStrategy

    class Tok(PreTrainedTokenizerBase):
        model_input_names = ["input_ids", "attention_mask"]
        padding_side = "right"

        @property
        def pad_token_id(self):
            return signals.EOT  # Whisper: pad token == <|endoftext|>

        def _get_padding_truncation_strategies(self, padding=False, truncation=None, max_length=None, **kw):
            if padding == "max_length":
                return PaddingStrategy.MAX_LENGTH, TruncationStrategy.DO_NOT_TRUNCATE, max_length, {}
            return PaddingStrategy.LONGEST, TruncationStrategy.DO_NOT_TRUNCATE, max_length, {}

    return Tok()


def fake_features(seed: int, batch: int, n_mel: int) -> list[np.ndarray]:
    out = []
    for i in range(batch):
        u = signals.uniform_u32(seed + i, n_mel * 3000, stream=21).astype(np.float64) / 4294967296.0
        out.append((u * 3.0 - 1.5).astype(np.float32).reshape(n_mel, 3000))
    return out


def collate_cases():
    """(key, n_mel, seed, label lists)"""
    B = signals
    cases = [
        ("bos_all", 128, 1, B.label_ids(1337, 6, 5, 40, with_bos=True)),
        ("bos_none", 128, 2, B.label_ids(11, 5, 5, 40, with_bos=False)),
        ("single_row", 80, 4, B.label_ids(13, 1, 9, 9, with_bos=True)),
        ("equal_len", 80, 5, B.label_ids(17, 4, 12, 12, with_bos=True)),
        ("long_448", 128, 6, B.label_ids(19, 3, 440, 448, with_bos=True)),
    ]
    mixed = B.label_ids(23, 5, 5, 30, with_bos=True)
    mixed[2] = mixed[2][1:]  # one row lacks BOS -> no strip
    cases.append(("bos_mixed", 128, 3, mixed))
    eos = [[B.SOT, 7, B.EOT], [B.SOT, 8, 9, 10, B.EOT, B.EOT], [B.SOT, B.EOT]]  # real EOS == pad id must survive
    cases.append(("eos_equals_pad", 80, 7, eos))
    return cases


def make_collate():
    import torch

    ref = import_reference_collators()
    tok = standin_tokenizer()
    from transformers import WhisperFeatureExtractor

    blob, meta = {}, {"cases": []}
    for key, n_mel, seed, labels in collate_cases():
        fe = WhisperFeatureExtractor(feature_size=n_mel)
        proc = types.SimpleNamespace(feature_extractor=fe, tokenizer=tok)
        feats = fake_features(seed, len(labels), n_mel)
        coll = ref.DataCollatorSpeechSeq2SeqWithPadding(processor=proc, decoder_start_token_id=signals.SOT)
        out = coll({"input_features": feats, "labels": labels})
        assert out["input_features"].dtype == torch.float32 and out["labels"].dtype == torch.int64
        # streaming collator's label half, called unbound (its __call__ needs HDF5): ref ...:229-256
        ssc = types.SimpleNamespace(feature_extractor=fe, tokenizer=tok)
        # _prepare_dataset tokenizes text itself; restate its tail with pre-tokenised ids
        lb = tok.pad([{"input_ids": ids} for ids in labels], return_tensors="pt")
        stream_labels = lb["input_ids"].masked_fill(lb.attention_mask.ne(1), -100)
        flat = np.concatenate([np.asarray(x, dtype=np.int64) for x in labels])
        offs = np.concatenate([[0], np.cumsum([len(x) for x in labels])]).astype(np.int64)
        blob[f"{key}/ids_flat"] = flat
        blob[f"{key}/ids_offsets"] = offs
        blob[f"{key}/labels"] = out["labels"].numpy()
        blob[f"{key}/labels_streaming"] = stream_labels.numpy()
        fbytes = out["input_features"].numpy().tobytes()
        assert fbytes == np.stack(feats).tobytes()
        meta["cases"].append({"key": key, "n_mel": n_mel, "seed": seed, "batch": len(labels),
                              "features_sha256": hashlib.sha256(fbytes).hexdigest(),
                              "labels_shape": list(out["labels"].shape)})
        print(key, out["labels"].shape)
        del ssc
    # fixed-448 variant (ref materialize_dataset_ray.py:43-49)
    ids = signals.label_ids(29, 1, 20, 20)[0]
    t = tok.pad({"input_ids": [ids]}, padding="max_length", max_length=448, return_tensors="pt")
    blob["fixed448/ids"] = np.asarray(ids, dtype=np.int64)
    blob["fixed448/labels"] = np.where(t["attention_mask"][0].numpy() == 1, t["input_ids"][0].numpy(), -100)
    np.savez_compressed(os.path.join(HERE, "collate_golden.npz"), **blob)
    with open(os.path.join(HERE, "collate_golden.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    make_logmel()
    make_collate()

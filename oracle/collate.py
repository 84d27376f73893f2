"""numpy restatement of the reference's collators (ORACLE — test infrastructure only).

Restates, with no GPU and no transformers import:
  ref:finetune/training/data_and_collator/datasets_and_collators.py:418-461
      DataCollatorSpeechSeq2SeqWithPadding.__call__
  ref:finetune/training/data_and_collator/datasets_and_collators.py:229-256
      SimpleStreamingCollator._prepare_dataset (same, but no BOS strip)
  ref:finetune/training/data_and_collator/datasets_and_collators.py:279-294
      collate_parquet (plain stack)
  ref:finetune/prepare_dataset/materialize_dataset_ray.py:43-49
      fixed-448 label padding, np.where(mask == 1, ids, -100)
and underneath them HF:tokenization_utils_base.py:2694-2783 (tokenizer.pad: pad to the
batch-longest, right side, pad_token_id, attention_mask 1/0 by LENGTH, not token value) and
HF:feature_extraction_sequence_utils.py:196-219 (feature_extractor.pad "longest" = stack).
Pinned by tests/golden/collate_*.npz (made from the unmodified reference class).
"""
from __future__ import annotations

import numpy as np

IGNORE_INDEX = -100


def pad_label_ids(id_lists, pad_token_id: int, max_length: int | None = None):
    """tokenizer.pad(..., padding=longest | max_length) -> (input_ids int64 (B,L), attention_mask int64 (B,L))."""
    lens = [len(x) for x in id_lists]
    width = max(lens) if max_length is None else max_length
    ids = np.full((len(id_lists), width), pad_token_id, dtype=np.int64)
    mask = np.zeros((len(id_lists), width), dtype=np.int64)
    for i, row in enumerate(id_lists):
        n = len(row)
        if n > width:
            raise ValueError("label longer than max_length (the reference does not truncate here)")
        ids[i, :n] = np.asarray(row, dtype=np.int64)
        mask[i, :n] = 1
    return ids, mask


def mask_labels(ids: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """labels_batch["input_ids"].masked_fill(attention_mask.ne(1), -100)  (ref ...:452, :252-254)."""
    return np.where(mask != 1, np.int64(IGNORE_INDEX), ids)


def stack_features(features) -> np.ndarray:
    """[{"input_features": np.vstack(list(f))}...] -> feature_extractor.pad('longest','pt') (ref ...:444-445).

    All items are (n_mel, 3000): pad() pads along axis 0 to the longest, i.e. a bit-exact stack.
    float64 inputs are cast to float32 (HF:feature_extraction_sequence_utils.py:215-216)."""
    mats = [np.vstack(list(f)) for f in features]
    longest = max(m.shape[0] for m in mats)
    out = []
    for m in mats:
        if m.dtype == np.float64:
            m = m.astype(np.float32)
        if m.shape[0] < longest:  # never happens for Whisper features; kept for fidelity
            m = np.pad(m, ((0, longest - m.shape[0]), (0, 0)), "constant", constant_values=0.0)
        out.append(m)
    return np.stack(out, axis=0)


def collate_padding(features, label_lists, pad_token_id: int, decoder_start_token_id: int):
    """DataCollatorSpeechSeq2SeqWithPadding.__call__ -> (input_features fp32 (B,n_mel,3000), labels int64)."""
    feats = stack_features(features)
    ids, mask = pad_label_ids(label_lists, pad_token_id)
    labels = mask_labels(ids, mask)
    if labels.shape[1] > 0 and bool((labels[:, 0] == decoder_start_token_id).all()):  # ref ...:456-457
        labels = labels[:, 1:]
    return feats, labels


def collate_streaming(features, label_lists, pad_token_id: int):
    """SimpleStreamingCollator._prepare_dataset: same padding/masking, NO BOS strip (ref ...:229-256)."""
    feats = stack_features(features)
    ids, mask = pad_label_ids(label_lists, pad_token_id)
    return feats, mask_labels(ids, mask)


def labels_fixed_length(id_list, pad_token_id: int, max_length: int = 448) -> np.ndarray:
    """materialize_dataset_ray.HDF5Worker.process_sample label half (ref ...materialize_dataset_ray.py:43-49)."""
    ids, mask = pad_label_ids([id_list], pad_token_id, max_length=max_length)
    return np.where(mask[0] == 1, ids[0], np.int64(IGNORE_INDEX))

"""Materialised (pre-computed) dataset path of the reference, fed by the GPU frontend.

The reference can pre-compute `input_features` / `labels` once and train from the stored arrays:
  * `finetune/prepare_dataset/materialize_dataset.py:63-183` maps a collator over index batches, adds
    `batch_dict["input_features"]` (fp32 (B, n_mel, 3000)) and `batch_dict["labels"]` (int64 (B, L)) as numpy and
    writes the batches as Parquet with the columns `idx, input_features, labels`;
  * `finetune/prepare_dataset/materialize_dataset_ray.py:27-63` (`HDF5Worker.process_sample`) produces per-sample
    records with the arrays as bytes + shape + dtype strings and labels padded to a fixed 448 with -100;
  * training reads them back through `collate_parquet` (`.../datasets_and_collators.py:279-294`).

Here the arrays come from the sm_100a kernels (`wfe_logmel`, `wfe_collate`); storage is plain pyarrow (no Ray): one row
per sample, `input_features` / `labels` as Arrow FIXED-SHAPE TENSOR columns (`arrow.fixed_shape_tensor`, the canonical
extension type; Ray Data writes ndarray columns the same way, as tensors with a fixed shape), so that any Arrow reader
gets (n_mel, 3000) float32 / (L,) int64 arrays per row -- which is what the reference's unmodified `collate_parquet`
stacks (tests/test_materialize_cpu.py feeds it rows read back with plain pyarrow).  HDF5 decode and tokenisation stay
outside (the callers pass decoded PCM and token ids), as everywhere in this package.
"""
from __future__ import annotations

import json
from typing import Iterable, Iterator, Optional, Sequence

import numpy as np
import torch

from .collator import _as_extractor, collate_labels_and_features

MAX_LABEL_LENGTH = 448  # ref:finetune/prepare_dataset/materialize_dataset_ray.py:33


def materialize_batch(feature_extractor, audio: Sequence[np.ndarray], label_ids: Sequence[Sequence[int]],
                      idx: Optional[Sequence[int]] = None, max_label_length: Optional[int] = MAX_LABEL_LENGTH) -> dict:
    """One `process_batch` of the reference (materialize_dataset.py:63-103) on the GPU.

    audio: decoded 16 kHz clips (float32 or int16); label_ids: token ids per clip (already tokenised).
    max_label_length: fixed label width (448, the Ray materialiser's choice) or None for batch-longest (the streaming
    collator's choice).  Returns host numpy arrays: {"idx" (B,) int64, "input_features" (B, n_mel, 3000) float32,
    "labels" (B, L) int64 with -100 on the padding}.
    """
    fe = _as_extractor(feature_extractor)
    if len(audio) == 0:
        raise RuntimeError("No valid data in batch")  # ref:.../datasets_and_collators.py:186-187
    if len(audio) != len(label_ids):
        raise ValueError("audio and label_ids must have the same length")
    feats = fe(list(audio), sampling_rate=fe.sampling_rate, return_tensors="pt")["input_features"]
    _, labels = collate_labels_and_features(fe, [list(x) for x in label_ids], None, width=max_label_length,
                                            decoder_start_token_id=-1, strip_bos=False)
    ids = np.arange(len(audio), dtype=np.int64) if idx is None else np.asarray(list(idx), dtype=np.int64)
    return {"idx": ids, "input_features": feats.numpy(), "labels": labels.cpu().numpy()}


def sample_records(batch: dict) -> list:
    """The per-sample serialisable dicts of `HDF5Worker.process_sample` (materialize_dataset_ray.py:52-60)."""
    out = []
    for i in range(len(batch["idx"])):
        f, lab = np.ascontiguousarray(batch["input_features"][i]), np.ascontiguousarray(batch["labels"][i])
        out.append({"idx": int(batch["idx"][i]), "input_features": f.tobytes(), "input_features_shape": f.shape,
                    "input_features_dtype": str(f.dtype), "labels": lab.tobytes(), "labels_shape": lab.shape,
                    "labels_dtype": str(lab.dtype)})
    return out


def record_to_arrays(rec: dict) -> dict:
    """Inverse of one `sample_records` entry."""
    f = np.frombuffer(rec["input_features"], dtype=np.dtype(rec["input_features_dtype"])).reshape(rec["input_features_shape"])
    lab = np.frombuffer(rec["labels"], dtype=np.dtype(rec["labels_dtype"])).reshape(rec["labels_shape"])
    return {"idx": rec["idx"], "input_features": f, "labels": lab}


def write_parquet(path: str, batches: Iterable[dict], row_group_size: int = 64) -> int:
    """Write materialised batches (dicts as returned by `materialize_batch`) to one Parquet file with the reference's
    column names `idx, input_features, labels`.  All batches must share the feature shape and the label width.
    Returns the number of rows written."""
    import pyarrow as pa
    import pyarrow.parquet as pq

    writer, rows, shapes = None, 0, None
    try:
        for b in batches:
            f = np.ascontiguousarray(b["input_features"], dtype=np.float32)
            lab = np.ascontiguousarray(b["labels"], dtype=np.int64)
            ids = np.asarray(b["idx"], dtype=np.int64)
            if f.shape[0] != lab.shape[0] or f.shape[0] != ids.shape[0]:
                raise ValueError("idx, input_features and labels must have the same number of rows")
            cur = (tuple(f.shape[1:]), tuple(lab.shape[1:]))
            if shapes is None:
                shapes = cur
                f_type = pa.fixed_shape_tensor(pa.float32(), list(cur[0]))
                l_type = pa.fixed_shape_tensor(pa.int64(), list(cur[1]))
                schema = pa.schema(
                    [pa.field("idx", pa.int64()), pa.field("input_features", f_type), pa.field("labels", l_type)],
                    metadata={"input_features_shape": json.dumps(cur[0]), "labels_shape": json.dumps(cur[1]),
                              "producer": "asr_finetune_b200"})
                writer = pq.ParquetWriter(path, schema)
            elif cur != shapes:
                raise ValueError(f"batch shapes {cur} differ from the file's {shapes} (use a fixed label width)")
            n = f.shape[0]
            table = pa.Table.from_arrays(
                [pa.array(ids), pa.FixedShapeTensorArray.from_numpy_ndarray(f),
                 pa.FixedShapeTensorArray.from_numpy_ndarray(lab)],
                schema=schema)
            writer.write_table(table, row_group_size=row_group_size)
            rows += n
    finally:
        if writer is not None:
            writer.close()
    if writer is None:
        raise ValueError("no batches to write")
    return rows


def iter_parquet(path: str, batch_size: int, columns: Sequence[str] = ("idx", "input_features", "labels")) -> Iterator[dict]:
    """Yield `{"idx", "input_features": [ (n_mel, 3000) float32 ... ], "labels": [ (L,) int64 ... ]}` batches — the dict
    of per-sample numpy arrays that `iter_torch_batches(collate_fn=collate_parquet)` hands to `collate_parquet`
    (ref:finetune/training/trainers/trainers.py:594-596)."""
    import pyarrow.parquet as pq

    pf = pq.ParquetFile(path)
    meta = pf.schema_arrow.metadata or {}
    fshape = tuple(json.loads(meta.get(b"input_features_shape", b"[]")))
    lshape = tuple(json.loads(meta.get(b"labels_shape", b"[]")))
    def rows(col, shape):
        if hasattr(col, "to_numpy_ndarray"):  # fixed-shape tensor column
            arr = col.to_numpy_ndarray()
        else:  # files written by round 1: flat fixed-size lists + the shape in the schema metadata
            arr = col.flatten().to_numpy(zero_copy_only=False).reshape((len(col),) + (shape if shape else (-1,)))
        return [arr[i] for i in range(arr.shape[0])]

    for rb in pf.iter_batches(batch_size=batch_size, columns=list(columns)):
        out: dict = {}
        if "idx" in columns:
            out["idx"] = rb.column(rb.schema.get_field_index("idx")).to_numpy(zero_copy_only=False)
        if "input_features" in columns:
            out["input_features"] = rows(rb.column(rb.schema.get_field_index("input_features")), fshape)
        if "labels" in columns:
            out["labels"] = rows(rb.column(rb.schema.get_field_index("labels")), lshape)
        yield out


def to_host_batch(batch: dict) -> dict:
    """CUDA tensors of a collator's output -> the numpy dict `write_parquet` takes."""
    out = {}
    for k, v in batch.items():
        out[k] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    if "idx" not in out:
        out["idx"] = np.arange(len(out["input_features"]), dtype=np.int64)
    return out

"""Materialised-dataset format (no GPU): Parquet round trip with the reference's column names, the per-sample record
format of `HDF5Worker.process_sample`, and the dict-of-arrays the Parquet reader hands to `collate_parquet`."""
import os

import numpy as np
import pytest

import asr_finetune_b200 as pkg

pa = pytest.importorskip("pyarrow")


def _batch(n, n_mel=80, width=448, seed=0, start=0):
    rng = np.random.default_rng(seed)
    labels = rng.integers(0, 50257, size=(n, width)).astype(np.int64)
    labels[:, width // 2:] = -100
    return {"idx": np.arange(start, start + n, dtype=np.int64),
            "input_features": rng.standard_normal((n, n_mel, 3000), dtype=np.float32),
            "labels": labels}


def test_parquet_round_trip_is_bit_exact_and_uses_reference_columns(tmp_path):
    import pyarrow.parquet as pq

    path = os.path.join(tmp_path, "train.parquet")
    b0, b1 = _batch(5, seed=1), _batch(3, seed=2, start=5)
    assert pkg.write_parquet(path, [b0, b1]) == 8
    # column names of ref:finetune/prepare_dataset/materialize_dataset.py:100-103
    assert pq.ParquetFile(path).schema_arrow.names == ["idx", "input_features", "labels"]
    got = list(pkg.iter_parquet(path, batch_size=4))
    assert [len(g["idx"]) for g in got] == [4, 4]
    feats = np.stack([f for g in got for f in g["input_features"]])
    labels = np.stack([x for g in got for x in g["labels"]])
    np.testing.assert_array_equal(np.concatenate([g["idx"] for g in got]), np.arange(8))
    np.testing.assert_array_equal(feats, np.concatenate([b0["input_features"], b1["input_features"]]))
    np.testing.assert_array_equal(labels, np.concatenate([b0["labels"], b1["labels"]]))
    assert feats.dtype == np.float32 and labels.dtype == np.int64
    assert got[0]["input_features"][0].shape == (80, 3000) and got[0]["labels"][0].shape == (448,)


def test_parquet_rejects_ragged_label_widths_and_empty_input(tmp_path):
    path = os.path.join(tmp_path, "x.parquet")
    with pytest.raises(ValueError):
        pkg.write_parquet(path, [_batch(2, width=448), _batch(2, width=100)])
    with pytest.raises(ValueError):
        pkg.write_parquet(os.path.join(tmp_path, "y.parquet"), [])


def test_sample_records_match_process_sample_format():
    b = _batch(3, n_mel=128, seed=3)
    recs = pkg.sample_records(b)
    # keys of ref:finetune/prepare_dataset/materialize_dataset_ray.py:52-60
    assert sorted(recs[0]) == sorted(["idx", "input_features", "input_features_shape", "input_features_dtype", "labels",
                                      "labels_shape", "labels_dtype"])
    assert recs[1]["input_features_shape"] == (128, 3000) and recs[1]["input_features_dtype"] == "float32"
    assert recs[1]["labels_shape"] == (448,) and recs[1]["labels_dtype"] == "int64"
    back = pkg.record_to_arrays(recs[2])
    np.testing.assert_array_equal(back["input_features"], b["input_features"][2])
    np.testing.assert_array_equal(back["labels"], b["labels"][2])
    assert back["idx"] == 2


def test_to_host_batch_adds_idx_and_converts_tensors():
    import torch

    out = pkg.to_host_batch({"input_features": torch.zeros(2, 80, 3000), "labels": torch.full((2, 7), -100)})
    assert out["input_features"].shape == (2, 80, 3000) and out["labels"].dtype == np.int64
    np.testing.assert_array_equal(out["idx"], [0, 1])


def test_plain_pyarrow_rows_feed_the_unmodified_reference_collate_parquet(tmp_path):
    """The on-disk format is readable without this package: rows read with plain pyarrow (Arrow fixed-shape tensors)
    go straight into the reference's own `collate_parquet` (ref ...datasets_and_collators.py:279-294), imported
    unmodified from /root/reference (build container only)."""
    import sys

    import pyarrow.parquet as pq
    import torch

    if not os.path.isdir("/root/reference/finetune/training"):
        pytest.skip("reference tree not mounted (build container only)")
    from conftest import GOLDEN_DIR

    sys.path.insert(0, GOLDEN_DIR)
    import make_golden as mg

    ref = mg.import_reference_collators()
    path = os.path.join(tmp_path, "train.parquet")
    b = _batch(6, n_mel=128, seed=11)
    pkg.write_parquet(path, [b])
    table = pq.read_table(path)  # no asr_finetune_b200 reader involved
    assert str(table.schema.field("input_features").type).startswith("extension<arrow.fixed_shape_tensor")
    feats = table.column("input_features").combine_chunks().to_numpy_ndarray()
    labels = table.column("labels").combine_chunks().to_numpy_ndarray()
    assert feats.shape == (6, 128, 3000) and labels.shape == (6, 448)
    out = ref.collate_parquet({"input_features": [np.array(x) for x in feats], "labels": [np.array(x) for x in labels]})
    assert out["input_features"].dtype == torch.float32 and out["labels"].dtype == torch.int64
    assert torch.equal(out["input_features"], torch.from_numpy(b["input_features"]))
    assert torch.equal(out["labels"], torch.from_numpy(b["labels"]))

"""Host-side logic of the drop-in (no GPU): constructor surface, filter bank, argument checks, `pad`, sharding,
and the N>1 per-rank layout over `gloo` (world_size 2)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import collate as ocollate
from oracle import signals

import asr_finetune_b200 as pkg


def test_constructor_attributes_match_hf_defaults(logmel_golden):
    for n_mel in (80, 128):
        fe = pkg.WhisperFeatureExtractor(feature_size=n_mel)
        assert (fe.feature_size, fe.sampling_rate, fe.hop_length, fe.n_fft, fe.chunk_length) == (n_mel, 16000, 160, 400, 30)
        assert fe.n_samples == 480000 and fe.nb_max_frames == 3000
        assert fe.padding_value == 0.0 and fe.padding_side == "right" and fe.return_attention_mask is False
        assert fe.model_input_names == ["input_features"]
        assert fe.mel_filters.dtype == np.float64 and fe.mel_filters.shape == (201, n_mel)
        np.testing.assert_allclose(fe.mel_filters, logmel_golden[f"mel_filters_{n_mel}"], rtol=0, atol=1e-15)


def test_matches_live_transformers_constructor_if_present():
    tr = pytest.importorskip("transformers")
    ref = tr.WhisperFeatureExtractor(feature_size=128)
    ours = pkg.WhisperFeatureExtractor(feature_size=128)
    for k in ("feature_size", "sampling_rate", "hop_length", "n_fft", "chunk_length", "n_samples", "nb_max_frames",
              "padding_value", "padding_side", "return_attention_mask", "dither"):
        assert getattr(ours, k) == getattr(ref, k), k
    np.testing.assert_allclose(ours.mel_filters, ref.mel_filters, rtol=0, atol=1e-15)
    assert ours.model_input_names == ref.model_input_names


def test_from_pretrained_reads_preprocessor_config_and_swallows_kwargs(tmp_path):
    d = tmp_path / "feature_extractor"
    d.mkdir()
    cfg = {"chunk_length": 30, "feature_extractor_type": "WhisperFeatureExtractor", "feature_size": 128, "hop_length": 160,
           "n_fft": 400, "n_samples": 480000, "nb_max_frames": 3000, "padding_side": "right", "padding_value": 0.0,
           "processor_class": "WhisperProcessor", "return_attention_mask": False, "sampling_rate": 16000}
    (d / "preprocessor_config.json").write_text(json.dumps(cfg))
    # exactly the reference call: ref:finetune/training/models/whisper_models.py:39
    fe = pkg.WhisperFeatureExtractor.from_pretrained(str(d), local_files_only=True, load_in_8bit=False)
    assert fe.feature_size == 128 and fe.n_samples == 480000
    out = fe.save_pretrained(str(tmp_path / "out"))
    fe2 = pkg.WhisperFeatureExtractor.from_pretrained(str(tmp_path / "out"))
    assert fe2.to_dict() == fe.to_dict() and os.path.exists(out[0])
    with pytest.raises(OSError):
        pkg.WhisperFeatureExtractor.from_pretrained(str(tmp_path / "nope"))


def test_sampling_rate_mismatch_raises_like_hf():
    fe = pkg.WhisperFeatureExtractor(feature_size=80)
    with pytest.raises(ValueError, match="sampling rate of 16000"):
        fe(np.zeros(16000, np.float32), sampling_rate=8000)
    with pytest.raises(ValueError, match="Only mono-channel"):
        fe(np.zeros((2, 2, 100), np.float32), sampling_rate=16000)


def test_resolve_length_rules():
    fe = pkg.WhisperFeatureExtractor(feature_size=80)
    assert fe._resolve_length([3, 640000], True, "max_length", None, None) == (480000, [3, 480000])
    assert fe._resolve_length([16000, 32000], True, "longest", None, None) == (32000, [16000, 32000])
    assert fe._resolve_length([16000], True, "max_length", 160000, None) == (160000, [16000])
    # the rules of HF's SequenceFeatureExtractor.pad (probed against the installed extractor, see the live test below)
    assert fe._resolve_length([16001], True, "longest", None, None) == (16001, [16001])  # any length: frames = n // 160
    assert fe._resolve_length([48077], True, "do_not_pad", None, None) == (48077, [48077])
    assert fe._resolve_length([48077], False, "max_length", 32000, None) == (48077, [48077])  # longer than the target: kept
    assert fe._resolve_length([48077], True, "max_length", 32077, None) == (32077, [32077])
    assert fe._resolve_length([100, 300], True, "longest", None, 256) == (512, [100, 300])
    with pytest.raises(ValueError, match="axes don't match array"):  # ragged batch: numpy's error inside HF
        fe._resolve_length([48077, 20000], True, "do_not_pad", None, None)
    with pytest.raises(ValueError, match="axes don't match array"):
        fe._resolve_length([48077, 20000], False, "max_length", 30000, None)
    with pytest.raises(NotImplementedError):
        fe._resolve_length([150], True, "longest", None, None)  # shorter than the centred reflect pad
    with pytest.raises(ValueError):
        fe._resolve_length([16000], True, "bogus", None, None)


def test_resolve_length_agrees_with_the_installed_extractor():
    """Frame counts / errors of the drop-in's length rules vs transformers' own padding logic (CPU, live)."""
    tr = pytest.importorskip("transformers")
    hf = tr.WhisperFeatureExtractor(feature_size=80)
    fe = pkg.WhisperFeatureExtractor(feature_size=80)
    x = np.random.default_rng(0).standard_normal(48077).astype(np.float32)
    cases = [([x], dict(padding="do_not_pad")), ([x], dict(padding=False)), ([x], dict(padding="longest")),
             ([x], dict(truncation=False, max_length=32000)), ([x], dict(padding="max_length", max_length=32077)),
             ([x, x[:20000]], dict(padding="longest")), ([x, x[:20000]], dict(padding="do_not_pad")),
             ([x, x[:20000]], dict(truncation=False, max_length=30000)), ([x[:777]], dict(padding="longest", pad_to_multiple_of=256))]
    for clips, kw in cases:
        args = ([len(c) for c in clips], kw.get("truncation", True), kw.get("padding", "max_length"), kw.get("max_length"),
                kw.get("pad_to_multiple_of"))
        try:
            ref = hf(clips if len(clips) > 1 else clips[0], sampling_rate=16000, return_attention_mask=True, **kw)
        except ValueError as e:
            with pytest.raises(ValueError, match=str(e)[:10]):
                fe._resolve_length(*args)
            continue
        n, used = fe._resolve_length(*args)
        assert ref["input_features"].shape[-1] == n // 160, (kw, n)
        assert ref["attention_mask"].shape[-1] == n // 160
        assert ref["attention_mask"].sum(-1).tolist() == [len(range(0, u, 160)) if u < n else n // 160 for u in used]


def test_pad_is_a_bit_exact_stack_like_the_reference(collate_golden):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden import fake_features

    fe = pkg.WhisperFeatureExtractor(feature_size=80)
    feats = fake_features(5, 4, 80)
    # the reference call: ref ...datasets_and_collators.py:236-240
    out = fe.pad([{"input_features": f} for f in feats], padding="longest", return_tensors="pt")
    assert out["input_features"].dtype == torch.float32 and tuple(out["input_features"].shape) == (4, 80, 3000)
    assert torch.equal(out["input_features"], torch.from_numpy(ocollate.stack_features(feats)))
    out64 = fe.pad({"input_features": [f.astype(np.float64) for f in feats]}, padding="longest", return_tensors="pt")
    assert out64["input_features"].dtype == torch.float32
    with pytest.raises(ValueError):
        fe.pad([{"labels": [1]}])


def test_rank_shard_is_equal_disjoint_cover():
    for n, w in [(100000, 8), (1024, 4), (17, 2), (5, 8)]:
        shards = [pkg.rank_shard(n, r, w) for r in range(w)]
        assert len({len(s) for s in shards}) == 1  # equal size (Ray's equal split), remainder dropped
        flat = [i for s in shards for i in s]
        assert len(flat) == len(set(flat)) == (n // w) * w
        full = [i for r in range(w) for i in pkg.rank_shard(n, r, w, drop_remainder=False)]
        assert full == list(range(n))
    assert [list(b) for b in pkg.shard_batches(10, 2, 1, 2)] == [[5, 6], [7, 8], [9]]
    assert [list(b) for b in pkg.shard_batches(10, 2, 1, 2, drop_last=True)] == [[5, 6], [7, 8]]
    with pytest.raises(ValueError):
        pkg.rank_shard(10, 2, 2)


def _gloo_worker(rank, world, port, n_clips, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    import asr_finetune_b200 as p

    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = p.rank_shard(n_clips)  # rank/world from the env, as under torchrun / Ray Train
    # per-rank collation width is the LOCAL batch max, exactly as each DDP rank collates its own batch
    labels = signals.label_ids(1337, n_clips, 5, 60)
    local = [labels[i] for i in shard]
    width = max(len(x) for x in local)
    # the only cross-rank traffic on this path is the timing/statistics gather for the report
    stats = torch.tensor([len(shard), shard.start, shard.stop, width], dtype=torch.int64)
    gathered = [torch.zeros_like(stats) for _ in range(world)]
    dist.all_gather(gathered, stats)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # bench.py's max-over-ranks timing
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        q.put(([g.tolist() for g in gathered], float(t.item())))


def test_two_rank_gloo_sharding_layout():
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 101, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0
    (n0, a0, b0, w0), (n1, a1, b1, w1) = gathered
    assert n0 == n1 == 50 and (a0, b0, a1, b1) == (0, 50, 50, 100)
    labels = signals.label_ids(1337, 101, 5, 60)
    assert w0 == max(len(x) for x in labels[:50]) and w1 == max(len(x) for x in labels[50:100])


def test_tools_compile():
    # the timing / diagnostic / fuzz scripts under tools/ only run on a GPU box: at least keep them syntactically alive
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    scripts = sorted(glob.glob(os.path.join(root, "tools", "*.py"))) + [os.path.join(root, "bench.py"),
                                                                         os.path.join(root, "__graft_entry__.py")]
    assert len(scripts) > 10
    for path in scripts:
        with open(path) as f:
            compile(f.read(), path, "exec")


def test_staging_copy_pool_stress_and_thread_sanitizer(tmp_path):
    # the host entry's staging threads (asr-finetune_b200/csrc/wfe_copy_pool.h, plain C++): random batches back to back,
    # every destination byte checked; then the same under ThreadSanitizer (an earlier lock-free piece claim could pair an
    # index of the previous batch with the new piece table when a worker woke up late)
    import ctypes
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "tests", "host", "copy_pool_host.cpp")
    gxx = shutil.which("g++")
    assert gxx, "g++ is part of the image"
    lib_path = str(tmp_path / "libcopy_pool_host.so")
    subprocess.run([gxx, "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", "-o", lib_path, src], check=True)
    lib = ctypes.CDLL(lib_path)
    lib.copy_pool_stress.restype = ctypes.c_int
    lib.copy_pool_stress.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_int]
    for threads in (0, 1, 3, 8):
        assert lib.copy_pool_stress(threads, 120, 1000 + threads, 0) == 0
        assert lib.copy_pool_stress(threads, 120, 2000 + threads, 1) == 0
    exe = str(tmp_path / "copy_pool_tsan")
    res = subprocess.run([gxx, "-O1", "-g", "-std=c++17", "-pthread", "-fsanitize=thread", "-DCOPY_POOL_MAIN", "-o", exe, src],
                         capture_output=True, text=True)
    if res.returncode != 0:
        pytest.skip("ThreadSanitizer runtime not available: " + res.stderr[-200:])
    run = subprocess.run([exe, "6", "25"], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0 and "0 bad batches" in run.stdout and "ThreadSanitizer" not in run.stderr, run.stderr[-2000:]

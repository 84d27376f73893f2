"""Drop-in speech seq2seq collators backed by ONE batched sm_100a kernel (`wfe_collate`).

  DataCollatorSpeechSeq2SeqWithPadding   ref:finetune/training/data_and_collator/datasets_and_collators.py:418-461
  StreamingFrontendCollator              GPU form of SimpleStreamingCollator's tail, ...:191-195 + 229-256
                                         (extractor loop + `_prepare_dataset`; HDF5 fetch and BPE stay on host)
  labels_fixed_length                    ref:finetune/prepare_dataset/materialize_dataset_ray.py:43-49

Semantics kept bit-exact: labels are right-padded to the batch-longest and the padding is replaced by -100 BY
LENGTH (Whisper's pad id == EOS id, the true EOS survives); the padding collator drops column 0 iff every row
starts with `decoder_start_token_id`; the streaming collator never does.  Returned tensors live on the CUDA
device (the reference's `data_collator_id` then has nothing left to move,
ref:finetune/training/trainers/utils.py:97-112); pass `device="cpu"` for host tensors.
"""
from __future__ import annotations

import ctypes as C
import itertools
import queue
import threading
from dataclasses import dataclass, field
from typing import Any, Callable, Iterable, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .feature_extraction import BatchFeature, WhisperFeatureExtractor, _cur_stream_ptr

IGNORE_INDEX = -100


def _as_extractor(fe: Any) -> WhisperFeatureExtractor:
    """Our extractor, or a B200 twin of a foreign (HF) one with the same config."""
    if isinstance(fe, WhisperFeatureExtractor):
        return fe
    twin = getattr(fe, "_b200_twin", None)
    if twin is None:
        twin = WhisperFeatureExtractor(feature_size=getattr(fe, "feature_size", 80),
                                       sampling_rate=getattr(fe, "sampling_rate", 16000),
                                       hop_length=getattr(fe, "hop_length", 160),
                                       chunk_length=getattr(fe, "chunk_length", 30), n_fft=getattr(fe, "n_fft", 400))
        try:
            fe._b200_twin = twin
        except Exception:
            pass
    return twin


def _pack_ids(label_lists: Sequence[Sequence[int]]):
    lens = np.fromiter((len(x) for x in label_lists), dtype=np.int64, count=len(label_lists))
    offs = np.zeros(len(label_lists) + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    total = int(offs[-1])
    # ids and offsets share one pinned buffer -> one H2D copy
    buf = torch.empty(total + len(offs), dtype=torch.int64, pin_memory=True)
    flat = buf.numpy()
    flat[:len(offs)] = offs
    if total:
        if all(isinstance(x, list) for x in label_lists):  # the common case: one pass over all ids
            flat[len(offs):] = np.fromiter(itertools.chain.from_iterable(label_lists), dtype=np.int64, count=total)
        else:
            pos = len(offs)
            for ids in label_lists:
                n = len(ids)
                if n:
                    flat[pos:pos + n] = ids.cpu().numpy() if torch.is_tensor(ids) else np.asarray(ids, dtype=np.int64)
                pos += n
    return buf, lens


def _all_rows_start_with(label_lists, token_id: int) -> bool:
    """The BOS-strip decision of the padding collator (ref ...datasets_and_collators.py:456-457), taken on the host:
    the ids are host lists, so there is no reason to synchronise with the device for it."""
    if len(label_lists) == 0:
        return False
    for ids in label_lists:
        if len(ids) == 0 or int(ids[0]) != token_id:
            return False
    return True


def collate_labels_and_features(fe: WhisperFeatureExtractor, label_lists, features, *, width: Optional[int],
                                decoder_start_token_id: int, strip_bos: bool, device: Optional[torch.device] = None,
                                feature_dtype: Optional[torch.dtype] = None):
    """-> (input_features (B, ...) CUDA or None, labels int64 CUDA).  One `wfe_collate` launch, no host sync."""
    dev = device or fe.cuda_device()
    h = fe._handle(None, dev)
    B = len(label_lists)
    packed, lens = _pack_ids(label_lists)
    if width is None:
        width = int(lens.max()) if B else 0
    elif B and int(lens.max()) > width:
        raise ValueError("label longer than max_length (the reference does not truncate labels)")
    with torch.cuda.device(dev):
        d_packed = packed.to(dev, non_blocking=True)
        d_offs, d_ids = d_packed[:B + 1], d_packed[B + 1:]
        labels = torch.empty((B, width), dtype=torch.int64, device=dev)
        feat_out, srcs_ptr, feat_elems, keep = None, None, 0, []
        if features is not None:
            mats = []
            for f in features:
                if torch.is_tensor(f):
                    t = f
                else:
                    # np.vstack(list(feature)) (ref ...:444): rows of a (n_mel, T) matrix stacked back
                    a = f if isinstance(f, np.ndarray) and f.ndim == 2 else np.vstack(list(f))
                    t = torch.from_numpy(np.ascontiguousarray(a))
                if t.dtype != torch.float32:
                    t = t.to(torch.float32)  # fp64 -> fp32 (HF:feature_extraction_sequence_utils.py:215-216)
                mats.append(t)
            shape = tuple(mats[0].shape)
            if any(tuple(m.shape) != shape for m in mats):
                raise NotImplementedError("feature matrices of different shapes in one batch")
            feat_elems = int(mats[0].numel())
            feat_out = torch.empty((B,) + shape, dtype=torch.float32, device=dev)
            if all(m.is_cuda for m in mats):
                keep = [m.contiguous() for m in mats]
                ptrs = torch.tensor([m.data_ptr() for m in keep], dtype=torch.int64).pin_memory().to(dev, non_blocking=True)
                keep.append(ptrs)
                srcs_ptr = ptrs.data_ptr()
            else:
                # host features: each clip's H2D copy lands directly in its slot of the batch tensor
                for i, m in enumerate(mats):
                    feat_out[i].copy_(m, non_blocking=True)
        _lib.check(h.lib.wfe_collate(h.ptr, d_ids.data_ptr() if d_ids.numel() else None, d_offs.data_ptr(), B, width,
                                     int(decoder_start_token_id), IGNORE_INDEX, labels.data_ptr() if width else None,
                                     None, srcs_ptr, feat_elems,
                                     feat_out.data_ptr() if srcs_ptr is not None else None, _cur_stream_ptr(dev)),
                   "wfe_collate")
        if strip_bos and width > 0 and _all_rows_start_with(label_lists, int(decoder_start_token_id)):  # ref ...:456-457
            labels = labels[:, 1:]
        if feat_out is not None and feature_dtype is not None and feat_out.dtype != feature_dtype:
            feat_out = feat_out.to(feature_dtype)
    del keep
    return feat_out, labels


@dataclass
class DataCollatorSpeechSeq2SeqWithPadding:
    """Same constructor and call contract as the reference class (ref ...datasets_and_collators.py:418-461).

    processor: anything with `.feature_extractor` and `.tokenizer` (only used to find the extractor config).
    decoder_start_token_id: BOS id removed from column 0 when every row starts with it.
    device: "cuda" (default: outputs stay on the GPU) or "cpu".
    """
    processor: Any
    decoder_start_token_id: int
    device: Optional[str] = None
    _fe: Any = field(default=None, init=False, repr=False)

    def __call__(self, features) -> BatchFeature:
        if self._fe is None:
            self._fe = _as_extractor(getattr(self.processor, "feature_extractor", self.processor))
        feats, labels = collate_labels_and_features(
            self._fe, list(features["labels"]), list(features["input_features"]), width=None,
            decoder_start_token_id=self.decoder_start_token_id, strip_bos=True)
        batch = BatchFeature({"input_features": feats})
        batch["labels"] = labels
        if self.device is not None and str(self.device) == "cpu":
            batch["input_features"] = batch["input_features"].cpu()
            batch["labels"] = batch["labels"].cpu()
        return batch


class _Worker:
    """One long-lived helper thread (the label collate overlaps the audio pipeline in the host-tensor path); a
    `ThreadPoolExecutor` built per call cost more than the work it ran at the reference's batch size of 8."""

    def __init__(self):
        self._q: "queue.Queue" = queue.Queue()
        self._t = threading.Thread(target=self._run, name="wfe-collate", daemon=True)
        self._t.start()

    def _run(self):
        while True:
            fn, box, done = self._q.get()
            try:
                box.append((True, fn()))
            except BaseException as e:  # handed back to the caller
                box.append((False, e))
            done.set()

    def submit(self, fn):
        box, done = [], threading.Event()
        self._q.put((fn, box, done))

        def result():
            done.wait()
            ok, val = box[0]
            if not ok:
                raise val
            return val

        return result


class StreamingFrontendCollator:
    """Raw audio + transcriptions -> {"input_features", "labels"} entirely on the GPU.

    GPU form of `SimpleStreamingCollator.__call__`/_prepare_dataset (ref ...:133-256) with the HDF5 fetch
    factored out: `batch` carries the decoded clips.  Keys: "audio" (list of 1-D float32/int16/float16 arrays) and either
    "labels" (list of id lists) or "transcription" (list of str, tokenised by `tokenizer`).  No BOS strip (ref quirk).

    Samples whose fetch failed are DROPPED like the reference does (`(idx, None, None)` entries are filtered out,
    ref ...:86-97,184): an entry whose audio is None (or empty) or whose label / transcription is None is skipped;
    a batch with nothing left raises `RuntimeError("No valid data in batch")` (ref ...:186-187).

    feature_dtype: torch.float32 (default) or float16 / bfloat16 — features rounded once in the kernel's epilogue, for
    the autocast consumer (`fp16 = True`, ref:finetune/training/configs/largev3_debug.config:8).
    """

    def __init__(self, feature_extractor, tokenizer: Optional[Callable] = None, device: Optional[str] = None,
                 feature_dtype: Optional[torch.dtype] = None):
        self.feature_extractor = _as_extractor(feature_extractor)
        self.tokenizer = tokenizer
        self.device = device
        self.feature_dtype = feature_dtype
        self.dropped = 0  # samples skipped so far
        self._worker: Optional[_Worker] = None

    def __call__(self, batch) -> dict:
        audio = list(batch["audio"])
        texts = batch["labels"] if "labels" in batch else batch.get("transcription")
        if texts is None:
            raise ValueError("need `labels` or `transcription`")
        texts = list(texts)
        if len(texts) != len(audio):
            raise ValueError("audio and labels must have the same length")
        keep = [i for i, (a, t) in enumerate(zip(audio, texts)) if a is not None and t is not None and len(a) > 0]
        if len(keep) != len(audio):
            self.dropped += len(audio) - len(keep)
            audio = [audio[i] for i in keep]
            texts = [texts[i] for i in keep]
        if len(audio) == 0:
            raise RuntimeError("No valid data in batch")  # ref ...:186-187
        if "labels" in batch:
            label_lists = [x if isinstance(x, list) else list(x) for x in texts]
        else:
            if self.tokenizer is None:
                raise ValueError("need `labels` or a tokenizer for `transcription`")
            label_lists = [self.tokenizer(t if isinstance(t, str) else str(t)).input_ids for t in texts]
        fe = self.feature_extractor
        if self.device is not None and str(self.device) == "cpu":
            # host tensors wanted: the audio goes through the pipelined host entry (chunked H2D / kernels / D2H on three
            # streams, GIL released) while the helper thread packs and collates the labels
            dev = fe.cuda_device()
            if self._worker is None:
                self._worker = _Worker()
            fut = self._worker.submit(lambda: collate_labels_and_features(fe, label_lists, None, width=None,
                                                                          decoder_start_token_id=-1, strip_bos=False,
                                                                          device=dev)[1].cpu())
            feats = fe(audio, sampling_rate=fe.sampling_rate, return_tensors="pt",
                       output_dtype=self.feature_dtype)["input_features"]
            return {"input_features": feats, "labels": fut()}
        # CUDA tensors wanted (the in-loop consumer): the same overlap -- the helper thread packs, uploads and collates
        # the labels while this one sits in the extractor's C entry (GIL released)
        dev = fe.cuda_device()
        if self._worker is None:
            self._worker = _Worker()
        fut = self._worker.submit(lambda: collate_labels_and_features(fe, label_lists, None, width=None,
                                                                      decoder_start_token_id=-1, strip_bos=False,
                                                                      device=dev)[1])
        out = fe(audio, sampling_rate=fe.sampling_rate, return_tensors="pt", output_device="cuda",
                 output_dtype=self.feature_dtype)
        return {"input_features": out["input_features"], "labels": fut()}


def labels_fixed_length(fe, id_list: Sequence[int], max_length: int = 448) -> torch.Tensor:
    """`tokenizer(text, padding="max_length", max_length=448)` + `np.where(mask == 1, ids, -100)`
    (ref:finetune/prepare_dataset/materialize_dataset_ray.py:43-49) -> int64 (max_length,) CUDA."""
    _, labels = collate_labels_and_features(_as_extractor(fe), [list(id_list)], None, width=max_length,
                                            decoder_start_token_id=-1, strip_bos=False)
    return labels[0]


def collate_parquet(batch, device: Optional[torch.device] = None) -> dict:
    """`collate_parquet` of the reference (ref ...datasets_and_collators.py:279-294): stack pre-materialised
    features/labels — here straight into CUDA batch tensors (pinned staging, one H2D per column)."""
    feats = [np.asarray(x, dtype=np.float32) for x in batch["input_features"]]
    labels = [np.asarray(x, dtype=np.int64) for x in batch["labels"]]
    dev = device or torch.device("cuda", torch.cuda.current_device())
    f = torch.empty((len(feats),) + feats[0].shape, dtype=torch.float32, pin_memory=True)
    lab = torch.empty((len(labels),) + labels[0].shape, dtype=torch.int64, pin_memory=True)
    fn, ln = f.numpy(), lab.numpy()
    for i, (a, b) in enumerate(zip(feats, labels)):  # (Parquet readers hand out read-only arrays: copy through numpy)
        np.copyto(fn[i], a)
        np.copyto(ln[i], b)
    return {"input_features": f.to(dev, non_blocking=True), "labels": lab.to(dev, non_blocking=True)}

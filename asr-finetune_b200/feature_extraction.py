"""Drop-in `WhisperFeatureExtractor` whose arithmetic runs in hand-written sm_100a kernels (libwfe.so).

Mirrors the call surface the reference uses:
  construct   HF:models/whisper/feature_extraction_whisper.py:69-103
              (`WhisperFeatureExtractor.from_pretrained(dir, local_files_only=True, load_in_8bit=...)` at
               ref:finetune/training/models/whisper_models.py:39,66)
  call        `fe(audio_1d_float32, sampling_rate=16000).input_features[0]`
              (ref:finetune/training/data_and_collator/datasets_and_collators.py:194-195,
               ref:finetune/prepare_dataset/materialize_dataset_ray.py:39-40) — full signature of
              HF:...feature_extraction_whisper.py:189-202 is accepted
  pad         `fe.pad(list_of_{"input_features"}, padding="longest", return_tensors="pt")`
              (ref ...datasets_and_collators.py:236-240,445)

Same arguments, same exceptions, same returned `input_features` (fp32 (B, n_mel, 3000)) and `attention_mask`
(int32 (B, 3000)).  Host code is Python/PyTorch plumbing only; there is NO CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import json
import logging
import os
import threading
from typing import Any, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib

logger = logging.getLogger(__name__)

try:  # same container type as the reference when transformers is present (it is a reference dependency)
    from transformers.feature_extraction_utils import BatchFeature  # type: ignore
except Exception:  # pragma: no cover - minimal stand-in with the same behaviour for our keys
    from collections import UserDict

    class BatchFeature(UserDict):  # type: ignore
        def __init__(self, data=None, tensor_type=None):
            super().__init__(data or {})
            if tensor_type is not None:
                self.convert_to_tensors(tensor_type)

        def __getattr__(self, item):
            try:
                return self.data[item]
            except KeyError:
                raise AttributeError(item)

        def convert_to_tensors(self, tensor_type=None):
            if tensor_type in ("pt", "torch"):
                for k, v in self.data.items():
                    if not torch.is_tensor(v):
                        self.data[k] = torch.as_tensor(np.asarray(v))
            elif tensor_type in ("np", "numpy"):
                for k, v in self.data.items():
                    self.data[k] = np.asarray(v)
            elif tensor_type is not None:
                raise ValueError(f"unsupported tensor type {tensor_type}")
            return self

        def to(self, *args, **kwargs):
            for k, v in self.data.items():
                if torch.is_tensor(v):
                    self.data[k] = v.to(*args, **kwargs)
            return self


# ---- slaney mel filter bank (HF:audio_utils.py:453-544 with norm="slaney", mel_scale="slaney") --------
def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) * (27.0 / np.log(6.4)), 3.0 * f / 200.0)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), 200.0 * m / 3.0)


def slaney_mel_filter_bank(num_frequency_bins: int, num_mel_filters: int, min_frequency: float, max_frequency: float,
                           sampling_rate: int) -> np.ndarray:
    """(num_frequency_bins, num_mel_filters) float64 triangular filters, slaney scale + slaney area norm."""
    mel_pts = np.linspace(_hz_to_mel(min_frequency), _hz_to_mel(max_frequency), num_mel_filters + 2)
    filt = _mel_to_hz(mel_pts)
    fft_hz = np.linspace(0, sampling_rate // 2, num_frequency_bins)
    diff = np.diff(filt)
    slopes = filt[None, :] - fft_hz[:, None]
    fb = np.maximum(0.0, np.minimum(-slopes[:, :-2] / diff[:-1], slopes[:, 2:] / diff[1:]))
    fb *= (2.0 / (filt[2:num_mel_filters + 2] - filt[:num_mel_filters]))[None, :]
    return fb


class _Handle:
    """Owns one `wfe_handle*` (constant tables for one (n_mel, n_samples, device))."""

    def __init__(self, n_mel: int, n_samples: int, mel_filters_f32: np.ndarray, device: int, sampling_rate: int):
        lib = _lib.load()
        cfg = _lib.WfeConfig(n_mel=n_mel, n_fft=400, hop_length=160, n_samples=n_samples, sampling_rate=sampling_rate,
                             device=device)
        filt = np.ascontiguousarray(mel_filters_f32, dtype=np.float32)
        ptr = C.c_void_p()
        _lib.check(lib.wfe_create(C.byref(cfg), filt.ctypes.data_as(C.c_void_p), C.byref(ptr)), "wfe_create")
        self.ptr, self.lib, self.device = ptr, lib, device
        self.n_mel, self.n_samples = n_mel, n_samples
        self.n_frames = int(lib.wfe_n_frames(ptr))

    def scratch_bytes(self, batch: int) -> int:
        return int(self.lib.wfe_logmel_scratch_bytes(self.ptr, batch))

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                self.lib.wfe_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


_DIRECT_DTYPES = (np.dtype(np.float32), np.dtype(np.int16), np.dtype(np.float16))
_OUT_DTYPES = {torch.float32: _lib.WFE_OUT_F32, torch.float16: _lib.WFE_OUT_F16, torch.bfloat16: _lib.WFE_OUT_BF16}


def _out_dtype(dtype) -> torch.dtype:
    """`output_dtype` argument -> torch dtype (None = float32, what the reference extractor returns)."""
    if dtype is None:
        return torch.float32
    if isinstance(dtype, str):
        dtype = {"float32": torch.float32, "fp32": torch.float32, "float16": torch.float16, "fp16": torch.float16,
                 "half": torch.float16, "bfloat16": torch.bfloat16, "bf16": torch.bfloat16}.get(dtype.lower(), dtype)
    if dtype not in _OUT_DTYPES:
        raise TypeError(f"output_dtype must be float32, float16 or bfloat16, got {dtype!r}")
    return dtype


def _cur_stream_ptr(device: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class WhisperFeatureExtractor:
    r"""B200-native Whisper feature extractor (log-mel spectrogram of 30-s, 16 kHz chunks).

    Arguments are those of `transformers.WhisperFeatureExtractor` (same names, defaults and meaning); unknown
    kwargs are swallowed like HF does (the reference passes `load_in_8bit=` to `from_pretrained`).

    Extra, optional knobs (all default to the reference behaviour):
        cuda_device   CUDA ordinal to run on; default `LOCAL_RANK` if set, else the current device
                      (mirrors ref:finetune/training/trainers/utils.py:108).
    """

    model_input_names = ["input_features"]

    def __init__(self, feature_size=80, sampling_rate=16000, hop_length=160, chunk_length=30, n_fft=400,
                 padding_value=0.0, dither=0.0, return_attention_mask=False, cuda_device: Optional[int] = None,
                 **kwargs):
        self.feature_size = feature_size
        self.sampling_rate = sampling_rate
        self.padding_value = padding_value
        self.padding_side = kwargs.pop("padding_side", "right")
        self.return_attention_mask = return_attention_mask
        self.do_normalize = kwargs.pop("do_normalize", False)
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.chunk_length = chunk_length
        self.n_samples = chunk_length * sampling_rate
        self.nb_max_frames = self.n_samples // hop_length
        self.dither = dither
        self.mel_filters = slaney_mel_filter_bank(1 + n_fft // 2, feature_size, 0.0, 8000.0, sampling_rate)
        self._extra = {k: v for k, v in kwargs.items() if k not in ("processor_class", "feature_extractor_type")}
        self._cuda_device = cuda_device
        self._handles: dict = {}
        self._lock = threading.Lock()
        self.last_transfer_bytes = (0, 0)  # (h2d, d2h) of the most recent host-buffer call

    # ---- construction helpers (HF FeatureExtractionMixin surface used by the reference) ----------------
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, **kwargs):
        path = str(pretrained_model_name_or_path)
        cfg_file = path if os.path.isfile(path) else os.path.join(path, "preprocessor_config.json")
        if not os.path.isfile(cfg_file):
            raise OSError(f"Can't load feature extractor for '{path}': no preprocessor_config.json found there "
                          "(this drop-in only loads local files; the reference uses local_files_only=True).")
        with open(cfg_file, "r", encoding="utf-8") as f:
            cfg = json.load(f)
        return cls.from_dict(cfg, **kwargs)

    @classmethod
    def from_dict(cls, cfg: dict, **kwargs):
        cfg = dict(cfg)
        for k in ("feature_extractor_type", "processor_class", "nb_max_frames", "n_samples", "mel_filters"):
            cfg.pop(k, None)
        return_unused = kwargs.pop("return_unused_kwargs", False)
        known = ("feature_size", "sampling_rate", "hop_length", "chunk_length", "n_fft", "padding_value", "dither",
                 "return_attention_mask", "cuda_device", "padding_side", "do_normalize")
        cfg.update({k: kwargs.pop(k) for k in list(kwargs) if k in known})
        obj = cls(**cfg)  # remaining kwargs (local_files_only, load_in_8bit, cache_dir, ...) are ignored like HF
        return (obj, kwargs) if return_unused else obj

    def to_dict(self) -> dict:
        return {"chunk_length": self.chunk_length, "dither": self.dither, "feature_extractor_type": "WhisperFeatureExtractor",
                "feature_size": self.feature_size, "hop_length": self.hop_length, "n_fft": self.n_fft,
                "n_samples": self.n_samples, "nb_max_frames": self.nb_max_frames, "padding_side": self.padding_side,
                "padding_value": self.padding_value, "return_attention_mask": self.return_attention_mask,
                "sampling_rate": self.sampling_rate}

    def save_pretrained(self, save_directory, **kwargs):
        os.makedirs(save_directory, exist_ok=True)
        out = os.path.join(save_directory, "preprocessor_config.json")
        with open(out, "w", encoding="utf-8") as f:
            json.dump(self.to_dict(), f, indent=2, sort_keys=True)
        return [out]

    def __repr__(self):
        return f"{self.__class__.__name__} {json.dumps(self.to_dict(), indent=2, sort_keys=True)}"

    # ---- device plumbing -------------------------------------------------------------------------------
    def cuda_device(self) -> torch.device:
        if self._cuda_device is not None:
            return torch.device("cuda", int(self._cuda_device))
        if not torch.cuda.is_available():
            raise RuntimeError("asr_finetune_b200.WhisperFeatureExtractor needs a CUDA device (B200, sm_100a); "
                               "there is no CPU fallback.")
        lr = os.environ.get("LOCAL_RANK")
        if lr is not None and int(lr) < torch.cuda.device_count():
            return torch.device("cuda", int(lr))
        return torch.device("cuda", torch.cuda.current_device())

    def _handle(self, n_samples: Optional[int] = None, device: Optional[torch.device] = None) -> _Handle:
        n_samples = self.n_samples if n_samples is None else int(n_samples)
        device = device or self.cuda_device()
        key = (n_samples, device.index)
        with self._lock:
            h = self._handles.get(key)
            if h is None:
                if self.n_fft != 400 or self.hop_length != 160:
                    raise NotImplementedError("sm_100a kernels are specialised for n_fft=400, hop_length=160")
                h = _Handle(self.feature_size, n_samples, self.mel_filters.astype(np.float32), device.index,
                            self.sampling_rate)
                self._handles[key] = h
            return h

    # ---- the hot path -----------------------------------------------------------------------------------
    def logmel_device(self, pcm: torch.Tensor, offsets: torch.Tensor, batch: int, *, n_samples: Optional[int] = None,
                      pcm_scale: float = 1.0, do_normalize: bool = False, return_attention_mask: bool = False,
                      out: Optional[torch.Tensor] = None, lengths: Optional[torch.Tensor] = None, out_dtype=None):
        """Device-resident entry: ragged `pcm` (float32, int16 or float16 CUDA tensor) + int64 `offsets` (CUDA; B+1
        entries, or B clip starts when `lengths` (B, int64, CUDA) is given — starts on 16-byte boundaries take the TMA /
        128-bit load paths) -> (input_features (B, n_mel, n_frames) CUDA, attention_mask (B, n_frames) int32 CUDA or
        None).  `out_dtype`: float32 (default, the reference's), float16 or bfloat16 — the fp32 result rounded once in
        the kernel's epilogue (the cast an autocast consumer applies anyway).  Runs on the current torch stream; no host
        synchronisation."""
        dev = pcm.device
        h = self._handle(n_samples, dev)
        if pcm.dtype == torch.float32:
            dt = _lib.WFE_PCM_F32
        elif pcm.dtype == torch.int16:
            dt = _lib.WFE_PCM_I16
        elif pcm.dtype == torch.float16:
            dt = _lib.WFE_PCM_F16
        else:
            raise TypeError(f"pcm must be float32, int16 or float16, got {pcm.dtype}")
        assert offsets.dtype == torch.int64 and offsets.is_cuda
        assert offsets.numel() == (batch + 1 if lengths is None else batch)
        assert lengths is None or (lengths.dtype == torch.int64 and lengths.is_cuda and lengths.numel() == batch)
        len_ptr = lengths.data_ptr() if lengths is not None else None
        if out is None:
            out = torch.empty((batch, h.n_mel, h.n_frames), dtype=_out_dtype(out_dtype), device=dev)
        elif out_dtype is not None and out.dtype != _out_dtype(out_dtype):
            raise TypeError("`out` does not have the requested out_dtype")
        if out.dtype not in _OUT_DTYPES or not out.is_contiguous():
            raise TypeError("`out` must be a contiguous float32 / float16 / bfloat16 tensor")
        mask = torch.empty((batch, h.n_frames), dtype=torch.int32, device=dev) if return_attention_mask else None
        scratch = torch.empty(max(h.scratch_bytes(batch), 4), dtype=torch.uint8, device=dev)
        if os.environ.get("WFE_DEBUG_POISON_SCRATCH"):
            # debugging aid: the allocator hands back the previous call's scratch, so a tile word the kernel forgot to
            # write would still hold the RIGHT value of an identical earlier run
            scratch.fill_(0xFF)
        stream = _cur_stream_ptr(dev)
        stats_ptr = None
        if do_normalize:
            stats = torch.empty((batch, 2), dtype=torch.float32, device=dev)
            _lib.check(h.lib.wfe_clip_stats(h.ptr, pcm.data_ptr(), dt, pcm_scale, offsets.data_ptr(), len_ptr, batch,
                                            stats.data_ptr(), stream), "wfe_clip_stats")
            stats_ptr = stats.data_ptr()
        _lib.check(h.lib.wfe_logmel_ex(h.ptr, pcm.data_ptr(), dt, pcm_scale, offsets.data_ptr(), len_ptr, batch, stats_ptr,
                                       out.data_ptr(), _OUT_DTYPES[out.dtype], mask.data_ptr() if mask is not None else None,
                                       scratch.data_ptr(), stream), "wfe_logmel")
        self._last_scratch = (h, scratch, batch)
        return out, mask

    def debug_kernel_error(self) -> int:
        """After a synchronize: non-zero if the last `logmel_device` launch gave up on an internal barrier (a bug)."""
        h, scratch, batch = self._last_scratch
        return int(h.lib.wfe_debug_scratch_error(h.ptr, scratch.data_ptr(), batch))

    def uses_tensor_cores(self, n_samples: Optional[int] = None) -> bool:
        """True when this configuration runs on the tcgen05 kernel (480000-sample window, slaney 80 / 128 mel bank)."""
        h = self._handle(n_samples)
        return bool(h.lib.wfe_uses_tensor_cores(h.ptr))

    def _extract_host(self, clips: Sequence[np.ndarray], n_samples: int, do_normalize: bool, want_mask: bool,
                      out_dtype=None, to_device: bool = False):
        """Host numpy clips -> host (pinned) torch tensors through the pipelined C entry point.  The PCM crosses PCIe in
        its own width only when EVERY clip has that 2-byte dtype (int16 as-is, like HF; float16 widened exactly);
        mixed batches have already been converted to float32 by `__call__`."""
        dev = self.cuda_device()
        h = self._handle(n_samples, dev)
        B = len(clips)
        kinds = {c.dtype for c in clips}
        if kinds == {np.dtype(np.int16)}:
            dt = _lib.WFE_PCM_I16
        elif kinds == {np.dtype(np.float16)}:
            dt = _lib.WFE_PCM_F16
        else:
            assert kinds == {np.dtype(np.float32)}, kinds
            dt = _lib.WFE_PCM_F32
        # pointer / length tables through numpy (`ndarray.ctypes` builds a ctypes object per clip: 1 us each, and a
        # training epoch makes this call for every batch)
        ptr_tab = np.fromiter((c.__array_interface__["data"][0] for c in clips), dtype=np.uint64, count=B)
        len_tab = np.fromiter((c.shape[0] for c in clips), dtype=np.int64, count=B)
        ptrs, lens = ptr_tab.ctypes.data, len_tab.ctypes.data  # plain addresses (the tables live until the call returns)
        odt = _out_dtype(out_dtype)
        if to_device:  # the in-loop training consumer: the kernels write straight into CUDA tensors, nothing comes back
            out = torch.empty((B, h.n_mel, h.n_frames), dtype=odt, device=dev)
            mask = torch.empty((B, h.n_frames), dtype=torch.int32, device=dev) if want_mask else None
        else:
            out = torch.empty((B, h.n_mel, h.n_frames), dtype=odt, pin_memory=True)
            mask = torch.empty((B, h.n_frames), dtype=torch.int32, pin_memory=True) if want_mask else None
        up, down = C.c_uint64(0), C.c_uint64(0)
        _lib.check(h.lib.wfe_extract_host_ex(h.ptr, ptrs, lens, B, dt, 1.0, int(bool(do_normalize)), out.data_ptr(),
                                             _OUT_DTYPES[odt], mask.data_ptr() if mask is not None else None,
                                             C.byref(up), C.byref(down)), "wfe_extract_host")
        self.last_transfer_bytes = (int(up.value), int(down.value))
        return out, mask

    def __call__(self, raw_speech, truncation: bool = True, pad_to_multiple_of: Optional[int] = None,
                 return_tensors: Optional[str] = None, return_attention_mask: Optional[bool] = None,
                 padding: Optional[str] = "max_length", max_length: Optional[int] = None,
                 sampling_rate: Optional[int] = None, do_normalize: Optional[bool] = None,
                 device: Optional[str] = "cpu", **kwargs) -> BatchFeature:
        """Same contract as HF:models/whisper/feature_extraction_whisper.py:189-342.

        `device` is accepted for signature compatibility; the STFT always runs on the CUDA device. Pass
        `output_device="cuda"` (extension) to keep the returned tensors on the GPU (`return_tensors="pt"` implied).
        """
        output_device = kwargs.pop("output_device", None)
        output_dtype = kwargs.pop("output_dtype", None)  # extension: float16 / bfloat16 features (fused cast)
        if sampling_rate is not None:
            if sampling_rate != self.sampling_rate:
                raise ValueError(
                    f"The model corresponding to this feature extractor: {self.__class__.__name__} was trained using a"
                    f" sampling rate of {self.sampling_rate}. Please make sure that the provided `raw_speech` input"
                    f" was sampled with {self.sampling_rate} and not {sampling_rate}.")
        else:
            logger.warning(f"It is strongly recommended to pass the `sampling_rate` argument to "
                           f"`{self.__class__.__name__}()`. Failing to do so can result in silent errors that might "
                           "be hard to debug.")
        if self.dither != 0.0:
            # HF adds `dither * randn` to the PADDED waveform (HF ...:146-147): the padded batch is built on the device
            return self._call_dithered(raw_speech, truncation, pad_to_multiple_of, return_tensors,
                                       return_attention_mask, padding, max_length, do_normalize, output_device,
                                       output_dtype)

        # ---- batched / unbatched normalisation (HF ...:274-290) ----
        if torch.is_tensor(raw_speech):
            if raw_speech.dim() > 2:
                raise ValueError(f"Only mono-channel audio is supported for input to {self}")
            clips_t = [raw_speech] if raw_speech.dim() == 1 else list(raw_speech)
            return self._call_device_tensors(clips_t, truncation, pad_to_multiple_of, return_tensors,
                                             return_attention_mask, padding, max_length, do_normalize, output_device,
                                             output_dtype)
        is_batched_numpy = isinstance(raw_speech, np.ndarray) and raw_speech.ndim > 1
        if is_batched_numpy and raw_speech.ndim > 2:
            raise ValueError(f"Only mono-channel audio is supported for input to {self}")
        is_batched = is_batched_numpy or (isinstance(raw_speech, (list, tuple)) and len(raw_speech) > 0 and
                                          isinstance(raw_speech[0], (np.ndarray, tuple, list, torch.Tensor)))
        if is_batched and len(raw_speech) > 0 and torch.is_tensor(raw_speech[0]):
            return self._call_device_tensors(list(raw_speech), truncation, pad_to_multiple_of, return_tensors,
                                             return_attention_mask, padding, max_length, do_normalize, output_device,
                                             output_dtype)
        seqs = list(raw_speech) if is_batched else [raw_speech]
        d0 = seqs[0].dtype if type(seqs[0]) is np.ndarray else None
        if d0 in _DIRECT_DTYPES and all(type(s) is np.ndarray and s.dtype == d0 and s.ndim == 1 and s.flags.c_contiguous
                                        for s in seqs):
            clips = seqs  # what the reference's loader yields: nothing to convert, nothing to copy
        else:
            clips = [np.asarray(s) for s in seqs]
            # int16 / float16 PCM travels in its own width only when the WHOLE batch has that dtype (the kernels convert
            # on load); anything else becomes float32 like HF does (fp64 / lists -> fp32, HF ...:282-286)
            narrow = len({c.dtype for c in clips}) == 1 and clips[0].dtype in (np.int16, np.float16)
            clips = [np.ascontiguousarray((c if narrow or c.dtype == np.float32 else c.astype(np.float32)).reshape(-1))
                     for c in clips]

        n_samples, lengths = self._resolve_length([int(c.shape[0]) for c in clips], truncation, padding, max_length,
                                                  pad_to_multiple_of)
        want_mask = bool(return_attention_mask if return_attention_mask is not None else self.return_attention_mask)
        norm = bool(do_normalize) if do_normalize is not None else bool(self.do_normalize)
        if any(c.shape[0] > n for c, n in zip(clips, lengths)):
            clips = [c[:n] for c, n in zip(clips, lengths)]  # truncation
        if output_device is not None and str(output_device).startswith("cuda"):
            # host clips in, CUDA tensors out: the same pipelined upload (staging threads, three streams), no download
            feats, mask = self._extract_host(clips, n_samples, norm, want_mask, output_dtype, to_device=True)
            data = {"input_features": feats}
            if want_mask:
                data["attention_mask"] = mask
            return BatchFeature(data)
        feats, mask = self._extract_host(clips, n_samples, norm, want_mask, output_dtype)
        if return_tensors in ("pt", "torch"):
            data = {"input_features": feats}
        else:  # numpy has no bfloat16: hand such features out as torch tensors whatever `return_tensors` says
            data = {"input_features": feats.numpy() if feats.dtype != torch.bfloat16 else feats}
        if want_mask:
            data["attention_mask"] = mask if return_tensors in ("pt", "torch") else mask.numpy()
        return BatchFeature(data)

    # ---- helpers ----------------------------------------------------------------------------------------
    def _resolve_length(self, lens, truncation, padding, max_length, pad_to_multiple_of):
        """-> (n, per-clip lengths): the common sample count `n` the STFT runs on and how many samples of each clip are
        used.  Follows `SequenceFeatureExtractor.pad` (HF:feature_extraction_sequence_utils.py:51-334): truncate to
        `max_length` (rounded up to `pad_to_multiple_of`) if `truncation`; pad to the longest / to `max_length` (rounded
        up) unless `padding` is False / "do_not_pad"; clips already longer than the target are left alone.  HF then
        stacks the clips into one array — clips of different final lengths fail there with the numpy error reproduced
        below.  Any common length >= 201 works (frames = n // 160; HF drops the last of the n // 160 + 1 STFT frames)."""
        max_length = max_length if max_length else self.n_samples
        if padding is True or padding == "longest":
            strategy = "longest"
        elif padding == "max_length" or padding is None:
            strategy = "max_length"
        elif padding is False or padding == "do_not_pad":
            strategy = "do_not_pad"
        else:
            raise ValueError(f"unknown padding strategy {padding!r}")
        up = (lambda n: ((n // pad_to_multiple_of) + 1) * pad_to_multiple_of
              if pad_to_multiple_of and n % pad_to_multiple_of else n)
        cut = [min(n, up(max_length)) if truncation else n for n in lens]
        if strategy == "do_not_pad":
            final = cut
        else:
            target = up(max(cut) if strategy == "longest" else max_length)
            final = [max(n, target) for n in cut]
        if len(set(final)) != 1:
            raise ValueError("axes don't match array")  # what numpy raises inside HF for a ragged batch
        n = final[0]
        if n <= self.n_fft // 2:
            raise NotImplementedError(f"padded length {n} is too short for the centred reflect pad ({self.n_fft // 2})")
        return n, cut

    def _call_device_tensors(self, clips, truncation, pad_to_multiple_of, return_tensors, return_attention_mask, padding,
                             max_length, do_normalize, output_device, output_dtype=None):
        """Input given as torch tensors (CPU or CUDA): results stay on the GPU unless asked otherwise."""
        dev = self.cuda_device() if not clips[0].is_cuda else clips[0].device
        clips = [c.reshape(-1) for c in clips]
        n_samples, lens_used = self._resolve_length([int(c.numel()) for c in clips], truncation, padding, max_length,
                                                    pad_to_multiple_of)
        want_mask = bool(return_attention_mask if return_attention_mask is not None else self.return_attention_mask)
        norm = bool(do_normalize) if do_normalize is not None else bool(self.do_normalize)
        kinds = {c.dtype for c in clips}
        dt = clips[0].dtype if len(kinds) == 1 and clips[0].dtype in (torch.int16, torch.float16) else torch.float32
        B = len(clips)
        lens = np.array(lens_used, dtype=np.int64)
        meta = np.zeros(2 * B, dtype=np.int64)  # clip starts (16-byte aligned), then lengths: one H2D copy
        meta[B:] = lens
        if B > 1:
            np.cumsum((lens[:-1] + 7) & ~7, out=meta[1:B])
        total = int(meta[B - 1] + lens[-1])
        with torch.cuda.device(dev):
            pcm = torch.empty(total + 8, dtype=dt, device=dev)
            for c, o, n in zip(clips, meta[:B].tolist(), lens.tolist()):
                pcm[o:o + n].copy_(c[:n], non_blocking=True)  # copy_ converts a stray dtype to `dt`
            d_meta = torch.from_numpy(meta).to(dev, non_blocking=True)
            feats, mask = self.logmel_device(pcm, d_meta[:B], B, n_samples=n_samples, do_normalize=norm,
                                             return_attention_mask=want_mask, lengths=d_meta[B:], out_dtype=output_dtype)
        keep = output_device is not None and str(output_device).startswith("cuda") or (output_device is None and clips[0].is_cuda)
        if not keep:
            feats = feats.cpu()
            mask = mask.cpu() if mask is not None else None
        as_pt = keep or return_tensors in ("pt", "torch") or feats.dtype == torch.bfloat16
        data = {"input_features": feats if as_pt else feats.numpy()}
        if want_mask:
            data["attention_mask"] = mask if as_pt else mask.numpy()
        return BatchFeature(data)

    def _call_dithered(self, raw_speech, truncation, pad_to_multiple_of, return_tensors, return_attention_mask, padding,
                       max_length, do_normalize, output_device, output_dtype=None):
        """`dither != 0`: zero-pad to the target length on the device, add `dither * N(0, 1)` to every sample of the
        padded buffer (padding included, as HF does), then run the kernels on the full-length clips.  The noise comes
        from torch's CUDA generator, so results are reproducible under `torch.manual_seed` but not bit-equal to HF."""
        if torch.is_tensor(raw_speech):
            seqs = [raw_speech] if raw_speech.dim() == 1 else list(raw_speech)
        elif isinstance(raw_speech, np.ndarray) and raw_speech.ndim > 1:
            seqs = list(raw_speech)
        elif isinstance(raw_speech, (list, tuple)) and len(raw_speech) and isinstance(
                raw_speech[0], (np.ndarray, tuple, list, torch.Tensor)):
            seqs = list(raw_speech)
        else:
            seqs = [raw_speech]
        dev = self.cuda_device()
        clips = [(c if torch.is_tensor(c) else torch.from_numpy(np.asarray(c, dtype=np.float32))).reshape(-1).to(
            torch.float32) for c in seqs]
        n_samples, lens = self._resolve_length([int(c.numel()) for c in clips], truncation, padding, max_length,
                                               pad_to_multiple_of)
        want_mask = bool(return_attention_mask if return_attention_mask is not None else self.return_attention_mask)
        norm = bool(do_normalize) if do_normalize is not None else bool(self.do_normalize)
        B = len(clips)
        with torch.cuda.device(dev):
            pcm = torch.zeros((B, n_samples), dtype=torch.float32, device=dev)
            for i, (c, n) in enumerate(zip(clips, lens)):
                pcm[i, :n].copy_(c[:n], non_blocking=True)
            if norm:  # zero-mean / unit-variance over the REAL samples comes before the dither (HF ...:306-312)
                for i, n in enumerate(lens):
                    v = pcm[i, :n]
                    pcm[i, :n] = (v - v.mean()) / torch.sqrt(v.var(unbiased=False) + 1e-7)
            pcm.add_(torch.randn(pcm.shape, dtype=pcm.dtype, device=dev), alpha=float(self.dither))
            offs = torch.arange(B + 1, dtype=torch.int64, device=dev) * n_samples
            feats, _ = self.logmel_device(pcm.view(-1), offs, B, n_samples=n_samples, out_dtype=output_dtype)
            mask = None
            if want_mask:
                t = torch.arange(n_samples // self.hop_length, device=dev) * self.hop_length
                mask = (t[None, :] < torch.tensor(lens, device=dev)[:, None]).to(torch.int32)
        keep = output_device is not None and str(output_device).startswith("cuda")
        if not keep:
            feats = feats.cpu()
            mask = mask.cpu() if mask is not None else None
        as_pt = keep or return_tensors in ("pt", "torch")
        data = {"input_features": feats if as_pt else feats.numpy()}
        if want_mask:
            data["attention_mask"] = mask if as_pt else mask.numpy()
        return BatchFeature(data)

    # ---- SequenceFeatureExtractor.pad for already-extracted features (ref ...datasets_and_collators.py:236,445)
    def pad(self, processed_features, padding: Union[bool, str] = True, max_length: Optional[int] = None,
            truncation: bool = False, pad_to_multiple_of: Optional[int] = None,
            return_attention_mask: Optional[bool] = None, return_tensors: Optional[str] = None) -> BatchFeature:
        """Stack per-clip feature matrices; pads along axis 0 to the longest, like
        HF:feature_extraction_sequence_utils.py:51-219 does for items shaped (n_mel, n_frames)."""
        if isinstance(processed_features, (list, tuple)) and len(processed_features) and isinstance(
                processed_features[0], (dict, BatchFeature)):
            processed_features = {k: [ex[k] for ex in processed_features] for k in processed_features[0].keys()}
        name = self.model_input_names[0]
        if name not in processed_features:
            raise ValueError("You should supply an instance of `transformers.BatchFeature` or list of "
                             f"`transformers.BatchFeature` to this method that includes {name}, but you provided "
                             f"{list(processed_features.keys())}")
        items = processed_features[name]
        if len(items) == 0:
            return BatchFeature({name: []})
        tens = []
        for it in items:
            t = it if torch.is_tensor(it) else torch.from_numpy(np.asarray(it))
            if t.dtype == torch.float64:
                t = t.to(torch.float32)
            tens.append(t)
        if padding is False or padding == "do_not_pad":
            target = None
        elif padding in (True, "longest"):
            target = max(t.shape[0] for t in tens)
        else:
            target = max_length if max_length is not None else max(t.shape[0] for t in tens)
        if target is not None and pad_to_multiple_of is not None and target % pad_to_multiple_of:
            target = ((target // pad_to_multiple_of) + 1) * pad_to_multiple_of
        out, masks = [], []
        for t in tens:
            n = t.shape[0]
            if truncation and target is not None and n > target:
                t, n = t[:target], target
            if target is not None and n < target:
                pad_shape = (target - n,) + tuple(t.shape[1:])
                t = torch.cat([t, torch.full(pad_shape, self.padding_value, dtype=t.dtype, device=t.device)], 0)
            masks.append(torch.cat([torch.ones(n, dtype=torch.int32), torch.zeros(t.shape[0] - n, dtype=torch.int32)]))
            out.append(t)
        want_mask = return_attention_mask if return_attention_mask is not None else self.return_attention_mask
        data: dict = {}
        if return_tensors in ("pt", "torch"):
            data[name] = torch.stack(out, 0)
            if want_mask:
                data["attention_mask"] = torch.stack(masks, 0)
        else:
            data[name] = [t.cpu().numpy() for t in out] if return_tensors is None else np.stack([t.cpu().numpy() for t in out])
            if want_mask:
                data["attention_mask"] = [m.numpy() for m in masks]
        return BatchFeature(data)

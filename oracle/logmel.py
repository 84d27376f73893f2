"""numpy restatement of the Whisper log-mel frontend (ORACLE — test infrastructure only).

The arithmetic of this path lives in a third-party dependency of the reference,
`transformers` (pinned ==4.46.3 in ref:requirements.txt:8; container has 5.5.0):

  HF:models/whisper/feature_extraction_whisper.py   (WhisperFeatureExtractor)
  HF:audio_utils.py                                  (mel_filter_bank, window_function, spectrogram)
  HF:feature_extraction_sequence_utils.py            (pad/_pad/_truncate)

called from ref:finetune/training/data_and_collator/datasets_and_collators.py:191-195
and ref:finetune/prepare_dataset/materialize_dataset_ray.py:39-40.

Each function cites the lines it restates.  `tests/test_oracle_cpu.py` pins this
file against the committed golden vectors (made by `tests/golden/make_golden.py`
from the real extractor) and, when transformers is importable, against the live one.
"""
from __future__ import annotations

import numpy as np

SAMPLING_RATE = 16000
N_FFT = 400
HOP = 160
CHUNK_S = 30
N_SAMPLES = CHUNK_S * SAMPLING_RATE  # 480000  (HF:...whisper.py:91)
N_FRAMES = N_SAMPLES // HOP  # 3000    (HF:...whisper.py:92)
N_BINS = N_FFT // 2 + 1  # 201


# ---- HF:audio_utils.py:285-296 (hertz_to_mel, slaney) -----------------------------------
def hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    mel = 3.0 * f / 200.0
    logstep = 27.0 / np.log(6.4)
    hi = f >= 1000.0
    return np.where(hi, 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) * logstep, mel)


# ---- HF:audio_utils.py:321-332 (mel_to_hertz, slaney) -----------------------------------
def mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f = 200.0 * m / 3.0
    logstep = np.log(6.4) / 27.0
    hi = m >= 15.0
    return np.where(hi, 1000.0 * np.exp(logstep * (m - 15.0)), f)


# ---- HF:audio_utils.py:453-544 + :356-375 (mel_filter_bank, norm=slaney, mel_scale=slaney)
def mel_filter_bank(n_mel: int, n_bins: int = N_BINS, fmin: float = 0.0, fmax: float = 8000.0,
                    sampling_rate: int = SAMPLING_RATE) -> np.ndarray:
    """(n_bins, n_mel) float64, as WhisperFeatureExtractor.__init__ builds it (HF:...whisper.py:95-103)."""
    mel_pts = np.linspace(hz_to_mel_slaney(fmin), hz_to_mel_slaney(fmax), n_mel + 2)
    filt_hz = mel_to_hz_slaney(mel_pts)
    fft_hz = np.linspace(0, sampling_rate // 2, n_bins)
    diff = np.diff(filt_hz)
    slopes = filt_hz[None, :] - fft_hz[:, None]
    down = -slopes[:, :-2] / diff[:-1]
    up = slopes[:, 2:] / diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    fb *= (2.0 / (filt_hz[2:n_mel + 2] - filt_hz[:n_mel]))[None, :]
    return fb


# ---- HF:audio_utils.py:593-607 (window_function "hann", periodic) == torch.hann_window(400)
def hann_periodic(n: int = N_FFT) -> np.ndarray:
    k = np.arange(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)


# ---- HF:feature_extraction_sequence_utils.py:265-278,327-332 (truncate + right zero-pad) ---
def pad_or_truncate(clip: np.ndarray, n_samples: int = N_SAMPLES):
    """-> (x fp32 (n_samples,), sample_mask int32 (n_samples,))."""
    clip = np.asarray(clip)
    if clip.dtype != np.float32:
        clip = clip.astype(np.float32)  # HF:...whisper.py:282-286
    length = min(len(clip), n_samples)
    x = np.zeros(n_samples, dtype=np.float32)
    x[:length] = clip[:length]
    mask = np.zeros(n_samples, dtype=np.int32)
    mask[:length] = 1
    return x, mask


def frames_reflect(x: np.ndarray, dtype) -> np.ndarray:
    """Centered framing: reflect-pad n_fft//2 each side (HF:audio_utils.py:769-771; torch.stft
    center=True, pad_mode='reflect'), hop 160; the LAST frame is dropped (HF:...whisper.py:128,150)."""
    xp = np.pad(x.astype(dtype), (N_FFT // 2, N_FFT // 2), mode="reflect")
    n_frames = len(x) // HOP  # 1 + len//hop frames exist; drop the last
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(n_frames)[:, None]
    return xp[idx]


def logmel_clip(clip: np.ndarray, n_mel: int, precision: str = "fp64", n_samples: int = N_SAMPLES) -> np.ndarray:
    """One clip -> (n_mel, n_samples//160) fp32.

    precision="fp64": HF:...whisper.py:105-133 (_np_extract_fbank_features -> spectrogram, float64).
    precision="fp32": HF:...whisper.py:135-164 (_torch_extract_fbank_features, float32 STFT/mel/log).
    """
    x, _ = pad_or_truncate(clip, n_samples)
    dt = np.float64 if precision == "fp64" else np.float32
    fr = frames_reflect(x, dt) * hann_periodic().astype(dt)[None, :]
    spec = np.fft.rfft(fr, axis=1)  # (T, 201); numpy>=2 keeps float32 -> complex64
    power = (spec.real.astype(dt) ** 2 + spec.imag.astype(dt) ** 2)  # abs()**2
    fb = mel_filter_bank(n_mel).astype(dt)  # fp32 cast at use: HF:...whisper.py:152
    mel = fb.T @ power.T  # (n_mel, T)
    log_spec = np.log10(np.maximum(mel, dt(1e-10)))  # :155 / audio_utils.py:813,819
    log_spec = np.maximum(log_spec, log_spec.max() - dt(8.0))  # per-clip max: :156-158 / :129
    log_spec = (log_spec + dt(4.0)) / dt(4.0)  # :161 / :130
    return log_spec.astype(np.float32)


def logmel_batch(clips, n_mel: int, precision: str = "fp64", n_samples: int = N_SAMPLES) -> np.ndarray:
    return np.stack([logmel_clip(c, n_mel, precision, n_samples) for c in clips], axis=0)


# ---- HF:...whisper.py:328-337 (mask decimation) -------------------------------------------
def frame_attention_mask(lengths, n_samples: int = N_SAMPLES) -> np.ndarray:
    lengths = np.minimum(np.asarray(lengths, dtype=np.int64), n_samples)
    t = np.arange(n_samples // HOP, dtype=np.int64) * HOP
    return (t[None, :] < lengths[:, None]).astype(np.int32)


# ---- HF:...whisper.py:168-187 (zero_mean_unit_var_norm, do_normalize=True) ----------------
def zero_mean_unit_var(x_padded: np.ndarray, length: int, padding_value: float = 0.0) -> np.ndarray:
    v = x_padded.astype(np.float32)
    out = (v - v[:length].mean()) / np.sqrt(v[:length].var() + 1e-7)
    if length < out.shape[0]:
        out[length:] = padding_value
    return out.astype(np.float32)

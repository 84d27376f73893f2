"""Deterministic synthetic 16 kHz audio and label-id generators (test infrastructure).

Everything is integer-hash based (splitmix64 on a sample counter) so the same
(seed, index) gives the same bits on every numpy version; tests, golden-fixture
generation and bench.py all draw their inputs from here.  Signal families follow
SURVEY.md §8(c)/(d): gaussian-ish noise 0.1*N(0,1), pure tones, chirp, AM,
zero-tailed clips, int16-quantised audio.
"""
from __future__ import annotations

import numpy as np

SR = 16000
N_SAMPLES = 480000

# Whisper multilingual special ids used for synthetic labels
# (ref: finetune/training/trainers/trainers.py:328 forces language/task tokens).
SOT, LANG_DE, TRANSCRIBE, NOTIMESTAMPS, EOT = 50258, 50261, 50360, 50364, 50257


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform_u32(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """n uint32 values from counter-mode splitmix64."""
    with np.errstate(over="ignore"):
        ctr = np.arange(n, dtype=np.uint64) + np.uint64((seed * 0x100000001B3 + stream * 0x51_7C_C1_B7_27_22_0A_95) & 0xFFFFFFFFFFFFFFFF)
        return (_splitmix64(ctr) >> np.uint64(32)).astype(np.uint32)


def noise(seed: int, n: int = N_SAMPLES, amp: float = 0.1) -> np.ndarray:
    """Approximately N(0, amp^2): Irwin-Hall sum of 4 uniforms, fp32."""
    acc = np.zeros(n, dtype=np.float64)
    for s in range(4):
        acc += uniform_u32(seed, n, stream=s).astype(np.float64) * (1.0 / 4294967296.0)
    g = (acc - 2.0) * np.sqrt(3.0)  # var of sum of 4 U(0,1) = 1/3
    return (amp * g).astype(np.float32)


def tone(freq: float, n: int = N_SAMPLES, amp: float = 0.5) -> np.ndarray:
    t = np.arange(n, dtype=np.float64) / SR
    return (amp * np.sin(2.0 * np.pi * freq * t)).astype(np.float32)


def chirp(f0: float = 50.0, f1: float = 7500.0, n: int = N_SAMPLES, amp: float = 0.3) -> np.ndarray:
    t = np.arange(n, dtype=np.float64) / SR
    dur = max(n / SR, 1e-9)
    phase = 2.0 * np.pi * (f0 * t + 0.5 * (f1 - f0) / dur * t * t)
    return (amp * np.sin(phase)).astype(np.float32)


def am_tone(fc: float = 2000.0, fm: float = 3.0, n: int = N_SAMPLES, amp: float = 0.4) -> np.ndarray:
    t = np.arange(n, dtype=np.float64) / SR
    return (amp * (0.5 + 0.5 * np.sin(2.0 * np.pi * fm * t)) * np.sin(2.0 * np.pi * fc * t)).astype(np.float32)


def quantise_int16(x: np.ndarray) -> np.ndarray:
    """What an HDF5 int16 recording decoded to float32 looks like."""
    q = np.clip(np.round(x.astype(np.float64) * 32768.0), -32768, 32767)
    return (q / 32768.0).astype(np.float32)


def speechlike(seed: int, n: int = N_SAMPLES) -> np.ndarray:
    """Bursts of harmonics with silences in between: exercises the max-8 clamp."""
    t = np.arange(n, dtype=np.float64) / SR
    env_bits = uniform_u32(seed, (n + 3199) // 3200, stream=7)  # 0.2 s segments
    env = np.repeat((env_bits % 3 != 0).astype(np.float64), 3200)[:n]
    f0 = 110.0 + (seed % 7) * 13.0
    sig = np.zeros(n, dtype=np.float64)
    for h in range(1, 9):
        sig += np.sin(2.0 * np.pi * f0 * h * t + h) / h
    out = 0.2 * env * sig + 1e-4 * noise(seed + 1, n, amp=1.0).astype(np.float64)
    return out.astype(np.float32)


def bursty(seed: int, n: int = N_SAMPLES, span_db: float = 60.0, amp: float = 0.1) -> np.ndarray:
    """White noise under 0.2-s segments with gains spread over `span_db`: speech-like dynamics, so that the per-clip
    clamp (max - 8) has elements to clamp in most 128-frame tiles (the data-dependent part of the kernel)."""
    u = uniform_u32(seed, (n + 3199) // 3200, stream=19).astype(np.float64) / 4294967296.0
    gain = np.repeat(10.0 ** (-(span_db / 20.0) * u), 3200)[:n]
    return (noise(seed + 101, n, amp=1.0).astype(np.float64) * amp * gain).astype(np.float32)


def click_in_silence(n: int = N_SAMPLES, at: float = 0.5) -> np.ndarray:
    """A loud 5-sample click in near-silence (1e-4 noise): one tile holds the clip maximum, every other tile sits at or
    below the clamp floor."""
    x = 1e-4 * noise(23, n, amp=1.0).astype(np.float64)
    c = int(n * at)
    x[c:c + 5] += np.array([0.3, 0.9, -0.9, 0.6, -0.2])[:max(0, min(5, n - c))]
    return x.astype(np.float32)


def clip_lengths(seed: int, batch: int, lo: int = SR, hi: int = N_SAMPLES) -> np.ndarray:
    """L_i ~ U{lo..hi} (seed 1337 is the reference's random_seed)."""
    u = uniform_u32(seed, batch, stream=11).astype(np.uint64)
    return (lo + (u * np.uint64(hi - lo + 1) >> np.uint64(32))).astype(np.int64)


def label_ids(seed: int, batch: int, lo: int = 5, hi: int = 448, with_bos: bool = True) -> list[list[int]]:
    """Synthetic Whisper label id lists: [SOT, de, transcribe, notimestamps, text..., EOT]."""
    lens = clip_lengths(seed, batch, lo, hi)
    out = []
    for i, n in enumerate(lens.tolist()):
        body = (uniform_u32(seed + 17 * (i + 1), max(n - 5, 0), stream=13) % np.uint32(50257)).astype(np.int64).tolist()
        ids = ([SOT] if with_bos else []) + [LANG_DE, TRANSCRIBE, NOTIMESTAMPS] + body + [EOT]
        out.append(ids)
    return out


def named_case(name: str, n: int = N_SAMPLES) -> np.ndarray:
    """The known-answer signal set of SURVEY §8(c)."""
    if name == "zeros":
        return np.zeros(n, dtype=np.float32)
    if name == "ones":
        return np.ones(n, dtype=np.float32)
    if name == "tone1k":
        return tone(1000.0, n, 0.5)
    if name == "tone1k_quiet":
        return (tone(1000.0, n, 0.5) * np.float32(1e-3)).astype(np.float32)
    if name == "chirp":
        return chirp(n=n)
    if name == "am":
        return am_tone(n=n)
    if name == "noise":
        return noise(0, n)
    if name == "noise_q16":
        return quantise_int16(noise(3, n))
    if name == "speechlike":
        return speechlike(5, n)
    if name == "impulse":
        x = np.zeros(n, dtype=np.float32)
        x[n // 3] = 1.0
        return x
    raise KeyError(name)


NAMED_CASES = ["zeros", "ones", "tone1k", "tone1k_quiet", "chirp", "am", "noise", "noise_q16", "speechlike", "impulse"]

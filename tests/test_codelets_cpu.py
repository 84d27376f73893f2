"""The device FFT codelets (csrc/wfe_codelets.cuh) compiled as plain C++ and checked against numpy's FFT.

Same straight-line arithmetic the sm_100a kernel runs per thread (400 = 16 x 25 Cooley-Tukey, real input, folded
bins, in-place power), so the maths is proven on the CPU before any GPU time is spent."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import logmel as ologmel
from oracle import signals

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")


@pytest.fixture(scope="module")
def codelet_lib():
    src = os.path.join(ROOT, "tests", "host", "codelet_host.cpp")
    out = os.path.join(ROOT, "tests", "host", "libcodelet_host.so")
    hdr = os.path.join(ROOT, "asr-finetune_b200", "csrc", "wfe_codelets.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", out, src], check=True)
    lib = ctypes.CDLL(out)
    lib.codelet_tile_power.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.codelet_tile_power.restype = None
    return lib


def tile_power(lib, sig):
    sig = np.ascontiguousarray(sig, dtype=np.float32)
    assert sig.shape == (5360,)
    power = np.zeros((201, 32), dtype=np.float32)
    lib.codelet_tile_power(sig.ctypes.data, power.ctypes.data)
    return power.T  # (32 frames, 201 bins)


def ref_power(sig):
    fr = np.stack([sig[160 * t:160 * t + 400] for t in range(32)]).astype(np.float64) * ologmel.hann_periodic()
    return np.abs(np.fft.rfft(fr, axis=1)) ** 2


@pytest.mark.parametrize("name", ["noise", "tone1k", "chirp", "ones", "impulse", "speechlike"])
def test_tile_power_matches_numpy_fft(codelet_lib, name):
    start = 200000 // 3 - 2500 if name == "impulse" else 100000  # keep the impulse inside the tile
    sig = signals.named_case(name, 200000)[start:start + 5360]
    got, ref = tile_power(codelet_lib, sig), ref_power(sig)
    scale = max(ref.max(), 1e-30)
    # fp32 FFT: absolute error relative to the frame's peak power ~1e-7; bins that matter (> peak*1e-8) to 1%
    assert np.abs(got - ref).max() / scale < 2e-6
    big = ref > scale * 1e-7
    assert (np.abs(got - ref)[big] / ref[big]).max() < 2e-2


def test_zero_input_is_exactly_zero(codelet_lib):
    assert not tile_power(codelet_lib, np.zeros(5360, np.float32)).any()


def test_single_bin_tones_hit_every_bin(codelet_lib):
    # a complex exponential at bin k puts (almost) all energy in bins k-1..k+1 under the Hann window: checks the fold
    # of bins 13..24 mod 25 (conjugate symmetry) and bin_to_row for every k
    t = np.arange(5360)
    for k in list(range(0, 201, 7)) + [12, 13, 24, 25, 26, 187, 188, 199, 200]:
        sig = np.cos(2 * np.pi * k * t / 400.0).astype(np.float32)
        got, ref = tile_power(codelet_lib, sig), ref_power(sig)
        assert int(got[0].argmax()) == int(ref[0].argmax()) == k
        assert np.abs(got - ref).max() / ref.max() < 2e-6


def test_logmel_through_codelets_matches_oracle(codelet_lib):
    # full clip through the codelets + numpy mel/log: the kernel's algorithm end to end on the CPU (1 s clip)
    clip = signals.noise(5, 16000)
    n_samples = 16000
    x = np.pad(clip, (200, 200), mode="reflect")
    n_frames = n_samples // 160
    fb = ologmel.mel_filter_bank(128).astype(np.float32)
    out = np.zeros((128, n_frames), np.float32)
    for t0 in range(0, n_frames, 32):
        sig = np.zeros(5360, np.float32)
        seg = x[160 * t0:160 * t0 + 5360]
        sig[:len(seg)] = seg
        p = tile_power(codelet_lib, sig)
        nv = min(32, n_frames - t0)
        mel = fb.T @ p[:nv].T
        out[:, t0:t0 + nv] = np.log10(np.maximum(mel, 1e-10))
    out = np.maximum(out, out.max() - 8.0)
    out = (out + 4.0) / 4.0
    ref = ologmel.logmel_clip(clip, 128, "fp64", n_samples=n_samples)
    assert np.abs(out - ref).max() < 2e-4

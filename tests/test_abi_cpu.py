"""The C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol include/wfe.h declares.
No compute entry point is exercised here (that is the `-m gpu` suite); without a device they must fail loudly."""
import ctypes as C
import os
import re
import shutil

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"),
                                reason="nvcc not available")


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as g

    mod = g._load_build_module()
    mod.build()  # no-op when libwfe.so is newer than its sources
    import asr_finetune_b200

    return asr_finetune_b200


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "wfe.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wfe_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(pkg):
    lib = pkg._lib.load()
    syms = declared_symbols()
    assert len(syms) >= 11
    for s in syms:
        assert hasattr(lib, s), f"libwfe.so does not export {s} declared in include/wfe.h"
    assert sorted(pkg._lib.SYMBOLS) == syms, "asr_finetune_b200._lib.SYMBOLS is out of sync with include/wfe.h"
    assert lib.wfe_abi_version() == 3
    assert lib.wfe_launch_count() == 0 or lib.wfe_launch_count() > 0  # callable without a device


def test_sass_is_sm100a_only(pkg):
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", pkg._lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_device(pkg):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = pkg._lib.load()
    cfg = pkg._lib.WfeConfig(n_mel=128, n_fft=400, hop_length=160, n_samples=480000, sampling_rate=16000, device=0)
    filt = np.zeros((201, 128), np.float32)
    h = C.c_void_p()
    rc = lib.wfe_create(C.byref(cfg), filt.ctypes.data_as(C.c_void_p), C.byref(h))
    assert rc == -3 and not h.value  # WFE_ERR_CUDA
    assert b"no CPU fallback" in lib.wfe_last_error()
    fe = pkg.WhisperFeatureExtractor(feature_size=128)
    with pytest.raises((RuntimeError, pkg._lib.WfeError)):
        fe(np.zeros(16000, np.float32), sampling_rate=16000)


def test_bad_config_is_rejected_before_touching_cuda(pkg):
    lib = pkg._lib.load()
    filt = np.zeros((201, 128), np.float32)
    h = C.c_void_p()
    cfg = pkg._lib.WfeConfig(n_mel=128, n_fft=512, hop_length=160, n_samples=480000, sampling_rate=16000, device=0)
    assert lib.wfe_create(C.byref(cfg), filt.ctypes.data_as(C.c_void_p), C.byref(h)) == -2  # WFE_ERR_UNSUPPORTED
    assert lib.wfe_create(None, None, None) == -1
    assert lib.wfe_logmel(None, None, 0, 1.0, None, None, 1, None, None, None, None, None) == -1
    assert lib.wfe_collate(None, None, None, 1, 1, 0, -100, None, None, None, 0, None, None) == -1
    assert lib.wfe_extract_host(None, None, None, 1, 0, 1.0, 0, None, None, None, None) == -1


def test_product_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "asr-finetune_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"
                assert "torch.stft" not in text.replace("no cuFFT/`torch.stft`", "") or f.endswith(".md")

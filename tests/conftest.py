import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: larger CPU cases")


@pytest.fixture(scope="session")
def logmel_golden():
    return np.load(os.path.join(GOLDEN_DIR, "logmel_golden.npz"))


@pytest.fixture(scope="session")
def collate_golden():
    return np.load(os.path.join(GOLDEN_DIR, "collate_golden.npz"))


def golden_logmel_cases():
    """Same list as tests/golden/make_golden.py:logmel_cases (inputs are regenerated, outputs are stored)."""
    from oracle import signals

    cases = []
    for name in signals.NAMED_CASES:
        cases.append((f"{name}_128", 128, lambda name=name: signals.named_case(name)))
    for name in ["zeros", "tone1k", "noise", "speechlike", "chirp"]:
        cases.append((f"{name}_80", 80, lambda name=name: signals.named_case(name)))
    for n in [3, 1, 200, 201, 16000, 112123, 479999, 640000]:
        cases.append((f"noise_len{n}_128", 128, lambda n=n: signals.noise(100 + n % 97, n)))
    cases.append(("tone_len112123_80", 80, lambda: signals.tone(440.0, 112123, 0.3)))
    cases.append(("speechlike_len250000_128", 128, lambda: signals.speechlike(9, 250000)))
    cases.append(("bursty60_128", 128, lambda: signals.bursty(11)))
    cases.append(("bursty60_80", 80, lambda: signals.bursty(12)))
    cases.append(("bursty60_len300007_128", 128, lambda: signals.bursty(13, 300007)))
    cases.append(("click_128", 128, lambda: signals.click_in_silence()))
    return cases


def check_against_golden(out, golden, key, tag, tol):
    """out: (n_mel, 3000) fp32 vs the stored sub-sampled reference output; returns max-abs error."""
    sub = golden[f"{key}/{tag}/sub"]
    head = golden[f"{key}/{tag}/head"]
    tail = golden[f"{key}/{tag}/tail"]
    err = max(
        float(np.abs(out[:, ::25] - sub).max()),
        float(np.abs(out[:, :48] - head).max()),
        float(np.abs(out[:, -48:] - tail).max()),
    )
    stats = golden[f"{key}/{tag}/stats"]
    err = max(err, abs(float(out.max()) - stats[0]), abs(float(out.min()) - stats[1]))
    assert err <= tol, f"{key}/{tag}: max-abs-err {err} > {tol}"
    return err

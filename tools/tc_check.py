"""Quick GPU check of the tcgen05 log-mel kernel against the fp64 oracle and the CUDA-core kernel (diagnostic tool)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import asr_finetune_b200 as pkg  # noqa: E402
from oracle import logmel as O  # noqa: E402
from oracle import signals as S  # noqa: E402


def run(fe, clips, n_mel):
    dev = torch.device("cuda", 0)
    lens = np.array([min(len(c), 480000) for c in clips], dtype=np.int64)
    starts = np.zeros(len(clips), dtype=np.int64)
    if len(clips) > 1:
        starts[1:] = np.cumsum((lens[:-1] + 7) & ~7)
    pcm = torch.zeros(int(starts[-1] + lens[-1] + 8), dtype=torch.float32, device=dev)
    for c, o, n in zip(clips, starts, lens):
        pcm[o:o + n] = torch.from_numpy(c[:n]).to(dev)
    out, mask = fe.logmel_device(pcm, torch.from_numpy(starts).to(dev), len(clips), lengths=torch.from_numpy(lens).to(dev),
                                 return_attention_mask=True)
    torch.cuda.synchronize()
    return out.cpu().numpy(), mask.cpu().numpy(), fe.debug_kernel_error()


def main():
    n_mel = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    fe = pkg.WhisperFeatureExtractor(feature_size=n_mel, cuda_device=0)
    h = fe._handle(None, torch.device("cuda", 0))
    print("uses tensor cores:", h.lib.wfe_uses_tensor_cores(h.ptr))
    sets = {
        "smoke4": [S.noise(1, 48000), S.tone(440.0, 20000, 0.3), S.speechlike(3, 65000), S.noise(2, 480000)],
        "full": [S.noise(5), S.speechlike(5), S.tone(1000.0), S.chirp(), np.zeros(480000, np.float32), S.named_case("ones"),
                 S.named_case("impulse"), (S.noise(7) * 1e-3).astype(np.float32)],
        "ragged": [S.noise(10 + i, int(n)) for i, n in enumerate(S.clip_lengths(1337, 12))],
    }
    for name, clips in sets.items():
        ref = O.logmel_batch(clips, n_mel, "fp64")
        mref = O.frame_attention_mask([len(c) for c in clips])
        for mode in ("tc", "cc"):
            os.environ["WFE_DISABLE_TC"] = "1" if mode == "cc" else "0"
            out, mask, err = run(fe, clips, n_mel)
            per = np.abs(out - ref).reshape(len(clips), -1).max(axis=1)
            print(f"{name:8s} {mode}: max-abs-err {per.max():.3e}  per-clip {np.array2string(per, precision=1)}  "
                  f"mask {'ok' if np.array_equal(mask, mref) else 'BAD'}  kernel-err {err:#x}  finite {np.isfinite(out).all()}")
    # timing: 64 clips of noise, device-resident
    dev = torch.device("cuda", 0)
    B = 64
    pcm = (0.1 * torch.randn(B * 480000, device=dev)).contiguous()
    offs = torch.arange(B + 1, dtype=torch.int64, device=dev) * 480000
    out = torch.empty((B, n_mel, 3000), dtype=torch.float32, device=dev)
    for mode in ("tc", "cc"):
        os.environ["WFE_DISABLE_TC"] = "1" if mode == "cc" else "0"
        for _ in range(3):
            fe.logmel_device(pcm, offs, B, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fe.logmel_device(pcm, offs, B, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gb = B * (480000 * 4 + n_mel * 3000 * 4) / 1e9
        print(f"timing {mode}: {ms:.4f} ms / {B} clips  {B * 30 / ms * 1e3 / 1e6:.2f} M audio-s/s  {gb / ms * 1e3:.0f} GB/s  err {fe.debug_kernel_error():#x}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Diagnostic for tests/test_parity_gpu.py::test_config3_full_size_batch_1024_properties (1024 packed ragged clips):
runs the tensor-core kernel several times on the test's input, compares every clip of every run with the CUDA-core
kernel (WFE_DISABLE_TC=1: an independent implementation of the same arithmetic) on the GPU, and the suspicious clips with
the CPU oracle.  python tools/diag_config3.py [reps] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import asr_finetune_b200 as pkg
from oracle import logmel as ologmel, signals

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
fe = pkg.WhisperFeatureExtractor(feature_size=128)
dev = fe.cuda_device()
lens = signals.clip_lengths(1337, B)
starts = np.zeros(B, dtype=np.int64)
np.cumsum((lens[:-1] + 3) & ~3, out=starts[1:])
g = torch.Generator(device=dev)
g.manual_seed(1337)
pcm = 0.1 * torch.randn(int(starts[-1] + lens[-1]), device=dev, generator=g)
d_starts, d_lens = torch.from_numpy(starts).to(dev), torch.from_numpy(lens).to(dev)
print(f"pcm ptr % 512 = {pcm.data_ptr() % 512}, total samples {pcm.numel()}, last clip len {lens[-1]} start {starts[-1]}")

os.environ["WFE_DISABLE_TC"] = "1"
ref_cc, mask_cc = fe.logmel_device(pcm, d_starts, B, return_attention_mask=True, lengths=d_lens)
torch.cuda.synchronize()
os.environ["WFE_DISABLE_TC"] = "0"
assert fe.uses_tensor_cores()


def describe(b, f, tag):
    clip = pcm[starts[b]:starts[b] + lens[b]].cpu().numpy()
    ref = ologmel.logmel_clip(clip, 128, "fp64")
    d = np.abs(f - ref)
    bad = np.argwhere(d > 1e-3)
    print(f"  {tag} clip {b} len {lens[b]} ({lens[b] / 160:.2f} hops, last real tile {int((lens[b] + 200) // 160) // 128}) vs oracle: "
          f"max err {d.max():.5f}, {len(bad)} elements > 1e-3")
    if len(bad):
        fr, ml = bad[:, 1], bad[:, 0]
        print(f"    frames {fr.min()}..{fr.max()} (tiles {sorted(set((fr // 128).tolist()))}), mels {ml.min()}..{ml.max()}; "
              f"ours[min,max] = ({f.min():.5f}, {f.max():.5f}) oracle = ({ref.min():.5f}, {ref.max():.5f})")
        for m, t in bad[:4]:
            print(f"    (mel {m}, frame {t}): ours {f[m, t]:.5f} oracle {ref[m, t]:.5f}")
        cnt = np.bincount(fr // 32, minlength=94)
        print("    bad elements per 32-frame block:", {int(i): int(c) for i, c in enumerate(cnt) if c})


first = None
for rep in range(reps):
    feats, mask = fe.logmel_device(pcm, d_starts, B, return_attention_mask=True, lengths=d_lens)
    torch.cuda.synchronize()
    err = fe.debug_kernel_error()
    diff = (feats - ref_cc).abs().amax(dim=(1, 2))
    badclips = torch.nonzero(diff > 5e-4).flatten().tolist()
    same = "first" if first is None else str(bool(torch.equal(feats, first)))
    print(f"rep {rep}: kernel error word {err:#x}; identical to rep 0: {same}; clips differing from the CUDA-core kernel by > 5e-4: "
          f"{badclips[:16]}{'...' if len(badclips) > 16 else ''} (max {float(diff.max()):.5f}); mask equal {bool(torch.equal(mask, mask_cc))}")
    if first is None:
        first = feats.clone()
    for b in badclips[:3]:
        describe(b, feats[b].cpu().numpy(), "TC")
        describe(b, ref_cc[b].cpu().numpy(), "CC")
if reps:
    describe(B - 1, first[B - 1].cpu().numpy(), "TC rep 0")
print("done")

#!/usr/bin/env python
"""Stress the tile hand-out and the barrier protocol of the tcgen05 kernel: many launches on three inputs (full-length
noise, speech-like dynamics, a ragged batch with many silent tiles), every output compared bit for bit with the first
launch's and the kernel's time-out word read back after every launch:  python tools/stress_determinism.py [launches]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import asr_finetune_b200 as pkg
n_launch = int(sys.argv[1]) if len(sys.argv) > 1 else 300
fe = pkg.WhisperFeatureExtractor(feature_size=128)
dev = fe.cuda_device()
g = torch.Generator(device=dev); g.manual_seed(11)
B = 256
noise = 0.1 * torch.randn(B * 480000, device=dev, generator=g)
seg = torch.rand(B * 480000 // 3200, device=dev, generator=g)
speech = noise * torch.repeat_interleave(10.0 ** (-3.0 * seg), 3200)
offs = torch.arange(B + 1, dtype=torch.int64, device=dev) * 480000
rng = np.random.default_rng(7)
lens = rng.integers(1, 6 * 16000, size=192); lens[::7] = rng.integers(16000, 480001, size=len(lens[::7]))
starts = np.zeros(192, dtype=np.int64); np.cumsum((lens[:-1] + 7) & ~7, out=starts[1:])
ragged = 0.1 * torch.randn(int(starts[-1] + lens[-1]) + 8, device=dev, generator=g)
cases = {"noise": (noise, offs, B, None), "speechlike": (speech, offs, B, None),
         "ragged": (ragged, torch.from_numpy(starts).to(dev), 192, torch.from_numpy(lens).to(dev))}
first, bad, errs = {}, 0, 0
t0 = time.time()
for it in range(n_launch):
    for name, (pcm, o, b, l) in cases.items():
        out, mask = fe.logmel_device(pcm, o, b, lengths=l, return_attention_mask=True)
        e = fe.debug_kernel_error()
        errs += int(e != 0)
        if name not in first:
            first[name] = (out.clone(), mask.clone())
            assert torch.isfinite(out).all()
        elif not (torch.equal(out, first[name][0]) and torch.equal(mask, first[name][1])):
            bad += 1
            print(f"launch {it} {name}: differs in {int((out != first[name][0]).sum())} elements, kernel err {e:#x}")
print(f"{n_launch} launches x {len(cases)} inputs in {time.time() - t0:.1f} s: {bad} outputs differ from the first launch, {errs} kernel time-outs")
sys.exit(1 if bad or errs else 0)

#!/usr/bin/env python
"""Time wfe_logmel alone on device-resident noise (no checks): python tools/time_kernel.py [batch] [n_mel] [iters]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import asr_finetune_b200 as pkg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n_mel = int(sys.argv[2]) if len(sys.argv) > 2 else 128
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
kind = sys.argv[4] if len(sys.argv) > 4 else "noise"
fe = pkg.WhisperFeatureExtractor(feature_size=n_mel)
dev = fe.cuda_device()
g = torch.Generator(device=dev); g.manual_seed(0)
pcm = 0.1 * torch.randn(B * 480000, device=dev, generator=g)
if kind == "bursty":  # speech-like dynamics: 0.2 s segments with random gains over 60 dB
    seg = torch.rand(B * 480000 // 3200, device=dev, generator=g)
    pcm = pcm * torch.repeat_interleave(10.0 ** (-3.0 * seg), 3200)
if kind == "int16":  # device-resident int16 PCM
    pcm = (pcm.clamp(-1, 1) * 32767).to(torch.int16)
if kind == "fp16":
    pcm = pcm.to(torch.float16)
offs = torch.arange(B + 1, dtype=torch.int64, device=dev) * 480000
lengths = None
audio_s = B * 30.0
if kind == "ragged":  # BASELINE configs[2]: lengths U{1..30 s}; clips keep their 30-s slots, only `lengths` samples are read
    lengths = torch.randint(16000, 480001, (B,), device=dev, generator=g, dtype=torch.int64)
    offs = offs[:B].contiguous()
    audio_s = float(lengths.sum()) / 16000.0
out = torch.empty((B, n_mel, 3000), dtype=torch.float32, device=dev)
for _ in range(3):
    fe.logmel_device(pcm, offs, B, out=out, lengths=lengths)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    fe.logmel_device(pcm, offs, B, out=out, lengths=lengths)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
err = fe.debug_kernel_error()
print(f"{os.environ.get('WFE_LIB_OVERRIDE','libwfe.so'):40s} B={B} n_mel={n_mel} {kind}: {ms:.4f} ms/launch  {audio_s/ms*1e3/1e6:.2f} M audio-s/s  "
      f"{(audio_s*16000*4+B*n_mel*3000*4)/ms/1e6:.0f} GB/s  ({B/ms*1e3/1e6:.3f} M clips/s)" + (f"  KERNEL ERROR {err:#x}" if err else ""))

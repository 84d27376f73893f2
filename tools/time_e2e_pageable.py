#!/usr/bin/env python
"""e2e of the drop-in call on separately allocated PAGEABLE clips (what the reference's loader yields), float32 -> float32
and int16 -> float16, for the current WFE_HOST_THREADS / WFE_HOST_CHUNK: python tools/time_e2e_pageable.py [batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import asr_finetune_b200 as pkg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
fe = pkg.WhisperFeatureExtractor(feature_size=128)
rng = np.random.default_rng(0)
base = (0.1 * rng.standard_normal(480000 * 8)).astype(np.float32)
clips = [np.array(np.roll(base[:480000 + 0], 977 * i)[:480000], copy=True) for i in range(B)]
clips16 = [np.array((c * 32767).astype(np.int16), copy=True) for c in clips]
def t(f, n=12):
    for _ in range(3): f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort(); return ts[len(ts) // 2], ts[0]
tag = f"threads={os.environ.get('WFE_HOST_THREADS','default')} chunk={os.environ.get('WFE_HOST_CHUNK','default')}"
m, mn = t(lambda: fe(clips, sampling_rate=16000, return_tensors="pt"))
print(f"{tag}: fp32 pageable -> fp32 host: median {m:.2f} ms (min {mn:.2f})  {B*30/m*1e3/1e6:.3f} M audio-s/s")
m, mn = t(lambda: fe(clips16, sampling_rate=16000, return_tensors="pt", output_dtype=torch.float16))
print(f"{tag}: int16 pageable -> fp16 host: median {m:.2f} ms (min {mn:.2f})  {B*30/m*1e3/1e6:.3f} M audio-s/s")
# the C entry alone (prebuilt pointer arrays, preallocated pinned output): what the Python shim adds is the difference
import ctypes as C
h = fe._handle(None, fe.cuda_device()); lib = pkg._lib.load()
ptrs = (C.c_void_p * B)(*[c.ctypes.data for c in clips16]); lens = (C.c_int64 * B)(*[len(c) for c in clips16])
out = torch.empty((B, 128, 3000), dtype=torch.float16, pin_memory=True)
up, down = C.c_uint64(0), C.c_uint64(0)
m, mn = t(lambda: lib.wfe_extract_host_ex(h.ptr, ptrs, lens, B, 1, 1.0 / 32768.0, 0, out.data_ptr(), 1, None, C.byref(up), C.byref(down)))
print(f"{tag}: int16 -> fp16, wfe_extract_host_ex alone: median {m:.2f} ms (min {mn:.2f})  {B*30/m*1e3/1e6:.3f} M audio-s/s")

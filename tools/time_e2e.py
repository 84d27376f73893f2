#!/usr/bin/env python
"""Where the host-buffer (e2e) path spends its time: python tools/time_e2e.py [batch]"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import asr_finetune_b200 as pkg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
fe = pkg.WhisperFeatureExtractor(feature_size=128)
host = torch.empty((B, 480000), dtype=torch.float32, pin_memory=True)
host.normal_(0, 0.1)
clips = [host.numpy()[i] for i in range(B)]
def t(f, n=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
ms = t(lambda: fe(clips, sampling_rate=16000, return_tensors="pt"))
print(f"fe(list of {B} pinned clips) -> pinned host: {ms:.2f} ms  {B*30/ms*1e3/1e6:.3f} M audio-s/s  ({(B*480000*4 + B*128*3000*4)/ms/1e6:.1f} GB/s both directions)")
# raw C call without the Python argument handling
h = fe._handle(None, fe.cuda_device()); lib = pkg._lib.load()
ptrs = (C.c_void_p * B)(*[c.ctypes.data for c in clips]); lens = (C.c_int64 * B)(*[480000] * B)
out = torch.empty((B, 128, 3000), dtype=torch.float32, pin_memory=True)
up, down = C.c_uint64(0), C.c_uint64(0)
ms2 = t(lambda: lib.wfe_extract_host(h.ptr, ptrs, lens, B, 0, 1.0, 0, out.data_ptr(), None, C.byref(up), C.byref(down)))
print(f"wfe_extract_host alone: {ms2:.2f} ms")
# pure copies for reference
d_in = torch.empty((B, 480000), dtype=torch.float32, device="cuda"); d_out = torch.empty((B, 128, 3000), dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def copies():
    with torch.cuda.stream(s1): d_in.copy_(host, non_blocking=True)
    with torch.cuda.stream(s2): out.copy_(d_out, non_blocking=True)
    s1.synchronize(); s2.synchronize()
ms3 = t(copies)
print(f"H2D {B*1.92:.0f} MB || D2H {B*1.536:.0f} MB concurrently: {ms3:.2f} ms;  H2D alone: {t(lambda: (d_in.copy_(host, non_blocking=True), torch.cuda.synchronize())):.2f} ms;  D2H alone: {t(lambda: (out.copy_(d_out, non_blocking=True), torch.cuda.synchronize())):.2f} ms")

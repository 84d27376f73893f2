#!/usr/bin/env python
"""Where a batch-8 collate spends its time (the in-loop training consumer): python tools/time_small_batch.py [B]"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import asr_finetune_b200 as pkg
from oracle import signals
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
fe = pkg.WhisperFeatureExtractor(feature_size=128)
dev = fe.cuda_device()
rng = np.random.default_rng(8)
clips = [np.array(0.1 * rng.standard_normal(int(n)), dtype=np.float32) for n in rng.integers(3 * 16000, 480001, size=B)]
labels = signals.label_ids(1337, B, 5, 60)
coll = pkg.StreamingFrontendCollator(fe, device=dev, feature_dtype=torch.float16)
coll32 = pkg.StreamingFrontendCollator(fe, device=dev)
def t(f, n=30):
    for _ in range(3): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort(); return ts[len(ts) // 2], ts[0]
mb = sum(c.nbytes for c in clips) / 1e6
print(f"B={B}, {mb:.1f} MB of float32 PCM, chunk env {os.environ.get('WFE_HOST_CHUNK')}")
print("collator (fp16 features on device): median %.3f ms, min %.3f" % t(lambda: coll({"audio": clips, "labels": labels})))
print("collator (fp32 features on device): median %.3f ms, min %.3f" % t(lambda: coll32({"audio": clips, "labels": labels})))
print("fe(list, output_device=cuda):       median %.3f ms, min %.3f" % t(lambda: fe(clips, sampling_rate=16000, output_device="cuda")))
h = fe._handle(None, dev); lib = pkg._lib.load()
ptrs = (C.c_void_p * B)(*[c.ctypes.data for c in clips]); lens = (C.c_int64 * B)(*[len(c) for c in clips])
out = torch.empty((B, 128, 3000), dtype=torch.float32, device=dev)
up, down = C.c_uint64(0), C.c_uint64(0)
print("wfe_extract_host_ex alone:          median %.3f ms, min %.3f" % t(lambda: lib.wfe_extract_host_ex(h.ptr, ptrs, lens, B, 0, 1.0, 0, out.data_ptr(), 0, None, C.byref(up), C.byref(down))))
pin = torch.empty(sum(len(c) for c in clips), dtype=torch.float32, pin_memory=True)
d_in = torch.empty_like(pin, device=dev)
def stage():
    o = 0
    for c in clips:
        pin.numpy()[o:o + len(c)] = c; o += len(c)
print("single-thread staging memcpy:       median %.3f ms, min %.3f" % t(stage))
print("H2D of the staged PCM:              median %.3f ms, min %.3f" % t(lambda: d_in.copy_(pin, non_blocking=True)))
offs = torch.tensor(np.concatenate([[0], np.cumsum([len(c) for c in clips])]), dtype=torch.int64, device=dev)
print("logmel_device on resident PCM:      median %.3f ms, min %.3f" % t(lambda: fe.logmel_device(d_in, offs, B, out=out)))

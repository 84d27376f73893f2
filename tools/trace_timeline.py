#!/usr/bin/env python
"""Per-warp stage timeline from the tracing what-if build (WFE_EXP=256): WFE_LIB_OVERRIDE=exp_so/libwfe_trace.so python tools/trace_timeline.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import asr_finetune_b200 as pkg
B = 256
fe = pkg.WhisperFeatureExtractor(feature_size=128); dev = fe.cuda_device()
g = torch.Generator(device=dev); g.manual_seed(0)
pcm = 0.1 * torch.randn(B * 480000, device=dev, generator=g)
offs = torch.arange(B + 1, dtype=torch.int64, device=dev) * 480000
out = torch.empty((B, 128, 3000), dtype=torch.float32, device=dev)
for _ in range(3): fe.logmel_device(pcm, offs, B, out=out)
torch.cuda.synchronize()
lib = pkg._lib.load()
n = 4 * 8 * 8 * 12 + 4 * 8 * 8
buf = (C.c_ulonglong * n)()
lib.wfe_debug_read_trace.argtypes = [C.c_void_p, C.c_int]
assert lib.wfe_debug_read_trace(buf, n) == 0
raw = np.frombuffer(buf, dtype=np.uint64).astype(np.int64)
t = raw[:4 * 8 * 8 * 12].reshape(4, 8, 8, 12)  # cta, iter, warp, point
ts = raw[4 * 8 * 8 * 12:].reshape(4, 8, 8)  # cta, iter, scheduler step
names = ["top", "pre-S1", "post-S1", "s1 start", "s1 end", "post-S2", "s2c end(+sched)", "pre-S3", "post-S3", "pre-S4", "post-S4"]
for cta in (0, 1):
    for it in (2, 3):
        base = t[cta, it, :, 0].min()
        print(f"--- CTA {cta} iteration {8+it}: cycles relative to the first warp's top; iteration length {t[cta,it,:,10].max()-base}")
        for w in range(8):
            print(f"  warp {w}: " + "  ".join(f"{names[p]}={t[cta,it,w,p]-base:6d}" for p in range(11)))
d = t[:, 1:7]  # skip first/last traced iterations
seg = [("top->pre-S1 (prefetch issue, wait)", 0, 1), ("S1 wait", 1, 2), ("fixups", 2, 3), ("stage 1", 3, 4), ("S2 wait", 4, 5),
       ("stage 2 compute / sched", 5, 6), ("S2b wait + stores", 6, 7), ("S3 wait", 7, 8), ("stage 3", 8, 9), ("S4 wait", 9, 10)]
print("mean cycles per warp per tile:")
tot = 0
for name, a, b in seg:
    v = (d[..., b] - d[..., a]).mean(); tot += v
    print(f"  {name:36s} {v:8.0f}   per-warp means: " + " ".join(f"{(d[:, :, w, b]-d[:, :, w, a]).mean():6.0f}" for w in range(8)))
print(f"  total {tot:.0f}")

sn = ["fence", "ticket red + desc write", "request clip", "fix decisions", "publish max + ring push", "request states"]
print("scheduler lane, mean cycles per step:")
for i, name in enumerate(sn):
    print(f"  {name:28s} {(ts[:, 1:7, i + 1] - ts[:, 1:7, i]).mean():8.0f}")

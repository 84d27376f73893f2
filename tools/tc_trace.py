"""Print the role timeline of the tcgen05 kernel from a -DWFE_TC_TRACE build (diagnostic tool).
   WFE_LIB_OVERRIDE=exp_so/libwfe_trace.so python tools/tc_trace.py [kind]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import asr_finetune_b200 as pkg

kind = sys.argv[1] if len(sys.argv) > 1 else "noise"
B = 256
fe = pkg.WhisperFeatureExtractor(feature_size=128)
dev = fe.cuda_device()
pcm = 0.1 * torch.randn(B * 480000, device=dev)
if kind == "bursty":
    seg = torch.rand(B * 480000 // 3200, device=dev)
    pcm = pcm * torch.repeat_interleave(10.0 ** (-3.0 * seg), 3200)
offs = torch.arange(B + 1, dtype=torch.int64, device=dev) * 480000
out = torch.empty((B, 128, 3000), dtype=torch.float32, device=dev)
for _ in range(2):
    fe.logmel_device(pcm, offs, B, out=out)
torch.cuda.synchronize()
lib = pkg._lib.load()
T, R, P = 8, 5, 32
buf = np.zeros(T * R * P, dtype=np.uint64)
rc = lib.wfe_debug_read_tc_trace(buf.ctypes.data_as(C.c_void_p), buf.size)
assert rc == 0, rc
tr = buf.reshape(T, R, P).astype(np.int64)
for it_, name in ((0, "CTA 0"), (1, "CTA 77")):
    c0, g0, c1, g1 = [int(x) for x in tr[it_, 4, :4]]
    print(f"{name}: kernel span {c1 - c0} SM cycles, {g1 - g0} ns -> {(c1 - c0) / max(g1 - g0, 1):.3f} GHz")
tr[:, 4, :] = 0
t0 = tr[tr > 0].min()
names = {0: "prep  [start wait_raw, raw ok, staged, scaled, k0..k6 done]", 1: "epi   [start wait_d, d ok, released, done]",
         2: "mma   [start wait_dempty, ok, k0..k6 issued]", 3: "load  [start wait_rawempty, ok, issued]"}
for it in range(T):
    print(f"--- tile iteration {it + 4} (CTA 0; tile ids 0+148*it: tile-in-clip {(148 * (it + 4)) % 24}) ---")
    for r in range(4):
        v = tr[it, r]
        pts = [int(x - t0) for x in v[:16] if x > 0]
        print(f"  {names[r]:60s} {pts}")
        fine = [int(x - t0) for x in v[16:] if x > 0]
        if fine:
            print(f"      k-step 3 detail (prep: compute done, then per slot [empty ok, st done, arrived]; mma: per slot [start wait, full ok, committed]): {fine}")

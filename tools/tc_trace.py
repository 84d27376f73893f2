"""Print the role timeline of the tcgen05 kernel from a -DWFE_TC_TRACE build (diagnostic tool).
   WFE_LIB_OVERRIDE=exp_so/libwfe_trace.so python tools/tc_trace.py [kind] [first_it] [n_it]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import asr_finetune_b200 as pkg

kind = sys.argv[1] if len(sys.argv) > 1 else "noise"
show0 = int(sys.argv[2]) if len(sys.argv) > 2 else 4
shown = int(sys.argv[3]) if len(sys.argv) > 3 else 2
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
T, R, P = 8, 8, 64
buf = np.zeros(T * R * P, dtype=np.uint64)
rc = lib.wfe_debug_read_tc_trace(buf.ctypes.data_as(C.c_void_p), buf.size)
assert rc == 0, rc
tr = buf.reshape(T, R, P).astype(np.int64)
tiles = np.zeros(3 * 64, dtype=np.uint64)
assert lib.wfe_debug_read_tc_tiles(tiles.ctypes.data_as(C.c_void_p)) == 0
tiles = tiles.reshape(3, 64).astype(np.int64)
for c, name in enumerate(("CTA 0", "CTA 77", "CTA 147")):
    st = tiles[c, 63]
    v = [int(x - st) for x in tiles[c, :58] if x > 0]
    print(f"{name}: tile starts (cycles after kernel start): {v}")
    print(f"   deltas: {[b - a for a, b in zip(v, v[1:])]}")
    print(f"   workers done {int(tiles[c, 62] - st)}, mma warp done {int(tiles[c, 61] - st)}, loader done {int(tiles[c, 60] - st)}, clamp warp done {int(tiles[c, 59] - st)}")
for it_, name in ((0, "CTA 0"), (1, "CTA 77")):
    c0, g0, c1, g1 = [int(x) for x in tr[it_, 4, :4]]
    print(f"{name}: kernel span {c1 - c0} SM cycles, {g1 - g0} ns -> {(c1 - c0) / max(g1 - g0, 1):.3f} GHz")
cta = np.zeros(160 * 4, dtype=np.uint64)
assert lib.wfe_debug_read_tc_cta(cta.ctypes.data_as(C.c_void_p)) == 0
cta = cta.reshape(160, 4).astype(np.int64)
print("per CTA: blockIdx smid tiles cycles/tile (tile 1 .. last tile start)")
rows = []
for c in range(148):
    n = int(cta[c, 1]) - 1
    if n > 0:
        rows.append((c, int(cta[c, 0]), n + 1, (int(cta[c, 3]) - int(cta[c, 2])) / n))
for r in rows:
    print(f"  cta {r[0]:3d} smid {r[1]:3d} tiles {r[2]:2d} {r[3]:8.0f}")
ww = np.zeros(128, dtype=np.uint64)
assert lib.wfe_debug_read_tc_warps(ww.ctypes.data_as(C.c_void_p)) == 0
ww = ww.reshape(16, 8).astype(np.int64)
w0 = ww[:12][ww[:12] > 0].min()
print("per worker warp, traced CTA, tile iteration 6.  E half 0: [tile start, k2 ready, k2 deposited, scanned, k5 ready, k5 deposited, d_full ok, epilogue end]; E half 1: [start, scanned, k3 ready, k3 deposited, k6 ready, k6 deposited, d_full ok, epilogue end]; P: [start, k4 ready, k4 deposited, scan barrier passed, ws out, k0' ready, k0' deposited, k1' deposited]")
k0 = int(ww[12, 0])
print(f"start-up of the traced CTA (cycles after kernel entry): set-up done {int(ww[12,1])-k0}; "
      f"loader [extent done, (tile 0: got tile, TMA landed, meta out), (tile 1: ...)] {[int(x)-k0 for x in ww[13,:7]]}; "
      f"epilogue warp 0 [got tile 0, scanned, window table out, -, got tile 1, first accumulators ready] {[int(x)-k0 if x>0 else 0 for x in ww[14,:6]]}; "
      f"prep warp 8 [got tile 0, may start] {[int(x)-k0 for x in ww[15,:2]]}")
for w in range(12):
    print(f"   warp {w} (quarter {w % 4}, {'E half ' + str(w // 4) if w < 8 else 'P'}): {[int(x - w0) for x in ww[w]]}")
kstart = int(tr[0, 4, 0])
tr[:, 4, :] = 0
t0 = tr[tr > 0].min()
print(f"first traced stamp is {t0 - kstart} cycles after kernel start (tile iteration 4)")
names = {0: "prep  [tile start, k0..k6 deposited]", 1: "epi   [start wait_d, d ok, released, done]",
         2: "mma   [start wait_dempty, ok, k0..k6 issued]", 3: "load  [start wait_rawempty, ok, tma landed, meta out]"}
for it in range(T):
    print(f"--- tile iteration {it + 4} ---")
    for r in range(4):
        pts = [int(x - t0) for x in tr[it, r, :16] if x > 0]
        print(f"  {names[r]:60s} {pts}")
    if show0 <= it + 4 < show0 + shown:
        for it2 in range(2):
            v = tr[it, 5 + it2]
            for half, base in (("A", 0), ("B", 21)):
                seg = [int(x - t0) for x in v[base:base + 21]]
                d = [b - a for a, b in zip(seg, seg[1:])]
                print(f"    loop trip {it2} step {half}: entry {seg[0]}; per slot (+loads&wait, +st issue, +math, +wait::st, +arrive): "
                      f"{[d[5 * s:5 * s + 5] for s in range(4)]}")
            print(f"      after scan slice: {int(v[42] - t0)}, end of trip {int(v[43] - t0)}")
        v = tr[it, 7]
        for j in range(7):
            seg = [int(x - t0) for x in v[j * 8:j * 8 + 8]]
            print(f"    mma k-step {j}: per slot (full ok, issued): {seg}")

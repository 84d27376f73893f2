#!/usr/bin/env python
"""Fuzz of the host entry (`WhisperFeatureExtractor.__call__` on numpy clips -> `wfe_extract_host`: staging threads, three
streams, chunks of 16) against the device-resident entry (`logmel_device`) on the same samples: the two must agree bit for
bit (same kernel, same tiles; only the packing, the chunking and the copies differ).  Random batch sizes around the chunk
boundaries, ragged lengths, pageable per-clip arrays / views of one pinned buffer / a mix, float32 / int16 / float16 PCM,
host or CUDA outputs, float32 / float16 / bfloat16 features, `do_normalize`; every call is made twice.
python tools/fuzz_host.py [seconds] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import asr_finetune_b200 as pkg

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 40.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
fes = {80: pkg.WhisperFeatureExtractor(feature_size=80), 128: pkg.WhisperFeatureExtractor(feature_size=128)}
dev = fes[128].cuda_device()
TILE = 128 * 160


def length():
    k = int(rng.integers(0, 6))
    t = int(rng.integers(0, 24))
    if k == 0:
        n = t * TILE - 200 + 130 * 160 + int(rng.integers(-6, 6))
    elif k == 1:
        n = t * TILE - 200 + int(rng.integers(-4, 5))
    elif k == 2:
        n = 480000 + int(rng.integers(-300, 3000))
    elif k == 3:
        n = int(rng.integers(1, 500))
    else:
        n = int(rng.integers(1, 480001))
    return int(min(max(n, 1), 490000))


n_calls = n_fail = 0
t_end = time.time() + budget
while time.time() < t_end:
    n_mel = 128 if rng.random() < 0.6 else 80
    fe = fes[n_mel]
    B = int(rng.choice([1, 2, 5, 8, 15, 16, 17, 31, 33, 64, 120]))
    lens = [length() for _ in range(B)]
    kind = str(rng.choice(["f32", "f32", "i16", "f16"]))
    npdt = {"f32": np.float32, "i16": np.int16, "f16": np.float16}[kind]
    layout = str(rng.choice(["pageable", "pinned_views", "mixed"]))
    normalize = rng.random() < 0.15
    out_where = str(rng.choice(["host", "host", "cuda"]))
    out_dtype = [None, None, torch.float16, torch.bfloat16][int(rng.integers(0, 4))]
    pinned = torch.empty(sum(lens) + 64 * B, dtype={"f32": torch.float32, "i16": torch.int16, "f16": torch.float16}[kind],
                         pin_memory=True).numpy() if layout != "pageable" else None
    clips, pos = [], 0
    for i, n in enumerate(lens):
        x = rng.standard_normal(n).astype(np.float32) * 0.1 * float(10.0 ** (-2.0 * rng.random()))
        if kind == "i16":
            x = np.clip(x * 32767.0, -32768, 32767).astype(np.int16)
        elif kind == "f16":
            x = x.astype(np.float16)
        if layout == "pinned_views" or (layout == "mixed" and i % 2 == 0):
            pos += int(rng.integers(0, 5))
            pinned[pos:pos + n] = x
            x = pinned[pos:pos + n]
            pos += n
        clips.append(x)
    kw = dict(sampling_rate=16000, return_attention_mask=True, return_tensors="pt", do_normalize=normalize)
    if out_where == "cuda":
        kw["output_device"] = "cuda"
    if out_dtype is not None:
        kw["output_dtype"] = out_dtype
    r1 = fe(clips, **kw)
    r2 = fe(clips, **kw)
    f1, m1 = r1["input_features"].to(dev), r1["attention_mask"].to(dev)
    f2 = r2["input_features"].to(dev)
    # the device-resident entry on the same samples (clip starts on 16-byte boundaries)
    starts = np.zeros(B, dtype=np.int64)
    step = 4 if kind == "f32" else 8
    np.cumsum([(n + step - 1) // step * step for n in lens[:-1]], out=starts[1:])
    host = np.zeros(int(starts[-1] + lens[-1]), dtype=npdt)
    for c, s in zip(clips, starts):
        host[s:s + len(c)] = c
    pcm = torch.from_numpy(host).to(dev)
    ref, mref = fe.logmel_device(pcm, torch.from_numpy(starts).to(dev), B, return_attention_mask=True,
                                 lengths=torch.tensor(lens, dtype=torch.int64, device=dev), do_normalize=normalize,
                                 out_dtype=out_dtype)
    torch.cuda.synchronize()
    n_calls += 1
    same = torch.equal(f1, ref)
    ok = same and torch.equal(f1, f2) and torch.equal(m1.to(torch.int32), mref) and f1.dtype == ref.dtype
    if not ok:
        n_fail += 1
        d = (f1.float() - ref.float()).abs()
        d = torch.where(torch.isfinite(d), d, torch.full_like(d, 9.0)).amax(dim=(1, 2))
        bad = torch.nonzero(d > 0).flatten().tolist()
        print(f"FAIL call {n_calls}: n_mel {n_mel} B {B} pcm {kind} layout {layout} normalize {normalize} out {out_where} {out_dtype}: "
              f"equals the device entry {same}, second call identical {bool(torch.equal(f1, f2))}, mask equal "
              f"{bool(torch.equal(m1.to(torch.int32), mref))}, dtype {f1.dtype} / {ref.dtype}; differing clips "
              f"{[(b, lens[b], round(float(d[b]), 6)) for b in bad[:8]]} (clip, len, max diff)")
print(f"fuzz_host: {n_calls} calls, {n_fail} failures (seed {seed}, {budget:.0f} s)")

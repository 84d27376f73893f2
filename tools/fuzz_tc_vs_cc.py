#!/usr/bin/env python
"""Differential fuzz of the tensor-core kernel against the CUDA-core kernel (WFE_DISABLE_TC=1), all on the GPU.

Random batches through `logmel_device`: batch size, clip lengths biased towards tile / hop-row / box boundaries, clip
starts aligned or not, GARBAGE (NaN, Inf, 1e30, random bits) in everything of the PCM buffer that is not a clip, PCM
dtype (float32 / int16 / float16), 80 or 128 mel, `do_normalize`, speech-like gains.  Checks per batch: every clip
within 1e-3 of the CUDA-core kernel (both are within 2e-4 of the oracle on the golden set), masks equal, a second run
bit-identical, the kernel's error word zero.  python tools/fuzz_tc_vs_cc.py [seconds] [seed] [P(do_normalize)]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import asr_finetune_b200 as pkg

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
p_norm = float(sys.argv[3]) if len(sys.argv) > 3 else 0.15
rng = np.random.default_rng(seed)
fes = {80: pkg.WhisperFeatureExtractor(feature_size=80), 128: pkg.WhisperFeatureExtractor(feature_size=128)}
dev = fes[128].cuda_device()
TILE = 128 * 160


def special_length():
    t = int(rng.integers(0, 24))
    base = t * TILE - 200
    kind = int(rng.integers(0, 8))
    if kind == 0:
        n = base + int(rng.integers(-4, 5))                       # around a tile's first sample
    elif kind == 1:
        n = base + 130 * 160 + int(rng.integers(-6, 6))           # around the end of the TMA box (rows 129 / "130")
    elif kind == 2:
        n = base + 129 * 160 + 164 + int(rng.integers(-3, 3))     # box_end
    elif kind == 3:
        n = base + 160 * int(rng.integers(0, 131)) + int(rng.integers(-2, 3))  # hop-row boundaries
    elif kind == 4:
        n = 480000 - int(rng.integers(0, 420))                    # reflect pad at sample 480000
    elif kind == 5:
        n = int(rng.integers(1, 420))                             # shorter than the reflect pad / one frame
    elif kind == 6:
        n = 480000 + int(rng.integers(0, 5000))                   # longer than 30 s (truncated)
    else:
        n = base + int(rng.integers(0, TILE))
    return int(min(max(n, 1), 500000))


def garbage(n):
    kind = int(rng.integers(0, 4))
    if kind == 0:
        return torch.full((n,), float("nan"), device=dev)
    if kind == 1:
        return torch.full((n,), float("inf"), device=dev)
    if kind == 2:
        return torch.full((n,), 1e30, device=dev)
    return torch.randint(-2**31, 2**31 - 1, (n,), device=dev, dtype=torch.int32).view(torch.float32)


focus = len(sys.argv) > 4 and sys.argv[4] == "focus"  # replay a failing batch many times and dissect the bad runs
n_focus = 0


def scratch_words(fe, B):
    h, scratch, batch = fe._last_scratch
    w = scratch.view(torch.int32)
    key = w[:B * 24].view(B, 24).clone()
    mn = w[B * 24:B * 24 + B * 24 * 16].view(B, 24, 16).clone()
    return key, mn


def replay(fe, pcm, d_starts, B, kw, ref, lens, n_mel, reps=300):
    good = None
    n_bad = 0
    for rep in range(reps):
        out, _ = fe.logmel_device(pcm, d_starts, B, **kw)
        torch.cuda.synchronize()
        key, mn = scratch_words(fe, B)
        d = (out - ref).abs()
        d = torch.where(torch.isfinite(d), d, torch.full_like(d, 9.0))
        per_clip = d.amax(dim=(1, 2))
        bad = torch.nonzero(per_clip > 1e-3).flatten().tolist()
        if not bad:
            if good is None:
                good = (out.clone(), key, mn)
            continue
        n_bad += 1
        if n_bad > 3:
            continue
        print(f"  replay {rep}: bad clips {[(b, int(lens[b]), round(float(per_clip[b]), 5)) for b in bad[:8]]}")
        for b in bad[:2]:
            e = d[b] > 1e-3
            mels = torch.nonzero(e.any(dim=1)).flatten()
            frs = torch.nonzero(e.any(dim=0)).flatten()
            o, r = out[b], ref[b]
            at_floor = float((o[e] == o.min()).float().mean())
            ref_at_floor = float((r[e] == r.min()).float().mean())
            print(f"    clip {b}: {int(e.sum())} bad elements, mels {int(mels.min())}..{int(mels.max())} ({mels.numel()}), frames "
                  f"{int(frs.min())}..{int(frs.max())} ({frs.numel()}); ours == our clip minimum on {at_floor:.2f} of them, CC == its minimum on "
                  f"{ref_at_floor:.2f}; clip min/max ours ({float(o.min()):.5f}, {float(o.max()):.5f}) CC ({float(r.min()):.5f}, {float(r.max()):.5f})")
            idx = torch.nonzero(e)[:6].tolist()
            print("    samples (mel, frame, ours, CC):", [(m_, f_, round(float(o[m_, f_]), 5), round(float(r[m_, f_]), 5)) for m_, f_ in idx])
            per_blk = e.view(n_mel // 32 if n_mel % 32 == 0 else -1, 32, 3000).any(dim=1) if n_mel % 32 == 0 else None
            if per_blk is not None:
                fb = torch.nonzero(per_blk.any(dim=0)).flatten() // 32
                print("    bad (32-frame block: mel groups):", {int(k): sorted(set(torch.nonzero(per_blk[:, 32 * int(k):32 * int(k) + 32].any(dim=1)).flatten().tolist())) for k in sorted(set(fb.tolist()))})
            if good is not None:
                dk = torch.nonzero(key[b] != good[1][b]).flatten().tolist()
                dm = torch.nonzero((mn[b] != good[2][b]).any(dim=1)).flatten().tolist()
                print(f"    scratch vs a good run: tile keys differ at tiles {dk}, block minima differ at tiles {dm}")
                for t in dm[:3]:
                    print(f"      tile {t} minima bad  {[hex(x & 0xffffffff) for x in mn[b, t].tolist()]}")
                    print(f"      tile {t} minima good {[hex(x & 0xffffffff) for x in good[2][b, t].tolist()]}")
                for t in dk[:3]:
                    print(f"      tile {t} key bad {int(key[b, t]) & 0xffffffff:#x} good {int(good[1][b, t]) & 0xffffffff:#x}")
    print(f"  replay: {n_bad} bad runs of {reps}")


n_batches = n_fail = 0
t_end = time.time() + budget
while time.time() < t_end:
    n_mel = 128 if rng.random() < 0.7 else 80
    fe = fes[n_mel]
    B = int(rng.choice([1, 2, 3, 7, 24, 60, 150, 300]))
    lens = np.array([special_length() if rng.random() < 0.6 else int(rng.integers(1, 480001)) for _ in range(B)], dtype=np.int64)
    if rng.random() < 0.05:
        lens[:] = 480000  # the bench workload: every clip full length (no tails, no silent tiles)
    dtype = rng.choice(["f32", "f32", "f32", "i16", "f16"])
    align = int(rng.choice([4, 4, 8, 1, 2]))  # samples; 4 fp32 samples = 16 bytes (TMA path)
    extra = int(rng.integers(0, 3)) * align + (int(rng.integers(0, 2)) if align == 1 else 0)
    normalize = rng.random() < p_norm
    starts = np.zeros(B, dtype=np.int64)
    first = int(rng.integers(0, 3)) * align
    gaps = [(int(n) + align - 1) // align * align + extra for n in lens[:-1]]
    starts[0] = first
    if B > 1:
        starts[1:] = first + np.cumsum(gaps)
    total = int(starts[-1] + lens[-1]) + int(rng.integers(0, 2)) * int(rng.integers(0, 30000))
    buf = garbage(total)
    g = torch.Generator(device=dev)
    g.manual_seed(int(rng.integers(0, 2**31)))
    speech = rng.random() < 0.5
    for s, n in zip(starts, lens):
        x = 0.1 * torch.randn(int(n), device=dev, generator=g)
        if speech:
            seg = torch.rand((int(n) + 3199) // 3200, device=dev, generator=g)
            x = x * torch.repeat_interleave(10.0 ** (-3.0 * seg), 3200)[:int(n)]
        buf[int(s):int(s) + int(n)] = x
    scale = 1.0
    if dtype == "i16":
        pcm = (buf.nan_to_num(0.0, 30000.0, -30000.0).clamp(-1, 1) * 32767).to(torch.int16)
        # garbage for int16: anything representable
        mask_clip = torch.zeros(total, dtype=torch.bool, device=dev)
        for s, n in zip(starts, lens):
            mask_clip[int(s):int(s) + int(n)] = True
        pcm = torch.where(mask_clip, pcm, torch.randint(-32768, 32767, (total,), device=dev, dtype=torch.int16))
        scale = 1.0 / 32768.0
    elif dtype == "f16":
        pcm = buf.to(torch.float16)
    else:
        pcm = buf
    d_starts, d_lens = torch.from_numpy(starts).to(dev), torch.from_numpy(lens).to(dev)
    kw = dict(return_attention_mask=True, lengths=d_lens, pcm_scale=scale, do_normalize=normalize)
    os.environ["WFE_DISABLE_TC"] = "1"
    ref, mref = fe.logmel_device(pcm, d_starts, B, **kw)
    torch.cuda.synchronize()
    os.environ["WFE_DISABLE_TC"] = "0"
    out, m = fe.logmel_device(pcm, d_starts, B, **kw)
    torch.cuda.synchronize()
    err = fe.debug_kernel_error()
    out2, _ = fe.logmel_device(pcm, d_starts, B, **kw)
    torch.cuda.synchronize()
    diff = (out - ref).abs()
    diff = torch.where(torch.isfinite(diff), diff, torch.full_like(diff, 9.0)).amax(dim=(1, 2))
    bad = torch.nonzero(diff > 1e-3).flatten().tolist()
    ok = not bad and err == 0 and torch.equal(m, mref) and torch.equal(out, out2) and bool(torch.isfinite(out).all())
    n_batches += 1
    if not ok:
        n_fail += 1
        print(f"FAIL batch {n_batches}: n_mel {n_mel} B {B} dtype {dtype} align {align} extra {extra} first {first} normalize {normalize} "
              f"speech {speech} total {total}: err word {err:#x}, mask equal {bool(torch.equal(m, mref))}, rerun identical "
              f"{bool(torch.equal(out, out2))}, finite {bool(torch.isfinite(out).all())}, ref finite {bool(torch.isfinite(ref).all())}, bad clips "
              f"{[(b, int(lens[b]), int(starts[b]) % 4, round(float(diff[b]), 5)) for b in bad[:8]]} (clip, len, start % 4, max diff)")
        for b in bad[:2]:
            d = (out[b] - ref[b]).abs()
            d = torch.where(torch.isfinite(d), d, torch.full_like(d, 9.0))
            fr = torch.nonzero(d.amax(dim=0) > 1e-3).flatten()
            print(f"    clip {b}: bad frames {int(fr.min())}..{int(fr.max())} ({fr.numel()} frames; tiles {sorted(set((fr // 128).tolist()))[:10]}), "
                  f"len/160 = {lens[b] / 160:.2f}, (len + 200) % 20480 = {(int(lens[b]) + 200) % TILE}")
        if focus and n_focus < 2:
            n_focus += 1
            replay(fe, pcm, d_starts, B, kw, ref, lens, n_mel)
print(f"fuzz: {n_batches} batches, {n_fail} failures (seed {seed}, {budget:.0f} s)")

#!/usr/bin/env python
"""Generates asr-finetune_b200/csrc/wfe_tc_epilogue_gen.inc: the straight-line epilogue of the tcgen05 log-mel kernel.

The tensor-core stage leaves, per frame (= TMEM lane = thread), the four 100-point sub-DFTs Y_n1[k2] (n1 = 0..3,
k2 = 0..51) of the DIT-4 split 400 = 4 x 100.  Twiddle + DFT-4 over n1 gives four bins per k2:
    k1 = 0 -> bin k2        k1 = 1 -> bin 100 + k2        k1 = 2 -> |X|^2 of bin 200 - k2        k1 = 3 -> bin 100 - k2
so the power spectrum arrives as four "fronts" sweeping the 201 bins.  The slaney filters are contiguous bands, hence at
any time only a dozen mel accumulators are live: with the filter-bank STRUCTURE known at compile time they stay in
registers and every non-zero is one FFMA with an immediate weight (on sm_100 a constant-bank operand costs an extra
LDCU, an immediate costs nothing).  The weights are baked bit-exactly from the same function the Python extractor
uses (`slaney_mel_filter_bank`); `wfe_create` compares the filter bank it is given against the baked non-zeros and
only then selects the tensor-core kernel.

Arithmetic restated from HF:models/whisper/feature_extraction_whisper.py:152-161 (mel_filters.T @ magnitudes, log10,
scale) — see SURVEY.md Appendix A steps 8, 9, 11.

Usage: python tools/gen_tc_epilogue.py   (rewrites the .inc; the output is committed)
"""
from __future__ import annotations

import importlib.util
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "asr-finetune_b200", "csrc", "wfe_tc_epilogue_gen.inc")


def filter_bank(n_mel: int) -> np.ndarray:
    """(201, n_mel) float32 — the product's own filter-bank function, loaded without importing torch."""
    src = os.path.join(ROOT, "asr-finetune_b200", "feature_extraction.py")
    text = open(src).read()
    a = text.index("# ---- slaney mel filter bank")
    b = text.index("class _Handle")
    ns: dict = {"np": np}
    exec(compile(text[a:b], src, "exec"), ns)
    return ns["slaney_mel_filter_bank"](201, n_mel, 0.0, 8000.0, 16000).astype(np.float32)


def bits(x: np.float32) -> int:
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]


def schedule():
    """[(pair p, element e, front f, bin)] in the order the epilogue produces power values."""
    order = []
    for p in range(26):
        for e in range(2):
            k2 = 2 * p + e
            if k2 > 50:
                continue
            for f, b in enumerate((k2, 100 + k2, 200 - k2, 100 - k2)):
                if k2 == 0 and f == 3:
                    continue  # bin 100 again
                if k2 == 50 and f in (2, 3):
                    continue  # bins 150 and 50 again
                order.append((p, e, f, b))
    assert sorted(b for _, _, _, b in order) == list(range(201))
    return order


def emit(n_mel: int, w: np.ndarray, out: list) -> None:
    """Two instruction streams, one per epilogue half (two warps share the 32 frames of a TMEM lane quarter):
    half 0 handles k2 pairs 0..12, half 1 pairs 13..25.  A filter fed by both halves is OWNED by half 0; half 1 meets
    those filters in its FIRST pairs (the four fronts cross the half boundary there), parks the partial sums in shared
    memory (TC_PART_PUT) and signals; half 0 picks them up after its own last pair (TC_PART_GET) -- it never waits."""
    order = schedule()
    half_of = lambda p: 0 if p < 13 else 1
    nz = {b: [m for m in range(n_mel) if w[b, m] != 0.0] for b in range(201)}
    halves_of_mel = {m: set() for m in range(n_mel)}
    for (p, e, f, b) in order:
        for m in nz[b]:
            halves_of_mel[m].add(half_of(p))
    shared = sorted(m for m in range(n_mel) if halves_of_mel[m] == {0, 1})
    slot_of = {m: i for i, m in enumerate(shared)}
    empty = [m for m in range(n_mel) if not halves_of_mel[m]]
    out.append(f"#if !defined(WFE_TC_GEN_HOST_TABLES) && !defined(WFE_TC_GEN_WINDOW) && !defined(WFE_TC_GEN_COUNTS) && WFE_TC_GEN_NMEL == {n_mel}")
    out.append(f"// ---- {n_mel} mel: {sum(len(v) for v in nz.values())} non-zeros, {len(empty)} empty filters, "
               f"{len(shared)} filters fed by both halves: {shared} ----")
    for hh in (0, 1):
        out.append(f"if (hh == {hh}) {{")
        seq = [(p, e, f, b) for (p, e, f, b) in order if half_of(p) == hh]
        first_time, last_time = {}, {}
        for t, (_, _, _, b) in enumerate(seq):
            for m in nz[b]:
                last_time[m] = t
                first_time.setdefault(m, t)
        if hh == 0:
            for m in empty:
                out.append(f"TC_FIN_ZERO({m})")
        pairs = sorted({p for (p, _, _, _) in seq})
        # TMEM reads: groups of two pairs (8 columns per n1), double-buffered: the loads of group g + 1 are issued before
        # the arithmetic of group g, the single tcgen05.wait::ld that follows the arithmetic then finds them done
        n_groups = (len(pairs) + 1) // 2
        out.append(f"TC_LOAD(0, {4 * pairs[0]})")
        out.append("TC_LOAD_WAIT()")
        cur_pair, t = -1, 0
        pending_put = [m for m in shared] if hh == 1 else []
        for (p, e, f, b) in seq:
            if p != cur_pair:
                i = pairs.index(p)
                if i % 2 == 0:
                    g = i // 2
                    if i > 0:
                        out.append("TC_LOAD_WAIT()")
                    if g + 1 < n_groups:
                        out.append(f"TC_LOAD({(g + 1) % 2}, {4 * pairs[2 * (g + 1)]})")
                    else:
                        out.append("TC_RELEASE()  // this thread's last TMEM read is done")
                out.append(f"TC_PAIR({(i // 2) % 2}, {i % 2}, {p})")
                cur_pair = p
            comp = "x" if e == 0 else "y"
            for m in nz[b]:
                wb = bits(w[b, m])
                if first_time[m] == t:
                    out.append(f"TC_ACC_SET({m}, P{f}.v.{comp}, 0x{wb:08x}u)  // bin {b}")
                else:
                    out.append(f"TC_ACC({m}, P{f}.v.{comp}, 0x{wb:08x}u)  // bin {b}")
                if last_time[m] == t:
                    if m in slot_of:
                        if hh == 1:
                            out.append(f"TC_PART_PUT({slot_of[m]}, {m})")
                            pending_put.remove(m)
                            if not pending_put:
                                out.append("TC_PART_SIGNAL()  // every partial sum is parked: wake half 0")
                        # half 0 finishes shared filters after the exchange (below)
                    else:
                        out.append(f"TC_FIN({m})")
            t += 1
        if hh == 0 and shared:
            out.append("TC_PART_WAIT()")
            for m in shared:
                out.append(f"TC_PART_GET({slot_of[m]}, {m})")
                out.append(f"TC_FIN({m})")
        out.append("}")
    out.append("#endif")
    out.append(f"#if defined(WFE_TC_GEN_COUNTS) && WFE_TC_GEN_NMEL == {n_mel}")
    out.append(f"constexpr int kTcShared{n_mel} = {len(shared)};")
    out.append("#endif")
    # host-side check table
    out.append(f"#if defined(WFE_TC_GEN_HOST_TABLES) && !defined(WFE_TC_GEN_WINDOW) && WFE_TC_GEN_NMEL == {n_mel}")
    trip = [(b, m, bits(w[b, m])) for b in range(201) for m in nz[b]]
    out.append(f"static const uint32_t kTcNnz{n_mel}[{len(trip)}][3] = {{")
    for i in range(0, len(trip), 6):
        out.append("  " + " ".join(f"{{{b},{m},0x{x:08x}u}}," for b, m, x in trip[i:i + 6]))
    out.append("};")
    out.append("#endif")


def emit_window(out: list) -> None:
    """Periodic Hann window (HF:audio_utils.py:593-607 == torch.hann_window(400)) as fp32 literals: with the prep loop
    fully unrolled every window factor becomes an FMUL immediate."""
    n = np.arange(400, dtype=np.float64)
    w = (0.5 - 0.5 * np.cos(2.0 * np.pi * n / 400.0)).astype(np.float32)
    out.append("#if defined(WFE_TC_GEN_WINDOW)")
    out.append("__device__ constexpr __align__(16) uint32_t kWinBits[400] = {")
    for i in range(0, 400, 8):
        out.append("  " + " ".join(f"0x{bits(x):08x}u," for x in w[i:i + 8]))
    out.append("};")
    out.append("#endif")


def main() -> int:
    out = ["// GENERATED by tools/gen_tc_epilogue.py -- do not edit.  Straight-line mel epilogue of wfe::tc::logmel_tc_kernel.",
           "// Include with WFE_TC_GEN_NMEL = 80 or 128 and the TC_* macros defined (see wfe_logmel_tc.cuh)."]
    emit_window(out)
    for n_mel in (80, 128):
        emit(n_mel, filter_bank(n_mel), out)
    with open(OUT, "w") as f:
        f.write("\n".join(out) + "\n")
    print(f"wrote {OUT}: {len(out)} lines")
    return 0


if __name__ == "__main__":
    sys.exit(main())

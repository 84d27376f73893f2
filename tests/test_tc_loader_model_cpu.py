"""Exhaustive CPU model of the raw-tile geometry of the tcgen05 log-mel kernel (asr-finetune_b200/csrc/wfe_logmel_tc.cuh).

The kernel loads a tile (128 frames) as ONE 2-D TMA box of 130 rows x 164 floats over a 160-float row pitch and lets the
loader warp patch the clip's edges INSIDE the landed tile (`tile_info`: classification; loader role: kModeAsyncHead /
kModeAsyncTail patches).  The index arithmetic of those patches is restated here line by line in numpy and checked, for
EVERY clip length that can put a clip's end into a tile (and every length around the reflect pad at sample 480000), against
what the reference extractor's padding defines (HF:models/whisper/feature_extraction_whisper.py:135-164 via torch.stft
center=True, pad_mode="reflect": zero-pad to 480000, then reflect 200 samples at both ends):
  * every store lands inside the 130 x 164 box (numpy raises on an out-of-range row; the round-2 bug this guards against
    started the zero fill at "row 130" for clips that end inside the box's four pad columns);
  * every sample a valid frame reads afterwards is the padded / reflected signal, whatever lies behind the clip.
It is a model (the GPU tests and tools/fuzz_tc_vs_cc.py check the kernel itself): when the kernel's patch code changes,
this file changes with it.
"""
import numpy as np
import pytest

HOP, NFFT, NS, NFRAMES = 160, 400, 480000, 3000
TILE_M, ROWS, PITCH = 128, 130, 164          # kTileM, kRawRows, kRawPitch
TILE = TILE_M * HOP
GARBAGE = 1.0e30                             # what lies behind the clip in the caller's buffer


def classify(tile: int, length: int) -> str:
    """wfe_logmel_tc.cuh `tile_info` (aligned float32 PCM, no normalisation, something follows the clip in the buffer)."""
    s_begin = tile * TILE - NFFT // 2
    nvalid = min(TILE_M, NFRAMES - tile * TILE_M)
    s_hi = s_begin + (nvalid - 1) * HOP + NFFT - 1
    lowest = max(s_begin, 0)
    if s_hi >= NS:
        lowest = min(lowest, 2 * (NS - 1) - s_hi)
    if lowest >= length:
        return "silent"
    box_end = s_begin + (ROWS - 1) * HOP + PITCH
    if box_end <= length:
        return "async" if s_begin >= 0 else "head"
    return "tail" if s_begin >= 0 else "sync"


def tma_box(buf: np.ndarray, first: int) -> np.ndarray:
    """The landed tile: raw[r, c] = buf[first + 160 r + c] (out-of-range coordinates read as zero)."""
    idx = first + HOP * np.arange(ROWS)[:, None] + np.arange(PITCH)[None, :]
    ok = (idx >= 0) & (idx < len(buf))
    return np.where(ok, buf[np.clip(idx, 0, len(buf) - 1)], 0.0).astype(np.float64)


def patch_head(raw: np.ndarray, length: int) -> None:
    """kModeAsyncHead: raw[i] = x[200 - i] = raw[400 - i], zero where the clip is shorter."""
    i = np.arange(NFFT // 2)
    j = NFFT - i
    v = np.where(NFFT // 2 - i < length, raw[j // HOP, j % HOP], 0.0)
    raw[i // HOP, i % HOP] = v


def patch_tail(raw: np.ndarray, tile: int, length: int) -> None:
    """kModeAsyncTail: zero from the clip's end on, then the reflect pad beyond sample 480000."""
    s_begin = tile * TILE - NFFT // 2
    i0 = length - s_begin
    r0, c0 = divmod(i0, HOP)
    if r0 < ROWS:
        raw[r0, c0:HOP] = 0.0                # (an out-of-range row raises: numpy does not wrap positive indices)
    for r in range(r0 + 1, ROWS):
        raw[r, :HOP] = 0.0
    ir = NS - s_begin
    if ir < ROWS * HOP:
        k = np.arange(NFFT // 2)
        j = ir - 2 - k
        v = np.where(j >= 0, raw[np.maximum(j, 0) // HOP, np.maximum(j, 0) % HOP], 0.0)
        i = ir + k
        keep = i < ROWS * HOP
        raw[i[keep] // HOP, i[keep] % HOP] = v[keep]


def expected(tile: int, x: np.ndarray) -> np.ndarray:
    """The padded / reflected signal at the linear positions the tile's valid frames read."""
    s_begin = tile * TILE - NFFT // 2
    nvalid = min(TILE_M, NFRAMES - tile * TILE_M)
    s = s_begin + np.arange((nvalid - 1) * HOP + NFFT)
    s = np.where(s < 0, -s, s)
    s = np.where(s >= NS, 2 * (NS - 1) - s, s)
    return np.where(s < len(x), x[np.clip(s, 0, len(x) - 1)], 0.0)


def run_case(tile: int, length: int, x_full: np.ndarray) -> None:
    mode = classify(tile, length)
    if mode == "silent":  # never loaded, never computed: its valid frames must see nothing but padding
        assert not expected(tile, x_full[:length]).any(), (tile, length)
        return
    if mode == "sync":
        return
    s_begin = tile * TILE - NFFT // 2
    buf = np.concatenate([np.full(256, GARBAGE), x_full[:length], np.full(TILE + 1024, GARBAGE)])  # garbage all around
    raw = tma_box(buf, 256 + s_begin)
    if mode == "head":
        patch_head(raw, length)
    elif mode == "tail":
        patch_tail(raw, tile, length)
    want = expected(tile, x_full[:length])
    lin = np.arange(len(want))
    got = raw[lin // HOP, lin % HOP]
    assert np.array_equal(got, want), (tile, length, mode, int(np.argmax(got != want)))


@pytest.fixture(scope="module")
def signal():
    rng = np.random.default_rng(0)
    return np.round(rng.standard_normal(NS + 4096) * 1000.0)  # distinct integers: an index slip cannot go unnoticed


def test_every_clip_end_inside_a_tile(signal):
    # tile 1: every position of the clip's end relative to the tile, from "the tile holds one sample" to "the clip covers
    # the whole box" (the lengths 20800..20803 past the tile's first sample end inside the box's pad columns)
    s_begin = TILE - NFFT // 2
    for i0 in range(1, (ROWS - 1) * HOP + PITCH + 3):
        run_case(1, s_begin + i0, signal)


def test_clip_ends_in_the_pad_columns_are_tail_tiles_of_every_tile(signal):
    for tile in range(1, 23):
        s_begin = tile * TILE - NFFT // 2
        for i0 in range(ROWS * HOP - 3, ROWS * HOP + 6):
            assert classify(tile, s_begin + i0) == ("tail" if i0 < ROWS * HOP + 4 else "async")
            run_case(tile, s_begin + i0, signal)


def test_last_tile_reflect_pad_for_every_length(signal):
    # tile 23 (56 valid frames) reaches sample 480000: reflect pad from the tile itself, clip end anywhere in the tile
    s_begin = 23 * TILE - NFFT // 2
    for length in list(range(s_begin - 450, s_begin + 450)) + list(range(s_begin + 450, NS - 450, 37)) + list(range(NS - 450, NS + 1)):
        run_case(23, length, signal)
        run_case(22, length, signal)


def test_silent_tiles_see_only_padding(signal):
    # around every tile's first sample, and around the reflect pad of the last tile, for every tile
    for tile in range(1, 24):
        s_begin = tile * TILE - NFFT // 2
        for length in (1, s_begin - 1, s_begin, s_begin + 1):
            for t in range(tile, 24):
                if classify(t, length) == "silent":
                    run_case(t, length, signal)
    assert classify(23, 23 * TILE - NFFT // 2) == "silent" and classify(23, 23 * TILE - NFFT // 2 + 1) == "tail"
    # a clip of one sample: only the first tile is not silent
    assert [classify(t, 1) for t in range(24)] == ["sync"] + ["silent"] * 23


def test_first_tile_head_patch(signal):
    # a clip's first tile takes the TMA + head patch once the whole box lies inside the clip (20604 samples); shorter
    # clips are staged by the workers (the generic path, not modelled here)
    box_end = -NFFT // 2 + (ROWS - 1) * HOP + PITCH
    assert classify(0, box_end - 1) == "sync" and classify(0, box_end) == "head"
    for length in list(range(box_end, box_end + 330)) + [TILE, 100000, NS]:
        run_case(0, length, signal)

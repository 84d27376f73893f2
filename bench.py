#!/usr/bin/env python
"""bench.py — log-mel audio-seconds/sec of the B200-native Whisper frontend (+ label collate), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...     # the reference's CPU extractor on the host cores
    torchrun --nproc-per-node N bench.py --gpus N ...               # one rank per GPU, per-rank shard, no data collective

Workloads (every N, weak scaling): each rank owns a shard of synthetic 16 kHz clips, 30 s each, and a "step" is one
pass of the hot path over one per-rank batch of `--batch` (256) clips: `wfe_logmel` (pad / reflect / STFT / power / mel /
log10 / clamp / scale, 128 mel = large-v3) + `wfe_collate` (label pad, -100 fill, BOS flag).  Timed on two inputs --
white noise and speech-like dynamics (noise under 0.2-s segments with random gains over 60 dB: the per-clip clamp has work
to do, as on real recordings) -- and the line LEADS WITH THE LOWER ONE; `workloads` holds both, plus BASELINE configs[2]
(1024 ragged clips + masks + labels).  `--workload shard` runs configs[3] (100 k clips over the ranks, whole shard).
  value  device-resident PCM -> device-resident features, CUDA events, max over ranks.
  e2e    the public drop-in call (`StreamingFrontendCollator` = `WhisperFeatureExtractor(list_of_host_clips)` + labels)
         fed what the reference's loader yields -- one separately allocated PAGEABLE float32 numpy array per clip -- and
         returning host tensors: H2D, kernels, D2H inside the timed region; median of >= 20 steps.  `variants` adds the
         pinned-contiguous input, int16 in / fp16 out and the training path (features stay on device).
  e2e_reference_loop  the reference's unmodified per-clip loop + fe.pad + .to(cuda) at per-device batch 8, through the
         drop-in and through the CPU extractor.
  roofline  algorithmic bytes (4 B/sample read + 4 B/feature written) / live CUDA-event time of the logmel call (main
         kernel + clamp pass), against MEASURED_PEAKS.json's HBM copy bandwidth; `traffic` is RECORDED from the committed
         ncu capture (profiles/roofline_traffic.json), not measured in this run.
  cpu_baseline  the reference extractor timed on this box's host cores on a bounded sample (N=1, rank 0 only).
Multi-GPU plumbing (barrier, max over ranks) runs over gloo: the data path has no collective.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "log-mel audio-seconds/sec"
UNIT = "audio-seconds/sec"
SR, N_SAMPLES, N_FRAMES = 16000, 480000, 3000
CLIP_SECONDS = 30.0
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--batch", type=int, default=256, help="clips per rank per step")
    ap.add_argument("--n-mel", type=int, default=128, help="128 = large-v3 (headline), 80 = whisper-small (configs[1])")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed host-buffer steps per variant (default 20)")
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="CPU time budget per reference variant")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-clips", type=int, default=0, help="clips per reference step (default: sized from the cores)")
    ap.add_argument("--workload", choices=["all", "headline", "noise", "speechlike", "config3", "shard"], default="all",
                    help="all = noise + speech-like (headline = the lower) + configs[2] ragged batch; shard = configs[3]")
    ap.add_argument("--shard-clips", type=int, default=100000, help="total clips of the --workload shard run")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cross-check", action="store_true",
                    help="skip the CUDA-core-kernel comparison after each device-resident workload (profiling runs)")
    return ap.parse_args()


def workload_config(args, n_gpus):
    model = "large-v3" if args.n_mel == 128 else ("whisper-small" if args.n_mel == 80 else f"{args.n_mel}-mel")
    return {
        "workload": f"{model} {args.n_mel}-mel log-mel extraction + label collate, {args.batch} synthetic 30-s 16 kHz "
                    f"clips per GPU per step (per-rank shard of BASELINE configs[3], batch size of configs[1])",
        "n_mel": args.n_mel, "clips_per_gpu_per_step": args.batch, "global_clips_per_step": args.batch * n_gpus,
        "clip_seconds": CLIP_SECONDS, "parallelism": f"clip-sharded x{n_gpus}, no collective",
        "l2": "per-step input (491 MB) and output (393 MB at 128 mel) exceed the 126 MB L2; no flush needed",
    }


def synth_clip(i: int):
    import numpy as np

    return (0.1 * np.random.default_rng(i).standard_normal(N_SAMPLES, dtype=np.float32)).astype(np.float32)


def synth_labels(batch: int, seed: int = 1337):
    """[SOT, de, transcribe, notimestamps, text..., EOT] id lists, len ~ U{5..448}."""
    import numpy as np

    rng = np.random.default_rng(seed)
    out = []
    for n in rng.integers(5, 449, size=batch):
        body = rng.integers(0, 50257, size=int(n) - 5).tolist()
        out.append([50258, 50261, 50360, 50364] + body + [50257])
    return out


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (transformers.WhisperFeatureExtractor, per-clip call loop as in
# ref:finetune/training/data_and_collator/datasets_and_collators.py:191-195), on all the host cores it can use
# ------------------------------------------------------------------------------------------------------------------
_REF_FE = None
_REF_KIND = None


def _ref_worker_init(n_mel):
    global _REF_FE, _REF_KIND
    try:
        import torch

        torch.set_num_threads(1)
    except Exception:
        pass
    _REF_FE, _REF_KIND = _make_reference_extractor(n_mel)


def _make_reference_extractor(n_mel):
    """('reference') the installed transformers extractor = the module the reference calls; else ('port') the oracle."""
    try:
        from transformers import WhisperFeatureExtractor  # third-party dependency that holds the arithmetic

        fe = WhisperFeatureExtractor(feature_size=n_mel)
        return (lambda clip: fe(clip, sampling_rate=16000).input_features[0]), "reference"
    except Exception:
        from oracle import logmel as ologmel  # CPU restatement (allowed here: cpu_baseline / reference arm only)

        return (lambda clip: ologmel.logmel_clip(clip, n_mel, "fp32")), "port"


def _ref_worker_run(args):
    seed, count = args
    clip = synth_clip(seed)
    t0 = time.perf_counter()
    for _ in range(count):
        out = _REF_FE(clip)
    assert out.shape[-1] == N_FRAMES
    return time.perf_counter() - t0


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class ReferenceRunner:
    """Process pool (one single-threaded extractor per core, like the reference's Ray num_cpus fan-out,
    ref:finetune/prepare_dataset/materialize_dataset.py:165-170) + the in-process loop with default torch threads."""

    def __init__(self, n_mel: int, max_workers: int = 128):
        import multiprocessing as mp

        self.n_mel = n_mel
        self.workers = max(1, min(host_cores(), max_workers))
        self.pool = mp.get_context("spawn").Pool(self.workers, initializer=_ref_worker_init, initargs=(n_mel,))
        self.pool.map(_ref_worker_run, [(i, 1) for i in range(self.workers)])  # import + first-call warm-up
        self.fe, self.kind = _make_reference_extractor(n_mel)

    def pool_step(self, clips_per_worker: int) -> tuple:
        t0 = time.perf_counter()
        self.pool.map(_ref_worker_run, [(i, clips_per_worker) for i in range(self.workers)], chunksize=1)
        return time.perf_counter() - t0, self.workers * clips_per_worker

    def loop_step(self, clips) -> tuple:
        t0 = time.perf_counter()
        for c in clips:
            self.fe(c)
        return time.perf_counter() - t0, len(clips)

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # under torchrun only rank 0 measures the host; the others exit 0 without work
    n_gpus = max(1, args.gpus)
    runner = ReferenceRunner(args.n_mel)
    import torch

    try:
        # size one step to ~0.5-2 s of wall time: calibrate the per-worker clip count once
        t, n = runner.pool_step(1)
        per_worker = args.ref_clips // runner.workers if args.ref_clips else max(1, int(round(1.0 / max(t, 1e-3))))
        per_worker = max(1, min(per_worker, 64))
        for _ in range(max(args.warmup, 1)):
            runner.pool_step(per_worker)
        t_pool, clips_pool = 0.0, 0
        for _ in range(args.steps):
            t, n = runner.pool_step(per_worker)
            t_pool += t
            clips_pool += n
        pool_rate = clips_pool * CLIP_SECONDS / t_pool
        # reference-as-used: one process, per-clip loop, default torch intra-op threads
        clips = [synth_clip(i) for i in range(8)]
        runner.loop_step(clips[:2])
        t_loop, clips_loop, t_end = 0.0, 0, time.perf_counter() + args.cpu_seconds
        while time.perf_counter() < t_end:
            t, n = runner.loop_step(clips)
            t_loop += t
            clips_loop += n
        loop_rate = clips_loop * CLIP_SECONDS / t_loop
    finally:
        runner.close()
    use_pool = pool_rate >= loop_rate
    value = pool_rate if use_pool else loop_rate
    step_clips = runner.workers * per_worker
    ms_per_step = (t_pool / args.steps * 1e3) if use_pool else (step_clips * CLIP_SECONDS / loop_rate * 1e3)
    cores = runner.workers if use_pool else torch.get_num_threads()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, n_gpus),
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": cores, "kind": runner.kind,
            "sample": f"{step_clips} x 30-s clips per step ({runner.workers} single-thread worker processes x "
                      f"{per_worker} clips), {args.steps} steps; per-clip call fe(clip, sampling_rate=16000) exactly as "
                      f"the reference's collator loop does",
            "pool_audio_s_per_s": pool_rate, "pool_workers": runner.workers,
            "inprocess_loop_audio_s_per_s": loop_rate, "inprocess_torch_threads": torch.get_num_threads(),
            "host_cores": host_cores(),
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:
        import transformers

        line["versions"] = {"transformers": transformers.__version__, "torch": torch.__version__}
    except Exception:
        pass
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.thr, self.idx = [], None, None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thr = threading.Thread(target=self._pump, daemon=True)
        self.thr.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower() == "active":
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic(n_mel: int):
    """(dram bytes per logmel launch per clip, source) RECORDED from the committed ncu --set full capture under profiles/
    (not measured in this run: ncu cannot run inside the timed bench), else (None, None)."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return d.get(f"logmel_f32_{n_mel}mel_dram_bytes_per_clip"), "recorded: " + d.get("source", "profiles/roofline_traffic.json")
    except Exception:
        return None, None


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def _stats(times_s):
    """median / min / max of per-step wall times (seconds)."""
    t = sorted(times_s)
    return {"median_ms": t[len(t) // 2] * 1e3, "min_ms": t[0] * 1e3, "max_ms": t[-1] * 1e3, "steps": len(t)}


def _timed_steps(fn, steps, sync, warm=2):
    for _ in range(warm):
        fn()
    sync()
    out = []
    for _ in range(steps):
        t0 = time.perf_counter()
        fn()
        sync()
        out.append(time.perf_counter() - t0)
    return out


def pin_rank_to_cores(local_rank: int, world: int) -> dict:
    """Give every rank its own slice of the host cores, preferring the cores `nvidia-smi topo -m` lists as local to its
    GPU (NUMA), and size the library's staging-copy pool to the slice: eight unpinned ranks each starting eight copy
    threads oversubscribe a 32-core host.  Best effort; returns what was done for the JSON line."""
    info = {"pinned": False}
    try:
        allowed = sorted(os.sched_getaffinity(0))
        local = None
        try:
            txt = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
            for ln in txt.splitlines():
                f = ln.split()
                if f and f[0] == f"GPU{local_rank}":
                    for tok in f[1:]:
                        if tok and tok[0].isdigit() and all(ch.isdigit() or ch in "-," for ch in tok) and ("-" in tok or "," in tok):
                            cores = set()
                            for part in tok.split(","):
                                a, _, b = part.partition("-")
                                cores.update(range(int(a), int(b or a) + 1))
                            local = sorted(cores & set(allowed))
                            break
        except Exception:
            local = None
        per = max(1, len(allowed) // world)
        if local and len(local) >= per:
            # ranks whose GPUs share a NUMA node split that node's cores between them
            same = max(1, world * len(local) // max(1, len(allowed)))
            k = local_rank % same
            mine = local[k * per:(k + 1) * per] or local[:per]
            info["numa_local"] = True
        else:
            mine = allowed[local_rank * per:(local_rank + 1) * per] or allowed
            info["numa_local"] = False
        os.sched_setaffinity(0, mine)
        os.environ.setdefault("WFE_HOST_THREADS", str(max(1, min(8, len(mine)))))
        info.update({"pinned": True, "cores": len(mine), "first_core": mine[0], "copy_threads": os.environ["WFE_HOST_THREADS"]})
    except Exception as e:
        info["error"] = str(e)[:120]
    return info


def run_b200(args):
    import ctypes as C

    import numpy as np
    import torch

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = world if world > 1 else 1

    # the reference CPU extractor on this box's host cores, before CUDA is touched (rank 0, N=1 only)
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "5", "--warmup", "1",
                   "--n-mel", str(args.n_mel), "--batch", str(args.batch), "--cpu-seconds", str(args.cpu_seconds)]
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
            for ln in reversed(res.stdout.strip().splitlines()):
                if ln.startswith("{"):
                    cpu_baseline = json.loads(ln)["cpu_baseline"]
                    break
            if cpu_baseline is None:
                cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                "sample": "failed: " + (res.stderr.strip().splitlines() or ["no output"])[-1][:200]}
        except Exception as e:  # the GPU number must not depend on the CPU leg
            cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"failed: {e}"}

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm"
    pin = pin_rank_to_cores(local_rank, world) if world > 1 else {"pinned": False}
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        # plumbing only (barrier + max-over-ranks of the timings): gloo on the host -- the data path has no collective
        dist.init_process_group("gloo")

    import asr_finetune_b200 as pkg

    B, n_mel = args.batch, args.n_mel
    fe = pkg.WhisperFeatureExtractor(feature_size=n_mel, cuda_device=local_rank)
    h = fe._handle(None, dev)
    lib = pkg._lib.load()

    def sync():
        torch.cuda.synchronize(dev)

    def barrier():
        sync()
        if world > 1:
            dist.barrier()
            sync()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- per-rank shard of synthetic clips: pinned host PCM and a device-resident copy ----
    shard = pkg.rank_shard(B * world, rank, world)
    host_pcm = torch.empty((B, N_SAMPLES), dtype=torch.float32, pin_memory=True)
    distinct = min(B, 32)
    base = [synth_clip(shard.start + i) for i in range(distinct)]
    hp = host_pcm.numpy()
    for i in range(B):
        # distinct noise for the first `distinct` clips, then gain-scaled repeats (keeps set-up time bounded)
        np.multiply(base[i % distinct], np.float32(1.0 - 0.5 * (i // distinct) / max(1, B // distinct)), out=hp[i])
    labels = synth_labels(B, seed=1337 + rank)
    d_noise = host_pcm.to(dev).view(-1)
    # speech-like dynamics: the same noise under 0.2-s segments with random gains over 60 dB -- the per-clip
    # clamp (max - 8) then has something to do in most tiles, as it has on real recordings
    g = torch.Generator(device=dev)
    g.manual_seed(4242 + rank)
    seg = torch.rand(B * N_SAMPLES // 3200, device=dev, generator=g)
    d_speech = d_noise * torch.repeat_interleave(10.0 ** (-3.0 * seg), 3200)
    del seg
    d_offs = (torch.arange(B + 1, dtype=torch.int64) * N_SAMPLES).to(dev)
    d_out = torch.empty((B, n_mel, N_FRAMES), dtype=torch.float32, device=dev)
    packed, lens = pkg.collator._pack_ids(labels)
    d_packed = packed.to(dev)
    width = int(lens.max())
    d_labels = torch.empty((B, width), dtype=torch.int64, device=dev)
    d_flag = torch.zeros(1, dtype=torch.int32, device=dev)

    def collate_labels():
        pkg._lib.check(lib.wfe_collate(h.ptr, d_packed[B + 1:].data_ptr(), d_packed[:B + 1].data_ptr(), B, width, 50258,
                                       -100, d_labels.data_ptr(), d_flag.data_ptr(), None, 0, None,
                                       C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "wfe_collate")

    bytes_per_clip = N_SAMPLES * 4 + n_mel * N_FRAMES * 4
    peak, peak_src = measured_hbm_peak()

    def device_workload(d_pcm, steps, offs=None, lengths=None, batch=B, out=None, audio_s=None, with_mask=False):
        """K timed steps of logmel + label collate on device-resident PCM; CUDA events, max over ranks.  A measurement in
        which the kernel's watchdog fired (an internal wait gave up: never expected) is repeated once and flagged."""
        w = _device_workload(d_pcm, steps, offs, lengths, batch, out, audio_s, with_mask)
        if w["kernel_timeouts"]:
            w = _device_workload(d_pcm, steps, offs, lengths, batch, out, audio_s, with_mask)
            w["remeasured_after_kernel_timeout"] = True
        return w

    def _device_workload(d_pcm, steps, offs, lengths, batch, out, audio_s, with_mask):
        out = d_out if out is None else out
        offs = d_offs if offs is None else offs

        def step(ev=None):
            if ev is not None:
                ev[0].record()
            fe.logmel_device(d_pcm, offs, batch, out=out, lengths=lengths, return_attention_mask=with_mask)
            if ev is not None:
                ev[1].record()
            collate_labels()

        for _ in range(max(args.warmup, 3)):
            step()
        barrier()
        l0 = pkg._lib.launch_count()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for s_ in range(steps):
            step(evs[s_])
        e1.record()
        barrier()
        launches = pkg._lib.launch_count() - l0
        total_ms = max_over_ranks(e0.elapsed_time(e1))
        kern_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in evs) / steps)
        audio = (batch * CLIP_SECONDS if audio_s is None else audio_s) * world
        timeouts = int(max_over_ranks(float(fe.debug_kernel_error() != 0)))
        return {"value": audio * steps / (total_ms * 1e-3), "ms_per_step": total_ms / steps, "kernel_ms": kern_ms,
                "launches": int(launches), "steps": steps, "kernel_timeouts": timeouts}

    def cross_check(d_pcm, out, offs=None, lengths=None, batch=B):
        """Outside every timed region: the timed output against the CUDA-core kernel (WFE_DISABLE_TC=1: the product's
        second, independent implementation of the same arithmetic) on the same input, max |difference| over the whole
        batch.  Reporting only: never fails the bench; the environment switch is restored whatever happens."""
        if args.no_cross_check:
            return None
        prev = os.environ.get("WFE_DISABLE_TC")
        try:
            os.environ["WFE_DISABLE_TC"] = "1"
            ref, _ = fe.logmel_device(d_pcm, d_offs if offs is None else offs, batch, lengths=lengths)
            torch.cuda.synchronize()
            diff = float((out - ref).abs().max())
            del ref
            return diff
        except Exception as exc:  # noqa: BLE001
            return f"unavailable: {type(exc).__name__}: {exc}"[:200]
        finally:
            if prev is None:
                os.environ.pop("WFE_DISABLE_TC", None)
            else:
                os.environ["WFE_DISABLE_TC"] = prev

    def roofline_of(w, batch=B, alg_bytes=None):
        alg = batch * bytes_per_clip if alg_bytes is None else alg_bytes
        ach = alg / (w["kernel_ms"] * 1e-3) / 1e9
        return {"achieved_gbs": ach, "frac": ach / peak, "algorithmic_bytes_per_launch": alg, "kernel_ms_per_launch": w["kernel_ms"]}

    sampler = ClockSampler(local_rank)  # nvidia-smi samples every 100 ms from here to the end of the e2e region
    if rank == 0:
        sampler.start()

    workloads = {}
    if args.workload in ("all", "headline", "noise"):
        workloads["noise"] = device_workload(d_noise, args.steps)
        span = float((d_out.amax(dim=(1, 2)) - d_out.amin(dim=(1, 2))).max())
        assert torch.isfinite(d_out).all() and 0.0 < span <= 2.0 + 1e-5, span  # the timed output is real
        workloads["noise"]["max_abs_diff_vs_cuda_core_kernel"] = cross_check(d_noise, d_out)
        ref_out = d_out.clone() if args.workload != "noise" else None
    if args.workload in ("all", "headline", "speechlike"):
        workloads["speechlike"] = device_workload(d_speech, args.steps)
        span = float((d_out.amax(dim=(1, 2)) - d_out.amin(dim=(1, 2))).max())
        assert torch.isfinite(d_out).all() and 0.0 < span <= 2.0 + 1e-5, span
        workloads["speechlike"]["max_abs_diff_vs_cuda_core_kernel"] = cross_check(d_speech, d_out)
    if args.workload in ("all", "config3"):
        # BASELINE configs[2]: 1024 clips of 1-30 s, padded to 3000 frames, attention masks + labels
        Bc = 1024
        rngc = np.random.default_rng(1337 + rank)
        lens_c = rngc.integers(16000, N_SAMPLES + 1, size=Bc)
        starts = np.zeros(Bc, dtype=np.int64)
        np.cumsum((lens_c[:-1] + 3) & ~3, out=starts[1:])
        d_pcm_c = 0.1 * torch.randn(int(starts[-1] + lens_c[-1]), device=dev, generator=g)
        d_out_c = torch.empty((Bc, n_mel, N_FRAMES), dtype=torch.float32, device=dev)
        wc = device_workload(d_pcm_c, max(3, min(args.steps, 10)), offs=torch.from_numpy(starts).to(dev),
                             lengths=torch.from_numpy(lens_c).to(dev), batch=Bc, out=d_out_c,
                             audio_s=float(lens_c.sum()) / SR, with_mask=True)
        wc["clips_per_s"] = Bc * world / (wc["ms_per_step"] * 1e-3)
        wc["roofline"] = roofline_of(wc, alg_bytes=int(lens_c.sum()) * 4 + Bc * n_mel * N_FRAMES * 4)
        wc["note"] = "1024 ragged clips (1-30 s); every clip still writes 3000 frames; `value` counts real audio seconds"
        wc["max_abs_diff_vs_cuda_core_kernel"] = cross_check(d_pcm_c, d_out_c, offs=torch.from_numpy(starts).to(dev),
                                                             lengths=torch.from_numpy(lens_c).to(dev), batch=Bc)
        workloads["config3_ragged_1024"] = wc
        del d_pcm_c, d_out_c
    if args.workload == "shard":
        # BASELINE configs[3]: 100 k clips sharded over the ranks (ref:finetune/training/trainers/trainers.py:785-791),
        # audio generated on the device chunk by chunk, every chunk checked by device-side invariants
        ev_sum, n_done = 0.0, 0
        barrier()
        t_wall0 = time.perf_counter()
        for idx in pkg.shard_batches(args.shard_clips, B, rank, world):
            lo, n = idx.start, len(idx)
            gen = torch.Generator(device=dev)
            gen.manual_seed(lo)
            d_chunk = 0.1 * torch.randn(n * N_SAMPLES, device=dev, generator=gen)
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fe.logmel_device(d_chunk, d_offs[:n + 1], n, out=d_out[:n])
            b_.record()
            sync()
            ev_sum += a.elapsed_time(b_)
            mx, mn = d_out[:n].amax(dim=(1, 2)), d_out[:n].amin(dim=(1, 2))
            assert bool(torch.isfinite(mx).all()) and bool(((mx - mn) <= 2.0 + 1e-5).all()) and bool((mx - mn > 0).all())
            n_done += n
        wall = max_over_ranks(time.perf_counter() - t_wall0)
        hot_ms = max_over_ranks(ev_sum)
        workloads["shard"] = {"value": n_done * world * CLIP_SECONDS / (hot_ms * 1e-3), "clips": n_done * world,
                              "clips_per_rank": n_done, "kernel_ms": hot_ms / max(1, (n_done + B - 1) // B),
                              "ms_per_step": hot_ms / max(1, (n_done + B - 1) // B), "launches": 2 * ((n_done + B - 1) // B),
                              "steps": (n_done + B - 1) // B,
                              "wall_s_including_generation_and_checks": wall,
                              "note": "full per-rank shard iterated with shard_batches; PCM generated on the device per "
                                      "chunk (outside the CUDA-event region); per-chunk invariants asserted"}
    for k_, w_ in workloads.items():
        if "roofline" not in w_:
            w_["roofline"] = roofline_of(w_)

    # headline: the LOWER of noise and speech-like (the reference's data is speech)
    if args.workload in ("all", "headline"):
        head_name = min(("noise", "speechlike"), key=lambda k_: workloads[k_]["value"])
    else:
        head_name = {"config3": "config3_ragged_1024"}.get(args.workload, args.workload)
    head = workloads[head_name]

    # ---- e2e: the public drop-in call with HOST buffers; H2D + kernels + D2H inside the timed region ----
    e2e = None
    if args.workload in ("all", "headline") and not args.no_e2e:
        e2e_steps = args.e2e_steps or 20
        audio_s_per_step = B * CLIP_SECONDS * world
        # (1) what the reference's loader yields: one separately allocated PAGEABLE float32 numpy array per clip
        #     (`np.array(h5['audio'][idx]).copy()`, ref ...datasets_and_collators.py:87)
        pageable = [np.array(hp[i], copy=True) for i in range(B)]
        pinned_views = [hp[i] for i in range(B)]
        coll_host = pkg.StreamingFrontendCollator(fe, device="cpu")
        coll_host16 = pkg.StreamingFrontendCollator(fe, device="cpu", feature_dtype=torch.float16)
        coll_dev = pkg.StreamingFrontendCollator(fe)  # training path: features and labels stay on the device

        def run_variant(fn):
            ts = _timed_steps(fn, e2e_steps, sync)
            st = _stats(ts)
            med = max_over_ranks(st["median_ms"])
            st["value"] = audio_s_per_step / (med * 1e-3)
            st["value_best"] = audio_s_per_step / (max_over_ranks(st["min_ms"]) * 1e-3)
            return st

        res_holder = {}

        def f_pageable():
            res_holder["o"] = coll_host({"audio": pageable, "labels": labels})

        v_pageable = run_variant(f_pageable)
        h2d, d2h = fe.last_transfer_bytes
        feats_h, lab_h = res_holder["o"]["input_features"], res_holder["o"]["labels"]
        h2d += int(packed.numel()) * 8
        d2h += int(lab_h.numel()) * 8
        if "noise" in workloads and ref_out is not None:
            assert torch.equal(feats_h, ref_out.cpu()), "host-buffer path and device-resident path disagree"
        v_pinned = run_variant(lambda: coll_host({"audio": pinned_views, "labels": labels}))
        # (2) SURVEY 8 f-1 / f-2: int16 PCM in (what HDF5 stores before the reference's float32 cast), fp16 features out
        #     (the autocast consumer's dtype): 246 + 197 MB instead of 492 + 393 MB across PCIe per step
        pcm16 = np.clip(np.round(hp * 32767.0), -32768, 32767).astype(np.int16)
        pageable16 = [np.array(pcm16[i], copy=True) for i in range(B)]
        v_i16_f16 = run_variant(lambda: coll_host16({"audio": pageable16, "labels": labels}))
        h2d16, d2h16 = fe.last_transfer_bytes
        v_i16_f32 = run_variant(lambda: coll_host({"audio": pageable16, "labels": labels}))

        # (3) the training path: host clips in, features stay on the device for the model (H2D + kernels; the step result
        #     read back is a checksum of the batch)
        def f_train():
            o = coll_dev({"audio": pageable, "labels": labels})
            res_holder["chk"] = float(o["input_features"][:, 0, 0].sum().item())

        v_train = run_variant(f_train)
        # the host's ceiling for this step: nothing but the two PCIe copies (pinned 492 MB up || 393 MB down), all ranks at once
        host_out = torch.empty((B, n_mel, N_FRAMES), dtype=torch.float32, pin_memory=True)
        s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

        def bare_copies():
            with torch.cuda.stream(s_up):
                d_noise.view(B, N_SAMPLES).copy_(host_pcm, non_blocking=True)
            with torch.cuda.stream(s_dn):
                host_out.copy_(d_out, non_blocking=True)
            s_up.synchronize()
            s_dn.synchronize()

        barrier()
        v_copies = run_variant(bare_copies)
        del host_out
        e2e = {"value": v_pageable["value"], "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": e2e_steps, "median_ms": v_pageable["median_ms"], "min_ms": v_pageable["min_ms"],
               "value_best_step": v_pageable["value_best"],
               "api": "StreamingFrontendCollator(fe, device='cpu')({'audio': clips, 'labels': id_lists}) = "
                      "WhisperFeatureExtractor(list_of_host_clips) + label collate; inputs = one separately allocated "
                      "PAGEABLE float32 numpy array per clip (what the reference's loader yields), outputs = host "
                      "tensors; wall clock per step, median of the steps, max over ranks",
               "copy_ceiling": {"value": v_copies["value"], "median_ms": v_copies["median_ms"],
                                "note": "the two bare PCIe copies of a step from/to pinned memory, all ranks concurrently: what the "
                                        "host (PCIe root, DRAM) allows whatever the code does"},
               "frac_of_copy_ceiling": v_pageable["value"] / v_copies["value"],
               "variants": {
                   "fp32_pageable_in_fp32_host_out": v_pageable,
                   "fp32_pinned_contiguous_in_fp32_host_out": v_pinned,
                   "int16_pageable_in_fp16_host_out": dict(v_i16_f16, h2d_bytes_per_step=h2d16 + int(packed.numel()) * 8,
                                                           d2h_bytes_per_step=d2h16 + int(lab_h.numel()) * 8),
                   "int16_pageable_in_fp32_host_out": v_i16_f32,
                   "training_path_fp32_pageable_in_device_out": v_train}}

    # ---- the reference's UNMODIFIED call pattern at per-device batch 8 (ref ...datasets_and_collators.py:191-195,
    #      236-240 + trainers/utils.py:108-112): per-clip extractor call, fe.pad, .to(cuda) -- through the drop-in and
    #      through the CPU extractor, same clips ----
    ref_loop = None
    if args.workload in ("all", "headline") and rank == 0 and not args.no_e2e:
        rng8 = np.random.default_rng(8)
        clips8 = [np.array(hp[i][:int(n)], copy=True) for i, n in enumerate(rng8.integers(3 * SR, N_SAMPLES + 1, size=8))]

        def loop_with(fe_any):
            mel = []
            for audio in clips8:
                feats = fe_any(audio, sampling_rate=16000)
                mel.append({"input_features": feats.input_features[0]})
            padded = fe_any.pad(mel, padding="longest", return_tensors="pt")
            return padded.input_features.to(dev)

        t_ours = _stats(_timed_steps(lambda: loop_with(fe), 20, sync))
        ref_loop = {"batch": 8, "drop_in_ms": t_ours["median_ms"], "drop_in_min_ms": t_ours["min_ms"],
                    "drop_in_audio_s_per_s": sum(len(c) for c in clips8) / SR / (t_ours["median_ms"] * 1e-3)}
        t_batched = _stats(_timed_steps(lambda: coll_dev({"audio": clips8, "labels": labels[:8]}), 20, sync))
        ref_loop["batched_collator_ms"] = t_batched["median_ms"]
        try:
            from transformers import WhisperFeatureExtractor as HFExtractor

            hf = HFExtractor(feature_size=n_mel)
            t_hf = _stats(_timed_steps(lambda: loop_with(hf), 10, sync, warm=1))
            ref_loop["cpu_extractor_ms"] = t_hf["median_ms"]
            ref_loop["speedup"] = t_hf["median_ms"] / t_ours["median_ms"]
        except Exception as e:
            ref_loop["cpu_extractor_ms"] = None
            ref_loop["note"] = f"transformers extractor unavailable: {e}"

    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        traffic_per_clip, traffic_src = recorded_traffic(n_mel)
        rl = head["roofline"]
        model = "large-v3" if n_mel == 128 else ("whisper-small" if n_mel == 80 else f"{n_mel}-mel")
        cfg = workload_config(args, n_gpus)
        cfg["workload"] = {
            "noise": f"{model} {n_mel}-mel log-mel extraction + label collate, {B} synthetic 30-s 16 kHz white-noise clips per GPU per step",
            "speechlike": f"{model} {n_mel}-mel log-mel extraction + label collate, {B} synthetic 30-s 16 kHz clips per GPU per step, "
                          "speech-like dynamics (noise under 0.2-s segments with random gains over 60 dB)",
            "config3_ragged_1024": f"{model} variable-length clips (1-30 s) padded to 3000 frames + masks + labels, batch 1024 (BASELINE configs[2])",
            "shard": f"{model} per-rank sharded extraction, {args.shard_clips} synthetic clips over {n_gpus} GPU(s) (BASELINE configs[3])",
        }[head_name]
        cfg["headline_workload"] = head_name
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": n_gpus, "steps": head["steps"],
            "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "gpu_launches": head["launches"],
            "roofline": {"bound": "hbm", "kernel": "wfe::tc::logmel_tc_kernel<float> (+ wfe::tc::clamp_kernel, timed together)",
                         "achieved": rl["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": rl["frac"],
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": rl["algorithmic_bytes_per_launch"],
                         "kernel_ms_per_launch": rl["kernel_ms_per_launch"],
                         "traffic": (traffic_per_clip * B) if traffic_per_clip else None,
                         "traffic_source": traffic_src},
            "workloads": {k_: {"value": w_["value"], "per_gpu_value": w_["value"] / n_gpus, "ms_per_step": w_["ms_per_step"],
                               "kernel_ms_per_launch": w_["kernel_ms"], "roofline_frac": w_["roofline"]["frac"],
                               "achieved_gbs": w_["roofline"]["achieved_gbs"],
                               **{x: w_[x] for x in ("clips_per_s", "note", "clips", "clips_per_rank", "kernel_timeouts",
                                                     "remeasured_after_kernel_timeout", "max_abs_diff_vs_cuda_core_kernel",
                                                     "wall_s_including_generation_and_checks") if x in w_}}
                          for k_, w_ in workloads.items()},
            "clocks": clocks,
            "per_gpu_value": head["value"] / n_gpus,
            "plumbing": "gloo (barrier / max over ranks only; no collective on the data path)" if world > 1 else "single process",
            "host_pinning": pin,
        }
        if e2e is not None:
            line["e2e"] = e2e
        if ref_loop is not None:
            line["e2e_reference_loop"] = ref_loop
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

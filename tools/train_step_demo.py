#!/usr/bin/env python
"""SURVEY 8(f-2) / BASELINE configs[4]: a Whisper fine-tuning step with the GPU frontend in the loop.

Random-init `WhisperForConditionalGeneration` (large-v3 dims by default, 128 mel) with hand-rolled LoRA on q_proj / v_proj
(the reference uses AdaLoRA, ref:finetune/training/trainers/trainers.py:523-538; `peft` is not in the image), fp16
autocast (ref:finetune/training/configs/largev3_debug.config:8), per-device batch 8.  Each step: collate raw audio ->
forward -> backward -> AdamW on the adapters.  Compares the reference's CPU frontend (transformers extractor per clip +
`.to(cuda)`, ref:.../datasets_and_collators.py:191-195, trainers/utils.py:108-112) with the sm_100a frontend
(`StreamingFrontendCollator`, CUDA tensors straight into the step).

    python tools/train_step_demo.py [--size large-v3|small|tiny] [--batch 8] [--steps 5]
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_step_demo.py --ddp   # BASELINE configs[4]

--ddp: one process per GPU (ref:finetune/training/train_hyper.py:319-329, N Ray workers), the model wrapped in
DistributedDataParallel over NCCL (gradient all-reduce of the adapters: the only collective, and it is the trainer's, not
the frontend's), every rank collating ITS OWN shard of clips (ref:finetune/training/trainers/trainers.py:785-791, 826-828)
with `StreamingFrontendCollator(feature_dtype=float16)` as the collate_fn -- the fp16 cast fused into the kernel's
epilogue.  Every rank checks its in-loop batch against the numpy oracle; rank 0 prints one JSON line with the max over
ranks of the step times.
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn as nn
import transformers as tr
import asr_finetune_b200 as pkg

DIMS = {"large-v3": dict(d_model=1280, layers=32, heads=20, ffn=5120, n_mel=128),
        "small": dict(d_model=768, layers=12, heads=12, ffn=3072, n_mel=80),
        "tiny": dict(d_model=384, layers=4, heads=6, ffn=1536, n_mel=80)}


class LoRALinear(nn.Module):
    def __init__(self, base: nn.Linear, r=8, alpha=16):
        super().__init__()
        self.base, self.scale = base, alpha / r
        for p in base.parameters():
            p.requires_grad_(False)
        self.a = nn.Parameter(torch.randn(r, base.in_features, device=base.weight.device) * 0.01)
        self.b = nn.Parameter(torch.zeros(base.out_features, r, device=base.weight.device))

    def forward(self, x):
        return self.base(x) + (x @ self.a.t().to(x.dtype)) @ self.b.t().to(x.dtype) * self.scale


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="large-v3"); ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--ddp", action="store_true")
    a = ap.parse_args()
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if a.ddp:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank if a.ddp else 0)
    d = DIMS[a.size]
    cfg = tr.WhisperConfig(vocab_size=51866, num_mel_bins=d["n_mel"], d_model=d["d_model"], encoder_layers=d["layers"],
                           decoder_layers=d["layers"], encoder_attention_heads=d["heads"], decoder_attention_heads=d["heads"],
                           encoder_ffn_dim=d["ffn"], decoder_ffn_dim=d["ffn"], decoder_start_token_id=50258,
                           pad_token_id=50257, bos_token_id=50257, eos_token_id=50257)
    torch.manual_seed(0)
    model = tr.WhisperForConditionalGeneration(cfg).to(dev, dtype=torch.float16 if a.size == "large-v3" else torch.float32)
    for p in model.parameters():
        p.requires_grad_(False)
    for m in list(model.modules()):
        for name in ("q_proj", "v_proj"):
            if hasattr(m, name) and isinstance(getattr(m, name), nn.Linear):
                setattr(m, name, LoRALinear(getattr(m, name)))
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4)
    net = model
    if a.ddp:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank])
    rng = np.random.default_rng(1000 + rank)  # every rank draws its own shard of clips
    audio = [(0.1 * rng.standard_normal(int(rng.integers(16000, 480001)))).astype(np.float32) for _ in range(a.batch)]
    labels = [[50258, 50261, 50360, 50364] + rng.integers(0, 50257, int(rng.integers(5, 100))).tolist() + [50257]
              for _ in range(a.batch)]
    ours = pkg.WhisperFeatureExtractor(feature_size=d["n_mel"], cuda_device=dev.index)
    # autocast consumer: features leave the kernel already in fp16 (ref:finetune/training/configs/largev3_debug.config:8)
    gpu_collate = pkg.StreamingFrontendCollator(ours, feature_dtype=torch.float16 if a.ddp else None)
    ref_fe = tr.WhisperFeatureExtractor(feature_size=d["n_mel"])

    def cpu_collate(batch):  # the reference's loop + tokenizer.pad semantics + data_collator_id
        feats = torch.from_numpy(np.stack([ref_fe(x, sampling_rate=16000).input_features[0] for x in batch["audio"]]))
        w = max(len(x) for x in batch["labels"])
        lab = torch.full((len(batch["labels"]), w), -100, dtype=torch.int64)
        for i, x in enumerate(batch["labels"]):
            lab[i, :len(x)] = torch.tensor(x)
        return {"input_features": feats.to(dev), "labels": lab.to(dev)}

    def step(collate):
        t0 = time.perf_counter()
        b = collate({"audio": audio, "labels": labels})
        torch.cuda.synchronize(); t1 = time.perf_counter()
        with torch.autocast("cuda", dtype=torch.float16):
            loss = net(input_features=b["input_features"].to(model.dtype), labels=b["labels"]).loss
        loss.backward(); opt.step(); opt.zero_grad(set_to_none=True)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        return (t1 - t0) * 1e3, (t2 - t1) * 1e3, float(loss)

    res = {}
    for name, coll in (("cpu_frontend", cpu_collate), ("b200_frontend", gpu_collate)):
        step(coll)
        ts = [step(coll) for _ in range(a.steps)]
        res[name] = {"collate_ms": float(np.median([t[0] for t in ts])), "model_ms": float(np.median([t[1] for t in ts])),
                     "loss": ts[-1][2]}
    fa = cpu_collate({"audio": audio, "labels": labels}); fb = gpu_collate({"audio": audio, "labels": labels})
    res["frontends_agree"] = {"features_max_abs_err": float((fa["input_features"] - fb["input_features"].float()).abs().max()),
                              "labels_equal": bool(torch.equal(fa["labels"], fb["labels"]))}
    # the in-loop batch against the numpy oracle (test infrastructure; this tool is a measurement script, not product code)
    from oracle import collate as ocollate, logmel as ologmel
    ref = ologmel.logmel_batch(audio, d["n_mel"], "fp64")
    got = fb["input_features"].float().cpu().numpy()
    tol = 1e-3 + (2e-3 if fb["input_features"].dtype == torch.float16 else 0.0)  # + one fp16 rounding of values in [-1.5, 2]
    res["oracle"] = {"features_max_abs_err": float(np.abs(got - ref).max()), "tolerance": tol,
                     "labels_equal": bool(np.array_equal(fb["labels"].cpu().numpy(),
                                                         ocollate.mask_labels(*ocollate.pad_label_ids(labels, 50257)))),
                     "feature_dtype": str(fb["input_features"].dtype)}
    assert res["oracle"]["features_max_abs_err"] <= tol and res["oracle"]["labels_equal"], res["oracle"]
    res["config"] = {"size": a.size, "batch": a.batch, "steps": a.steps, "lora_params": sum(p.numel() for p in params),
                     "host_cores": len(os.sched_getaffinity(0)), "torch_threads": torch.get_num_threads(),
                     "world_size": world, "ddp": bool(a.ddp)}
    if a.ddp:
        # max over ranks of every timing; all ranks must have passed their oracle check to get here
        keys = [(n, k) for n in ("cpu_frontend", "b200_frontend") for k in ("collate_ms", "model_ms")]
        t = torch.tensor([res[n][k] for n, k in keys] + [res["oracle"]["features_max_abs_err"]], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        for (n, k), v in zip(keys, t.tolist()):
            res[n][k] = v
        res["oracle"]["features_max_abs_err_max_over_ranks"] = t.tolist()[-1]
        res["global_batch"] = a.batch * world
        dist.barrier()
        if rank == 0:
            print(json.dumps(res))
        dist.destroy_process_group()
    else:
        print(json.dumps(res))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""SURVEY 8(f-2) / BASELINE configs[4]: a Whisper fine-tuning step with the GPU frontend in the loop.

Random-init `WhisperForConditionalGeneration` (large-v3 dims by default, 128 mel) with hand-rolled LoRA on q_proj / v_proj
(the reference uses AdaLoRA, ref:finetune/training/trainers/trainers.py:523-538; `peft` is not in the image), fp16
autocast (ref:finetune/training/configs/largev3_debug.config:8), per-device batch 8.  Each step: collate raw audio ->
forward -> backward -> AdamW on the adapters.  Compares the reference's CPU frontend (transformers extractor per clip +
`.to(cuda)`, ref:.../datasets_and_collators.py:191-195, trainers/utils.py:108-112) with the sm_100a frontend
(`StreamingFrontendCollator`, CUDA tensors straight into the step).

    python tools/train_step_demo.py [--size large-v3|small|tiny] [--batch 8] [--steps 5]
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
    a = ap.parse_args()
    d = DIMS[a.size]
    cfg = tr.WhisperConfig(vocab_size=51866, num_mel_bins=d["n_mel"], d_model=d["d_model"], encoder_layers=d["layers"],
                           decoder_layers=d["layers"], encoder_attention_heads=d["heads"], decoder_attention_heads=d["heads"],
                           encoder_ffn_dim=d["ffn"], decoder_ffn_dim=d["ffn"], decoder_start_token_id=50258,
                           pad_token_id=50257, bos_token_id=50257, eos_token_id=50257)
    torch.manual_seed(0)
    model = tr.WhisperForConditionalGeneration(cfg).to("cuda", dtype=torch.float16 if a.size == "large-v3" else torch.float32)
    for p in model.parameters():
        p.requires_grad_(False)
    for m in list(model.modules()):
        for name in ("q_proj", "v_proj"):
            if hasattr(m, name) and isinstance(getattr(m, name), nn.Linear):
                setattr(m, name, LoRALinear(getattr(m, name)))
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4)
    rng = np.random.default_rng(0)
    audio = [(0.1 * rng.standard_normal(int(rng.integers(16000, 480001)))).astype(np.float32) for _ in range(a.batch)]
    labels = [[50258, 50261, 50360, 50364] + rng.integers(0, 50257, int(rng.integers(5, 100))).tolist() + [50257]
              for _ in range(a.batch)]
    ours = pkg.WhisperFeatureExtractor(feature_size=d["n_mel"])
    gpu_collate = pkg.StreamingFrontendCollator(ours)
    ref_fe = tr.WhisperFeatureExtractor(feature_size=d["n_mel"])

    def cpu_collate(batch):  # the reference's loop + tokenizer.pad semantics + data_collator_id
        feats = torch.from_numpy(np.stack([ref_fe(x, sampling_rate=16000).input_features[0] for x in batch["audio"]]))
        w = max(len(x) for x in batch["labels"])
        lab = torch.full((len(batch["labels"]), w), -100, dtype=torch.int64)
        for i, x in enumerate(batch["labels"]):
            lab[i, :len(x)] = torch.tensor(x)
        return {"input_features": feats.to("cuda:0"), "labels": lab.to("cuda:0")}

    def step(collate):
        t0 = time.perf_counter()
        b = collate({"audio": audio, "labels": labels})
        torch.cuda.synchronize(); t1 = time.perf_counter()
        with torch.autocast("cuda", dtype=torch.float16):
            loss = model(input_features=b["input_features"].to(model.dtype), labels=b["labels"]).loss
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
    res["frontends_agree"] = {"features_max_abs_err": float((fa["input_features"] - fb["input_features"]).abs().max()),
                              "labels_equal": bool(torch.equal(fa["labels"], fb["labels"]))}
    res["config"] = {"size": a.size, "batch": a.batch, "steps": a.steps, "lora_params": sum(p.numel() for p in params),
                     "host_cores": len(os.sched_getaffinity(0)), "torch_threads": torch.get_num_threads()}
    print(json.dumps(res))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Small ragged batch through the C ABI for compute-sanitizer (memcheck / racecheck / synccheck):
    python tools/sanitize_case.py [B]      (compute-sanitizer is closed on the shared pool; this is the case to run under it elsewhere)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import asr_finetune_b200 as pkg
from oracle import logmel as ologmel, signals
B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
lens = [480000, 3, 70001, 16000, 479999, 250000, 1, 123457][:B]
clips = [signals.noise(10 + i, n, amp=0.1 * 10.0 ** (-(i % 3))) for i, n in enumerate(lens)]
fe = pkg.WhisperFeatureExtractor(feature_size=128)
dev = fe.cuda_device()
pcm = torch.from_numpy(np.concatenate(clips)).to(dev)
offs = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).to(dev)
feats, mask = fe.logmel_device(pcm, offs, B, return_attention_mask=True)
torch.cuda.synchronize()
err = max(float(np.abs(feats[i].cpu().numpy() - ologmel.logmel_clip(clips[i], 128, "fp64")).max()) for i in range(B))
q = (np.clip(np.round(clips[0].astype(np.float64) * 32768.0), -32768, 32767)).astype(np.int16)
f16, _ = fe.logmel_device(torch.from_numpy(q).to(dev), torch.tensor([0, len(q)], device=dev), 1, pcm_scale=1.0 / 32768.0,
                          do_normalize=True)
labels = signals.label_ids(5, B, 5, 40)
coll = pkg.DataCollatorSpeechSeq2SeqWithPadding(processor=type("P", (), {"feature_extractor": fe})(), decoder_start_token_id=signals.SOT)
out = coll({"input_features": [f for f in feats], "labels": labels})
torch.cuda.synchronize()
print(f"ok: max-abs-err {err:.2e}, labels {tuple(out['labels'].shape)}, int16+normalize finite {bool(torch.isfinite(f16).all())}")

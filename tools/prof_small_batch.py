#!/usr/bin/env python
"""cProfile of one batch-8 collate (host-side overheads of the in-loop consumer): python tools/prof_small_batch.py"""
import os, sys, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import asr_finetune_b200 as pkg
from oracle import signals
B = 8
fe = pkg.WhisperFeatureExtractor(feature_size=128)
dev = fe.cuda_device()
rng = np.random.default_rng(8)
clips = [np.array(0.1 * rng.standard_normal(int(n)), dtype=np.float32) for n in rng.integers(3 * 16000, 480001, size=B)]
labels = signals.label_ids(1337, B, 5, 60)
coll = pkg.StreamingFrontendCollator(fe, device=dev, feature_dtype=torch.float16)
for _ in range(5): coll({"audio": clips, "labels": labels})
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(200): coll({"audio": clips, "labels": labels})
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])

#!/usr/bin/env python
"""Where does the ragged (config 3) case differ from the oracle?  python tools/diag_ragged.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import asr_finetune_b200 as pkg
from oracle import logmel as ologmel, signals
B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
lens = signals.clip_lengths(1337, B)
clips = [signals.noise(1000 + i, int(n)) if i % 3 else signals.speechlike(i, int(n)) for i, n in enumerate(lens)]
fe = pkg.WhisperFeatureExtractor(feature_size=128)
dev = fe.cuda_device()
pcm = torch.from_numpy(np.concatenate(clips)).to(dev)
offs = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).to(dev)
for rep in range(3):
    feats, mask = fe.logmel_device(pcm, offs, B, return_attention_mask=True)
    f = feats.cpu().numpy()
    for i in range(B):
        ref = ologmel.logmel_clip(clips[i], 128, "fp64")
        d = np.abs(f[i] - ref)
        if d.max() > 1e-3:
            bad = np.argwhere(d > 1e-3)
            tiles = sorted(set((bad[:, 1] // 32).tolist()))
            fr = bad[:, 1]
            print(f"rep {rep} clip {i} len {lens[i]} ({lens[i]/160:.1f} frames, off%4={int(offs[i])%4}) max err {d.max():.4f} nbad {len(bad)} "
                  f"tiles {tiles[:12]}{'...' if len(tiles)>12 else ''} mels {bad[:,0].min()}..{bad[:,0].max()} "
                  f"ours[min,max]=({f[i].min():.4f},{f[i].max():.4f}) ref=({ref.min():.4f},{ref.max():.4f})")
            m, t = bad[0]
            print(f"    first bad (mel {m}, frame {t}): ours {f[i][m,t]:.5f} ref {ref[m,t]:.5f}; ours row slice {f[i][m, t:t+4]} ref {ref[m, t:t+4]}")
print("done")

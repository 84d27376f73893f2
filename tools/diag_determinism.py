"""Diagnostic: which elements differ between two launches / violate the silent-frame invariant (ragged stress batch)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import asr_finetune_b200 as pkg
from oracle import signals

fe = pkg.WhisperFeatureExtractor(feature_size=128)
dev = fe.cuda_device()
B = 192
rng = np.random.default_rng(7)
lens = rng.integers(1, 6 * 16000, size=B)
lens[::7] = rng.integers(16000, 480001, size=len(lens[::7]))
clips = [signals.noise(500 + i, int(n), amp=0.1 * 10.0 ** (-(i % 5))) for i, n in enumerate(lens)]
pcm = torch.from_numpy(np.concatenate(clips)).to(dev)
offs = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)).to(dev)
outs = []
for rep in range(4):
    feats, mask = fe.logmel_device(pcm, offs, B, return_attention_mask=True)
    outs.append(feats.clone())
    print("launch", rep, "kernel err", hex(fe.debug_kernel_error()))
for rep in range(1, 4):
    d = (outs[rep] != outs[0])
    print(f"launch {rep}: {int(d.sum())} differing elements")
    if d.any():
        idx = d.nonzero()[:2000].cpu().numpy()
        clips_bad = np.unique(idx[:, 0])
        for b in clips_bad[:8]:
            fr = np.unique(idx[idx[:, 0] == b][:, 2])
            ml = np.unique(idx[idx[:, 0] == b][:, 1])
            print(f"   clip {b} len {lens[b]} (off {int(offs[b])}, off%4={int(offs[b]) % 4}) frames {fr.min()}..{fr.max()} ({len(fr)} distinct; tiles {np.unique(fr // 128)}) mels {ml.min()}..{ml.max()};"
                  f" values {outs[0][b, ml[0], fr[0]].item():.5f} vs {outs[rep][b, ml[0], fr[0]].item():.5f}")
first = outs[0]
gmax = first.amax(dim=(1, 2))
floor = torch.maximum(gmax - 2.0, torch.full_like(gmax, -1.5))
bad = first.amin(dim=(1, 2)) < gmax - 2.0
print("clips whose minimum is below the floor:", bad.nonzero().flatten().tolist()[:20])
t_sil = torch.from_numpy((lens + 200) // 160 + 1).to(dev)
silent = torch.arange(3000, device=dev)[None, :] >= t_sil[:, None]
viol = (first != floor[:, None, None]) & silent[:, None, :]
print("silent-frame violations:", int(viol.sum()))
if viol.any():
    idx = viol.nonzero()[:4000].cpu().numpy()
    for b in np.unique(idx[:, 0])[:8]:
        fr = np.unique(idx[idx[:, 0] == b][:, 2])
        print(f"   clip {b} len {lens[b]} t_sil {int(t_sil[b])} frames {fr.min()}..{fr.max()} tiles {np.unique(fr // 128)} value {first[b, 0, fr[0]].item():.5f} floor {floor[b].item():.5f}")

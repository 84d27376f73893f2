"""Back-to-back launches (diagnostic): python tools/tc_b2b.py B n_launches [sync_each]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import asr_finetune_b200 as pkg
B, n, sync_each = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 0
fe = pkg.WhisperFeatureExtractor(feature_size=128)
dev = fe.cuda_device()
pcm = 0.1 * torch.randn(B * 480000, device=dev)
offs = torch.arange(B + 1, dtype=torch.int64, device=dev) * 480000
out = torch.empty((B, 128, 3000), dtype=torch.float32, device=dev)
for i in range(n):
    fe.logmel_device(pcm, offs, B, out=out)
    if sync_each:
        torch.cuda.synchronize()
        print(f"launch {i}: err {fe.debug_kernel_error():#x}", flush=True)
try:
    torch.cuda.synchronize()
    print(f"B={B} n={n} sync_each={sync_each}: ok err {fe.debug_kernel_error():#x}", flush=True)
except Exception as e:
    print(f"B={B} n={n} sync_each={sync_each}: FAILED {str(e).splitlines()[0]}", flush=True)

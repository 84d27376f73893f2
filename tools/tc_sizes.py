"""One launch per batch size with the error word printed (diagnostic)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import asr_finetune_b200 as pkg
fe = pkg.WhisperFeatureExtractor(feature_size=128)
dev = fe.cuda_device()
for B in [int(a) for a in sys.argv[1:]] or [64, 128, 256]:
    pcm = 0.1 * torch.randn(B * 480000, device=dev)
    offs = torch.arange(B + 1, dtype=torch.int64, device=dev) * 480000
    out = torch.empty((B, 128, 3000), dtype=torch.float32, device=dev)
    t0 = time.time()
    fe.logmel_device(pcm, offs, B, out=out)
    torch.cuda.synchronize()
    print(f"B={B}: {time.time() - t0:.3f} s  err {fe.debug_kernel_error():#x}  finite {bool(torch.isfinite(out).all())}", flush=True)

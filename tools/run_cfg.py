#!/usr/bin/env python
"""Device-resident pipeline timing of ONE named BASELINE shape (profiling target).  usage: run_cfg.py cfg3 [iters] [tile]"""
import hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "corpus")):
    sys.path.insert(0, p)
import torch
import pycorpus
from birdnest.audio_b200 import _abi

CFG = {
    "cfg1": dict(ch=2, bps=16, sr=44100, seconds=60, bs=4096, lpc=8, maxpo=5, tile=60, seed=2026),
    "cfg2": dict(ch=2, bps=24, sr=96000, seconds=60, bs=4096, lpc=12, maxpo=6, stereo=1, search=1, tile=60, seed=2026),
    "cfg3": dict(ch=8, bps=24, sr=192000, samples=16384 * 12, bs=16384, lpc=32, minpo=8, maxpo=8, noise=19, search=0, tile=500, seed=5),
    "mono": dict(ch=1, bps=16, sr=48000, seconds=60, bs=1152, lpc=0, tile=120, seed=7),
    "special": dict(ch=6, bps=24, sr=48000, seconds=30, bs=1152, lpc=8, kind=1, period=1152, tile=120, seed=8),
}
name = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
kw = dict(CFG[name])
if len(sys.argv) > 3: kw["tile"] = int(sys.argv[3])
dev = torch.device("cuda:0")
s = pycorpus.make(**kw)
n_all = s.total_samples * s.channels
B = (s.bps + 7) // 8
d_in = torch.zeros(len(s.flac) + 256, dtype=torch.uint8, device=dev)
d_in[:len(s.flac)] = torch.frombuffer(bytearray(s.flac), dtype=torch.uint8).to(dev)
d_out = torch.empty(n_all * B + 256, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
with _abi.open_device(d_in.data_ptr(), len(s.flac), s.flac[:1 << 20], device=0, stream=torch.cuda.current_stream().cuda_stream, keep=d_in) as h:
    _, w = h.decode_device(d_out.data_ptr(), d_out.numel())
    ok = hashlib.md5(d_out[:w].cpu().numpy().tobytes()).digest() == s.md5
    best = None
    for _ in range(iters):
        h.decode_device(d_out.data_ptr(), d_out.numel()); t = h.timing()
        if best is None or t.total < best.total: best = t
t = best
alg = len(s.flac) + n_all * B
print(f"{name} {n_all/1e6:.1f} M samples {len(s.flac)/1e6:.1f} MB in {t.total:.3f} ms {n_all/t.total/1e6:.1f} G samples/s {alg/t.total/1e6:.1f} GB/s alg "
      f"[scan {t.scan:.3f} crc {t.crc:.3f} parse {t.parse:.3f} decode {t.decode:.3f}] md5 {'ok' if ok else 'MISMATCH'}", flush=True)
sys.exit(0 if ok else 1)

#!/usr/bin/env python
"""One device-resident GPU encode of a BASELINE-shaped PCM buffer (profiling target for ncu).  usage: enc_profile.py cfg2|cfg3 [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "corpus")):
    sys.path.insert(0, p)
import torch
import pycorpus
from birdnest.audio_b200 import _abi

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
if name == "cfg2":
    s = pycorpus.make(ch=2, bps=24, sr=96000, seconds=60, bs=4096, lpc=12, maxpo=6, stereo=1, search=1, seed=2026)
    enc, reps = dict(blocksize=4096, max_lpc_order=12, max_partition_order=6), 8
else:
    s = pycorpus.make(ch=8, bps=24, sr=192000, samples=16384 * 12, bs=16384, lpc=32, minpo=8, maxpo=8, noise=19, search=0, seed=5)
    enc, reps = dict(blocksize=16384, max_lpc_order=32, min_partition_order=8, max_partition_order=8, flags=_abi.ENC_FIXED_ORDER), 100
if len(sys.argv) > 2:
    reps = int(sys.argv[2])
dev = torch.device("cuda:0")
d_pcm = torch.frombuffer(bytearray(s.pcm), dtype=torch.uint8).to(dev).repeat(reps)
n = d_pcm.numel()
o = _abi.enc_opts(s.sample_rate, s.channels, s.bps, flags=_abi.ENC_NO_MD5 | enc.pop("flags", 0), device=0, **enc)
cap = _abi.encode_bound(n, o)
d_flac = torch.zeros(cap + 256, dtype=torch.uint8, device=dev)
best = None
for _ in range(3):
    w, st = _abi.encode_device(d_pcm.data_ptr(), n, o, d_flac.data_ptr(), cap)
    if best is None or st.total_ms < best.total_ms:
        best = st
B = (s.bps + 7) // 8
print(f"{name} encode: {n / B / 1e6:.1f} M samples, {n / 1e6:.0f} MB PCM -> {w / 1e6:.0f} MB ({w / n:.3f}), {best.frames} frames, "
      f"{best.total_ms:.3f} ms (plan {best.plan_ms:.3f}, write {best.write_ms:.3f}) = {n / B / best.total_ms / 1e6:.2f} G samples/s", flush=True)

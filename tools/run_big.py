#!/usr/bin/env python
"""Device-resident pass over a bench.py configuration that may exceed 2 GiB (run_cfg.py holds the stream as Python bytes): the same
corpus view and device_pass as bench.py's config_entry.  usage: run_big.py cfg3 [tiles] [steps]   (library via BNFLAC_LIB)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "corpus")):
    sys.path.insert(0, p)
import torch
import bench, pycorpus
from birdnest.audio_b200 import _abi
name = sys.argv[1]
kw = {"cfg1": bench.cfg1_kwargs, "cfg3": bench.cfg3_kwargs}[name](1.0)
if len(sys.argv) > 2: kw["tile"] = int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
ctx = (torch, _abi, dev, 0, torch.cuda.Stream(device=dev))
s = pycorpus.make(md5=False, view=True, **kw)
B = (s.bps + 7) // 8; n_all = s.total_samples * s.channels
r = bench.device_pass(ctx, s.flac, n_all * B, steps, 2)
tile_dev = torch.frombuffer(bytearray(s.pcm), dtype=torch.uint8).to(dev)
ok = r["written"] == n_all * B and bench.verify_periodic(torch, r["d_out"], r["written"], tile_dev, 0)
print(f"{name} tile {kw['tile']} frames {len(s.frame_bs)} {len(s.flac)/1e6:.1f} MB: {r['ms']:.3f} ms {n_all/r['ms']/1e6:.1f} G samples/s {r['stage_ms']} pcm {'ok' if ok else 'MISMATCH'}", flush=True)

#!/usr/bin/env python
"""cfg4 share (12,500 clips, host memory -> device PCM) through bnflac_decode_batch; BNFLAC_BATCH_RUNS / BNFLAC_TRACE from the environment."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "corpus")):
    sys.path.insert(0, p)
import torch
import bench
from birdnest.audio_b200 import _abi
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
ctx = (torch, _abi, dev, 0, torch.cuda.Stream(device=dev))
pool = bench.cfg4_pool()
e = bench.batch_entry(ctx, pool, range(0, 100000, 8), 5)
print(os.environ.get("BNFLAC_BATCH_RUNS", "1"), "runs:", e["ms_per_step"], "ms", round(e["samples_per_s"] / 1e9, 2), "G samples/s")

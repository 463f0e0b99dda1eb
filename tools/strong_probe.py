#!/usr/bin/env python
"""One rank's share of the strong job (bench.py corpus_job) on ONE GPU: shard (r, n) of every cfg5 stream.  usage: strong_probe.py [scale] [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "corpus")):
    sys.path.insert(0, p)
import torch
import pycorpus
import bench
from birdnest.audio_b200 import _abi
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda:0")
ctx = (torch, _abi, dev, 0, torch.cuda.current_stream())
kept = [pycorpus.make(md5=False, view=True, **kw) for _, _, kw in bench.cfg5_formats(scale)]
info = [(len(g.flac), g.total_samples * g.channels, (g.bps + 7) // 8, len(g.pcm)) for g in kept]
maps = [g.flac for g in kept]
for shard in (None, (0, n), (n // 2, n)):
    ms, outs, wr = bench.corpus_job(ctx, maps, info, shard, 5)
    del outs
    torch.cuda.empty_cache()
    print(f"shard {shard}: {ms:.3f} ms", flush=True)

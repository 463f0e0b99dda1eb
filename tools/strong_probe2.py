#!/usr/bin/env python
"""Where one rank's strong-job step goes (one GPU, shard r of n of every cfg5 stream): the 8-channel pass alone, the other five alone, all six."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "corpus")):
    sys.path.insert(0, p)
import torch
import pycorpus
import bench
from birdnest.audio_b200 import _abi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
ctx = (torch, _abi, dev, 0, torch.cuda.current_stream())
fm = bench.cfg5_formats(1.0)
kept = [pycorpus.make(md5=False, view=True, **kw) for _, _, kw in fm]
info = [(len(g.flac), g.total_samples * g.channels, (g.bps + 7) // 8, len(g.pcm)) for g in kept]
maps = [g.flac for g in kept]
big = max(range(len(kept)), key=lambda i: info[i][1])
def run(sel, label):
    ms, outs, wr = bench.corpus_job(ctx, [maps[i] for i in sel], [info[i] for i in sel], (0, n), 5)
    del outs; torch.cuda.empty_cache()
    print(f"{label}: {ms:.3f} ms", flush=True)
run([big], "8-channel shard alone")
run([i for i in range(len(kept)) if i != big], "the other five")
for i in range(len(kept)):
    if i != big: run([i], f"  alone {fm[i][0]}")
run(list(range(len(kept))), "all six")

#!/usr/bin/env python
"""End-to-end (host FLAC bytes -> host PCM) time of the cfg2 stream for several pipeline sub-shard sizes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "corpus")):
    sys.path.insert(0, p)
import torch
import bench
from birdnest.audio_b200 import _abi
secs = int(sys.argv[1]) if len(sys.argv) > 1 else 3600
s = bench.make_stream(secs)
host_in = torch.frombuffer(bytearray(s.flac), dtype=torch.uint8).pin_memory()
host_out = torch.empty(s.total_samples * s.channels * 3 + 256, dtype=torch.uint8).pin_memory()
for mb, first, growth in ((256, 16, 200), (256, 16, 170), (256, 16, 150), (256, 8, 170), (256, 8, 150), (256, 4, 170), (192, 8, 170), (128, 8, 150), (256, 8, 200), (256, 32, 150)):
    os.environ["BNFLAC_PIPE_MB"] = str(mb); os.environ["BNFLAC_PIPE_FIRST_MB"] = str(first); os.environ["BNFLAC_PIPE_GROWTH"] = str(growth)
    ts = []
    for it in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h = _abi.open_memory(host_in, device=0, flags=_abi.OPT_BORROW_INPUT)
        t1 = time.perf_counter()
        n = h.decode_all(host_out)
        t2 = time.perf_counter()
        tm = h.timing()
        h.close()
        t3 = time.perf_counter()
        ts.append((t3 - t0) * 1e3)
        if it == 4: print(f"   open {(t1-t0)*1e3:.2f} decode_all {(t2-t1)*1e3:.2f} close {(t3-t2)*1e3:.2f}")
    print(f"PIPE_MB={mb:6d} FIRST={first:3d} GROWTH={growth:3d}  e2e ms {min(ts[1:]):8.2f} (first {ts[0]:8.2f})  engine total {tm.total:7.2f} launches {tm.launches}  bytes {n}", flush=True)
import hashlib
print("md5 ok", hashlib.md5(host_out[:n].numpy().tobytes()).digest() == s.md5)

#!/usr/bin/env python
"""One traced end-to-end decode of the cfg2 stream (BNFLAC_TRACE=1 prints the sub-shard timeline) + raw PCIe copy rates."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "corpus")):
    sys.path.insert(0, p)
import torch
import bench
from birdnest.audio_b200 import _abi
s = bench.make_stream(3600)
bench.bind_to_gpu_cpus(0)
host_in = torch.frombuffer(bytearray(s.flac), dtype=torch.uint8).pin_memory()
host_out = torch.empty(s.total_samples * s.channels * 3 + 256, dtype=torch.uint8).pin_memory()
d = torch.empty(host_out.numel(), dtype=torch.uint8, device="cuda")
for name, src, dst in (("H2D", host_in, d[:host_in.numel()]), ("D2H", d, host_out)):
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); dst.copy_(src, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"raw {name}: {src.numel()/dt/1e9:.1f} GB/s ({dt*1e3:.1f} ms)")
for it in range(4):
    if it == 3: os.environ["BNFLAC_TRACE"] = "1"
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with _abi.open_memory(host_in, device=0, flags=_abi.OPT_BORROW_INPUT) as h:
        n = h.decode_all(host_out)
    print(f"e2e {1e3*(time.perf_counter()-t0):.2f} ms", flush=True)

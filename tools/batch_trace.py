#!/usr/bin/env python
"""cfg4 shape: 20,000 short clips through bnflac_decode_batch, C call timed alone (spans prebuilt) and through the Python wrapper."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "corpus")):
    sys.path.insert(0, p)
import torch
import pycorpus
from birdnest.audio_b200 import _abi
pool = []
for i in range(200):
    ch = 1 + (i & 1)
    kw = dict(ch=ch, bps=16, sr=44100, seconds=0.5 + (i * 37 % 26) / 10.0, lpc=0 if i % 4 < 2 else 8, seed=1000 + i)
    if i % 10 == 3: kw["var"] = (4096, 1152, 4080, 720, 16, 192, 2304)
    else: kw["bs"] = (576, 1152, 2304, 4096, 4608)[i % 5]
    pool.append(pycorpus.make(**kw))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
clips = [pool[(7 * k) % 200] for k in range(N)]
blobs = [c.flac for c in clips]
n_all = sum(c.total_samples * c.channels for c in clips)
out = torch.empty(n_all * 2 + 256, dtype=torch.uint8, device="cuda")
spans = (_abi.Span * N)()
for i, b in enumerate(blobs):
    spans[i].data = _abi._addr(b); spans[i].len = len(b)
res = (_abi.ClipResult * N)()
o = _abi._opts(0); w = C.c_uint64()
L = _abi.lib()
for it in range(4):
    if it == 3: os.environ["BNFLAC_TRACE"] = "1"
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rc = L.bnflac_decode_batch(spans, N, C.byref(o), out.data_ptr(), out.numel(), 1, res, C.byref(w))
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    print(f"C call: rc {rc} {dt:.1f} ms = {n_all/dt/1e6:.1f} G samples/s ({sum(len(b) for b in blobs)/1e6:.0f} MB in, {w.value/1e6:.0f} MB PCM)", flush=True)
# the same clips packed into ONE pinned buffer (a shard file read in one piece): no gather, one upload in place
total = sum(len(b) + 512 for b in blobs)
pk = torch.empty(total, dtype=torch.uint8).pin_memory()
off = 0
pspans = (_abi.Span * N)()
import numpy as np
pkn = pk.numpy()
for i, b in enumerate(blobs):
    off += 512
    pkn[off:off + len(b)] = np.frombuffer(b, dtype=np.uint8)
    pspans[i].data = pk.data_ptr() + off; pspans[i].len = len(b)
    off += len(b)
po = _abi._opts(0, flags=_abi.OPT_PACKED_INPUT)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rc = L.bnflac_decode_batch(pspans, N, C.byref(po), out.data_ptr(), out.numel(), 1, res, C.byref(w))
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    print(f"C call, packed pinned input: rc {rc} {dt:.1f} ms = {n_all/dt/1e6:.1f} G samples/s", flush=True)
os.environ.pop("BNFLAC_TRACE", None)
t0 = time.perf_counter(); n, r = _abi.decode_batch(blobs, device=0, dst=out, dst_is_device=True); torch.cuda.synchronize()
print(f"python wrapper: {(time.perf_counter()-t0)*1e3:.1f} ms")
import hashlib
bad = sum(1 for k in range(0, N, 997) if hashlib.md5(out[res[k].pcm_offset:res[k].pcm_offset + res[k].pcm_bytes].cpu().numpy().tobytes()).digest() != clips[k].md5)
print("spot md5 mismatches:", bad)

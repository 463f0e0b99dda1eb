#!/usr/bin/env python
"""Device-resident pipeline timing of the BASELINE.json shapes that are not the bench line (cfg1, cfg3, cfg4, cfg5 parts).
Prints samples/s, stage times and md5 verification for each.  Usage: python tools/other_configs.py"""
import hashlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "corpus")):
    sys.path.insert(0, p)
import torch
import pycorpus
from birdnest.audio_b200 import _abi

dev = torch.device("cuda:0")

def run_stream(name, **kw):
    s = pycorpus.make(**kw)
    n_all = s.total_samples * s.channels
    B = (s.bps + 7) // 8
    d_in = torch.zeros(len(s.flac) + 256, dtype=torch.uint8, device=dev)
    d_in[:len(s.flac)] = torch.frombuffer(bytearray(s.flac), dtype=torch.uint8).to(dev)
    d_out = torch.empty(n_all * B + 256, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    with _abi.open_device(d_in.data_ptr(), len(s.flac), s.flac[:1 << 20], device=0, stream=torch.cuda.current_stream().cuda_stream, keep=d_in) as h:
        for _ in range(3):
            _, w = h.decode_device(d_out.data_ptr(), d_out.numel())
        ok = hashlib.md5(d_out[:w].cpu().numpy().tobytes()).digest() == s.md5
        ts = []
        for _ in range(5):
            h.decode_device(d_out.data_ptr(), d_out.numel()); t = h.timing(); ts.append(t.total)
        t = h.timing()
    ms = min(ts)
    alg = len(s.flac) + n_all * B
    print(f"{name:34s} {n_all/1e6:9.1f} M samples  {len(s.flac)/1e6:8.1f} MB in  {ms:7.3f} ms  {n_all/ms/1e6:7.1f} G samples/s  {alg/ms/1e6:7.1f} GB/s alg  "
          f"[scan {t.scan:.3f} crc {t.crc:.3f} parse {t.parse:.3f} decode {t.decode:.3f}] md5 {'ok' if ok else 'MISMATCH'}", flush=True)

run_stream("cfg1 16b stereo 44.1k lpc8 (1 h)", ch=2, bps=16, sr=44100, seconds=60, bs=4096, lpc=8, maxpo=5, tile=60, seed=2026)
run_stream("cfg2 24b stereo 96k lpc12 (1 h)", ch=2, bps=24, sr=96000, seconds=60, bs=4096, lpc=12, maxpo=6, stereo=1, search=1, tile=60, seed=2026)
run_stream("cfg3 24b 8ch 192k lpc32 po8 (512s)", ch=8, bps=24, sr=192000, samples=16384 * 12, bs=16384, lpc=32, minpo=8, maxpo=8, noise=19, search=0, tile=500, seed=5)
run_stream("cfg5 16b mono 48k fixed bs1152 (2h)", ch=1, bps=16, sr=48000, seconds=60, bs=1152, lpc=0, tile=120, seed=7)
run_stream("cfg5 24b 6ch 48k special (1 h)", ch=6, bps=24, sr=48000, seconds=30, bs=1152, lpc=8, kind=1, period=1152, tile=120, seed=8)

# cfg4: batch of short clips (pool of 200 unique clips tiled to 20,000)
pool = []
for i in range(200):
    ch = 1 + (i & 1)
    kw = dict(ch=ch, bps=16, sr=44100, seconds=0.5 + (i * 37 % 26) / 10.0, lpc=0 if i % 4 < 2 else 8, seed=1000 + i)
    if i % 10 == 3: kw["var"] = (4096, 1152, 4080, 720, 16, 192, 2304)
    else: kw["bs"] = (576, 1152, 2304, 4096, 4608)[i % 5]
    pool.append(pycorpus.make(**kw))
clips = [pool[(7 * k) % 200] for k in range(20000)]
blobs = [c.flac for c in clips]
n_all = sum(c.total_samples * c.channels for c in clips)
out = torch.empty(n_all * 2 + 256, dtype=torch.uint8, device=dev)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n, res = _abi.decode_batch(blobs, device=0, dst=out, dst_is_device=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
bad = sum(r.status != 0 for r in res)
print(f"cfg4 batch of 20,000 clips: {n_all/1e6:.1f} M samples, {sum(len(b) for b in blobs)/1e6:.1f} MB in, host->device-PCM wall {dt:.1f} ms = {n_all/dt/1e6:.1f} G samples/s (includes staging + H2D), damaged {bad}")

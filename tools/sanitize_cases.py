#!/usr/bin/env python
"""Small invocations of every kernel path for `compute-sanitizer --tool memcheck` (one tool per gpurun call):
intact + damaged streams, 8-channel LPC-32 (k_decode<32>, predicated refill), generic-format kernel, VERBATIM/escape
lanes, batch (gathered and packed), streaming Read with tiny sub-shards, frame-range shards.  Exits non-zero on mismatch."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "corpus")):
    sys.path.insert(0, p)
os.environ["BNFLAC_READ_FIRST_KB"] = "16"; os.environ["BNFLAC_READ_MB"] = "1"
import pycorpus, pyoracle
from birdnest.audio_b200 import _abi
mk = pycorpus.make
streams = {
    "stereo24": mk(ch=2, bps=24, sr=96000, seconds=0.5, bs=4096, lpc=12, maxpo=6),
    "ch8_lpc32": mk(ch=8, bps=24, sr=192000, samples=16384 * 2, bs=16384, lpc=32, minpo=8, maxpo=8, noise=19, search=0),
    "special6": mk(ch=6, bps=24, sr=48000, seconds=0.3, bs=1152, lpc=8, kind=1, period=1152),
    "esc_verb": mk(ch=1, bps=16, sr=44100, seconds=0.5, bs=576, lpc=0, kind=1, period=576, esc=5, verb=7),
    "ch3_8bit": mk(ch=3, bps=8, sr=8000, seconds=1, bs=256, lpc=4, noise=3),
    "var16": mk(ch=2, bps=16, sr=44100, seconds=0.5, lpc=8, var=(4096, 1152, 4080, 720, 16, 192, 2304)),
}
bad = 0
rng = random.Random(1)
for name, s in streams.items():
    for lean in ("0", "1"):
        os.environ["BNFLAC_PARSE_LEAN"] = lean
        for damage in (0, 1):
            b = bytearray(s.flac)
            if damage:
                for _ in range(3): b[rng.randrange(s.frame_off[0], len(b))] ^= 1 << rng.randrange(8)
                del b[len(b) - 37:]
            b = bytes(b)
            want, _, _, oerrs = pyoracle.decode(b)
            with _abi.open_memory(b) as h:
                out = bytearray(len(s.pcm) + (1 << 16)); k = h.decode_all(out); errs = h.errors()
            ok = bytes(out[:k]) == want and errs == oerrs
            bad += not ok
            print(name, "lean", lean, "damaged" if damage else "intact", "ok" if ok else "MISMATCH", flush=True)
os.environ.pop("BNFLAC_PARSE_LEAN")
s = streams["stereo24"]
with _abi.open_memory(s.flac) as h:                       # streaming Read, many tiny sub-shards
    got = bytearray(); buf = bytearray(50000)
    while True:
        n = h.read_into(buf); got += buf[:n]
        if n < len(buf): break
ok = bytes(got) == s.pcm; bad += not ok; print("streamed read", "ok" if ok else "MISMATCH")
parts = []
for i in range(3):                                        # frame-range shards
    with _abi.open_memory(s.flac, shard_index=i, shard_count=3) as h:
        out = bytearray(len(s.pcm) + 64); parts.append(bytes(out[:h.decode_all(out)]))
ok = b"".join(parts) == s.pcm; bad += not ok; print("shards", "ok" if ok else "MISMATCH")
clips = [streams["var16"].flac, streams["esc_verb"].flac, streams["stereo24"].flac, b"junk", streams["var16"].flac]
pcm, res = _abi.decode_batch(clips)
packed = bytearray()
spans = []
for c in clips:
    packed += b"\xff\xf8" * 100; spans.append((len(packed), len(c))); packed += c
mv = memoryview(packed)
pcm2, res2 = _abi.decode_batch([mv[o:o + n] for o, n in spans], packed=True)
ok = pcm == pcm2 and [r.status for r in res] == [r.status for r in res2] == [0, 0, 0, 4, 0]
bad += not ok; print("batch gathered/packed", "ok" if ok else "MISMATCH")
print("FAILED" if bad else "all ok")
sys.exit(1 if bad else 0)

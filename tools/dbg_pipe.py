import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "corpus")):
    sys.path.insert(0, p)
import pycorpus, pyoracle
from birdnest.audio_b200 import _abi
s = pycorpus.make(ch=2, bps=24, sr=96000, seconds=4, bs=4096, lpc=12, maxpo=6, tile=3, seed=7)
b = bytearray(s.flac); n = len(b)
for pos, mask in ((n // 5, 0x10), (n // 2 + 3, 0xFF), (n - 40000, 0x01)):
    b[pos] ^= mask
b = bytes(b)
want, oframes, _, oerrs = pyoracle.decode(b, want_frames=True)
print("oracle", len(want), len(oframes), oerrs)
for mb in ("4096", "1"):
    os.environ["BNFLAC_PIPE_MB"] = mb
    with _abi.open_memory(b) as h:
        out = bytearray(len(s.pcm) * s.tiles + 64)
        k = h.decode_all(out)
        fr = h.frames(); er = h.errors()
    print("PIPE", mb, k, len(fr), er, bytes(out[:k]) == want)
    of = {o.offset: o for o in oframes}
    gf = {f.offset: f for f in fr}
    for off in sorted(set(of) ^ set(gf)):
        print("  only in", "oracle" if off in of else "gpu", off, (of.get(off) or gf.get(off)).length, (of.get(off) or gf.get(off)).number)
    for off in sorted(set(of) & set(gf)):
        if (of[off].length, of[off].status) != (gf[off].length, gf[off].status):
            print("  differ", off, of[off].length, of[off].status, gf[off].length, gf[off].status)

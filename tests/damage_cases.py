"""Seeded damaged streams shared by the CPU and GPU suites (test infrastructure).

SHAPES + damage() generate the streams; golden_damage_records() pairs the first `trials` of every (shape, kind) with what the
REFERENCE's LibFlac.dll produced for them (tests/golden/golden_damage.json, written by oracle/fuzz_vs_ref.py --write-golden)."""
import hashlib
import json
import os
import random
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SHAPES = {
    "s16_lpc8": dict(ch=2, bps=16, sr=44100, seconds=1.5, bs=1152, lpc=8, seed=21),
    "s16_fixed_small_k": dict(ch=1, bps=16, sr=44100, seconds=1.5, bs=576, lpc=0, noise=3, seed=22),
    "s24_lpc12": dict(ch=2, bps=24, sr=96000, seconds=0.7, bs=4096, lpc=12, maxpo=6, seed=23),
    "s16_var": dict(ch=2, bps=16, sr=44100, seconds=1.5, lpc=8, var=(4096, 1152, 4080, 720, 16, 192, 2304), seed=24),
}
KINDS = ["flip", "run", "ones", "zeros", "delete", "insert", "truncate"]


def damage(flac: bytes, first: int, rng: random.Random, kind: str) -> bytes:
    b = bytearray(flac)
    n = len(b)
    if kind == "flip":
        for _ in range(rng.randint(1, 4)):
            b[rng.randrange(first, n)] ^= 1 << rng.randrange(8)
    elif kind == "run":
        p = rng.randrange(first, n - 64)
        for k in range(rng.randint(2, 48)):
            b[p + k] = rng.getrandbits(8)
    elif kind == "ones":
        p = rng.randrange(first, n - 64)
        for k in range(rng.randint(2, 40)):
            b[p + k] = 0xFF
    elif kind == "zeros":
        p = rng.randrange(first, n - 64)
        for k in range(rng.randint(2, 40)):
            b[p + k] = 0
    elif kind == "delete":
        p = rng.randrange(first, n - 64)
        del b[p:p + rng.randint(1, 30)]
    elif kind == "insert":
        p = rng.randrange(first, n - 64)
        b[p:p] = bytes(rng.getrandbits(8) for _ in range(rng.randint(1, 30)))
    elif kind == "truncate":
        del b[rng.randrange(first + 100, n):]
    return bytes(b)


def streams(shape, kind, trials):
    """(trial, blob) for the seeded sequence of one shape and kind; the base stream comes from the corpus generator"""
    import pycorpus
    s = pycorpus.make(**SHAPES[shape])
    rng = random.Random(zlib.crc32(f"{shape}/{kind}".encode()))
    for t in range(trials):
        yield t, damage(s.flac, s.frame_off[0], rng, kind), s


def golden_damage_records():
    """(key, blob, reference-DLL record) for the 224 committed records; asserts that the regenerated stream is the one the DLL saw"""
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_damage.json")))
    rec = g["records"]
    for shape in sorted(SHAPES):
        for kind in KINDS:
            for t, blob, _ in streams(shape, kind, g["trials"]):
                r = rec[f"{shape}/{kind}/{t}"]
                assert hashlib.md5(blob).hexdigest() == r["blob_md5"], "the damaged stream is not the one the reference DLL decoded"
                yield f"{shape}/{kind}/{t}", blob, r

"""Damaged streams: PCM, frame table and error-callback events of the engine against the CPU oracle (the restatement of
the reference's sequential decoder) under seeded random damage: bit flips, byte runs overwritten, bytes deleted and
inserted, truncation.  What the reference does after a damaged frame depends on where its parse of the damaged bits
stops, so these cases exercise k_parse's walk of CRC-failed frames and k_resync."""
import os
import random
import zlib

import pytest

pytestmark = pytest.mark.gpu

TRIALS = 25

SHAPES = {
    "s16_lpc8": dict(ch=2, bps=16, sr=44100, seconds=1.5, bs=1152, lpc=8, seed=21),
    "s16_fixed_small_k": dict(ch=1, bps=16, sr=44100, seconds=1.5, bs=576, lpc=0, noise=3, seed=22),
    "s24_lpc12": dict(ch=2, bps=24, sr=96000, seconds=0.7, bs=4096, lpc=12, maxpo=6, seed=23),
    "s16_var": dict(ch=2, bps=16, sr=44100, seconds=1.5, lpc=8, var=(4096, 1152, 4080, 720, 16, 192, 2304), seed=24),
}


def _damage(flac: bytes, first: int, rng: random.Random, kind: str) -> bytes:
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


@pytest.mark.parametrize("shape", sorted(SHAPES))
@pytest.mark.parametrize("kind", ["flip", "run", "ones", "zeros", "delete", "insert", "truncate"])
def test_random_damage_matches_the_oracle(shape, kind):
    import pycorpus
    import pyoracle
    from birdnest.audio_b200 import _abi
    s = pycorpus.make(**SHAPES[shape])
    first = s.frame_off[0]
    rng = random.Random(zlib.crc32(f"{shape}/{kind}".encode()))
    bad = []
    for trial in range(TRIALS):
        blob = _damage(s.flac, first, rng, kind)
        want, oframes, _, oerrs = pyoracle.decode(blob, want_frames=True)
        with _abi.open_memory(blob) as h:
            out = bytearray(len(s.pcm) + (1 << 20))
            k = h.decode_all(out)
            frames = h.frames()
            errs = h.errors()
        got = bytes(out[:k])
        same_frames = [(f.offset, f.length, f.status != 0) for f in frames] == [(o.offset, o.length, o.status != 0) for o in oframes]
        if got != want or not same_frames or errs != oerrs:
            bad.append((trial, len(got), len(want), len(frames), len(oframes), errs[:8], oerrs[:8]))
            dump = os.environ.get("BNFLAC_FUZZ_DUMP")
            if dump:
                os.makedirs(dump, exist_ok=True)
                with open(os.path.join(dump, f"fuzz_{shape}_{kind}_{trial}.flac"), "wb") as f:
                    f.write(blob)
    assert not bad, f"{shape}/{kind}: {len(bad)} of {TRIALS} differ: {bad[:3]}"

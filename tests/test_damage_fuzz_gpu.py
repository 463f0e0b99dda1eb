"""Damaged streams: PCM, frame table and error-callback events of the engine against the CPU oracle (the restatement of
the reference's sequential decoder) under seeded random damage: bit flips, byte runs overwritten, bytes deleted and
inserted, truncation.  What the reference does after a damaged frame depends on where its parse of the damaged bits
stops, so these cases exercise k_parse's walk of CRC-failed frames and k_resync."""
import os

import pytest

pytestmark = pytest.mark.gpu

from damage_cases import KINDS, SHAPES, streams as damaged_streams

TRIALS = 25


@pytest.mark.parametrize("shape", sorted(SHAPES))
@pytest.mark.parametrize("kind", KINDS)
def test_random_damage_matches_the_oracle(shape, kind):
    import pyoracle
    from birdnest.audio_b200 import _abi
    bad = []
    for trial, blob, s in damaged_streams(shape, kind, TRIALS):
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

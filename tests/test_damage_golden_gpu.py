"""The engine against the REFERENCE on damaged streams: the 224 records of tests/golden/golden_damage.json hold what the
reference's own LibFlac.dll (libFLAC 1.2.1, run in the build container by oracle/fuzz_vs_ref.py --write-golden) produced for
seeded damaged streams -- md5 of the PCM it delivered, number of frames delivered, the complete list of error-callback
statuses (LibFLACSharp.cs:262-268; raised as exceptions by FLACDecoder.cs:590-594).  The streams are regenerated from the
seeds (their md5 is in the record) and decoded through the C ABI; the oracle is not involved."""
import hashlib
import json
import os

import pytest

pytestmark = pytest.mark.gpu

import damage_cases


def _decode(blob, cap):
    from birdnest.audio_b200 import _abi
    with _abi.open_memory(blob) as h:
        out = bytearray(cap)
        k = h.decode_all(out)
        frames = h.frames()
        errs = h.errors()
    return bytes(out[:k]), len(frames), errs


def test_engine_matches_the_reference_dll_records():
    seen, bad = 0, []
    for key, blob, r in damage_cases.golden_damage_records():
        pcm, nframes, errs = _decode(blob, 4 << 20)
        got = (hashlib.md5(pcm).hexdigest(), nframes, errs[:64], len(errs))
        want = (r["pcm_md5"], r["frames"], r["errors"], r["n_errors"])
        if got != want:
            bad.append((key, got[1:], want[1:]))
        seen += 1
    log = os.environ.get("BNFLAC_DAMAGE_LOG")
    if log:
        with open(log, "w") as f:
            json.dump({"seen": seen, "bad": bad}, f, indent=1)
    assert seen == 224
    assert not bad, f"{len(bad)} of {seen} damaged streams differ from the reference DLL: {bad[:4]}"

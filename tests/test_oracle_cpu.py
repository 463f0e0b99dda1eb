"""CPU suite: the oracle (oracle/flac_oracle.c) against what the REFERENCE decoder produced (tests/golden/, written by
oracle/make_golden.py running LibFlac.dll), plus known-answer tests for its CRC/MD5 and for the stream generator."""
import hashlib
import json
import os
import shutil
import subprocess
import sys

import pytest

from conftest import CASES, ROOT

GOLD = os.path.join(ROOT, "tests", "golden")
golden = json.load(open(os.path.join(GOLD, "golden.json")))


def test_known_answers():
    import pyoracle
    L = pyoracle.lib()
    assert L.fo_crc8(b"123456789", 9) == 0xF4            # CRC-8 poly 0x07 init 0
    assert L.fo_crc16(b"123456789", 9) == 0xFEE8         # CRC-16 poly 0x8005 init 0, MSB first
    assert pyoracle.md5(b"abc").hex() == "900150983cd24fb0d6963f7d28e17f72"
    assert pyoracle.md5(b"") == hashlib.md5(b"").digest()
    blob = bytes(range(256)) * 37
    assert pyoracle.md5(blob) == hashlib.md5(blob).digest()


@pytest.mark.parametrize("name", sorted(golden["fixtures"]))
def test_oracle_matches_reference_on_reference_encoded_fixtures(name):
    """Streams ENCODED by the reference's encoder; the md5 recorded is the one the reference DECODER produced."""
    import pyoracle
    g = golden["fixtures"][name]
    flac = open(os.path.join(GOLD, name + ".flac"), "rb").read()
    pcm, nframes, _, errs = pyoracle.decode(flac)
    si = pyoracle.streaminfo(flac)
    assert hashlib.md5(pcm).hexdigest() == g["pcm_md5"] == g["si_md5"] == bytes(si.md5).hex()
    assert (nframes, len(pcm), errs) == (g["frames"], g["bytes"], [])
    assert (si.channels, si.bits_per_sample, si.sample_rate, si.total_samples) == (g["channels"], g["bps"], g["sample_rate"], g["total_samples"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_on_synthetic_cases(streams, name):
    import pyoracle
    s = streams(name)
    g = golden["cases"][name]
    pcm, nframes, _, errs = pyoracle.decode(s.flac)
    assert pcm == s.pcm * s.tiles and errs == []
    assert hashlib.md5(pcm).digest() == s.md5 == bytes(pyoracle.streaminfo(s.flac).md5)
    # the generator is integer / IEEE-basic-operation only (no libm): the same bytes on every host as the ones the reference decoded
    assert hashlib.md5(s.flac).hexdigest() == g["flac_md5"], "corpus generator output differs from the golden run"
    assert hashlib.md5(pcm).hexdigest() == g["ref_pcm_md5"] and nframes == g["frames"]


@pytest.mark.parametrize("name", sorted(golden["faults"]))
def test_oracle_matches_reference_on_damaged_streams(streams, name):
    """CRC mismatch -> error 2 + zero-filled frame; bad header -> errors 1,0 and the frame is dropped."""
    import pyoracle
    g = golden["faults"][name]
    s = streams(g["case"])
    if hashlib.md5(s.flac).hexdigest() != golden["cases"][g["case"]]["flac_md5"]:
        pytest.skip("corpus bytes differ from the golden run on this host")
    b = bytearray(s.flac)
    b[g["pos"]] ^= g["mask"]
    pcm, nframes, _, errs = pyoracle.decode(bytes(b))
    assert (hashlib.md5(pcm).hexdigest(), len(pcm), nframes, errs) == (g["ref_pcm_md5"], g["ref_bytes"], g["frames"], g["errors"])


def test_reference_binary_agrees_when_it_can_run_here(streams):
    """Where oracle/_ref exists (the build container), run the reference decoder itself against the oracle."""
    exe, dll = os.path.join(ROOT, "oracle", "_ref", "refflac"), os.path.join(ROOT, "oracle", "_ref", "LibFlac.dll")
    if not (os.path.exists(exe) and os.path.exists(dll)):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    try:
        if subprocess.run([exe], capture_output=True, timeout=10).returncode != 2:
            pytest.skip("32-bit binaries do not run on this host")
    except OSError:
        pytest.skip("32-bit binaries do not run on this host")
    import pyoracle
    import tempfile
    for name in ("cfg2_24bit_stereo_lpc12", "cfg5_6ch_special", "cfg4_clip_stereo_var"):
        s = streams(name)
        with tempfile.TemporaryDirectory() as d:
            open(os.path.join(d, "i.flac"), "wb").write(s.flac)
            subprocess.check_call([exe, "dec", dll, os.path.join(d, "i.flac"), os.path.join(d, "o.pcm")], stdout=subprocess.DEVNULL)
            assert open(os.path.join(d, "o.pcm"), "rb").read() == pyoracle.decode(s.flac)[0]


def test_oracle_edge_cases():
    import pyoracle
    with pytest.raises(ValueError):
        pyoracle.streaminfo(b"RIFF....WAVE")
    with pytest.raises(ValueError):
        pyoracle.streaminfo(b"fLa")
    # metadata only, no frames: decodes to nothing
    import pycorpus
    s = pycorpus.make(ch=2, bps=16, sr=44100, samples=4096, bs=4096)
    head = s.flac[:s.frame_off[0]]
    pcm, nframes, _, errs = pyoracle.decode(head)
    assert (pcm, nframes, errs) == (b"", 0, [])
    # truncated in the middle of the only frame: reference reaches END_OF_STREAM without delivering it
    pcm, nframes, _, _ = pyoracle.decode(s.flac[:len(s.flac) - 50])
    assert (pcm, nframes) == (b"", 0)


@pytest.mark.parametrize("name", sorted(golden["metadata"]))
def test_oracle_matches_reference_on_metadata_variants(name):
    """Container breadth: every metadata block type, an ID3v2 tag, a PADDING block larger than the reference's
    16 KiB read buffer, a STREAMINFO without length / md5.  The md5 recorded is the one the reference DECODER produced."""
    import pyoracle
    g = golden["metadata"][name]
    flac = open(os.path.join(GOLD, name + ".flac"), "rb").read()
    pcm, nframes, _, errs = pyoracle.decode(flac)
    assert hashlib.md5(pcm).hexdigest() == g["ref_pcm_md5"]
    assert (nframes, len(pcm), errs) == (g["frames"], g["bytes"], g["errors"])


def test_oracle_matches_the_reference_dll_on_random_damage():
    """224 seeded damaged streams (bit flips, overwritten / 0xFF / 0x00 runs, deleted and inserted bytes, truncation) decoded
    by the reference's LibFlac.dll in the build container (oracle/fuzz_vs_ref.py --write-golden ->
    tests/golden/golden_damage.json).  The oracle reproduces PCM, frame count and the complete event list of every one
    (tests/test_damage_golden_gpu.py holds the engine to the same records)."""
    import damage_cases
    seen = 0
    for key, blob, r in damage_cases.golden_damage_records():
        pcm, nfr, _, errs = pyoracle_decode(blob)
        assert (hashlib.md5(pcm).hexdigest(), nfr, errs[:64], len(errs)) == (r["pcm_md5"], r["frames"], r["errors"], r["n_errors"]), key
        seen += 1
    assert seen == 224


def pyoracle_decode(blob):
    import pyoracle
    return pyoracle.decode(blob)


@pytest.mark.skipif(not (os.path.exists(os.path.join(ROOT, "oracle", "_ref", "refflac")) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "LibFlac.dll"))),
                    reason="oracle/_ref (the reference DLL under the PE loader) is not built here")
def test_fuzz_oracle_against_the_live_reference_dll():
    """oracle/fuzz_vs_ref.py as a test: the reference DLL is RUN here on 3 trials per shape and kind (84 streams) and compared
    with the oracle in PCM md5, frame count and events -- the live counterpart of the committed records above.  Skipped where
    the box cannot execute the 32-bit loader."""
    probe = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "refflac")], capture_output=True)
    if probe.returncode in (126, 127) or probe.returncode < 0:
        pytest.skip("this box does not execute 32-bit binaries")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "fuzz_vs_ref.py"), "3"], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    last = r.stdout.strip().splitlines()[-1]
    assert last.startswith("total 84 mismatch 0"), r.stdout[-2000:]

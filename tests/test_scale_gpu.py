"""GPU parity at sizes the CPU oracle does not finish quickly: checked through size-independent properties
(md5(PCM) == STREAMINFO md5, tiling linearity, shard concatenation) on BASELINE.json-shaped streams, plus the C++ host
mirror (flacdecoder_demo) against the oracle."""
import hashlib
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _decode(flac, **kw):
    from birdnest.audio_b200 import _abi
    with _abi.open_memory(flac, **kw) as h:
        info = h.info()
        out = bytearray(info.pcm_bytes + 64)
        n = h.decode_all(out)
        errs = h.errors()
    return bytes(out[:n]), info, errs


def test_cfg2_ten_minutes_md5_and_tiling():
    """cfg2 shape, 600 s: 20 s of unique audio tiled x30 -> the PCM is the 20 s PCM repeated 30 times and its md5 is STREAMINFO's."""
    import pycorpus
    s = pycorpus.make(ch=2, bps=24, sr=96000, seconds=20, bs=4096, lpc=12, maxpo=6, stereo=1, search=1, tile=30, seed=11)
    pcm, info, errs = _decode(s.flac)
    assert errs == []
    assert len(pcm) == info.pcm_bytes == len(s.pcm) * 30
    assert hashlib.md5(pcm).digest() == bytes(info.md5) == s.md5
    unit = len(s.pcm)
    assert all(pcm[k * unit:(k + 1) * unit] == s.pcm for k in (0, 1, 13, 29))


def test_cfg3_8ch_lpc32_rice2_one_minute_md5():
    """cfg3 shape: 24-bit 8-channel 192 kHz, blocksize 16384, LPC order 32, Rice2 with partition order 8 on every subframe."""
    import pycorpus
    s = pycorpus.make(ch=8, bps=24, sr=192000, samples=16384 * 12, bs=16384, lpc=32, minpo=8, maxpo=8, noise=19, search=0, tile=60, seed=5)
    pcm, info, errs = _decode(s.flac)
    assert errs == [] and len(pcm) == info.pcm_bytes
    assert hashlib.md5(pcm).digest() == bytes(info.md5) == s.md5
    # the same stream in 4 frame-range shards (cfg5's partitioning) concatenates to the same bytes
    parts = [_decode(s.flac, shard_index=i, shard_count=4)[0] for i in range(4)]
    assert b"".join(parts) == pcm


def test_cfg4_batch_of_2000_clips():
    """cfg4 shape: short 16-bit mono/stereo clips, mixed FIXED/LPC and block sizes, one pipeline pass per format group."""
    import pycorpus
    from birdnest.audio_b200 import _abi
    pool = []
    for i in range(40):
        ch = 1 + (i & 1)
        kw = dict(ch=ch, bps=16, sr=44100, seconds=0.5 + (i * 37 % 26) / 10.0, lpc=0 if i % 4 < 2 else 8, seed=100 + i)
        if i % 10 == 3:
            kw["var"] = (4096, 1152, 4080, 720, 16, 192, 2304)
        else:
            kw["bs"] = (576, 1152, 2304, 4096, 4608)[i % 5]
        pool.append(pycorpus.make(**kw))
    clips = [pool[(7 * k) % 40] for k in range(2000)]
    pcm, res = _abi.decode_batch([c.flac for c in clips])
    assert len(pcm) == sum(len(c.pcm) for c in clips)
    for c, r in zip(clips, res):
        assert r.status == 0 and r.pcm_bytes == len(c.pcm) and r.channels == c.channels
    for k in list(range(0, 2000, 97)) + [1999]:
        r = res[k]
        assert pcm[r.pcm_offset:r.pcm_offset + r.pcm_bytes] == clips[k].pcm, k
    seen = {}
    for c, r in zip(clips, res):                     # every clip's md5 matches its STREAMINFO
        d = seen.setdefault(id(c), hashlib.md5(pcm[r.pcm_offset:r.pcm_offset + r.pcm_bytes]).digest())
        assert d == c.md5


def test_cpp_host_mirror_demo(tmp_path):
    """csrc/flac_decoder.hpp (C++ mirror of FLACDecoder) through its Program.cs-shaped demo, against the oracle."""
    import pyoracle
    exe = os.path.join(ROOT, "birdnest", "audio_b200", "flacdecoder_demo")
    if not os.path.exists(exe):
        pytest.skip("flacdecoder_demo not built (make host)")
    src = os.path.join(ROOT, "tests", "golden", "ref_24bit_stereo_lpc12.flac")
    out = tmp_path / "out.pcm"
    r = subprocess.run([exe, src, str(out)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "BitsPerSample 24" in r.stdout and "Channels 2" in r.stdout
    with open(src, "rb") as f:
        want = pyoracle.decode(f.read())[0]
    assert out.read_bytes() == want


def test_encoder_ten_minutes_round_trip_and_reference_spot_check():
    """SURVEY 8f-4 at size: 600 s of the cfg2 shape (20 s of unique PCM x 30) encoded on the GPU from host memory, md5 in STREAMINFO,
    decoded back by the GPU decoder (whole) and -- where 32-bit binaries run -- by the reference DLL (the first 20 s re-encoded alone)."""
    import pycorpus
    from birdnest.audio_b200 import _abi
    from test_encode_emu_cpu import _ref_decode
    s = pycorpus.make(ch=2, bps=24, sr=96000, seconds=20, bs=4096, lpc=12, maxpo=6, seed=11)
    pcm = s.pcm * 30
    o = _abi.enc_opts(96000, 2, 24, max_lpc_order=12)
    flac, st = _abi.encode(pcm, o, want_stats=True)
    assert st.frames == (len(pcm) // 6 + 4095) // 4096 and len(flac) < 0.62 * len(pcm)
    back, info, errs = _decode(flac)
    assert errs == [] and back == pcm
    assert bytes(info.md5) == hashlib.md5(pcm).digest() and info.total_samples == len(pcm) // 6
    # frames are independent: the first tile's frames are the same bytes whether 20 s or 600 s are encoded
    one = _abi.encode(s.pcm, o)
    nfull = (len(s.pcm) // 6) // 4096
    with _abi.open_memory(one) as h:
        h.decode_all()
        fr = h.frames()
    end = fr[nfull - 1].offset + fr[nfull - 1].length
    assert flac[42:end] == one[42:end]
    ref = _ref_decode(one)
    if ref is not None:
        assert ref == s.pcm

"""Ogg FLAC (container breadth): the pages are taken apart on the host (`ogg_depage`, csrc/engine.cu) and the
native stream inside goes down the same GPU pipeline.  The reference's C# binds only the native-FLAC entry points
(LibFLACSharp.cs:42-85), but its LibFlac.dll exports FLAC__stream_decoder_init_ogg_stream: the records at the end of this
file were produced by that binary (oracle/make_golden_ogg.py) and pin the container layer; the other tests check that an
Ogg-wrapped stream decodes to exactly what the oracle produces for the native stream it was made from -- through every
page / packet layout the mapping allows."""
import hashlib
import random

import pytest

pytestmark = pytest.mark.gpu

from oggmux import mux, native_packets, page          # noqa: E402  (tests/oggmux.py)


SHAPES = {
    "stereo16": dict(ch=2, bps=16, sr=44100, seconds=4, bs=4096, lpc=8, maxpo=5, seed=51),
    "stereo24_msw": dict(ch=2, bps=24, sr=96000, seconds=2, bs=4096, lpc=12, maxpo=6, stereo=1, search=1, seed=52),
    "mono_small_frames": dict(ch=1, bps=16, sr=48000, seconds=3, bs=576, lpc=0, seed=53),
    "ch8_big_frames": dict(ch=8, bps=24, sr=192000, samples=16384 * 3, bs=16384, lpc=32, minpo=8, maxpo=8, noise=19, search=0, seed=54),
}


@pytest.mark.parametrize("shape", sorted(SHAPES))
@pytest.mark.parametrize("max_segs", [255, 17, 1], ids=["full_pages", "small_pages", "one_segment_pages"])
def test_ogg_flac_decodes_like_the_native_stream(shape, max_segs):
    import pycorpus
    import pyoracle
    from birdnest.audio_b200 import _abi
    s = pycorpus.make(**SHAPES[shape])
    want, nframes, _, oerrs = pyoracle.decode(s.flac)
    assert want == s.pcm and oerrs == []
    if max_segs == 1 and len(s.flac) > 600000:
        pytest.skip("one 255-byte segment per page: too many pages for a Python muxer")
    blob = b"".join(mux(s, random.Random(1), max_segs=max_segs, other_serial=0x77 if max_segs == 17 else None))
    with _abi.open_memory(blob) as h:
        info = h.info()
        assert (info.channels, info.bits_per_sample) == (SHAPES[shape]["ch"], SHAPES[shape]["bps"])
        assert info.total_samples == s.total_samples
        out = bytearray(len(want) + 64)
        n = h.decode_all(out)
        frames, errs = h.frames(), h.errors()
    assert bytes(out[:n]) == want and hashlib.md5(bytes(out[:n])).digest() == s.md5
    assert len(frames) == nframes and errs == []
    # frame offsets are those of the de-paged (native) stream
    assert [f.offset for f in frames] == list(s.frame_off)[:nframes]


def test_ogg_through_the_flacdecoder_mirror_and_lazy_pull():
    import io
    import pycorpus
    from birdnest.audio_b200 import FLACDecoder, FLACPacketQueue, EmptyStubLogger
    s = pycorpus.make(**SHAPES["stereo16"])
    blob = b"".join(mux(s, random.Random(2), max_segs=40))
    dec = FLACDecoder(io.BytesIO(blob), FLACPacketQueue(), EmptyStubLogger())
    assert (dec.Channels, dec.BitsPerSample, dec.SampleRate) == (2, 16, 44100)
    ms = io.BytesIO()
    dec.CopyTo(ms)
    dec.Dispose()
    assert ms.getvalue() == s.pcm


def test_ogg_with_lost_and_damaged_pages():
    """A page that fails its CRC, and a page that is missing altogether, cost the packets they carry (and the packets they
    interrupt); everything else decodes.  Checked against the oracle on the native stream with those frames cut out."""
    import pycorpus
    import pyoracle
    from birdnest.audio_b200 import _abi
    s = pycorpus.make(**SHAPES["mono_small_frames"])
    headers, frames = native_packets(s)
    # one frame per page (no spanning), so that what a lost page costs is exactly its frame
    serial, pages = 9, []
    pages.append(page(serial, 0, 2, 0, [len(headers[0])], headers[0]))
    seq = 1
    for hp in headers[1:]:
        pages.append(page(serial, seq, 0, 0, [255] * (len(hp) // 255) + [len(hp) % 255], hp)); seq += 1
    first_audio = len(pages)
    for i, f in enumerate(frames):
        pages.append(page(serial, seq, 4 if i + 1 == len(frames) else 0, i, [255] * (len(f) // 255) + [len(f) % 255], f)); seq += 1
    lost, bad = 10, 40
    dmg = list(pages)
    b = bytearray(dmg[first_audio + bad]); b[len(b) // 2] ^= 0x20; dmg[first_audio + bad] = bytes(b)      # CRC mismatch
    del dmg[first_audio + lost]                                                                            # missing page
    blob = b"".join(dmg)
    native = s.flac[:s.frame_off[0]] + b"".join(f for i, f in enumerate(frames) if i not in (lost, bad))
    want, nf, _, oerrs = pyoracle.decode(native)
    with _abi.open_memory(blob) as h:
        out = bytearray(len(s.pcm) + 64)
        n = h.decode_all(out)
        got_frames, errs = h.frames(), h.errors()
    assert bytes(out[:n]) == want and len(got_frames) == nf == len(frames) - 2
    assert errs == oerrs


def test_not_ogg_flac():
    from birdnest.audio_b200 import _abi
    vorbis_like = page(1, 0, 2, 0, [30], b"\x01vorbis" + bytes(23)) + page(1, 1, 4, 0, [10], bytes(10))
    with pytest.raises(_abi.BnflacError) as e:
        _abi.open_memory(vorbis_like)
    assert e.value.code == _abi.ERR_NOT_FLAC
    v2 = bytearray(b"\x7fFLAC\x02\x00\x00\x00fLaC" + bytes(38))
    with pytest.raises(_abi.BnflacError) as e:
        _abi.open_memory(page(1, 0, 2, 0, [len(v2)], bytes(v2)))
    assert e.value.code == _abi.ERR_UNSUPPORTED


import json
import os

_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_ogg_golden = json.load(open(os.path.join(_GOLD, "golden_ogg.json")))


@pytest.mark.parametrize("name", sorted(k for k in _ogg_golden if k != "note"))
def test_ogg_golden_records_of_the_reference_dll(name):
    """Pages decoded by the reference's own LibFlac.dll through init_ogg_stream (oracle/make_golden_ogg.py): same PCM, frame
    count, end state and error events from the GPU path -- this pins the container layer to the reference binary."""
    from birdnest.audio_b200 import _abi
    g = _ogg_golden[name]
    blob = open(os.path.join(_GOLD, name + ".oga"), "rb").read()
    with _abi.open_memory(blob) as h:
        out = bytearray(g["bytes"] + 4096)
        n = h.decode_all(out)
        frames, errs, state = h.frames(), h.errors(), h.state()
        info = h.info()
    assert n == g["bytes"] and hashlib.md5(bytes(out[:n])).hexdigest() == g["pcm_md5"]
    assert len(frames) == g["frames"] and errs == g["errors"] and state == g["state"]
    assert bytes(info.md5).hex() == g["si_md5"]

"""GPU parity against what the REFERENCE decoder itself produced (tests/golden/, see oracle/make_golden.py), through the C ABI
and through the FLACDecoder mirror of the reference class."""
import hashlib
import io
import json
import os

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")
golden = json.load(open(os.path.join(GOLD, "golden.json")))


@pytest.mark.parametrize("name", sorted(golden["fixtures"]))
def test_reference_encoded_fixture_decodes_to_the_reference_pcm(name):
    from birdnest.audio_b200 import _abi
    g = golden["fixtures"][name]
    flac = open(os.path.join(GOLD, name + ".flac"), "rb").read()
    with _abi.open_memory(flac, flags=_abi.OPT_VERIFY_MD5) as h:
        pcm = h.decode_all()
        info = h.info()
        frames = h.frames()
        errs = h.errors()
    assert hashlib.md5(pcm).hexdigest() == g["pcm_md5"] == bytes(info.md5).hex()
    assert (len(frames), len(pcm), errs) == (g["frames"], g["bytes"], [])
    assert (info.channels, info.bits_per_sample, info.sample_rate, info.total_samples) == (g["channels"], g["bps"], g["sample_rate"], g["total_samples"])


@pytest.mark.parametrize("name", sorted(golden["faults"]))
def test_damaged_streams_behave_like_the_reference(streams, name):
    """CRC mismatch -> frame delivered zero-filled + FrameCrcMismatch; damaged header -> frame dropped, BadHeader + LostSync."""
    import pyoracle
    from birdnest.audio_b200 import _abi
    g = golden["faults"][name]
    s = streams(g["case"])
    b = bytearray(s.flac)
    b[g["pos"]] ^= g["mask"]
    want, nframes, _, oerrs = pyoracle.decode(bytes(b))
    with _abi.open_memory(bytes(b)) as h:
        pcm = h.decode_all()
        frames = h.frames()
        errs = h.errors()
    assert pcm == want and len(frames) == nframes and errs == oerrs
    if hashlib.md5(s.flac).hexdigest() == golden["cases"][g["case"]]["flac_md5"]:
        assert (hashlib.md5(pcm).hexdigest(), len(pcm), len(frames), errs) == (g["ref_pcm_md5"], g["ref_bytes"], g["frames"], g["errors"])


def test_flacdecoder_stream_surface_like_openaldemo(streams):
    """OpenALDemo/Program.cs:26-38: new FLACDecoder(fs, queue, logger); SampleRate/Duration/Format; CopyTo(MemoryStream)."""
    from birdnest.audio_b200 import FLACDecoder, FLACPacketQueue, EmptyStubLogger, ALFormat
    s = streams("cfg1_16bit_stereo_lpc8")
    src = io.BytesIO(s.flac)
    reader = FLACDecoder(src, FLACPacketQueue(), EmptyStubLogger())
    assert (reader.SampleRate, reader.Channels, reader.BitsPerSample, reader.Format) == (44100, 2, 16, ALFormat.Stereo16)
    assert reader.Length == s.total_samples * 4 and abs(reader.Duration.total_seconds() - s.total_samples / 44100) < 1e-6
    assert reader.CanRead and not reader.CanSeek and not reader.CanWrite
    for fn in (reader.Flush, lambda: reader.Seek(0, 0), lambda: reader.SetLength(1), lambda: reader.Write(b"", 0, 0), lambda: reader.Position):
        with pytest.raises(NotImplementedError):
            fn()
    ms = io.BytesIO()
    reader.CopyTo(ms)
    assert ms.getvalue() == s.pcm
    assert reader.Read(bytearray(16), 0, 16) == 0            # end of stream
    reader.Dispose()
    assert src.closed                                          # the decoder owns and closes the inner stream (FLACDecoder.cs:308)


def test_flacdecoder_read_granularity(streams):
    """Read returns exactly `count` until the stream ends (FLACDecoder.cs:131-187), for odd sizes and offsets."""
    from birdnest.audio_b200 import FLACDecoder, FLACPacketQueue, EmptyStubLogger
    s = streams("cfg4_clip_mono_fixed")
    reader = FLACDecoder(io.BytesIO(s.flac), FLACPacketQueue(), EmptyStubLogger(), bytearray(777))   # tiny in-stream buffer
    out = bytearray()
    buf = bytearray(5000)
    sizes = [1, 2, 3, 997, 4096, 81920, 7]
    i = 0
    while True:
        want = min(sizes[i % len(sizes)], len(buf) - 11)
        n = reader.Read(buf, 11, want)
        out += buf[11:11 + n]
        if n < want:
            break
        i += 1
    assert bytes(out) == s.pcm
    reader.Dispose()


def test_flacdecoder_24bit_and_strict_reference_mode(streams):
    from birdnest.audio_b200 import FLACDecoder, FLACPacketQueue, EmptyStubLogger, ALFormat, ApplicationException

    class Log:
        def __init__(self):
            self.m = []

        def Warning(self, t):
            self.m.append(t)
    s = streams("cfg2_24bit_stereo_lpc12")
    log = Log()
    r = FLACDecoder(io.BytesIO(s.flac), FLACPacketQueue(), log)
    assert r.Format == ALFormat.Unmapped and log.m == ["FLAC: Unsupported sample bit size: 24\n"]    # FLACDecoder.cs:462-465
    ms = io.BytesIO()
    r.CopyTo(ms)
    assert ms.getvalue() == s.pcm                     # FLACFileReader's 3-byte layout (FLACFileReader.cs:222-236)
    r.Dispose()
    # the shipped class aborts on anything but 16-bit (FLACDecoder.cs:526-530): reproduced on request
    r = FLACDecoder(io.BytesIO(s.flac), FLACPacketQueue(), Log(), strict_reference=True)
    with pytest.raises(ApplicationException) as e:
        r.Read(bytearray(100), 0, 100)
    assert str(e.value) == "FLAC: Could not process single - Aborted!"
    r.Dispose()


def test_flacdecoder_error_callback_text(streams):
    """ErrorCallback throws from the decode call (FLACDecoder.cs:590-594) with the reference's status/state names."""
    from birdnest.audio_b200 import FLACDecoder, FLACPacketQueue, EmptyStubLogger, ApplicationException
    g = golden["faults"]["payload_bit_mid_frame"]
    s = streams(g["case"])
    b = bytearray(s.flac)
    b[g["pos"]] ^= g["mask"]
    r = FLACDecoder(io.BytesIO(bytes(b)), FLACPacketQueue(), EmptyStubLogger())
    with pytest.raises(ApplicationException) as e:
        r.Read(bytearray(4096), 0, 4096)
    assert str(e.value) == "FLAC: Could not decode frame: FrameCrcMismatch - ReadFrame!"
    r.Dispose()


def test_empty_and_truncated_streams(streams):
    from birdnest.audio_b200 import _abi
    s = streams("short_single_frame")
    head = s.flac[:s.frame_off[0]]
    with _abi.open_memory(head) as h:           # metadata only
        assert h.decode_all() == b"" and h.frames() == []
    with _abi.open_memory(s.flac[:len(s.flac) - 3]) as h:    # the only frame is cut short: nothing is delivered (reference: END_OF_STREAM)
        assert h.decode_all() == b""


@pytest.mark.parametrize("name", sorted(golden["metadata"]))
@pytest.mark.parametrize("how", ["decode_all", "read", "callbacks"])
def test_metadata_variants_decode_to_the_reference_pcm(name, how):
    """APPLICATION / SEEKTABLE / VORBIS_COMMENT / CUESHEET / PICTURE / PADDING / reserved blocks (with sync-code
    look-alikes inside them), an ID3v2 prefix, 70 kB of PADDING and a STREAMINFO that knows neither length nor md5: the
    engine must deliver what the reference decoder delivered (tests/golden/golden.json, written by oracle/make_golden.py)."""
    from birdnest.audio_b200 import _abi
    g = golden["metadata"][name]
    flac = open(os.path.join(GOLD, name + ".flac"), "rb").read()
    if how == "callbacks":
        # FLACDecoder.ReadCallback exactly (FLACDecoder.cs:325-363): at most 16 KiB per call whatever was asked for, and
        # end-of-stream only when the inner stream returned less than that capped length
        import ctypes as C
        pos = [0]

        def _cb(user, buf, nbytes):
            want = min(nbytes[0], 16384)
            chunk = flac[pos[0]:pos[0] + want]
            pos[0] += len(chunk)
            C.memmove(buf, chunk, len(chunk))
            nbytes[0] = len(chunk)
            return 1 if len(chunk) < want else 0
        cb = _abi.READ_CB(_cb)
        o = _abi._opts(-1)
        hp = C.c_void_p()
        assert _abi.lib().bnflac_open_callbacks(cb, None, C.byref(o), C.byref(hp)) == 0
        h = _abi.Handle(hp.value)
    else:
        h = _abi.open_memory(flac)
    with h:
        info = h.info()
        if how == "read":
            pcm = bytearray()
            buf = bytearray(81920)
            while True:
                n = h.read_into(buf)
                pcm += buf[:n]
                if n < len(buf):
                    break
            pcm = bytes(pcm)
        else:
            pcm = h.decode_all()
        frames = h.frames()
        errs = h.errors()
    assert hashlib.md5(pcm).hexdigest() == g["ref_pcm_md5"]
    assert (len(frames), len(pcm), errs) == (g["frames"], g["bytes"], g["errors"])
    assert (info.channels, info.bits_per_sample, info.sample_rate, info.total_samples) == (g["ch"], g["bps"], g["sr"], g["total"])
    if g["total"]:
        assert bytes(info.md5).hex() == g["si_md5"] == g["ref_pcm_md5"]

"""Ogg FLAC (container breadth): the pages are taken apart on the host (`ogg_depage`, csrc/engine.cu) and the
native stream inside goes down the same GPU pipeline.  The reference's C# binds only the native-FLAC entry points
(LibFLACSharp.cs:42-85), so there is no reference surface to pin this against: **parity unpinned for the container layer**;
the payload decode is the pinned native path, and these tests check that an Ogg-wrapped stream decodes to exactly what the
oracle produces for the native stream it was made from -- through every page / packet layout the mapping allows."""
import binascii
import hashlib
import random
import struct

import pytest

pytestmark = pytest.mark.gpu

_REV8 = bytes(int(f"{i:08b}"[::-1], 2) for i in range(256))


def ogg_crc(data: bytes) -> int:
    """CRC-32 with polynomial 0x04C11DB7, no reflection, initial value 0, no final XOR -- computed with zlib's reflected
    CRC on bit-reversed bytes (an independent route from the table-driven one in the library)."""
    r = binascii.crc32(data.translate(_REV8), 0xFFFFFFFF) ^ 0xFFFFFFFF
    return int(f"{r:032b}"[::-1], 2)


def page(serial, seq, flags, granule, lacing, body):
    hdr = b"OggS\0" + bytes([flags]) + struct.pack("<qIII", granule, serial, seq, 0) + bytes([len(lacing)]) + bytes(lacing)
    crc = ogg_crc(hdr + body)
    return hdr[:22] + struct.pack("<I", crc) + hdr[26:] + body


def native_packets(s):
    """Ogg FLAC packets of a native stream: header packet, one packet per further metadata block, one per frame."""
    flac = s.flac
    assert flac[:4] == b"fLaC"
    blocks, pos, last = [], 4, False
    while not last:
        last = bool(flac[pos] & 0x80)
        n = int.from_bytes(flac[pos + 1:pos + 4], "big")
        blocks.append(flac[pos:pos + 4 + n])
        pos += 4 + n
    assert pos == s.frame_off[0]
    offs = list(s.frame_off)                          # frame starts + the end of the stream
    assert offs[-1] == len(flac)
    frames = [flac[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]
    first = b"\x7fFLAC\x01\x00" + struct.pack(">H", len(blocks) - 1) + b"fLaC" + blocks[0]
    return [first] + blocks[1:], frames


def mux(s, rng=None, max_segs=255, serial=0x1234, other_serial=None, split_pages=True):
    """Pages as an Ogg muxer would write them: the header packet alone on the first page, the other header packets on the
    next, then audio packets packed into pages of up to `max_segs` segments, packets spanning pages where they do not fit."""
    headers, frames = native_packets(s)
    out, seq = [], 0
    out.append(page(serial, seq, 2, 0, [len(headers[0])], headers[0])); seq += 1
    if other_serial is not None:                      # a second logical stream multiplexed in (its pages must be ignored)
        out.append(page(other_serial, 0, 2, 0, [8], b"\x01garbage"))
    lacing, body, pending = [], bytearray(), []

    def flush(cont, last=False):
        nonlocal lacing, body, seq
        if not lacing and not last:
            return
        out.append(page(serial, seq, (1 if cont else 0) | (4 if last else 0), seq, lacing, bytes(body)))
        seq += 1
        if other_serial is not None and seq % 5 == 0:
            out.append(page(other_serial, seq // 5, 0, 0, [255, 3], bytes(258)))
        lacing, body = [], bytearray()

    cont = False
    def add_packet(pkt):
        nonlocal cont, lacing, body
        segs = [255] * (len(pkt) // 255) + [len(pkt) % 255]       # a multiple of 255 ends with a 0 lacing value
        at = 0
        for L in segs:
            if len(lacing) >= max_segs:
                flush(cont)
                cont = at > 0                                   # the next page continues this packet
            lacing.append(L)
            body += pkt[at:at + L]
            at += L
        if rng is not None and split_pages and rng.random() < 0.3:
            flush(cont); cont = False

    for hp in headers[1:]:
        add_packet(hp)
    flush(cont); cont = False
    for f in frames:
        add_packet(f)
    flush(cont, last=True)
    return out


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

"""Streaming Read: `bnflac_read` on a large host-resident stream decodes sub-shards AHEAD of the reader
instead of the whole stream at the first call (FLACDecoder.Read, FLACDecoder.cs:124-224, fed to the buffer ring that
StreamingPlayer.cs:8-19,424-464 sketches).  The bytes a sequence of Reads returns must be exactly what the one-shot
decode and the CPU oracle produce, for any read sizes, on intact and damaged streams; the error events and frame table
must be those of the whole stream."""
import hashlib
import os
import random
import time

import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def tiny_subshards(monkeypatch):
    """Sub-shards of 16 KiB, 32 KiB ... 1 MiB so that streams of a few MB run through many of them."""
    monkeypatch.setenv("BNFLAC_READ_FIRST_KB", "16")
    monkeypatch.setenv("BNFLAC_READ_MB", "1")


def _read_all(h, sizes):
    out = bytearray()
    k = 0
    while True:
        n = sizes[k % len(sizes)]
        k += 1
        buf = bytearray(n)
        r = h.read_into(buf)
        out += buf[:r]
        if r < n:
            break
    return bytes(out)


SHAPES = {
    "cfg2_24bit_stereo": dict(ch=2, bps=24, sr=96000, seconds=6, bs=4096, lpc=12, maxpo=6, stereo=1, search=1, seed=31),
    "cfg1_16bit_stereo": dict(ch=2, bps=16, sr=44100, seconds=20, bs=4096, lpc=8, maxpo=5, seed=32),
    "var_blocksize": dict(ch=2, bps=16, sr=44100, seconds=12, lpc=8, var=(4096, 1152, 4080, 720, 16, 192, 2304), seed=33),
    "ch8_24bit_lpc32": dict(ch=8, bps=24, sr=192000, samples=16384 * 10, bs=16384, lpc=32, minpo=8, maxpo=8, noise=19, search=0, seed=34),
    "mono_fixed_small_frames": dict(ch=1, bps=16, sr=48000, seconds=20, bs=576, lpc=0, seed=35),
}


@pytest.mark.parametrize("shape", sorted(SHAPES))
@pytest.mark.parametrize("sizes", [(81920,), (7777, 1, 300001), (3,)], ids=["copyto", "ragged", "tiny"])
def test_streamed_reads_equal_the_oracle(tiny_subshards, shape, sizes):
    import pycorpus
    import pyoracle
    from birdnest.audio_b200 import _abi
    s = pycorpus.make(**SHAPES[shape])
    if sizes == (3,):
        s = pycorpus.make(**dict(SHAPES[shape], seconds=1) if "seconds" in SHAPES[shape] else dict(SHAPES[shape], samples=16384 * 2))
    want, nframes, _, oerrs = pyoracle.decode(s.flac)
    assert want == s.pcm
    with _abi.open_memory(s.flac) as h:
        got = _read_all(h, sizes)
        assert h.read_into(bytearray(64)) == 0           # stays at end of stream
        assert h.state() == 4                            # StreamDecoderState.EndOfStream
        frames = h.frames()
        errs = h.errors()
    assert got == want
    assert hashlib.md5(got).digest() == s.md5
    assert len(frames) == nframes and errs == oerrs == []
    offs = [f.pcm_offset for f in frames]
    assert offs == sorted(offs) and offs[0] == 0        # frame table in stream coordinates across sub-shards


def test_streamed_read_of_damaged_stream_matches_the_oracle(tiny_subshards):
    import pycorpus
    import pyoracle
    from birdnest.audio_b200 import _abi
    s = pycorpus.make(ch=2, bps=16, sr=44100, seconds=20, bs=1152, lpc=8, seed=36)
    rng = random.Random(5)
    b = bytearray(s.flac)
    first = s.frame_off[0]
    for _ in range(12):                                  # bit flips, an overwritten run and a deleted run, spread over the stream
        b[rng.randrange(first, len(b))] ^= 1 << rng.randrange(8)
    p = rng.randrange(first, len(b) - 100)
    b[p:p + 40] = bytes(rng.getrandbits(8) for _ in range(40))
    p = rng.randrange(first, len(b) - 100)
    del b[p:p + 17]
    blob = bytes(b)
    want, oframes, _, oerrs = pyoracle.decode(blob, want_frames=True)
    with _abi.open_memory(blob) as h:
        got = _read_all(h, (65536, 12345))
        frames = h.frames()
        errs = h.errors()
    with _abi.open_memory(blob) as h2:                   # the one-shot decode of the same bytes
        out = bytearray(len(s.pcm) + (1 << 20))
        k = h2.decode_all(out)
    assert got == want == bytes(out[:k])
    assert errs == oerrs and len(oerrs) > 0
    assert [(f.offset, f.length, f.status != 0) for f in frames] == [(o.offset, o.length, o.status != 0) for o in oframes]


def test_diagnostics_in_the_middle_of_a_streamed_read(tiny_subshards):
    """Asking for the frame table half way decodes what is left; the reads that follow still return the right bytes."""
    import pycorpus
    from birdnest.audio_b200 import _abi
    s = pycorpus.make(**SHAPES["cfg1_16bit_stereo"])
    with _abi.open_memory(s.flac) as h:
        head = bytearray(1 << 20)
        assert h.read_into(head) == len(head)
        frames = h.frames()
        assert sum(f.blocksize for f in frames) == s.total_samples
        rest = _read_all(h, (81920,))
        assert bytes(head) + rest == s.pcm
        # a one-shot decode on the same handle afterwards restarts from the beginning
        out = bytearray(len(s.pcm) + 64)
        assert h.decode_all(out) == len(s.pcm) and bytes(out[:len(s.pcm)]) == s.pcm


def test_read_chunk_frames_option():
    """bnflac_opts.read_chunk_frames: look-ahead batches of that many frames."""
    import pycorpus
    from birdnest.audio_b200 import _abi
    s = pycorpus.make(**SHAPES["cfg1_16bit_stereo"])
    with _abi.open_memory(s.flac, read_chunk_frames=8) as h:
        assert _read_all(h, (81920,)) == s.pcm
        assert len(h.frames()) == len(s.frame_bs)


def test_flacdecoder_mirror_streams(tiny_subshards):
    """The reference-facing surface: FLACDecoder(stream, queue, logger).Read(buf, 0, 81920) until 0 (Program.cs:26-38)."""
    import io
    import pycorpus
    from birdnest.audio_b200 import FLACDecoder, FLACPacketQueue, EmptyStubLogger
    s = pycorpus.make(**SHAPES["cfg1_16bit_stereo"])
    dec = FLACDecoder(io.BytesIO(s.flac), FLACPacketQueue(), EmptyStubLogger())
    out = bytearray()
    buf = bytearray(81920)
    while True:
        n = dec.Read(buf, 0, len(buf))
        if n == 0:
            break
        out += buf[:n]
    dec.Dispose()
    assert bytes(out) == s.pcm


def test_first_read_does_not_wait_for_the_whole_stream():
    """cfg2 shape, 10 minutes (200 MB compressed, 346 MB PCM): the first Read returns after the first 4 MiB sub-shard, long
    before a one-shot decode of the stream would; the streamed bytes hash to STREAMINFO's md5."""
    import pycorpus
    import torch
    from birdnest.audio_b200 import _abi
    s = pycorpus.make(ch=2, bps=24, sr=96000, seconds=20, bs=4096, lpc=12, maxpo=6, stereo=1, search=1, tile=30, seed=11)
    host = torch.frombuffer(bytearray(s.flac), dtype=torch.uint8).pin_memory()
    out = torch.empty(len(s.pcm) * 30 + 256, dtype=torch.uint8).pin_memory()
    for _ in range(2):                                   # warm the block pools: both paths then run allocation-free
        with _abi.open_memory(host, flags=_abi.OPT_BORROW_INPUT) as h:
            h.decode_all(out)
        with _abi.open_memory(host, flags=_abi.OPT_BORROW_INPUT) as h:
            _read_all(h, (1 << 22,))
    with _abi.open_memory(host, flags=_abi.OPT_BORROW_INPUT) as h:
        t0 = time.perf_counter()
        h.decode_all(out)
        t_all = time.perf_counter() - t0
    md5 = hashlib.md5()
    buf = bytearray(81920)
    with _abi.open_memory(host, flags=_abi.OPT_BORROW_INPUT) as h:
        t0 = time.perf_counter()
        n = h.read_into(buf)
        t_first = time.perf_counter() - t0
        assert n == len(buf)
        md5.update(buf)
        big = bytearray(1 << 22)
        while True:
            r = h.read_into(big)
            md5.update(memoryview(big)[:r])
            if r < len(big):
                break
    print(f"first Read(81920) {t_first * 1e3:.2f} ms; one-shot decode_all of the stream {t_all * 1e3:.2f} ms")
    assert md5.digest() == s.md5
    assert t_first < 0.5 * t_all

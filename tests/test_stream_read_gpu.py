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
    monkeypatch.setenv("BNFLAC_PULL_KB", "16")          # lazy pull: 16 KiB per callback, the reference's buffer (FLACDecoder.cs:21)


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


# ---- BNFLAC_OPT_LAZY_PULL: the source is a pull callback and only the metadata is read at open (FLACDecoder.cs:49-88);
# the rest is pulled as the reader advances (ReadCallback, FLACDecoder.cs:325-363), one sub-shard ahead of the decode.
class _CountingSource:
    def __init__(self, blob, fail_after=None):
        self.blob, self.pos, self.calls, self.fail_after = blob, 0, 0, fail_after

    def __call__(self, n):
        self.calls += 1
        if self.fail_after is not None and self.pos >= self.fail_after:
            return None                                       # ReadStatusAbort
        chunk = self.blob[self.pos:self.pos + n]
        self.pos += len(chunk)
        return chunk


@pytest.mark.parametrize("shape", sorted(SHAPES))
@pytest.mark.parametrize("sizes", [(81920,), (7777, 1, 300001)], ids=["copyto", "ragged"])
def test_lazy_pull_reads_equal_the_oracle(tiny_subshards, shape, sizes):
    import pycorpus
    import pyoracle
    from birdnest.audio_b200 import _abi
    s = pycorpus.make(**SHAPES[shape])
    want, nframes, _, oerrs = pyoracle.decode(s.flac)
    src = _CountingSource(s.flac)
    with _abi.open_callbacks(src, flags=_abi.OPT_LAZY_PULL) as h:
        at_open = src.pos
        info = h.info()
        assert info.channels == SHAPES[shape]["ch"] and info.bits_per_sample == SHAPES[shape]["bps"]
        buf = bytearray(4096)
        assert h.read_into(buf) == len(buf) and bytes(buf) == want[:4096]
        after_first = src.pos
        got = bytes(buf) + _read_all(h, sizes)
        assert h.read_into(bytearray(64)) == 0 and h.state() == 4
        frames, errs = h.frames(), h.errors()
    assert at_open <= (64 << 10) < len(s.flac)                   # metadata only
    assert after_first < len(s.flac)                             # the first Read did not pull the whole stream either
    assert src.pos == len(s.flac)
    assert got == want and hashlib.md5(got).digest() == s.md5
    assert len(frames) == nframes and errs == oerrs == []
    offs = [f.pcm_offset for f in frames]
    assert offs == sorted(offs) and offs[0] == 0


def test_lazy_pull_of_damaged_stream_matches_the_oracle(tiny_subshards):
    import pycorpus
    import pyoracle
    from birdnest.audio_b200 import _abi
    s = pycorpus.make(ch=2, bps=16, sr=44100, seconds=20, bs=1152, lpc=8, seed=37)
    rng = random.Random(6)
    b = bytearray(s.flac)
    first = s.frame_off[0]
    for _ in range(12):
        b[rng.randrange(first, len(b))] ^= 1 << rng.randrange(8)
    p = rng.randrange(first, len(b) - 100)
    b[p:p + 40] = bytes(rng.getrandbits(8) for _ in range(40))
    p = rng.randrange(first, len(b) - 100)
    del b[p:p + 17]
    del b[len(b) - 333:]                                        # and the stream ends inside a frame
    blob = bytes(b)
    want, oframes, _, oerrs = pyoracle.decode(blob, want_frames=True)
    with _abi.open_callbacks(_CountingSource(blob), flags=_abi.OPT_LAZY_PULL) as h:
        got = _read_all(h, (65536, 12345))
        frames, errs = h.frames(), h.errors()
    assert got == want
    assert errs == oerrs and len(oerrs) > 0
    assert [(f.offset, f.length, f.status != 0) for f in frames] == [(o.offset, o.length, o.status != 0) for o in oframes]


def test_lazy_pull_other_entry_points_pull_what_they_need(tiny_subshards):
    """Diagnostics or a one-shot decode on a lazily pulled handle pull the rest of the stream; a streamed Read session
    interrupted by either carries on / restarts with the right bytes."""
    import pycorpus
    from birdnest.audio_b200 import _abi
    s = pycorpus.make(**SHAPES["cfg1_16bit_stereo"])
    out = bytearray(len(s.pcm) + 64)
    # the tables are those of the last decode: empty before any Read (as on every handle); the reads that follow are right
    src = _CountingSource(s.flac)
    with _abi.open_callbacks(src, flags=_abi.OPT_LAZY_PULL) as h:
        assert src.pos < len(s.flac)
        assert h.frames() == [] and h.errors() == []
        assert _read_all(h, (81920,)) == s.pcm and src.pos == len(s.flac)
        assert sum(f.blocksize for f in h.frames()) == s.total_samples
    # decode_all straight after open
    with _abi.open_callbacks(_CountingSource(s.flac), flags=_abi.OPT_LAZY_PULL) as h:
        assert h.decoded_size() == len(s.pcm)
        assert h.decode_all(out) == len(s.pcm) and bytes(out[:len(s.pcm)]) == s.pcm
    # frames() in the middle of a session, the session carries on; then decode_all restarts from the beginning
    with _abi.open_callbacks(_CountingSource(s.flac), flags=_abi.OPT_LAZY_PULL) as h:
        head = bytearray(1 << 20)
        assert h.read_into(head) == len(head)
        assert sum(f.blocksize for f in h.frames()) == s.total_samples
        assert bytes(head) + _read_all(h, (81920,)) == s.pcm
        out[:] = bytes(len(out))
        assert h.decode_all(out) == len(s.pcm) and bytes(out[:len(s.pcm)]) == s.pcm
        assert h.errors() == []
    # decode_all in the middle of a session (nothing but the head pulled so far)
    src = _CountingSource(s.flac)
    with _abi.open_callbacks(src, flags=_abi.OPT_LAZY_PULL) as h:
        head = bytearray(100000)
        assert h.read_into(head) == len(head) and src.pos < len(s.flac)
        out[:] = bytes(len(out))
        assert h.decode_all(out) == len(s.pcm) and bytes(out[:len(s.pcm)]) == s.pcm
        assert len(h.frames()) == len(s.frame_bs)


def test_lazy_pull_abort_and_tiny_streams(tiny_subshards):
    import pycorpus
    from birdnest.audio_b200 import _abi
    s = pycorpus.make(**SHAPES["cfg1_16bit_stereo"])
    # the source aborts half way: the Read that needs those bytes reports ERR_ABORTED (ReadStatusAbort, FLACDecoder.cs:330-333)
    src = _CountingSource(s.flac, fail_after=len(s.flac) // 2)
    with _abi.open_callbacks(src, flags=_abi.OPT_LAZY_PULL) as h:
        with pytest.raises(_abi.BnflacError) as ei:
            _read_all(h, (81920,))
        assert ei.value.code == _abi.ERR_ABORTED
    # a stream shorter than the first request: open sees end of stream, everything still works
    t = pycorpus.make(ch=1, bps=16, sr=8000, samples=700, bs=256, lpc=4, seed=38)
    with _abi.open_callbacks(_CountingSource(t.flac), flags=_abi.OPT_LAZY_PULL) as h:
        assert _read_all(h, (100,)) == t.pcm
        assert len(h.frames()) == len(t.frame_bs)
    # metadata only, no frames
    with _abi.open_callbacks(_CountingSource(t.flac[:t.frame_off[0]]), flags=_abi.OPT_LAZY_PULL) as h:
        assert h.read_into(bytearray(100)) == 0 and h.state() == 4
        assert h.frames() == [] and h.errors() == []
    # a source that hands out 1000 bytes per call until the end (short counts only at end of stream are EOF: the mirror
    # classes loop until the request is full, as here)
    class Dribble(_CountingSource):
        def __call__(self, n):
            out = bytearray()
            while len(out) < n:
                c = super().__call__(min(1000, n - len(out)))
                out += c
                if len(c) < min(1000, n - len(out) + len(c)):
                    break
            return bytes(out)
    with _abi.open_callbacks(Dribble(s.flac), flags=_abi.OPT_LAZY_PULL) as h:
        assert _read_all(h, (50001,)) == s.pcm


def test_lazy_pull_with_read_chunk_frames(monkeypatch):
    """bnflac_opts.read_chunk_frames on a lazily pulled source: sub-shards (and pulls) of about that many frames."""
    import pycorpus
    from birdnest.audio_b200 import _abi
    monkeypatch.setenv("BNFLAC_PULL_KB", "16")
    s = pycorpus.make(**SHAPES["cfg1_16bit_stereo"])
    src = _CountingSource(s.flac)
    with _abi.open_callbacks(src, flags=_abi.OPT_LAZY_PULL, read_chunk_frames=8) as h:
        buf = bytearray(4096)
        assert h.read_into(buf) == len(buf) and bytes(buf) == s.pcm[:4096]
        assert src.pos < len(s.flac) // 2                            # a few 8-frame batches, not the 4 MiB default
        assert bytes(buf) + _read_all(h, (81920,)) == s.pcm
        assert len(h.frames()) == len(s.frame_bs)


def test_flacdecoder_mirror_pulls_lazily(tiny_subshards):
    """FLACDecoder(stream, ...) reads the metadata in its constructor and the stream as Read advances (FLACDecoder.cs:49-88,207-224)."""
    import io
    import pycorpus
    from birdnest.audio_b200 import FLACDecoder, FLACPacketQueue, EmptyStubLogger
    s = pycorpus.make(**dict(SHAPES["cfg2_24bit_stereo"], seconds=12))
    stream = io.BytesIO(s.flac)
    dec = FLACDecoder(stream, FLACPacketQueue(), EmptyStubLogger())
    assert stream.tell() < len(s.flac) and dec.Channels == 2 and dec.BitsPerSample == 24
    buf = bytearray(81920)
    n = dec.Read(buf, 0, len(buf))
    assert n == len(buf) and bytes(buf) == s.pcm[:n] and stream.tell() < len(s.flac)
    out = bytearray(buf[:n])
    while True:
        n = dec.Read(buf, 0, len(buf))
        if n == 0:
            break
        out += buf[:n]
    assert bytes(out) == s.pcm and stream.tell() == len(s.flac)
    dec.Dispose()                                               # closes the input stream too (FLACDecoder.cs:308)
    assert stream.closed


def _damage_late(s, seed):
    """A bit flip in the payload of a frame in the last quarter of the stream."""
    rng = random.Random(seed)
    k = rng.randrange(3 * len(s.frame_off) // 4, len(s.frame_off) - 1)
    b = bytearray(s.flac)
    b[(s.frame_off[k] + s.frame_off[k + 1]) // 2] ^= 0x10
    return bytes(b), k


@pytest.mark.parametrize("lazy", [False, True], ids=["memory", "lazy_pull"])
def test_errors_so_far_grow_with_the_reader(tiny_subshards, lazy):
    """bnflac_errors_so_far describes what the session has decoded so far and decodes / pulls nothing ahead; at the end it
    equals bnflac_errors and the oracle's event list."""
    import pycorpus
    import pyoracle
    from birdnest.audio_b200 import _abi
    s = pycorpus.make(ch=2, bps=16, sr=44100, seconds=60, bs=1152, lpc=8, seed=41)
    blob, k = _damage_late(s, 1)
    want, _, _, oerrs = pyoracle.decode(blob)
    assert oerrs
    src = _CountingSource(blob)
    h = _abi.open_callbacks(src, flags=_abi.OPT_LAZY_PULL) if lazy else _abi.open_memory(blob)
    with h:
        assert h.errors_so_far() == []                       # nothing decoded yet
        buf = bytearray(81920)
        got = bytearray()
        seen_at = None
        while True:
            n = h.read_into(buf)
            got += buf[:n]
            pulled = src.pos
            e = h.errors_so_far()
            assert src.pos == pulled                         # polling pulls nothing
            if e and seen_at is None:
                seen_at = len(got)
                assert e == oerrs[:len(e)]
            if n < len(buf):
                break
        assert h.errors_so_far() == oerrs == h.errors()
        frames = h.frames()
    assert bytes(got) == want
    # not before the reader was within the look-ahead (3 sub-shards of <= 1 MiB) of the damaged frame, not after it passed it
    first_bad = next(f for f in frames if f.status != 0)
    assert seen_at is not None and seen_at <= first_bad.pcm_offset + len(buf) + 81920
    assert seen_at > len(want) // 8


def test_flacdecoder_mirror_raises_when_the_reader_gets_there(tiny_subshards):
    """ErrorCallback (FLACDecoder.cs:590-594) fires from the Read that decodes the damaged frame, not from the first one."""
    import io
    import pycorpus
    from birdnest.audio_b200 import FLACDecoder, FLACPacketQueue, EmptyStubLogger, ApplicationException
    s = pycorpus.make(ch=2, bps=16, sr=44100, seconds=60, bs=1152, lpc=8, seed=42)
    blob, k = _damage_late(s, 2)
    stream = io.BytesIO(blob)
    dec = FLACDecoder(stream, FLACPacketQueue(), EmptyStubLogger())
    buf = bytearray(81920)
    out = bytearray()
    with pytest.raises(ApplicationException) as e:
        while True:
            n = dec.Read(buf, 0, len(buf))
            if n == 0:
                break
            out += buf[:n]
    assert str(e.value) == "FLAC: Could not decode frame: FrameCrcMismatch - ReadFrame!"
    assert len(out) > len(s.pcm) // 8 and bytes(out) == s.pcm[:len(out)]       # everything delivered before it is intact
    dec.Dispose()


def test_lazy_pull_of_a_ten_minute_stream():
    """cfg2 shape, 10 minutes (200 MB compressed, 346 MB PCM) behind a pull callback: the first Read needs only the head of
    the source, the pulls stay a bounded look-ahead in front of the reader, and the streamed bytes hash to STREAMINFO's md5."""
    import pycorpus
    from birdnest.audio_b200 import _abi
    s = pycorpus.make(ch=2, bps=24, sr=96000, seconds=20, bs=4096, lpc=12, maxpo=6, stereo=1, search=1, tile=30, seed=11)
    with _abi.open_callbacks(_CountingSource(s.flac), flags=_abi.OPT_LAZY_PULL) as h:      # warm the block pools (as the test above)
        _read_all(h, (1 << 22,))
    src = _CountingSource(s.flac)
    md5 = hashlib.md5()
    with _abi.open_callbacks(src, flags=_abi.OPT_LAZY_PULL) as h:
        assert src.pos <= (1 << 20)
        pcm_total = h.info().pcm_bytes                       # (s.pcm is the un-tiled 20 s)
        buf = bytearray(81920)
        t0 = time.perf_counter()
        assert h.read_into(buf) == len(buf)
        t_first = time.perf_counter() - t0
        md5.update(buf)
        assert src.pos < len(s.flac) // 8                    # 4 + 8 MiB sub-shards and their look-ahead, not the stream
        big = bytearray(1 << 22)
        total = len(buf)
        ahead = 0
        while True:
            r = h.read_into(big)
            md5.update(memoryview(big)[:r])
            total += r
            ahead = max(ahead, src.pos - int(total / pcm_total * len(s.flac)))
            if r < len(big):
                break
        assert h.errors() == [] and len(h.frames()) == len(s.frame_bs)
    print(f"lazy pull: first Read(81920) {t_first * 1e3:.2f} ms; source at most {ahead / 2**20:.0f} MiB ahead of the reader")
    assert md5.digest() == s.md5 and src.pos == len(s.flac)
    assert (1 << 20) < ahead < (230 << 20)                   # <= 3 sub-shards of <= 64 MiB in flight + half of one pulled beyond

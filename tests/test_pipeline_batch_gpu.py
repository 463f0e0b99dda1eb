"""GPU parity of the two host-facing fast paths: the pipelined (sub-sharded, multi-stream) host decode that large
streams take through bnflac_decode_all / bnflac_read, and bnflac_decode_batch (BASELINE cfg4: many clips, one pass per
format group).  Both are compared bit-exactly with the CPU oracle."""
import hashlib
import os

import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def small_pipe_shards(monkeypatch):
    monkeypatch.setenv("BNFLAC_PIPE_MB", "1")     # force the pipelined path on test-sized streams (default shard: 96 MiB)
    yield
    monkeypatch.delenv("BNFLAC_PIPE_MB", raising=False)


def _stream():
    import pycorpus
    return pycorpus.make(ch=2, bps=24, sr=96000, seconds=4, bs=4096, lpc=12, maxpo=6, tile=3, seed=7)


def test_pipelined_host_decode_matches_oracle(small_pipe_shards):
    import pyoracle
    from birdnest.audio_b200 import _abi
    s = _stream()
    assert len(s.flac) > 3 << 20
    want, oframes, _, oerrs = pyoracle.decode(s.flac, want_frames=True)
    with _abi.open_memory(s.flac) as h:
        out = bytearray(len(want) + 64)
        n = h.decode_all(out)
        frames, errs, info = h.frames(), h.errors(), h.info()
        t = h.timing()
    assert n == len(want) and bytes(out[:n]) == want
    assert hashlib.md5(out[:n]).digest() == bytes(info.md5)
    assert t.launches > 10, "expected several sub-shard passes"
    assert [(f.offset, f.length, f.number, f.pcm_offset) for f in frames] == \
           [(o.offset, o.length, o.number, sum(x.blocksize for x in oframes[:i]) * 6) for i, o in enumerate(oframes)]
    assert errs == oerrs == []


def test_pipelined_decode_of_damaged_stream_reports_like_one_pass(small_pipe_shards, monkeypatch):
    """Same bytes, same events and same frame table whether the stream is decoded in one pass or in pipelined sub-shards."""
    import pyoracle
    from birdnest.audio_b200 import _abi
    s = _stream()
    b = bytearray(s.flac)
    n = len(b)
    # damage inside Rice remainder bits (frame payloads): the frames keep their length and fail CRC-16
    hits = 0
    for f in (len(s.frame_off) // 5, len(s.frame_off) // 2, len(s.frame_off) - 3):
        b[s.frame_off[f] + 4000] ^= 0x01
        hits += 1
    b = bytes(b)
    want, nframes, _, oerrs = pyoracle.decode(b)
    with _abi.open_memory(b) as h:
        out = bytearray(len(want) + 64)
        k = h.decode_all(out)
        got_frames = [(f.offset, f.length, f.status, f.pcm_offset) for f in h.frames()]
        got_errs = h.errors()
    assert bytes(out[:k]) == want and got_errs == oerrs and len(got_frames) == nframes
    monkeypatch.setenv("BNFLAC_PIPE_MB", "4096")
    with _abi.open_memory(b) as h:
        one = h.decode_all()
        assert [(f.offset, f.length, f.status, f.pcm_offset) for f in h.frames()] == got_frames
        assert h.errors() == got_errs
    assert one == want


def test_header_and_payload_damage_across_sub_shards(small_pipe_shards, monkeypatch):
    """Damage that makes the reference drop a frame and resynchronise (a flipped bit in a subframe header, a wiped byte):
    the pipelined decode and the one-pass decode both match the oracle's PCM, frame list and events."""
    import pyoracle
    from birdnest.audio_b200 import _abi
    s = _stream()
    b = bytearray(s.flac)
    n = len(b)
    for pos, mask in ((n // 5, 0x10), (n // 2 + 3, 0xFF), (n - 40000, 0x01)):
        b[pos] ^= mask
    b = bytes(b)
    want, oframes, _, oerrs = pyoracle.decode(b, want_frames=True)
    for mb in ("1", "4096"):
        monkeypatch.setenv("BNFLAC_PIPE_MB", mb)
        with _abi.open_memory(b) as h:
            out = bytearray(len(s.pcm) * s.tiles + 64)
            k = h.decode_all(out)
            frames, errs = h.frames(), h.errors()
        assert bytes(out[:k]) == want, mb
        assert [(f.offset, f.length) for f in frames] == [(o.offset, o.length) for o in oframes], mb
        assert errs == oerrs, mb


def test_pipelined_read_stream_semantics(small_pipe_shards):
    from birdnest.audio_b200 import _abi
    s = _stream()
    want = s.pcm * s.tiles
    got = bytearray()
    with _abi.open_memory(s.flac) as h:
        buf = bytearray(81920)                     # Stream.CopyTo's buffer size (Program.cs:33)
        while True:
            k = h.read_into(buf)
            if k == 0:
                break
            got += buf[:k]
        assert h.state() == 4                      # EndOfStream
    assert bytes(got) == want


def _clips():
    import pycorpus
    mk = pycorpus.make
    return [
        mk(ch=1, bps=16, sr=44100, seconds=0.7, bs=576, lpc=0, seed=1),
        mk(ch=2, bps=16, sr=44100, seconds=1.1, lpc=8, var=(4096, 1152, 4080, 720, 16, 192, 2304), seed=2),
        mk(ch=2, bps=16, sr=44100, seconds=0.5, bs=1152, lpc=8, seed=3),
        mk(ch=1, bps=16, sr=44100, seconds=2.9, bs=4608, lpc=8, seed=4),
        mk(ch=2, bps=24, sr=96000, seconds=0.6, bs=4096, lpc=12, maxpo=6, seed=5),
        mk(ch=1, bps=16, sr=22050, seconds=0.5, bs=2304, lpc=0, seed=6),
        mk(ch=2, bps=16, sr=48000, samples=100, bs=4096, lpc=8, seed=7),
    ]


def test_batch_matches_per_clip_oracle():
    import pyoracle
    from birdnest.audio_b200 import _abi
    clips = _clips()
    blobs = [c.flac for c in clips] * 3 + [b"not a flac stream at all"]
    pcm, res = _abi.decode_batch(blobs)
    assert len(res) == len(blobs)
    total = 0
    for blob, r in zip(blobs[:-1], res[:-1]):
        want = pyoracle.decode(blob)[0]
        assert r.status == 0 and r.pcm_bytes == len(want)
        assert pcm[r.pcm_offset:r.pcm_offset + r.pcm_bytes] == want
        total += len(want)
    assert res[-1].status == 4 and res[-1].pcm_bytes == 0
    assert len(pcm) == total
    # clips of one format group are laid out in input order
    offs = [r.pcm_offset for b, r in zip(blobs[:-1], res[:-1])]
    assert len(set(offs)) == len(offs)


def test_batch_flags_a_damaged_clip_and_leaves_the_others_alone():
    import pyoracle
    from birdnest.audio_b200 import _abi
    clips = [c.flac for c in _clips()]
    bad = bytearray(clips[2])
    bad[len(bad) // 2] ^= 0x20
    clips[2] = bytes(bad)
    pcm, res = _abi.decode_batch(clips)
    for i, (blob, r) in enumerate(zip(clips, res)):
        want = pyoracle.decode(blob)[0]
        assert pcm[r.pcm_offset:r.pcm_offset + r.pcm_bytes] == want, i
        assert (r.status != 0) == (i == 2)


def test_batch_to_device_buffer():
    import torch
    from birdnest.audio_b200 import _abi
    clips = [c.flac for c in _clips()]
    pcm, res = _abi.decode_batch(clips)
    out = torch.zeros(len(pcm) + 64, dtype=torch.uint8, device="cuda:0")
    n, res2 = _abi.decode_batch(clips, device=0, dst=out, dst_is_device=True)
    torch.cuda.synchronize()
    assert n == len(pcm) and bytes(out[:n].cpu().numpy()) == pcm
    assert [(r.pcm_offset, r.pcm_bytes) for r in res] == [(r.pcm_offset, r.pcm_bytes) for r in res2]


def test_batch_table_is_reusable():
    """_abi.BatchTable: the bnflac_span table built once, decoded more than once (what bench.py's cfg4 / by_file legs do)."""
    from birdnest.audio_b200 import _abi
    clips = [c.flac for c in _clips()] * 2
    want_pcm, want_res = _abi.decode_batch(clips)
    table = _abi.BatchTable(clips)
    for _ in range(3):
        pcm, res = _abi.decode_batch(table)
        assert pcm == want_pcm
        assert [(r.pcm_offset, r.pcm_bytes, r.status) for r in res] == [(r.pcm_offset, r.pcm_bytes, r.status) for r in want_res]


def test_batch_gathered_and_uploaded_in_runs():
    """Large batches are gathered into staging memory in runs, each uploaded while the next is gathered (engine.cu, BNFLAC_BATCH_RUNS;
    by default one run per 128 MB).  Forced here on a small batch -- the variable is read once per process, hence the subprocess --
    with run boundaries falling between clips of different formats and sizes."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    code = r"""
import sys
sys.path[:0] = [%r, %r, %r, %r]
import pyoracle
from test_pipeline_batch_gpu import _clips
from birdnest.audio_b200 import _abi
clips = [c.flac for c in _clips()] * 40
pcm, res = _abi.decode_batch(clips)
ok = all(pcm[r.pcm_offset:r.pcm_offset + r.pcm_bytes] == pyoracle.decode(b)[0] and r.status == 0 for b, r in zip(clips[:len(_clips())], res))
ok = ok and all((res[i].pcm_bytes, res[i].status) == (res[i %% len(_clips())].pcm_bytes, 0) for i in range(len(clips)))
import hashlib
print(hashlib.md5(pcm).hexdigest(), "ok" if ok else "FAIL")
""" % (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "corpus"), os.path.join(ROOT, "tests"))
    outs = []
    for runs in ("1", "3", "7"):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=dict(os.environ, BNFLAC_BATCH_RUNS=runs))
        assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (runs, r.stdout[-500:], r.stderr[-1500:])
        outs.append(r.stdout.split()[0])
    assert len(set(outs)) == 1, outs


def test_packed_batch_is_uploaded_in_place_and_matches():
    """Clips that lie in ascending order inside ONE host buffer (a shard file read in one piece, tar-like 512-byte headers
    full of sync-code look-alikes between them) take the no-gather path: same PCM, same per-clip results as separate buffers."""
    from birdnest.audio_b200 import _abi
    clips = _clips()
    blobs = [c.flac for c in clips] * 4
    want_pcm, want_res = _abi.decode_batch(blobs)
    packed = bytearray()
    spans = []
    for i, b in enumerate(blobs):
        packed += (b"\xff\xf8\xc9\x18" * 128)[:512 - (i % 3)]          # junk between the clips, odd alignments
        spans.append((len(packed), len(b)))
        packed += b
    packed += bytes(100)
    mv = memoryview(packed)
    views = [mv[o:o + n] for o, n in spans]
    got_pcm, got_res = _abi.decode_batch(views, packed=True)
    assert _abi.decode_batch(views)[0] == want_pcm            # without the flag the same views are gathered: same result
    assert got_pcm == want_pcm
    assert [(r.pcm_offset, r.pcm_bytes, r.channels, r.bits_per_sample, r.status) for r in got_res] == \
           [(r.pcm_offset, r.pcm_bytes, r.channels, r.bits_per_sample, r.status) for r in want_res]
    for c, r in zip(clips * 4, got_res):
        assert got_pcm[r.pcm_offset:r.pcm_offset + r.pcm_bytes] == c.pcm
    # a damaged clip in the middle of the packed buffer is reported, the others are untouched
    o, n = spans[5]
    packed[o + n // 2] ^= 0x20
    bad_pcm, bad_res = _abi.decode_batch([mv[o:o + n] for o, n in spans], packed=True)
    assert bad_res[5].status != 0 and all(r.status == 0 for k, r in enumerate(bad_res) if k != 5)

"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.  Bit-exact."""
import hashlib

import pytest

from conftest import CASES

pytestmark = pytest.mark.gpu


def _first_diff(a: bytes, b: bytes) -> int:
    n = min(len(a), len(b))
    if a[:n] == b[:n]:
        return n
    lo, hi = 0, n
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if a[:mid] == b[:mid]:
            lo = mid
        else:
            hi = mid
    return lo


@pytest.mark.parametrize("name", sorted(CASES))
def test_pcm_bit_exact(streams, name):
    import pyoracle
    from birdnest.audio_b200 import _abi
    s = streams(name)
    want, oframes, osubs, oerrs = pyoracle.decode(s.flac, want_frames=True)
    assert want == s.pcm * s.tiles, "oracle disagrees with the synthesis PCM"
    with _abi.open_memory(s.flac) as h:
        got = h.decode_all()
        frames = h.frames()
        subs = h.subframes()
        errs = h.errors()
        info = h.info()
    if got != want:
        d = _first_diff(got, want)
        B = info.bytes_per_sample * info.channels
        pytest.fail(f"{name}: PCM differs at byte {d} (sample {d // B}) len got {len(got)} want {len(want)}; frames got {len(frames)} want {len(oframes)}")
    # STREAMINFO MD5 == md5(PCM)
    assert hashlib.md5(got).digest() == bytes(info.md5) == s.md5
    # K1 frame table == oracle frame list
    assert len(frames) == len(oframes)
    for f, o in zip(frames, oframes):
        assert (f.offset, f.length, f.blocksize, f.assignment, f.bits_per_sample, f.number, f.status) == \
               (o.offset, o.length, o.blocksize, o.channel_assignment, o.bits_per_sample, o.number, 0)
    # K2 subframe table == oracle
    for i, (f, osf) in enumerate(zip(frames, osubs)):
        for c, o in enumerate(osf):
            g = subs[8 * i + c]
            assert (g.type, g.order, g.wasted, g.bit_offset) == (o.type, o.order, o.wasted, o.bit_offset - 8 * f.offset), (name, i, c)
    assert errs == oerrs == []


@pytest.mark.parametrize("name", ["cfg1_16bit_stereo_lpc8", "cfg3_24bit_8ch_lpc32_rice2_po8", "cfg4_clip_stereo_var", "mono_special_escape_verbatim", "tiled_fixed"])
@pytest.mark.parametrize("nshards", [2, 3, 8])
def test_frame_range_shards_concatenate(streams, name, nshards):
    """frame-range shards decoded independently concatenate to the whole stream (no collective)."""
    from birdnest.audio_b200 import _abi
    s = streams(name)
    want = s.pcm * s.tiles
    parts = []
    nframes = 0
    for i in range(nshards):
        with _abi.open_memory(s.flac, shard_index=i, shard_count=nshards) as h:
            parts.append(h.decode_all())
            nframes += len(h.frames())
    assert b"".join(parts) == want
    assert nframes == len(s.frame_bs)


def test_device_resident_input_and_output(streams):
    import torch
    from birdnest.audio_b200 import _abi
    s = streams("cfg2_24bit_stereo_lpc12")
    want = s.pcm * s.tiles
    dev = torch.device("cuda:0")
    buf = torch.zeros(len(s.flac) + 64, dtype=torch.uint8, device=dev)
    buf[:len(s.flac)] = torch.frombuffer(bytearray(s.flac), dtype=torch.uint8).to(dev)
    out = torch.empty(len(want) + 16, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    with _abi.open_device(buf.data_ptr(), len(s.flac), s.flac[:65536], keep=buf, stream=torch.cuda.current_stream().cuda_stream) as h:
        for _ in range(3):     # repeated passes over resident input must give the same bytes
            out.zero_()
            ptr, n = h.decode_device(out.data_ptr(), out.numel())
            torch.cuda.synchronize()
            assert ptr == out.data_ptr() and n == len(want)
            assert bytes(out[:n].cpu().numpy()) == want
        t = h.timing()
        assert t.total > 0


def test_capacity_error(streams):
    from birdnest.audio_b200 import _abi
    s = streams("cfg1_16bit_stereo_lpc8")
    with _abi.open_memory(s.flac) as h:
        small = bytearray(1000)
        with pytest.raises(_abi.BnflacError) as e:
            h.decode_all(small)
        assert e.value.code == _abi.ERR_CAPACITY


@pytest.mark.parametrize("lean", ["0", "1"])
@pytest.mark.parametrize("name", ["cfg2_24bit_stereo_lpc12", "cfg3_24bit_8ch_lpc32_rice2_po8", "cfg5_6ch_special", "mono_special_escape_verbatim",
                                  "tiled_variable", "bps20_4ch_odd_bs_zeropart", "tiny_blocks"])
def test_both_parse_kernel_variants(streams, monkeypatch, name, lean):
    """k_parse exists with a looped and with a branch-free predicated ring refill; the launcher picks by frame count (the
    predicated one below ~38,000 frames).  Both must produce the oracle's PCM and subframe table on the same streams, intact
    and damaged."""
    import pyoracle
    from birdnest.audio_b200 import _abi
    monkeypatch.setenv("BNFLAC_PARSE_LEAN", lean)
    s = streams(name)
    for damage in (False, True):
        blob = bytearray(s.flac)
        if damage:
            n = len(blob)
            for pos, mask in ((s.frame_off[0] + (n - s.frame_off[0]) // 3, 0x04), (n - (n - s.frame_off[0]) // 4, 0x80)):
                blob[pos] ^= mask
        blob = bytes(blob)
        want, oframes, _, oerrs = pyoracle.decode(blob, want_frames=True)
        with _abi.open_memory(blob) as h:
            out = bytearray(len(want) + (1 << 20))
            k = h.decode_all(out)
            frames, errs = h.frames(), h.errors()
        assert bytes(out[:k]) == want
        assert errs == oerrs
        assert [(f.offset, f.length, f.status != 0) for f in frames] == [(o.offset, o.length, o.status != 0) for o in oframes]


@pytest.mark.parametrize("spec", ["0", "1"])
@pytest.mark.parametrize("name", ["cfg2_24bit_stereo_lpc12", "cfg3_24bit_8ch_lpc32_rice2_po8", "cfg5_6ch_special", "stereo_escape_lpc", "force_side_right",
                                  "tiled_variable", "bps20_4ch_odd_bs_zeropart", "ch5_24bit", "ch7_16bit", "tiny_blocks", "bps12_sihdr_padding"])
def test_speculative_parse_equals_the_serial_parse(streams, monkeypatch, name, spec):
    """Streams of few frames: subframe starts are guessed, walked in parallel and kept when they chain up from channel 0's
    known start (kernels.cu, "speculative parse"); what does not chain up goes to the serial walk.  Forced on and off, intact
    and damaged, both must give the oracle's PCM, frame table, subframe table and events."""
    import pyoracle
    from birdnest.audio_b200 import _abi
    monkeypatch.setenv("BNFLAC_PARSE_SPEC", spec)
    s = streams(name)
    for damage in (False, True):
        blob = bytearray(s.flac)
        if damage:
            n = len(blob)
            for pos, mask in ((s.frame_off[0] + (n - s.frame_off[0]) // 3, 0x10), (n - (n - s.frame_off[0]) // 5, 0x01)):
                blob[pos] ^= mask
        blob = bytes(blob)
        want, oframes, osubs, oerrs = pyoracle.decode(blob, want_frames=True)
        with _abi.open_memory(blob) as h:
            out = bytearray(len(want) + (1 << 20))
            k = h.decode_all(out)
            frames, subs, errs = h.frames(), h.subframes(), h.errors()
        assert bytes(out[:k]) == want
        assert errs == oerrs
        assert [(f.offset, f.length, f.status != 0) for f in frames] == [(o.offset, o.length, o.status != 0) for o in oframes]
        for fi, (f, of, osl) in enumerate(zip(frames, oframes, osubs)):
            if of.status:
                continue
            got = [(x.bit_offset, x.type, x.order, x.wasted) for x in subs[8 * fi:8 * fi + f.channels]]
            assert got == [(o.bit_offset - of.offset * 8, o.type, o.order, o.wasted) for o in osl], (name, fi)

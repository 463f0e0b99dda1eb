"""CPU suite for the host-only entry points of the C ABI (no device, no decode): `bnflac_probe` -- the metadata parse behind
SetupStreamInfo + MetadataCallback (FLACDecoder.cs:66-70,431-473) -- against the reference-DLL golden records and the oracle,
and `bnflac_ogg_to_native` -- the Ogg FLAC de-pager -- against streams muxed by tests/oggmux.py."""
import json
import os
import random

import pytest

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden")
golden = json.load(open(os.path.join(GOLD, "golden.json")))


@pytest.mark.parametrize("name", sorted(golden["fixtures"]) + sorted(golden["metadata"]))
def test_probe_matches_the_reference_and_the_oracle(name):
    import pyoracle
    from birdnest.audio_b200 import _abi
    blob = open(os.path.join(GOLD, name + ".flac"), "rb").read()
    g = golden["fixtures"].get(name) or golden["metadata"][name]
    info = _abi.probe(blob)
    si = pyoracle.streaminfo(blob)
    # what the reference DLL reported through its metadata callback (oracle/make_golden.py)
    assert info.channels == g.get("channels", g.get("ch")) and info.bits_per_sample == g["bps"]
    assert info.sample_rate == g.get("sample_rate", g.get("sr")) and info.total_samples == g.get("total_samples", g.get("total"))
    assert bytes(info.md5).hex() == g["si_md5"] or g["si_md5"] in (None, "")
    # and the oracle's view of the same STREAMINFO block
    for f in ("min_blocksize", "max_blocksize", "min_framesize", "max_framesize", "sample_rate", "channels", "bits_per_sample",
              "total_samples", "first_frame_offset"):
        assert getattr(info, f) == getattr(si, f), f
    assert bytes(info.md5) == bytes(si.md5)
    # derived fields (FLACDecoder.cs:448-452; FLACFileReader's byte layout)
    assert info.bytes_per_sample == (info.bits_per_sample + 7) // 8
    assert info.pcm_bytes == info.total_samples * info.channels * info.bytes_per_sample
    assert abs(info.duration_seconds - info.total_samples / info.sample_rate) < 1e-9
    if info.bits_per_sample == 16:                                # FLACDecoder.cs:454-465
        want_al = 0x1103 if info.channels == 2 else 0x1101
    elif info.bits_per_sample == 8:
        want_al = 0x1102 if info.channels == 2 else 0x1100
    else:
        want_al = 0
    assert info.al_format == want_al
    # the metadata alone is enough; one byte less than that is reported as truncated
    head = blob[:info.first_frame_offset]
    assert _abi.probe(head).total_samples == info.total_samples
    with pytest.raises(_abi.BnflacError) as e:
        _abi.probe(head[:-1])
    assert e.value.code == _abi.ERR_TRUNCATED


def test_probe_rejects_what_is_not_flac():
    from birdnest.audio_b200 import _abi
    for blob in (b"RIFF0000WAVEfmt " * 4, b"OggS" + bytes(60), b"\x00" * 64):
        with pytest.raises(_abi.BnflacError) as e:
            _abi.probe(blob)
        assert e.value.code == _abi.ERR_NOT_FLAC


@pytest.mark.parametrize("max_segs", [255, 40, 3, 1])
def test_ogg_pages_come_apart_into_the_native_stream(streams, max_segs):
    """Whatever the page layout (full pages, packets spanning pages, one segment per page, a foreign logical stream in
    between), what is left is byte for byte the native stream the packets were cut from."""
    from oggmux import mux
    from birdnest.audio_b200 import _abi
    for case in ("cfg4_clip_stereo_var", "cfg1_16bit_stereo_lpc8", "bps8_3ch"):
        s = streams(case)
        if max_segs == 1 and len(s.flac) > 300000:
            continue
        pages = mux(s, random.Random(max_segs), max_segs=max_segs, other_serial=0xABCD if max_segs == 40 else None)
        blob = b"".join(pages)
        assert _abi.ogg_to_native(blob) == s.flac
        info, want = _abi.probe(blob), _abi.probe(s.flac)
        assert (info.channels, info.bits_per_sample, info.total_samples, bytes(info.md5)) == (want.channels, want.bits_per_sample, want.total_samples, bytes(want.md5))
        # the header pages alone are enough for the probe
        assert _abi.probe(b"".join(pages[:3])).total_samples == want.total_samples


def test_ogg_lost_corrupt_and_out_of_order_pages(streams):
    from oggmux import native_packets, page
    from birdnest.audio_b200 import _abi
    s = streams("cfg4_clip_mono_fixed")
    headers, frames = native_packets(s)
    assert len(frames) >= 8
    lac = lambda b: [255] * (len(b) // 255) + [len(b) % 255]
    pages = [page(7, 0, 2, 0, lac(headers[0]), headers[0])]
    for hp in headers[1:]:
        pages.append(page(7, len(pages), 0, 0, lac(hp), hp))
    first_audio = len(pages)
    for i, f in enumerate(frames):
        pages.append(page(7, len(pages), 4 if i + 1 == len(frames) else 0, i, lac(f), f))
    head = s.flac[:s.frame_off[0]]
    assert _abi.ogg_to_native(b"".join(pages)) == s.flac
    # a missing page and a page with a flipped payload bit each cost exactly their frame
    dmg = list(pages)
    b = bytearray(dmg[first_audio + 5]); b[-3] ^= 1; dmg[first_audio + 5] = bytes(b)
    del dmg[first_audio + 2]
    assert _abi.ogg_to_native(b"".join(dmg)) == head + b"".join(f for i, f in enumerate(frames) if i not in (2, 5))
    # garbage between pages is skipped; a truncated last page costs its frame
    junk = list(pages)
    junk.insert(first_audio + 1, b"Ogg but not a page, then OggS\x01 with a bad version" + bytes(40))
    blob = b"".join(junk)
    assert _abi.ogg_to_native(blob) == s.flac
    assert _abi.ogg_to_native(blob[:-5]) == head + b"".join(frames[:-1])
    # a packet continued over two pages whose first half is lost: the rest of it is dropped, the next packet survives
    f0, f1 = frames[0], frames[1]
    assert len(f0) > 300
    cut = 255 * max(1, len(f0) // 510)
    p_a = page(7, first_audio, 0, 0, [255] * (cut // 255), f0[:cut])
    p_b = page(7, first_audio + 1, 1, 0, lac(f0[cut:]) + lac(f1), f0[cut:] + f1)
    assert _abi.ogg_to_native(b"".join(pages[:first_audio] + [p_a, p_b])) == head + f0 + f1
    assert _abi.ogg_to_native(b"".join(pages[:first_audio] + [p_b])) == head + f1
    # capacity protocol
    import ctypes as C
    n = C.c_size_t()
    small = bytearray(10)
    assert _abi.lib().bnflac_ogg_to_native(blob, len(blob), _abi._addr(small), len(small), C.byref(n)) == _abi.ERR_CAPACITY
    assert n.value == len(s.flac) and bytes(small) == bytes(10)
    assert _abi.lib().bnflac_ogg_to_native(s.flac, len(s.flac), None, 0, C.byref(n)) == _abi.ERR_NOT_FLAC


def test_depager_and_probe_survive_mutated_input(streams):
    """The container layer reads untrusted bytes: random mutations (bit flips, overwritten runs, deletions, insertions,
    truncation) of an Ogg FLAC stream and of a native header must yield an error code or a well-formed result, never a
    crash; what comes out of the de-pager can only be bytes that went in."""
    from oggmux import mux
    from birdnest.audio_b200 import _abi
    s = streams("cfg4_clip_stereo_var")
    ogg = b"".join(mux(s, random.Random(3), max_segs=9, other_serial=5))
    native = _abi.ogg_to_native(ogg)
    assert native == s.flac
    rng = random.Random(99)
    outcomes = {"ok": 0, "err": 0}
    for it in range(3000):
        src = ogg if it % 3 else s.flac[:s.frame_off[0] + 200]
        b = bytearray(src)
        for _ in range(rng.randrange(1, 4)):
            kind = rng.randrange(5)
            p = rng.randrange(len(b))
            if kind == 0:
                b[p] ^= 1 << rng.randrange(8)
            elif kind == 1:
                n = rng.randrange(1, 40); b[p:p + n] = bytes(rng.getrandbits(8) for _ in range(n))
            elif kind == 2:
                del b[p:p + rng.randrange(1, 300)]
            elif kind == 3:
                b[p:p] = bytes(rng.getrandbits(8) for _ in range(rng.randrange(1, 300)))
            else:
                del b[p:]
            if not b:
                b = bytearray(b"O")
        blob = bytes(b)
        try:
            info = _abi.probe(blob)
            assert 1 <= info.channels <= 8 and info.first_frame_offset <= len(blob) + (1 << 24)
            outcomes["ok"] += 1
        except _abi.BnflacError as e:
            assert e.code in (_abi.ERR_NOT_FLAC, _abi.ERR_TRUNCATED, _abi.ERR_UNSUPPORTED)
            outcomes["err"] += 1
        if blob[:4] == b"OggS":
            try:
                out = _abi.ogg_to_native(blob)
                assert out[:4] == b"fLaC" and len(out) <= len(blob)
            except _abi.BnflacError as e:
                assert e.code in (_abi.ERR_NOT_FLAC, _abi.ERR_UNSUPPORTED)
    assert outcomes["ok"] > 300 and outcomes["err"] > 50


ogg_golden = json.load(open(os.path.join(GOLD, "golden_ogg.json")))


@pytest.mark.parametrize("name", sorted(k for k in ogg_golden if k != "note"))
def test_depager_plus_oracle_equal_the_reference_dll_on_ogg(name):
    """The reference's LibFlac.dll decoded these pages through FLAC__stream_decoder_init_ogg_stream (oracle/make_golden_ogg.py);
    the library's de-pager followed by the CPU oracle must produce the same PCM, frame count and (absence of) error events --
    including the stream with a missing page, a corrupt page and junk between pages."""
    import hashlib
    import pyoracle
    from birdnest.audio_b200 import _abi
    g = ogg_golden[name]
    blob = open(os.path.join(GOLD, name + ".oga"), "rb").read()
    native = _abi.ogg_to_native(blob)
    pcm, nframes, _, errs = pyoracle.decode(native)
    assert hashlib.md5(pcm).hexdigest() == g["pcm_md5"] and len(pcm) == g["bytes"]
    assert nframes == g["frames"] and errs == g["errors"]
    info = _abi.probe(blob)
    assert (info.channels, info.bits_per_sample, info.sample_rate, info.total_samples) == (g["channels"], g["bps"], g["sample_rate"], g["total_samples"])
    assert bytes(info.md5).hex() == g["si_md5"]

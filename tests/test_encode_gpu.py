"""the GPU FLAC encoder (bnflac_encode / bnflac_encode_device, csrc/encoder.cu) -- the encoder half of the native codec
that LibFLACSharp.cs:322-387 declares.  An encoder is free in its choices, so parity is the round trip: every stream it writes must
decode back to the input PCM bit for bit through (1) the oracle, (2) the GPU decoder, and (3) where 32-bit binaries run, the
reference's own LibFlac.dll; STREAMINFO must carry md5(PCM); and it must not compress worse than the CPU corpus encoder."""
import hashlib
import os
import sys

import pytest

from conftest import CASES, ROOT, has_gpu
from test_encode_emu_cpu import CASES as SMALL, _ref_decode

pytestmark = pytest.mark.gpu

# pycorpus shapes -> encoder settings (the PCM of the decode corpus, encoded again on the GPU)
SHAPES = ["cfg1_16bit_stereo_lpc8", "cfg2_24bit_stereo_lpc12", "cfg3_24bit_8ch_lpc32_rice2_po8", "cfg4_clip_mono_fixed", "cfg5_6ch_special",
          "bps8_3ch", "bps12_sihdr_padding", "bps20_4ch_odd_bs_zeropart", "ch5_24bit", "ch7_16bit", "tiny_blocks", "silence_mono", "short_single_frame"]


def _roundtrip(pcm, ch, bps, sr, **kw):
    import pyoracle
    from birdnest.audio_b200 import _abi
    o = _abi.enc_opts(sr, ch, bps, **kw)
    flac, st = _abi.encode(pcm, o, want_stats=True)
    got, nframes, _, errs = pyoracle.decode(flac)
    assert got == pcm and not list(errs), "oracle"
    with _abi.open_memory(flac) as h:
        assert bytes(h.decode_all()) == pcm, "GPU decoder"
        info = h.info()
    assert bytes(info.md5) == hashlib.md5(pcm).digest()
    assert (info.channels, info.bits_per_sample, info.sample_rate) == (ch, bps, sr)
    assert st.frames == nframes and st.bytes == len(flac)
    assert info.min_framesize == st.min_framesize and info.max_framesize == st.max_framesize
    ref = _ref_decode(flac)
    if ref is not None:
        assert ref == pcm, "reference decoder (LibFlac.dll)"
    return flac, st


@pytest.mark.parametrize("name", sorted(SMALL))
def test_gpu_encoder_small_cases(name):
    if not has_gpu():
        pytest.skip("no CUDA device")
    pcm, ch, bps, sr, kw = SMALL[name]
    tr = dict(bs="blocksize", lpc="max_lpc_order", minpo="min_partition_order", maxpo="max_partition_order", prec="qlp_precision")
    args = {tr[k]: v for k, v in kw.items() if k in tr}
    from birdnest.audio_b200 import _abi
    flags = _abi.ENC_FIXED_ORDER if kw.get("search", 1) == 0 else 0
    _roundtrip(pcm, ch, bps, sr, flags=flags, **args)


@pytest.mark.parametrize("name", SHAPES)
def test_gpu_encoder_on_corpus_shapes(streams, name):
    if not has_gpu():
        pytest.skip("no CUDA device")
    s = streams(name)
    c = CASES[name]
    bs = c.get("bs", 4096)
    if bs > 16384:
        pytest.skip("encoder blocksize limit")
    flac, st = _roundtrip(s.pcm, s.channels, s.bps, s.sample_rate, blocksize=bs, max_lpc_order=c.get("lpc", 8),
                          min_partition_order=c.get("minpo", 0), max_partition_order=min(8, c.get("maxpo", 5)))
    # never worse than the CPU corpus encoder on the same PCM and settings (three Rice parameters per partition instead of two)
    assert len(flac) <= len(s.flac) * 1.01 + 64, (len(flac), len(s.flac))


def test_gpu_encoder_empty_input_is_a_metadata_only_stream():
    if not has_gpu():
        pytest.skip("no CUDA device")
    flac, st = _roundtrip(b"", 2, 16, 44100)
    assert len(flac) == 42 and st.frames == 0


def test_gpu_encoder_compression_levels_and_int32_input():
    if not has_gpu():
        pytest.skip("no CUDA device")
    import pycorpus
    import numpy as np
    from birdnest.audio_b200 import _abi
    s = pycorpus.make(ch=2, bps=16, sr=44100, seconds=3, bs=4096, lpc=8)
    sizes = []
    for lvl in (0, 3, 5, 8):
        flac, st = _roundtrip(s.pcm, 2, 16, 44100, compression_level=lvl, blocksize=0)
        sizes.append(len(flac))
    assert sizes[-1] <= sizes[0]
    # int32 container input (the layout of FLAC__stream_encoder_process_interleaved) gives the same stream
    pcm32 = np.frombuffer(s.pcm, dtype="<i2").astype("<i4").tobytes()
    o = _abi.enc_opts(44100, 2, 16, compression_level=5, blocksize=0, flags=_abi.ENC_INPUT_INT32)
    assert _abi.encode(pcm32, o) == _abi.encode(s.pcm, _abi.enc_opts(44100, 2, 16, compression_level=5, blocksize=0))


def test_gpu_encoder_device_to_device_and_decode_device():
    if not has_gpu():
        pytest.skip("no CUDA device")
    import torch
    import pycorpus
    import pyoracle
    from birdnest.audio_b200 import _abi
    s = pycorpus.make(ch=2, bps=24, sr=96000, seconds=5, bs=4096, lpc=12, maxpo=6)
    d_pcm = torch.frombuffer(bytearray(s.pcm), dtype=torch.uint8).cuda()
    o = _abi.enc_opts(96000, 2, 24, max_lpc_order=12)
    cap = _abi.encode_bound(len(s.pcm), o)
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    n, st = _abi.encode_device(d_pcm.data_ptr(), len(s.pcm), o, d_out.data_ptr(), cap)
    flac = bytes(d_out[:n].cpu().numpy())
    assert pyoracle.decode(flac)[0] == s.pcm
    assert bytes(pyoracle.streaminfo(flac).md5) == hashlib.md5(s.pcm).digest()
    assert st.total_ms > 0 and st.frames == (len(s.pcm) // 6 + 4095) // 4096
    # too small a destination is reported, not overrun
    with pytest.raises(_abi.BnflacError) as e:
        _abi.encode_device(d_pcm.data_ptr(), len(s.pcm), o, d_out.data_ptr(), 4096)
    assert e.value.code == _abi.ERR_CAPACITY

"""CPU suite: the C-ABI library loads, exports exactly what include/bnflac.h declares, and its host-side logic
(metadata parse, error vocabulary, no-device behaviour) works without a GPU.  No compute calls here."""
import ctypes as C
import io
import os
import re
import subprocess

import pytest

from conftest import ROOT, has_gpu


def test_header_symbols_are_exported_and_bound():
    from birdnest.audio_b200 import _abi
    hdr = open(os.path.join(ROOT, "include", "bnflac.h")).read()
    declared = set(re.findall(r"\b(bnflac_[a-z_0-9]+)\s*\(", hdr)) - {"bnflac_read_cb"}
    out = subprocess.check_output(["nm", "-D", "--defined-only", _abi.lib_path()], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert declared <= exported, declared - exported
    assert declared == set(_abi._PROTOS), declared ^ set(_abi._PROTOS)
    L = _abi.lib()
    assert L.bnflac_abi_version() == 1
    assert L.bnflac_strerror(-4).decode().startswith("no CUDA device")
    assert [L.bnflac_state_name(i).decode() for i in (2, 3, 4, 7)] == ["SearchForFrameSync", "ReadFrame", "EndOfStream", "Aborted"]
    assert [L.bnflac_frame_status_name(i).decode() for i in range(5)] == ["Ok", "LostSync", "BadHeader", "FrameCrcMismatch", "UnparsableStream"]


def test_no_torch_types_in_the_abi():
    hdr = open(os.path.join(ROOT, "include", "bnflac.h")).read()
    assert "torch" not in hdr and "at::" not in hdr and 'extern "C"' in hdr


def test_struct_layouts_match_the_header():
    from birdnest.audio_b200 import _abi
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "bnflac.h"
    int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(bnflac_opts), sizeof(bnflac_info_t), sizeof(bnflac_frame_t), sizeof(bnflac_subframe_t),
        sizeof(bnflac_timing), sizeof(bnflac_clip_result), offsetof(bnflac_info_t, md5), offsetof(bnflac_frame_t, pcm_offset)); return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])   # the header is plain C
        got = [int(x) for x in subprocess.check_output([os.path.join(d, "t")], text=True).split()]
    want = [C.sizeof(_abi.Opts), C.sizeof(_abi.Info), C.sizeof(_abi.FrameRec), C.sizeof(_abi.SubframeRec), C.sizeof(_abi.Timing),
            C.sizeof(_abi.ClipResult), _abi.Info.md5.offset, _abi.FrameRec.pcm_offset.offset]
    assert got == want


def test_open_rejects_non_flac_before_touching_a_device():
    from birdnest.audio_b200 import _abi
    for blob, code in ((b"RIFF0000WAVEfmt ", _abi.ERR_NOT_FLAC), (b"fLaC\x00\x00\x00\x22" + b"\x00" * 10, _abi.ERR_TRUNCATED), (b"fL", _abi.ERR_TRUNCATED)):
        with pytest.raises(_abi.BnflacError) as e:
            _abi.open_memory(blob)
        assert e.value.code == code


def test_valid_stream_without_device_fails_loudly(streams):
    from birdnest.audio_b200 import _abi
    if has_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(_abi.BnflacError) as e:
        _abi.open_memory(streams("cfg1_16bit_stereo_lpc8").flac)
    assert e.value.code == _abi.ERR_NO_DEVICE      # there is no CPU fallback to fall into


def test_flacdecoder_mirror_error_texts(streams):
    """FLACDecoder ctor failure texts (FLACDecoder.cs:66-70,98-105) and the closed/NotImplemented surface."""
    from birdnest.audio_b200 import FLACDecoder, FLACPacketQueue, EmptyStubLogger, ApplicationException
    with pytest.raises(ApplicationException) as e:
        FLACDecoder(io.BytesIO(b"this is not flac at all" * 10), FLACPacketQueue(), EmptyStubLogger())
    assert str(e.value) == "FLAC: Could not Could not process until end of metadata - EndOfStream!"


def test_library_has_no_oracle_or_cpu_decode_symbols():
    """The product must not link the oracle (a CPU fallback would void the parity claims)."""
    from birdnest.audio_b200 import _abi
    out = subprocess.check_output(["nm", "-D", _abi.lib_path()], text=True)
    assert "fo_decode" not in out and "fo_read_streaminfo" not in out and "bnc_encode" not in out


def test_legacy_shim_exports_the_symbols_the_csharp_binds():
    """include/bnflac_legacy.h == the decoder DllImports of LibFLACSharp.cs:42-85,175-185; libLibFlac.so exports every one,
    and the struct views have the offsets the C# marshals against (LibFLACSharp.cs:216-234, 295-319)."""
    shim = os.path.join(ROOT, "birdnest", "audio_b200", "libLibFlac.so")
    if not os.path.exists(shim):
        pytest.skip("libLibFlac.so not built (make shim)")
    hdr = open(os.path.join(ROOT, "include", "bnflac_legacy.h")).read()
    declared = set(re.findall(r"\b(FLAC__stream_decoder_[a-z_]+)\s*\(", hdr))
    assert {"FLAC__stream_decoder_new", "FLAC__stream_decoder_init_stream", "FLAC__stream_decoder_process_until_end_of_metadata",
            "FLAC__stream_decoder_process_single", "FLAC__stream_decoder_get_state", "FLAC__stream_decoder_finish", "FLAC__stream_decoder_delete",
            "FLAC__stream_decoder_init_file", "FLAC__stream_decoder_get_total_samples", "FLAC__stream_decoder_seek_absolute"} <= declared
    out = subprocess.check_output(["nm", "-D", "--defined-only", shim], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert declared <= exported, declared - exported
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "bnflac_legacy.h"
    int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(FLAC__FrameHeader), offsetof(FLAC__FrameHeader, bits_per_sample), offsetof(FLAC__FrameHeader, number),
        offsetof(FLAC__StreamMetadata, stream_info), offsetof(FLAC__StreamMetadata, stream_info.sample_rate), offsetof(FLAC__StreamMetadata, stream_info.bits_per_sample),
        offsetof(FLAC__StreamMetadata, stream_info.total_samples), offsetof(FLAC__StreamMetadata, stream_info.md5sum)); return 0; }'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        got = [int(x) for x in subprocess.check_output([os.path.join(d, "t")], text=True).split()]
    assert got == [40, 16, 24, 16, 32, 40, 48, 56]
    L = C.CDLL(shim)
    L.FLAC__stream_decoder_new.restype = C.c_void_p
    L.FLAC__stream_decoder_get_state.argtypes = [C.c_void_p]
    L.FLAC__stream_decoder_delete.argtypes = [C.c_void_p]
    dec = L.FLAC__stream_decoder_new()
    assert L.FLAC__stream_decoder_get_state(dec) == 9      # Uninitialized (LibFLACSharp.cs:36)
    L.FLAC__stream_decoder_delete(dec)


def test_encoder_entry_points_without_a_device():
    """argument checking and the size bound are host-only; encoding itself needs the device (no CPU encoder in the product)."""
    from birdnest.audio_b200 import _abi
    o = _abi.enc_opts(44100, 2, 16)
    assert _abi.encode_bound(4 * 4096 * 10, o) >= 4 * 4096 * 10
    for bad in (dict(channels=9), dict(bits_per_sample=32), dict(blocksize=65535), dict(max_lpc_order=33), dict(max_partition_order=9)):
        kw = dict(sample_rate=44100, channels=2, bits_per_sample=16)
        kw.update(bad)
        with pytest.raises(_abi.BnflacError) as e:
            _abi.encode_bound(1024, _abi.enc_opts(**kw))
        assert e.value.code == _abi.ERR_UNSUPPORTED
    with pytest.raises(_abi.BnflacError) as e:
        _abi.encode_bound(1023, o)          # not a whole number of stereo 16-bit samples
    assert e.value.code == _abi.ERR_ARG
    if not has_gpu():
        with pytest.raises(_abi.BnflacError) as e:
            _abi.encode(b"\0" * 4096, o)
        assert e.value.code == _abi.ERR_NO_DEVICE


def test_legacy_shim_exports_the_encoder_symbols_the_csharp_declares():
    """every FLAC__stream_encoder_* DllImport of LibFLACSharp.cs:322-373 is declared in include/bnflac_legacy.h and exported by
    libLibFlac.so; setters and init argument checks work without a device (encoding itself happens in finish(), on the GPU)."""
    shim = os.path.join(ROOT, "birdnest", "audio_b200", "libLibFlac.so")
    if not os.path.exists(shim):
        pytest.skip("libLibFlac.so not built (make shim)")
    want = {"FLAC__stream_encoder_" + n for n in (
        "new", "finish", "delete", "set_channels", "set_bits_per_sample", "set_sample_rate", "set_compression_level", "set_blocksize", "init_stream",
        "init_file", "process_interleaved", "process", "set_verify", "set_streamable_subset", "set_do_mid_side_stereo", "set_loose_mid_side_stereo", "get_state")}
    ref = "/root/reference/Library/LibFLACSharp/LibFLACSharp.cs"
    if os.path.exists(ref):      # the build container: the list above IS the reference's
        assert set(re.findall(r"\b(FLAC__stream_encoder_[a-z_]+)\s*\(", open(ref).read())) == want
    hdr = open(os.path.join(ROOT, "include", "bnflac_legacy.h")).read()
    assert set(re.findall(r"\b(FLAC__stream_encoder_[a-z_]+)\s*\(", hdr)) == want
    out = subprocess.check_output(["nm", "-D", "--defined-only", shim], text=True)
    assert want <= {l.split()[-1] for l in out.splitlines() if " T " in l}
    L = C.CDLL(shim)
    L.FLAC__stream_encoder_new.restype = C.c_void_p
    for n in ("get_state", "finish", "delete"):
        getattr(L, "FLAC__stream_encoder_" + n).argtypes = [C.c_void_p]
    for n in ("set_channels", "set_bits_per_sample", "set_sample_rate", "set_compression_level", "set_blocksize"):
        getattr(L, "FLAC__stream_encoder_" + n).argtypes = [C.c_void_p, C.c_uint]
    L.FLAC__stream_encoder_init_stream.argtypes = [C.c_void_p] + [C.c_void_p] * 5
    e = L.FLAC__stream_encoder_new()
    assert L.FLAC__stream_encoder_get_state(e) == 1                       # FLAC__STREAM_ENCODER_UNINITIALIZED
    assert L.FLAC__stream_encoder_set_channels(e, 9) and L.FLAC__stream_encoder_init_stream(e, None, None, None, None, None) == 3   # no write callback
    WRITE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint, C.c_uint, C.c_void_p)
    cb = WRITE(lambda *a: 0)
    assert L.FLAC__stream_encoder_init_stream(e, C.cast(cb, C.c_void_p), None, None, None, None) == 4    # INVALID_NUMBER_OF_CHANNELS
    assert L.FLAC__stream_encoder_set_channels(e, 2) and L.FLAC__stream_encoder_set_bits_per_sample(e, 16) and L.FLAC__stream_encoder_set_sample_rate(e, 44100)
    assert L.FLAC__stream_encoder_init_stream(e, C.cast(cb, C.c_void_p), None, None, None, None) == 0
    assert L.FLAC__stream_encoder_get_state(e) == 0 and not L.FLAC__stream_encoder_set_channels(e, 1)  # setters only before init
    assert L.FLAC__stream_encoder_init_stream(e, C.cast(cb, C.c_void_p), None, None, None, None) == 13   # ALREADY_INITIALIZED
    L.FLAC__stream_encoder_delete(e)


def test_encoder_struct_layouts_match_the_header():
    from birdnest.audio_b200 import _abi
    src = r"""
    #include <stdio.h>
    #include <stddef.h>
    #include "bnflac.h"
    int main(void){ printf("%zu %zu %zu %zu %zu\n", sizeof(bnflac_enc_opts), offsetof(bnflac_enc_opts, flags), offsetof(bnflac_enc_opts, first_frame_number),
        sizeof(bnflac_enc_stats), offsetof(bnflac_enc_stats, frame_sizes)); return 0; }"""
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        got = [int(x) for x in subprocess.check_output([os.path.join(d, "t")], text=True).split()]
    assert got == [C.sizeof(_abi.EncOpts), _abi.EncOpts.flags.offset, _abi.EncOpts.first_frame_number.offset, C.sizeof(_abi.EncStats), _abi.EncStats.frame_sizes.offset]

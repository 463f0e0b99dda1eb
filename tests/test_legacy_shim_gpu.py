"""libLibFlac.so -- the libFLAC 1.2.1 stream-decoder symbols the unmodified C# P/Invokes -- driven with exactly
FLACDecoder.cs's / FLACFileReader.cs's call sequences and struct offsets, against the oracle."""
import ctypes as C
import io
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "birdnest", "audio_b200", "libLibFlac.so")

READ_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_size_t), C.c_void_p)
SEEK_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_void_p)
TELL_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p)
LEN_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p)
EOF_CB = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p)
WRITE_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.POINTER(C.c_int32)), C.c_void_p)
META_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p)
ERR_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_void_p)


def _lib():
    if not os.path.exists(SHIM):
        pytest.skip("libLibFlac.so not built (make shim)")
    L = C.CDLL(SHIM)
    L.FLAC__stream_decoder_new.restype = C.c_void_p
    for name in ("delete", "finish", "process_single", "process_until_end_of_metadata", "process_until_end_of_stream", "get_state", "get_channels",
                 "get_bits_per_sample", "get_sample_rate", "reset"):
        getattr(L, "FLAC__stream_decoder_" + name).argtypes = [C.c_void_p]
    L.FLAC__stream_decoder_get_total_samples.argtypes = [C.c_void_p]
    L.FLAC__stream_decoder_get_total_samples.restype = C.c_uint64
    L.FLAC__stream_decoder_seek_absolute.argtypes = [C.c_void_p, C.c_uint64]
    L.FLAC__stream_decoder_init_stream.argtypes = [C.c_void_p, READ_CB, SEEK_CB, TELL_CB, LEN_CB, EOF_CB, WRITE_CB, META_CB, ERR_CB, C.c_void_p]
    L.FLAC__stream_decoder_init_file.argtypes = [C.c_void_p, C.c_char_p, WRITE_CB, META_CB, ERR_CB, C.c_void_p]
    return L


class Client:
    """What FLACDecoder.cs does in its callbacks, in Python: ReadCallback :325-363, MetadataCallback :431-473 (struct offsets of
    LibFLACSharp.cs:295-319), WriteCallback :520-580 generalised to N channels like FLACFileReader.cs:208-254."""

    def __init__(self, data: bytes, bytes_per_sample: int):
        self.src = io.BytesIO(data)
        self.B = bytes_per_sample
        self.pcm = bytearray()
        self.errors = []
        self.meta = None
        self.headers = []
        self.buf = bytearray(16384)        # mInstreamBuffer (DEFAULT_MAX_BUFFER_SIZE)

        def read(dec, buf, nbytes, client):
            want = min(nbytes[0], len(self.buf))
            chunk = self.src.read(want)
            C.memmove(buf, chunk, len(chunk))
            nbytes[0] = len(chunk)
            return 1 if len(chunk) < want else 0       # short read => EndOfStream

        def write(dec, frame, planes, client):
            hdr = (C.c_uint32 * 6).from_address(frame)   # blocksize@0 sample_rate@4 channels@8 assignment@12 bps@16 number_type@20
            bs, sr, ch, assign, bps, ntype = list(hdr)
            num = C.c_uint64.from_address(frame + 24).value
            self.headers.append((bs, sr, ch, assign, bps, ntype, num))
            out = bytearray(bs * ch * self.B)
            for c in range(ch):
                col = planes[c]
                for t in range(bs):
                    v = col[t] & 0xFFFFFFFF
                    o = (t * ch + c) * self.B
                    out[o:o + self.B] = v.to_bytes(4, "little")[:self.B]
            self.pcm += out
            return 0

        def meta(dec, md, client):
            raw = (C.c_uint8 * 72).from_address(md)
            u32 = lambda o: int.from_bytes(bytes(raw[o:o + 4]), "little")
            self.meta = dict(type=u32(0), min_bs=u32(16), max_bs=u32(20), sample_rate=u32(32), channels=u32(36), bps=u32(40),
                             total=int.from_bytes(bytes(raw[48:56]), "little"), md5=bytes(raw[56:72]))

        def err(dec, status, client):
            self.errors.append(status)

        self.cbs = (READ_CB(read), SEEK_CB(lambda *a: 1), TELL_CB(lambda *a: 1), LEN_CB(lambda *a: 1), EOF_CB(lambda *a: 0), WRITE_CB(write), META_CB(meta), ERR_CB(err))


def _run_like_flacdecoder(L, blob, B):
    cl = Client(blob, B)
    dec = L.FLAC__stream_decoder_new()
    assert L.FLAC__stream_decoder_get_state(dec) == 9                         # Uninitialized
    assert L.FLAC__stream_decoder_init_stream(dec, *cl.cbs, None) == 0       # SetupFLACStream (:58-64)
    assert L.FLAC__stream_decoder_process_until_end_of_metadata(dec) == 1    # SetupStreamInfo (:66-70)
    guard = 0
    while L.FLAC__stream_decoder_get_state(dec) < 4:                         # RequestAnotherFLACPacket (:207-224)
        assert L.FLAC__stream_decoder_process_single(dec) == 1
        guard += 1
        assert guard < 100000
    assert L.FLAC__stream_decoder_get_state(dec) == 4                         # EndOfStream
    assert L.FLAC__stream_decoder_finish(dec) == 1                           # Dispose (:296-300)
    L.FLAC__stream_decoder_delete(dec)
    return cl


@pytest.mark.parametrize("name", ["cfg1_16bit_stereo_lpc8", "cfg2_24bit_stereo_lpc12", "cfg5_6ch_special", "tiled_variable", "bps12_sihdr_padding"])
def test_unmodified_call_sequence_gets_the_oracle_pcm(streams, name):
    import pyoracle
    L = _lib()
    s = streams(name)
    want, oframes, _, _ = pyoracle.decode(s.flac, want_frames=True)
    cl = _run_like_flacdecoder(L, s.flac, (s.bps + 7) // 8)
    assert bytes(cl.pcm) == want and cl.errors == []
    assert cl.meta["type"] == 0 and (cl.meta["sample_rate"], cl.meta["channels"], cl.meta["bps"], cl.meta["total"]) == (s.sample_rate, s.channels, s.bps, s.total_samples)
    assert cl.meta["md5"] == s.md5
    assert [(h[0], h[2], h[4]) for h in cl.headers] == [(o.blocksize, o.channels, o.bits_per_sample) for o in oframes]
    assert [h[3] for h in cl.headers] == [0 if o.channel_assignment < 8 else o.channel_assignment - 7 for o in oframes]


def test_error_callbacks_fire_where_the_reference_fires_them(streams):
    import pyoracle
    L = _lib()
    s = streams("cfg1_16bit_stereo_lpc8")
    b = bytearray(s.flac)
    b[41491] ^= 128          # golden fault crc16_byte: frame delivered zero-filled + FrameCrcMismatch
    b[16548] ^= 64           # golden fault header_blocksize_bits: frame dropped, BadHeader + LostSync
    want, nframes, _, oerrs = pyoracle.decode(bytes(b))
    cl = _run_like_flacdecoder(L, bytes(b), 2)
    assert bytes(cl.pcm) == want and len(cl.headers) == nframes and cl.errors == oerrs


def test_init_file_total_samples_and_seek_absolute(streams, tmp_path):
    """FLACFileReader.cs:53-77 (init_file, get_total_samples) and :298 (seek_absolute delivers from the target sample on)."""
    L = _lib()
    s = streams("cfg1_16bit_stereo_lpc8")
    path = tmp_path / "a.flac"
    path.write_bytes(s.flac)
    cl = Client(b"", 2)
    dec = L.FLAC__stream_decoder_new()
    assert L.FLAC__stream_decoder_init_file(dec, str(path).encode(), cl.cbs[5], cl.cbs[6], cl.cbs[7], None) == 0
    assert L.FLAC__stream_decoder_process_until_end_of_metadata(dec) == 1
    assert L.FLAC__stream_decoder_get_total_samples(dec) == s.total_samples
    assert (L.FLAC__stream_decoder_get_channels(dec), L.FLAC__stream_decoder_get_bits_per_sample(dec), L.FLAC__stream_decoder_get_sample_rate(dec)) == (2, 16, 44100)
    target = 3 * 4096 + 1234
    assert L.FLAC__stream_decoder_seek_absolute(dec, target) == 1
    assert cl.headers[0][0] == 4096 - 1234 and cl.headers[0][5] == 1 and cl.headers[0][6] == target
    assert L.FLAC__stream_decoder_process_until_end_of_stream(dec) == 1
    assert bytes(cl.pcm) == s.pcm[target * 4:]
    assert L.FLAC__stream_decoder_seek_absolute(dec, s.total_samples + 5) == 0 and L.FLAC__stream_decoder_get_state(dec) == 6   # SeekError
    L.FLAC__stream_decoder_finish(dec)
    L.FLAC__stream_decoder_delete(dec)


# ---------------------------------------------------------------------------------------------- encoder half
ENC_WRITE_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint8), C.c_size_t, C.c_uint, C.c_uint, C.c_void_p)
ENC_META_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p)


def _enc_lib():
    L = _lib()
    L.FLAC__stream_encoder_new.restype = C.c_void_p
    for n in ("get_state", "finish", "delete"):
        getattr(L, "FLAC__stream_encoder_" + n).argtypes = [C.c_void_p]
    for n in ("set_channels", "set_bits_per_sample", "set_sample_rate", "set_compression_level", "set_blocksize"):
        getattr(L, "FLAC__stream_encoder_" + n).argtypes = [C.c_void_p, C.c_uint]
    for n in ("set_verify", "set_streamable_subset", "set_do_mid_side_stereo", "set_loose_mid_side_stereo"):
        getattr(L, "FLAC__stream_encoder_" + n).argtypes = [C.c_void_p, C.c_int32]
    L.FLAC__stream_encoder_init_stream.argtypes = [C.c_void_p, ENC_WRITE_CB, C.c_void_p, C.c_void_p, ENC_META_CB, C.c_void_p]
    L.FLAC__stream_encoder_init_file.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p]
    L.FLAC__stream_encoder_process_interleaved.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_uint]
    L.FLAC__stream_encoder_process.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_int32)), C.c_uint]
    return L


@pytest.mark.parametrize("planar", [False, True])
def test_legacy_encoder_symbols_drive_the_gpu_encoder(planar, tmp_path):
    """The libFLAC call sequence the declarations of LibFLACSharp.cs:322-387 are for: new, setters, init_stream, process* in pieces,
    finish, delete.  The bytes handed to the write callback are a FLAC stream that decodes to the samples, in libFLAC's call pattern."""
    import hashlib
    import numpy as np
    import pycorpus
    import pyoracle
    L = _enc_lib()
    s = pycorpus.make(ch=2, bps=16, sr=44100, samples=4096 * 5 + 777, bs=4096, lpc=8)
    x = np.frombuffer(s.pcm, dtype="<i2").astype(np.int32)            # interleaved int32, what process_interleaved takes
    calls, meta = [], []

    def write(enc, buf, n, samples, frame, client):
        calls.append((bytes(bytearray(buf[:n])), samples, frame))
        return 0

    def metadata(enc, m, client):
        f = (C.c_uint32 * 12).from_address(m)
        meta.append((f[0], f[8], f[9], f[10], C.c_uint64.from_address(m + 48).value, bytes((C.c_uint8 * 16).from_address(m + 56))))
    wcb, mcb = ENC_WRITE_CB(write), ENC_META_CB(metadata)
    e = L.FLAC__stream_encoder_new()
    assert L.FLAC__stream_encoder_set_channels(e, 2) and L.FLAC__stream_encoder_set_bits_per_sample(e, 16) and L.FLAC__stream_encoder_set_sample_rate(e, 44100)
    assert L.FLAC__stream_encoder_set_compression_level(e, 5) and L.FLAC__stream_encoder_set_verify(e, 1) and L.FLAC__stream_encoder_set_do_mid_side_stereo(e, 1)
    assert L.FLAC__stream_encoder_init_stream(e, wcb, None, None, mcb, None) == 0
    step = 1000                                                        # inter-channel samples per call
    for at in range(0, len(x) // 2, step):
        n = min(step, len(x) // 2 - at)
        if planar:
            l = np.ascontiguousarray(x[2 * at:2 * (at + n):2]); r = np.ascontiguousarray(x[2 * at + 1:2 * (at + n):2])
            ptrs = (C.POINTER(C.c_int32) * 2)(l.ctypes.data_as(C.POINTER(C.c_int32)), r.ctypes.data_as(C.POINTER(C.c_int32)))
            assert L.FLAC__stream_encoder_process(e, ptrs, n)
        else:
            chunk = np.ascontiguousarray(x[2 * at:2 * (at + n)])
            assert L.FLAC__stream_encoder_process_interleaved(e, chunk.ctypes.data_as(C.POINTER(C.c_int32)), n)
    assert L.FLAC__stream_encoder_finish(e) and L.FLAC__stream_encoder_get_state(e) == 1
    L.FLAC__stream_encoder_delete(e)
    assert calls[0] == (b"fLaC", 0, 0) and len(calls[1][0]) == 38 and calls[1][1:] == (0, 0)
    assert [c[1:] for c in calls[2:]] == [(4096, i) for i in range(5)] + [(777, 5)]
    flac = b"".join(c[0] for c in calls)                 # no seek callback: the STREAMINFO in the stream stays the provisional one of init
    got, nframes, _, errs = pyoracle.decode(flac)
    assert got == s.pcm and nframes == 6 and not list(errs)
    assert meta == [(0, 44100, 2, 16, 4096 * 5 + 777, hashlib.md5(s.pcm).digest())]
    # init_file writes the same frames to disk, with the final STREAMINFO in place (it can seek)
    e = L.FLAC__stream_encoder_new()
    L.FLAC__stream_encoder_set_channels(e, 2); L.FLAC__stream_encoder_set_bits_per_sample(e, 16); L.FLAC__stream_encoder_set_sample_rate(e, 44100)
    p = str(tmp_path / "o.flac").encode()
    assert L.FLAC__stream_encoder_init_file(e, p, None, None) == 0
    assert L.FLAC__stream_encoder_process_interleaved(e, np.ascontiguousarray(x).ctypes.data_as(C.POINTER(C.c_int32)), len(x) // 2)
    assert L.FLAC__stream_encoder_finish(e)
    L.FLAC__stream_encoder_delete(e)
    disk = open(p, "rb").read()
    assert disk[42:] == flac[42:] and disk[:4] == b"fLaC"
    si = pyoracle.streaminfo(disk)
    assert bytes(si.md5) == hashlib.md5(s.pcm).digest() and si.total_samples == 4096 * 5 + 777 and si.min_framesize > 0


def test_legacy_encoder_emits_while_processing_and_rewrites_streaminfo_through_seek():
    """Frames leave through the write callback during process() (every 1024 blocks are encoded in one GPU call, numbered on from the last), and a
    sink that can seek gets the final STREAMINFO written over the provisional one -- libFLAC's behaviour."""
    import hashlib
    import numpy as np
    import pycorpus
    import pyoracle
    L = _enc_lib()
    SEEK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_void_p)
    L.FLAC__stream_encoder_init_stream.argtypes = [C.c_void_p, ENC_WRITE_CB, SEEK, C.c_void_p, ENC_META_CB, C.c_void_p]
    s = pycorpus.make(ch=1, bps=16, sr=8000, samples=256 * 2500 + 10, bs=256, lpc=4, maxpo=3)
    x = np.frombuffer(s.pcm, dtype="<i2").astype(np.int32)
    sink, pos, progress = bytearray(), [0], []

    def write(enc, buf, n, samples, frame, client):
        data = bytes(bytearray(buf[:n]))
        sink[pos[0]:pos[0] + n] = data
        pos[0] += n
        return 0

    def seek(enc, off, client):
        pos[0] = off
        return 0
    wcb, scb, mcb = ENC_WRITE_CB(write), SEEK(seek), ENC_META_CB(lambda *a: None)
    e = L.FLAC__stream_encoder_new()
    L.FLAC__stream_encoder_set_channels(e, 1); L.FLAC__stream_encoder_set_bits_per_sample(e, 16); L.FLAC__stream_encoder_set_sample_rate(e, 8000)
    L.FLAC__stream_encoder_set_blocksize(e, 256); L.FLAC__stream_encoder_set_compression_level(e, 3)
    assert L.FLAC__stream_encoder_init_stream(e, wcb, scb, None, mcb, None) == 0
    for at in range(0, len(x), 50000):
        chunk = np.ascontiguousarray(x[at:at + 50000])
        assert L.FLAC__stream_encoder_process_interleaved(e, chunk.ctypes.data_as(C.POINTER(C.c_int32)), len(chunk))
        progress.append(len(sink))
    assert progress[0] == 42 and progress[-1] > 42       # frames left through the callback while samples were still being handed in
    assert L.FLAC__stream_encoder_finish(e)
    L.FLAC__stream_encoder_delete(e)
    flac = bytes(sink)
    got, nframes, _, errs = pyoracle.decode(flac)
    assert got == s.pcm and nframes == 2501 and not list(errs)
    si = pyoracle.streaminfo(flac)
    assert bytes(si.md5) == hashlib.md5(s.pcm).digest() and si.total_samples == 256 * 2500 + 10
    from birdnest.audio_b200 import _abi
    with _abi.open_memory(flac) as h:
        assert bytes(h.decode_all()) == s.pcm

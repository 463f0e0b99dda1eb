"""ctypes binding of include/bnflac.h -- the same entry points the C# P/Invoke shim binds.

The library is loaded lazily from this directory (libbnflac.so, built in-tree by `make lib` /
__graft_entry__.build()).  A missing library is a hard error: there is no Python or CPU decode path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB: Optional[C.CDLL] = None

STATE_NAMES = ["SearchForMetadata", "ReadMetadata", "SearchForFrameSync", "ReadFrame", "EndOfStream", "OggError",
               "SeekError", "Aborted", "MemoryAllocationError", "Uninitialized"]
AL_FORMAT_NAMES = {0: "None", 0x1100: "Mono8", 0x1101: "Mono16", 0x1102: "Stereo8", 0x1103: "Stereo16"}

ERR_ARG, ERR_NOT_FLAC, ERR_TRUNCATED, ERR_NO_DEVICE, ERR_CUDA, ERR_MEMORY, ERR_CAPACITY, ERR_ABORTED, ERR_UNSUPPORTED, ERR_STATE = range(-1, -11, -1)
OPT_VERIFY_MD5 = 1
OPT_LAZY_PULL = 4      # open_callbacks pulls the metadata only; bnflac_read pulls the rest on demand
OPT_BORROW_INPUT = 2
OPT_PACKED_INPUT = 8   # decode_batch: the clips are views into ONE caller-owned buffer, uploaded in place


class Opts(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("stream", C.c_void_p), ("shard_index", C.c_uint32),
                ("shard_count", C.c_uint32), ("flags", C.c_uint32), ("read_chunk_frames", C.c_uint32)]


class Info(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("channels", C.c_uint32), ("bits_per_sample", C.c_uint32),
                ("min_blocksize", C.c_uint32), ("max_blocksize", C.c_uint32), ("min_framesize", C.c_uint32), ("max_framesize", C.c_uint32),
                ("block_align", C.c_uint32), ("al_format", C.c_uint32), ("bytes_per_sample", C.c_uint32),
                ("total_samples", C.c_uint64), ("pcm_bytes", C.c_uint64), ("length_reference", C.c_uint64),
                ("duration_seconds", C.c_double), ("md5", C.c_uint8 * 16), ("first_frame_offset", C.c_uint64)]


class FrameRec(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("length", C.c_uint32), ("blocksize", C.c_uint32), ("channels", C.c_uint8),
                ("bits_per_sample", C.c_uint8), ("assignment", C.c_uint8), ("status", C.c_uint8), ("pad", C.c_uint32),
                ("number", C.c_uint64), ("pcm_offset", C.c_uint64)]


class SubframeRec(C.Structure):
    _fields_ = [("bit_offset", C.c_uint32), ("type", C.c_uint8), ("order", C.c_uint8), ("wasted", C.c_uint8), ("flags", C.c_uint8)]


class Timing(C.Structure):
    _fields_ = [("total", C.c_float), ("scan", C.c_float), ("crc", C.c_float), ("link", C.c_float), ("parse", C.c_float),
                ("decode", C.c_float), ("launches", C.c_uint32), ("pad", C.c_uint32)]


class Span(C.Structure):
    _fields_ = [("data", C.c_void_p), ("len", C.c_size_t)]


class ClipResult(C.Structure):
    _fields_ = [("pcm_offset", C.c_uint64), ("pcm_bytes", C.c_uint64), ("sample_rate", C.c_uint32), ("channels", C.c_uint32),
                ("bits_per_sample", C.c_uint32), ("status", C.c_uint32), ("total_samples", C.c_uint64)]


class EncOpts(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("stream", C.c_void_p), ("sample_rate", C.c_uint32), ("channels", C.c_uint32),
                ("bits_per_sample", C.c_uint32), ("blocksize", C.c_uint32), ("max_lpc_order", C.c_uint32), ("qlp_precision", C.c_uint32),
                ("min_partition_order", C.c_uint32), ("max_partition_order", C.c_uint32), ("mid_side", C.c_uint32),
                ("compression_level", C.c_uint32), ("flags", C.c_uint32), ("first_frame_number", C.c_uint64)]


class EncStats(C.Structure):
    _fields_ = [("plan_ms", C.c_float), ("write_ms", C.c_float), ("total_ms", C.c_float), ("frames", C.c_uint32), ("bytes", C.c_uint64),
                ("min_framesize", C.c_uint32), ("max_framesize", C.c_uint32), ("frame_sizes", C.POINTER(C.c_uint32)), ("frame_sizes_cap", C.c_uint64)]


ENC_NO_MD5, ENC_INPUT_INT32, ENC_USE_LEVEL, ENC_FIXED_ORDER = 1, 2, 4, 8

READ_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_size_t))

# every symbol include/bnflac.h declares (tests check the export list against the header)
_PROTOS = {
    "bnflac_open_memory": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(Opts), C.POINTER(C.c_void_p)]),
    "bnflac_open_callbacks": (C.c_int, [READ_CB, C.c_void_p, C.POINTER(Opts), C.POINTER(C.c_void_p)]),
    "bnflac_open_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(Opts), C.POINTER(C.c_void_p)]),
    "bnflac_shard_range": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "bnflac_info": (C.c_int, [C.c_void_p, C.POINTER(Info)]),
    "bnflac_state": (C.c_int, [C.c_void_p]),
    "bnflac_close": (None, [C.c_void_p]),
    "bnflac_read": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "bnflac_decode_all": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64)]),
    "bnflac_decode_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "bnflac_decoded_size": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "bnflac_decode_batch": (C.c_int, [C.POINTER(Span), C.c_size_t, C.POINTER(Opts), C.c_void_p, C.c_size_t, C.c_int, C.POINTER(ClipResult), C.POINTER(C.c_uint64)]),
    "bnflac_encode_bound": (C.c_int, [C.c_size_t, C.POINTER(EncOpts), C.POINTER(C.c_uint64)]),
    "bnflac_encode": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(EncOpts), C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(EncStats)]),
    "bnflac_encode_device": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(EncOpts), C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(EncStats)]),
    "bnflac_frames": (C.c_int, [C.c_void_p, C.POINTER(C.POINTER(FrameRec)), C.POINTER(C.c_size_t)]),
    "bnflac_subframes": (C.c_int, [C.c_void_p, C.POINTER(C.POINTER(SubframeRec)), C.POINTER(C.c_size_t)]),
    "bnflac_probe": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(Info)]),
    "bnflac_ogg_to_native": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "bnflac_errors": (C.c_int, [C.c_void_p, C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_size_t)]),
    "bnflac_errors_so_far": (C.c_int, [C.c_void_p, C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_size_t)]),
    "bnflac_error_frames": (C.c_int, [C.c_void_p, C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_size_t)]),
    "bnflac_last_timing": (C.c_int, [C.c_void_p, C.POINTER(Timing)]),
    "bnflac_strerror": (C.c_char_p, [C.c_int]),
    "bnflac_state_name": (C.c_char_p, [C.c_int]),
    "bnflac_frame_status_name": (C.c_char_p, [C.c_int]),
    "bnflac_last_cuda_error": (C.c_char_p, []),
    "bnflac_abi_version": (C.c_int, []),
    "bnflac_device_count": (C.c_int, []),
    "bnflac_kernel_launches": (C.c_uint64, []),
    "bnflac_trim_pools": (None, []),
}


def lib_path() -> str:
    return os.environ.get("BNFLAC_LIB", os.path.join(_HERE, "libbnflac.so"))


def lib() -> C.CDLL:
    """Load libbnflac.so (fails loudly if it has not been built: no fallback exists)."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise RuntimeError(f"{path} not found: build it with `make lib` or __graft_entry__.build(); "
                               "birdnest.audio_b200 has no CPU/Python decode path")
        L = C.CDLL(path)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


class BnflacError(RuntimeError):
    def __init__(self, code: int, where: str):
        L = lib()
        msg = L.bnflac_strerror(code).decode()
        if code == ERR_CUDA:
            msg += ": " + L.bnflac_last_cuda_error().decode()
        super().__init__(f"{where}: {msg} ({code})")
        self.code = code


def _check(rc: int, where: str) -> None:
    if rc != 0:
        raise BnflacError(rc, where)


def _opts(device=-1, stream=0, shard_index=0, shard_count=0, flags=0, read_chunk_frames=0) -> Opts:
    o = Opts()
    o.struct_size = C.sizeof(Opts)
    o.device = device
    o.stream = stream or None
    o.shard_index = shard_index
    o.shard_count = shard_count
    o.flags = flags
    o.read_chunk_frames = read_chunk_frames
    return o


def _addr(buf) -> int:
    """Address of a writable/readable buffer: bytes, bytearray, numpy array, torch tensor (cpu) or int."""
    if isinstance(buf, int):
        return buf
    if hasattr(buf, "data_ptr"):
        return buf.data_ptr()
    if hasattr(buf, "ctypes"):
        return buf.ctypes.data
    if isinstance(buf, (bytes,)):
        return C.cast(C.c_char_p(buf), C.c_void_p).value
    return C.addressof((C.c_uint8 * len(buf)).from_buffer(buf))


class Handle:
    """Thin owner of a bnflac_t*."""

    def __init__(self, ptr: int, keep=None):
        self._p = C.c_void_p(ptr)
        self._keep = keep

    def close(self):
        if self._p:
            lib().bnflac_close(self._p)
            self._p = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def info(self) -> Info:
        i = Info()
        _check(lib().bnflac_info(self._p, C.byref(i)), "bnflac_info")
        return i

    def state(self) -> int:
        return lib().bnflac_state(self._p)

    def read_into(self, buf, count: Optional[int] = None) -> int:
        n = len(buf) if count is None else count
        r = lib().bnflac_read(self._p, _addr(buf), n)
        if r < 0:
            raise BnflacError(int(r), "bnflac_read")
        return int(r)

    def decode_all(self, buf=None) -> bytes | int:
        """Decode to host memory.  With buf=None allocates a bytearray of the needed size and returns it."""
        if buf is None:
            for attempt in range(2):
                size = self.decoded_size()
                out = bytearray(size)
                w = C.c_uint64()
                rc = lib().bnflac_decode_all(self._p, _addr(out), size, C.byref(w)) if size else 0
                if rc != ERR_CAPACITY or attempt:      # STREAMINFO understated the stream: decoded_size now scans (include/bnflac.h)
                    break
            _check(rc, "bnflac_decode_all")
            return bytes(out[:w.value])
        w = C.c_uint64()
        n = buf.numel() * buf.element_size() if hasattr(buf, "numel") else len(buf)
        _check(lib().bnflac_decode_all(self._p, _addr(buf), n, C.byref(w)), "bnflac_decode_all")
        return int(w.value)

    def decoded_size(self) -> int:
        w = C.c_uint64()
        _check(lib().bnflac_decoded_size(self._p, C.byref(w)), "bnflac_decoded_size")
        return int(w.value)

    def decode_device(self, d_dst: int = 0, cap: int = 0):
        """Decode leaving the PCM in device memory; returns (device_ptr, bytes)."""
        out = C.c_void_p()
        w = C.c_uint64()
        _check(lib().bnflac_decode_device(self._p, d_dst or None, cap, C.byref(out), C.byref(w)), "bnflac_decode_device")
        return out.value, int(w.value)

    def frames(self):
        p = C.POINTER(FrameRec)()
        n = C.c_size_t()
        _check(lib().bnflac_frames(self._p, C.byref(p), C.byref(n)), "bnflac_frames")
        return [FrameRec.from_buffer_copy(p[i]) for i in range(n.value)]

    def subframes(self):
        p = C.POINTER(SubframeRec)()
        n = C.c_size_t()
        _check(lib().bnflac_subframes(self._p, C.byref(p), C.byref(n)), "bnflac_subframes")
        return [SubframeRec.from_buffer_copy(p[i]) for i in range(n.value)]

    def errors(self):
        p = C.POINTER(C.c_uint32)()
        n = C.c_size_t()
        _check(lib().bnflac_errors(self._p, C.byref(p), C.byref(n)), "bnflac_errors")
        return [int(p[i]) for i in range(n.value)]

    def errors_so_far(self):
        """Events of what a streamed Read session has decoded so far (no decode / pull ahead); else == errors()."""
        p = C.POINTER(C.c_uint32)()
        n = C.c_size_t()
        _check(lib().bnflac_errors_so_far(self._p, C.byref(p), C.byref(n)), "bnflac_errors_so_far")
        return [int(p[i]) for i in range(n.value)]

    def timing(self) -> Timing:
        t = Timing()
        _check(lib().bnflac_last_timing(self._p, C.byref(t)), "bnflac_last_timing")
        return t


def open_memory(data, device=-1, stream=0, shard_index=0, shard_count=0, flags=0, read_chunk_frames=0) -> Handle:
    o = _opts(device, stream, shard_index, shard_count, flags, read_chunk_frames)
    h = C.c_void_p()
    n = data.numel() * data.element_size() if hasattr(data, "numel") else len(data)
    _check(lib().bnflac_open_memory(_addr(data), n, C.byref(o), C.byref(h)), "bnflac_open_memory")
    return Handle(h.value, keep=data if (flags & OPT_BORROW_INPUT) else None)


def open_device(d_ptr: int, length: int, header: bytes, device=-1, stream=0, shard_index=0, shard_count=0, keep=None) -> Handle:
    """bnflac_open_device: the stream's bytes are already in device memory.  The allocation must be readable for 64 bytes past
    d_ptr + length (allocate length + 64; contents irrelevant) -- see include/bnflac.h."""
    o = _opts(device, stream, shard_index, shard_count, 0)
    h = C.c_void_p()
    _check(lib().bnflac_open_device(d_ptr, length, _addr(header), len(header), C.byref(o), C.byref(h)), "bnflac_open_device")
    return Handle(h.value, keep=(keep, header))


def shard_range(stream_len: int, first_frame_offset: int, index: int, count: int):
    """bnflac_shard_range: [own_begin, own_end) of shard `index` of `count` (host-only; the engine's own arithmetic)."""
    b, e = C.c_uint64(), C.c_uint64()
    _check(lib().bnflac_shard_range(stream_len, first_frame_offset, index, count, C.byref(b), C.byref(e)), "bnflac_shard_range")
    return int(b.value), int(e.value)


def probe(data) -> Info:
    """bnflac_probe: host-only metadata parse (native FLAC, ID3v2-prefixed, or Ogg FLAC); no device needed."""
    info = Info()
    _check(lib().bnflac_probe(_addr(data), len(data), C.byref(info)), "bnflac_probe")
    return info


def ogg_to_native(data) -> bytes:
    """bnflac_ogg_to_native: the native FLAC stream inside an Ogg FLAC stream (host-only)."""
    n = C.c_size_t()
    rc = lib().bnflac_ogg_to_native(_addr(data), len(data), None, 0, C.byref(n))
    if rc not in (0, ERR_CAPACITY):
        _check(rc, "bnflac_ogg_to_native")
    out = bytearray(n.value)
    _check(lib().bnflac_ogg_to_native(_addr(data), len(data), _addr(out), len(out), C.byref(n)), "bnflac_ogg_to_native")
    return bytes(out[:n.value])


def open_callbacks(read_fn, device=-1, flags=0, read_chunk_frames=0) -> Handle:
    """read_fn(n) -> bytes (b'' at end of stream); mirrors the pull model of FLACDecoder.ReadCallback.
    flags=OPT_LAZY_PULL: only the metadata is pulled here, the rest as read_into() advances."""
    spill = [b""]          # what a read_fn returned beyond the room the engine offered, served first by the next call

    def _cb(user, buf, nbytes):
        want = nbytes[0]
        chunk = spill[0]
        if not chunk:
            try:
                chunk = read_fn(want)
            except Exception:
                return 2
            if chunk is None:
                return 2
        k = min(len(chunk), want)          # never write past the room the engine reserved
        spill[0] = bytes(chunk[k:])
        if k:
            C.memmove(buf, bytes(chunk[:k]) if k < len(chunk) else chunk, k)
        nbytes[0] = k
        return 1 if (k == 0 and want) else 0       # end of stream only on an empty read: sockets and pipes return short reads
    cb = READ_CB(_cb)
    o = _opts(device, flags=flags, read_chunk_frames=read_chunk_frames)
    h = C.c_void_p()
    _check(lib().bnflac_open_callbacks(cb, None, C.byref(o), C.byref(h)), "bnflac_open_callbacks")
    return Handle(h.value, keep=cb)        # the callback object must outlive the handle (lazy pull calls it from bnflac_read)


class BatchTable:
    """The bnflac_span table of a batch (and the result array the library fills), built once for a list of clips that is decoded
    more than once: filling 100,000 ctypes structs from Python takes ~4 us each -- host-language marshalling, not part of the C ABI
    (a C# or C++ caller hands over an array it already has)."""

    def __init__(self, clips):
        n = len(clips)
        self.n = n
        self.spans = (Span * max(1, n))()
        self.keep = list(clips)                       # the table points into these buffers
        for i, c in enumerate(clips):
            self.spans[i].data = _addr(c)
            self.spans[i].len = c.numel() * c.element_size() if hasattr(c, "numel") else len(c)
        self.res = (ClipResult * max(1, n))()


def decode_batch(clips, device=-1, dst=None, dst_is_device=False, packed=False):
    """bnflac_decode_batch: many independent clips in one pipeline pass per (channels, bits) group (BASELINE cfg4).
    clips: sequence of bytes-like, or a BatchTable built from one.  Returns (pcm, results) with pcm = bytes (dst=None), or the byte
    count written into `dst` (a writable host buffer / torch tensor; a device pointer or CUDA tensor with dst_is_device=True).
    packed=True: the clips are ascending views into ONE buffer the caller owns (BNFLAC_OPT_PACKED_INPUT: uploaded in place)."""
    tab = clips if isinstance(clips, BatchTable) else BatchTable(clips)
    n, spans, res = tab.n, tab.spans, tab.res
    o = _opts(device, flags=OPT_PACKED_INPUT if packed else 0)
    w = C.c_uint64()
    if dst is None:
        _check(lib().bnflac_decode_batch(spans, n, C.byref(o), None, 0, 0, res, C.byref(w)), "bnflac_decode_batch(size)")
        out = bytearray(int(w.value))
        if w.value:
            _check(lib().bnflac_decode_batch(spans, n, C.byref(o), _addr(out), len(out), 0, res, C.byref(w)), "bnflac_decode_batch")
        return bytes(out[:w.value]), [ClipResult.from_buffer_copy(res[i]) for i in range(n)]
    cap = dst.numel() * dst.element_size() if hasattr(dst, "numel") else len(dst)
    _check(lib().bnflac_decode_batch(spans, n, C.byref(o), _addr(dst), cap, 1 if dst_is_device else 0, res, C.byref(w)), "bnflac_decode_batch")
    return int(w.value), res[:n]      # views into the result array (no per-clip copies: 100k clips per call are the target)


def enc_opts(sample_rate, channels, bits_per_sample, blocksize=4096, max_lpc_order=8, qlp_precision=0, min_partition_order=0,
             max_partition_order=6, mid_side=True, compression_level=None, flags=0, device=-1, stream=0, first_frame_number=0) -> EncOpts:
    """Settings of the GPU encoder -- what the reference's FLAC__stream_encoder_set_* calls carry (LibFLACSharp.cs:333-369)."""
    o = EncOpts()
    o.struct_size = C.sizeof(EncOpts)
    o.device, o.stream = device, stream or None
    o.sample_rate, o.channels, o.bits_per_sample = sample_rate, channels, bits_per_sample
    o.blocksize, o.max_lpc_order, o.qlp_precision = blocksize, max_lpc_order, qlp_precision
    o.min_partition_order, o.max_partition_order, o.mid_side = min_partition_order, max_partition_order, 1 if mid_side else 0
    if compression_level is not None:
        o.compression_level, flags = compression_level, flags | ENC_USE_LEVEL
    o.flags = flags
    o.first_frame_number = first_frame_number
    return o


def encode_bound(pcm_bytes: int, opts: EncOpts) -> int:
    b = C.c_uint64()
    _check(lib().bnflac_encode_bound(pcm_bytes, C.byref(opts), C.byref(b)), "bnflac_encode_bound")
    return b.value


def encode(pcm, opts: EncOpts, want_stats=False):
    """Host PCM (interleaved LE, ceil(bps/8) bytes per sample -- what the decoder returns) -> FLAC stream bytes, encoded on the GPU."""
    n = len(pcm)
    cap = encode_bound(n, opts)
    out = bytearray(cap)
    w, st = C.c_uint64(), EncStats()
    _check(lib().bnflac_encode(_addr(pcm) if n else None, n, C.byref(opts), _addr(out), cap, C.byref(w), C.byref(st)), "bnflac_encode")
    del out[w.value:]
    return (bytes(out), st) if want_stats else bytes(out)


def encode_device(d_pcm: int, pcm_bytes: int, opts: EncOpts, d_dst: int, cap: int):
    """Device PCM -> device FLAC stream; returns (bytes written, EncStats)."""
    w, st = C.c_uint64(), EncStats()
    _check(lib().bnflac_encode_device(d_pcm, pcm_bytes, C.byref(opts), d_dst, cap, C.byref(w), C.byref(st)), "bnflac_encode_device")
    return w.value, st

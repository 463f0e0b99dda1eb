"""ctypes view of oracle/_build/liboracle.so (flac_oracle.c).  TEST INFRASTRUCTURE: import only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_L = None


class StreamInfo(C.Structure):
    _fields_ = [("min_blocksize", C.c_uint32), ("max_blocksize", C.c_uint32), ("min_framesize", C.c_uint32), ("max_framesize", C.c_uint32),
                ("sample_rate", C.c_uint32), ("channels", C.c_uint32), ("bits_per_sample", C.c_uint32), ("total_samples", C.c_uint64),
                ("md5", C.c_uint8 * 16), ("first_frame_offset", C.c_uint64)]


class Frame(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("length", C.c_uint32), ("blocksize", C.c_uint32), ("channels", C.c_uint32),
                ("bits_per_sample", C.c_uint32), ("channel_assignment", C.c_uint32), ("sample_rate", C.c_uint32),
                ("variable", C.c_uint32), ("number", C.c_uint64), ("status", C.c_uint32)]


class Subframe(C.Structure):
    _fields_ = [("type", C.c_uint8), ("order", C.c_uint8), ("wasted", C.c_uint8), ("precision", C.c_uint8), ("shift", C.c_uint8),
                ("rice_method", C.c_uint8), ("partition_order", C.c_uint8), ("pad", C.c_uint8), ("bit_offset", C.c_uint64)]


def build():
    subprocess.check_call(["make", "-s", "-C", _ROOT, "oracle"])


def lib():
    global _L
    if _L is None:
        if not os.path.exists(_SO):
            build()
        _L = _bind(C.CDLL(_SO))
    return _L


def _bind(L):
    L.fo_read_streaminfo.restype = C.c_int
    L.fo_read_streaminfo.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(StreamInfo)]
    L.fo_decode.restype = C.c_int64
    L.fo_decode.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(Frame), C.c_size_t, C.POINTER(C.c_size_t),
                            C.POINTER(Subframe), C.POINTER(C.c_uint32), C.c_size_t, C.POINTER(C.c_size_t)]
    L.fo_decode_range.restype = C.c_int64
    L.fo_decode_range.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(StreamInfo), C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.fo_crc8.restype = C.c_uint8
    L.fo_crc8.argtypes = [C.c_char_p, C.c_size_t]
    L.fo_crc16.restype = C.c_uint16
    L.fo_crc16.argtypes = [C.c_char_p, C.c_size_t]
    L.fo_md5.restype = None
    L.fo_md5.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
    return L


def streaminfo(data: bytes) -> StreamInfo:
    si = StreamInfo()
    rc = lib().fo_read_streaminfo(data, len(data), C.byref(si))
    if rc:
        raise ValueError(f"not a FLAC stream ({rc})")
    return si


def decode(data: bytes, want_frames=False):
    """-> (pcm bytes, frames, subframes, errors)"""
    L = lib()
    need = L.fo_decode(data, len(data), None, 0, None, 0, None, None, None, 0, None)
    if need < 0:
        raise ValueError(f"oracle decode failed ({need})")
    pcm = C.create_string_buffer(int(need) + 1)
    nf = C.c_size_t()
    ne = C.c_size_t()
    errs = (C.c_uint32 * 4096)()
    frames = subs = None
    cap = 0
    if want_frames:
        n0 = C.c_size_t()
        L.fo_decode(data, len(data), None, 0, None, 0, C.byref(n0), None, None, 0, None)
        cap = n0.value + 1
        frames = (Frame * cap)()
        subs = (Subframe * (8 * cap))()
    got = L.fo_decode(data, len(data), pcm, int(need), frames, cap, C.byref(nf), subs, errs, 4096, C.byref(ne))
    assert got == need
    fl = [frames[i] for i in range(nf.value)] if want_frames else nf.value
    sl = [[subs[8 * i + c] for c in range(frames[i].channels)] for i in range(nf.value)] if want_frames else None
    return pcm.raw[:int(got)], fl, sl, [int(errs[i]) for i in range(min(ne.value, 4096))]


def md5(data: bytes) -> bytes:
    out = C.create_string_buffer(16)
    lib().fo_md5(data, len(data), out)
    return out.raw

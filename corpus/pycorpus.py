"""ctypes view of corpus/_build/libbncorpus.so: deterministic synthetic PCM + FLAC streams of the BASELINE shapes."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SO = os.path.join(_HERE, "_build", "libbncorpus.so")
_L = None


class Params(C.Structure):
    _fields_ = [("channels", C.c_uint32), ("bps", C.c_uint32), ("sample_rate", C.c_uint32), ("blocksize", C.c_uint32),
                ("nvar", C.c_uint32), ("var_bs", C.c_uint32 * 16), ("max_lpc_order", C.c_uint32), ("qlp_precision", C.c_uint32),
                ("min_part_order", C.c_uint32), ("max_part_order", C.c_uint32), ("stereo_mode", C.c_uint32), ("search_order", C.c_uint32),
                ("escape_every", C.c_uint32), ("verbatim_every", C.c_uint32), ("allow_zero_part", C.c_uint32),
                ("streaminfo_in_frames", C.c_uint32), ("no_md5", C.c_uint32), ("padding_bytes", C.c_uint32)]


class Stream(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("len", C.c_size_t), ("frame_off", C.POINTER(C.c_uint64)), ("nframes", C.c_size_t),
                ("frame_bs", C.POINTER(C.c_uint32)), ("total_samples", C.c_uint64), ("md5", C.c_uint8 * 16), ("first_frame", C.c_size_t)]


def build():
    subprocess.check_call(["make", "-s", "-C", _ROOT, "corpus"])


def lib():
    global _L
    if _L is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.bnc_synth.restype = None
        L.bnc_synth.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.bnc_encode.restype = C.c_int
        L.bnc_encode.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(Params), C.POINTER(Stream)]
        L.bnc_tile.restype = C.c_int
        L.bnc_tile.argtypes = [C.POINTER(Stream), C.c_uint32, C.c_char_p, C.c_size_t, C.c_int, C.POINTER(Stream)]
        L.bnc_to_variable.restype = C.c_int
        L.bnc_to_variable.argtypes = [C.POINTER(Stream), C.POINTER(Stream)]
        L.bnc_free.restype = None
        L.bnc_free.argtypes = [C.POINTER(Stream)]
        L.bnc_pack_pcm.restype = None
        L.bnc_pack_pcm.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]
        _L = L
    return _L


class Synth:
    """Result of make(): .flac (bytes), .pcm (packed LE bytes of ONE tile), .tiles, .frame_off, .frame_bs, .md5, .total_samples"""
    _stream = None

    def free(self):
        if self._stream is not None:
            self.flac = None
            lib().bnc_free(C.byref(self._stream))
            self._stream = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def make(ch=2, bps=16, sr=44100, samples=None, seconds=1.0, bs=4096, lpc=8, prec=0, minpo=0, maxpo=5, stereo=1, search=1,
         noise=None, kind=0, period=None, seed=2026, esc=0, verb=0, var=(), zeropart=0, sihdr=0, pad=0, tile=1, tovar=0, threads=8,
         want_pcm=True, md5=True, view=False):
    """md5=False: the tiled stream's STREAMINFO md5 is left zero (hashing tens of gigabytes of PCM takes longer than generating the
    stream; the benchmark compares the decoded PCM with the unique tile instead).  view=True: .flac is a numpy uint8 VIEW of the
    generator's own buffer (no copy; streams beyond 2 GiB), released by .free() or with the object."""
    L = lib()
    n = int(samples if samples is not None else seconds * sr)
    if tile > 1 and not var:
        n -= n % bs
    p = Params()
    p.channels, p.bps, p.sample_rate, p.blocksize = ch, bps, sr, bs
    p.nvar = len(var)
    for i, v in enumerate(var):
        p.var_bs[i] = v
    p.max_lpc_order, p.qlp_precision, p.min_part_order, p.max_part_order = lpc, prec, minpo, maxpo
    p.stereo_mode, p.search_order, p.escape_every, p.verbatim_every = stereo, search, esc, verb
    p.allow_zero_part, p.streaminfo_in_frames, p.padding_bytes = zeropart, sihdr, pad
    pcm32 = (C.c_int32 * (n * ch))()
    L.bnc_synth(pcm32, n, ch, bps, sr, noise if noise is not None else (12 if bps > 16 else 6), kind, period or bs, seed)
    s = Stream()
    if L.bnc_encode(pcm32, n, C.byref(p), C.byref(s)):
        raise RuntimeError("bnc_encode failed")
    B = (bps + 7) // 8
    packed = C.create_string_buffer(n * ch * B + 1)
    L.bnc_pack_pcm(pcm32, n * ch, bps, packed)
    del pcm32
    if tovar:
        v = Stream()
        if L.bnc_to_variable(C.byref(s), C.byref(v)):
            raise RuntimeError("bnc_to_variable failed")
        L.bnc_free(C.byref(s))
        s = v
    if tile > 1:
        t = Stream()
        rc = L.bnc_tile(C.byref(s), tile, packed if md5 else None, n * ch * B, threads, C.byref(t))
        if rc:
            raise RuntimeError(f"bnc_tile failed {rc}")
        L.bnc_free(C.byref(s))
        s = t
    r = Synth()
    if view:
        import numpy as np
        r.flac = np.ctypeslib.as_array(s.data, shape=(s.len,))
    else:
        r.flac = C.string_at(s.data, s.len)
    r.pcm = packed.raw[:n * ch * B] if want_pcm else None
    r.tiles = tile
    r.frame_off = [s.frame_off[i] for i in range(s.nframes + 1)]
    r.frame_bs = [s.frame_bs[i] for i in range(s.nframes)]
    r.md5 = bytes(s.md5)
    r.total_samples = int(s.total_samples)
    r.channels, r.bps, r.sample_rate = ch, bps, sr
    if view:
        r._stream = s                      # owns .flac's memory
    else:
        L.bnc_free(C.byref(s))
    return r

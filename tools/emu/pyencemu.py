"""ctypes view of tools/emu/_build/libencemu.so: the encoder KERNELS run on the CPU through the CUDA emulation shim (test tool)."""
import ctypes as C
import hashlib
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_SO = os.path.join(_HERE, "_build", "libencemu.so")
_L = None


def lib():
    global _L
    if _L is None:
        if not os.path.exists(_SO):
            subprocess.check_call(["make", "-s", "-C", _ROOT, "emu"])
        _L = C.CDLL(_SO)
        _L.enc_emu.restype = C.c_int
        _L.enc_emu.argtypes = [C.c_char_p, C.c_uint64] + [C.c_uint32] * 11 + [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    return _L


def streaminfo(bs, minfs, maxfs, sr, ch, bps, total, md5):
    x = (sr << 44) | ((ch - 1) << 41) | ((bps - 1) << 36) | (total & 0xFFFFFFFFF)
    return (b"fLaC" + bytes([0x80, 0, 0, 34]) + bs.to_bytes(2, "big") * 2 + minfs.to_bytes(3, "big") + maxfs.to_bytes(3, "big")
            + x.to_bytes(8, "big") + md5)


def encode(pcm: bytes, ch, bps, sr, bs=4096, lpc=8, prec=0, minpo=0, maxpo=6, stereo=1, search=1):
    B = (bps + 7) // 8
    total = len(pcm) // (B * ch)
    cap = 42 + len(pcm) * 2 + 4096 + (total // bs + 1) * 64
    out = C.create_string_buffer(cap)
    n, mn, mx = C.c_uint64(), C.c_uint32(), C.c_uint32()
    rc = lib().enc_emu(pcm, total, ch, bps, B, bs, sr, lpc, prec, minpo, maxpo, 1 if (stereo and ch == 2) else 0, search, out, cap,
                       C.byref(n), C.byref(mn), C.byref(mx))
    if rc:
        raise RuntimeError(f"enc_emu {rc}")
    return streaminfo(bs, mn.value, mx.value, sr, ch, bps, total, hashlib.md5(pcm).digest()) + out.raw[42:n.value]

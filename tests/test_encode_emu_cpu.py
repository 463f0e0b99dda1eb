"""SURVEY 8f-4, CPU side: the GPU encoder's KERNELS (csrc/encoder_kernels.cuh), compiled for the CPU through tools/emu/cuda_emu.h
(one pthread per CUDA thread, barriers for __syncthreads and the warp collectives), produce streams that the oracle -- and, where
it can run, the reference's own LibFlac.dll -- decode back to the input PCM bit for bit.  An encoder is free in its choices, so this
round trip IS the parity statement; the same checks run on the real device in tests/test_encode_gpu.py."""
import hashlib
import os
import random
import subprocess
import sys
import tempfile

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools", "emu"))


def _pack(samples, bps):
    B = (bps + 7) // 8
    return b"".join((s & ((1 << (8 * B)) - 1)).to_bytes(B, "little") for s in samples)


def _ref_decode(flac):
    exe, dll = os.path.join(ROOT, "oracle", "_ref", "refflac"), os.path.join(ROOT, "oracle", "_ref", "LibFlac.dll")
    if not (os.path.exists(exe) and os.path.exists(dll)):
        return None
    try:
        if subprocess.run([exe], capture_output=True, timeout=10).returncode != 2:
            return None
    except OSError:
        return None
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "i.flac"), "wb").write(flac)
        subprocess.check_call([exe, "dec", dll, os.path.join(d, "i.flac"), os.path.join(d, "o.pcm")], stdout=subprocess.DEVNULL)
        return open(os.path.join(d, "o.pcm"), "rb").read()


def _cases():
    rnd = random.Random(7)
    import math
    music16 = [int(9000 * math.sin(i * 0.031)) + int(3000 * math.sin(i * 0.0071)) + rnd.randrange(-40, 40) for i in range(2 * (4096 + 700))]
    music24 = [int(2000000 * math.sin(i * 0.011)) + rnd.randrange(-3000, 3000) for i in range(2 * 4096)]
    return {
        "stereo16_lpc8_short_last_frame": (_pack(music16, 16), 2, 16, 44100, dict(bs=4096, lpc=8, maxpo=5)),
        "stereo24_lpc12_wide": (_pack(music24, 24), 2, 24, 96000, dict(bs=4096, lpc=12, maxpo=6)),
        "mono16_fixed_only": (_pack(music16[:3000], 16), 1, 16, 48000, dict(bs=1152, lpc=0, maxpo=3)),
        "silence_and_dc_constant": (_pack([0] * 3000 + [77, -5] * 1500, 16), 2, 16, 44100, dict(bs=1024)),
        "full_scale_noise_verbatim": (_pack([rnd.randrange(-(1 << 23), 1 << 23) for _ in range(3000)], 24), 2, 24, 96000, dict(bs=1024, lpc=12)),
        "wasted_bits": (_pack([rnd.randrange(-2000, 2000) * 16 for _ in range(3000)], 16), 1, 16, 44100, dict(bs=1024)),
        "one_sample": (_pack([5, -5], 16), 2, 16, 44100, dict()),
        "blocksize16_odd_rate": (_pack([rnd.randrange(-100, 100) for _ in range(16 * 5 + 3)], 16), 1, 16, 12345, dict(bs=16, lpc=4, maxpo=2)),
        "rail_to_rail_24": (_pack([(-(1 << 23) if (i // 5) % 2 else (1 << 23) - 1) for i in range(4000)], 24), 2, 24, 44100, dict(lpc=12)),
        "lpc32_rice2_8ch_po8": (_pack([rnd.randrange(-(1 << 20), 1 << 20) + int(3000000 * math.sin(i * 0.002)) for i in range(8 * 8192)], 24), 8, 24, 192000,
                                dict(bs=8192, lpc=32, minpo=8, maxpo=8, search=0)),
    }


CASES = _cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_encoder_kernels_round_trip_through_oracle_and_reference(name):
    import pyencemu
    import pyoracle
    pcm, ch, bps, sr, kw = CASES[name]
    flac = pyencemu.encode(pcm, ch, bps, sr, **kw)
    got, nframes, _, errs = pyoracle.decode(flac)
    assert got == pcm and not list(errs)
    si = pyoracle.streaminfo(flac)
    assert (si.channels, si.bits_per_sample, si.sample_rate, si.total_samples) == (ch, bps, sr, len(pcm) // (ch * ((bps + 7) // 8)))
    assert bytes(si.md5) == hashlib.md5(pcm).digest()
    assert len(flac) <= len(pcm) + 42 + 32 * nframes          # VERBATIM is the worst case
    ref = _ref_decode(flac)
    if ref is not None:
        assert ref == pcm, "the reference decoder (LibFlac.dll) disagrees"


def test_encoder_kernels_compress_no_worse_than_the_reference_encoder():
    """Same PCM, same settings, the reference DLL's own encoder (FLAC__stream_encoder_*, driven by oracle/refdll `refflac enc`) beside the
    encoder kernels: the stream written here is not larger (1 % tolerance).  Runs where 32-bit binaries run."""
    import pycorpus
    import pyencemu
    exe, dll = os.path.join(ROOT, "oracle", "_ref", "refflac"), os.path.join(ROOT, "oracle", "_ref", "LibFlac.dll")
    if not (os.path.exists(exe) and os.path.exists(dll)):
        pytest.skip("oracle/_ref not built")
    try:
        if subprocess.run([exe], capture_output=True, timeout=10).returncode != 2:
            pytest.skip("32-bit binaries do not run on this host")
    except OSError:
        pytest.skip("32-bit binaries do not run on this host")
    for kw, enc in ((dict(ch=2, bps=24, sr=96000, samples=4096 * 8, bs=4096, lpc=12, maxpo=6), ("2", "24", "96000", "4096", "12", "0", "6", "1")),
                    (dict(ch=2, bps=16, sr=44100, samples=4096 * 8, bs=4096, lpc=8, maxpo=5), ("2", "16", "44100", "4096", "8", "0", "5", "1"))):
        s = pycorpus.make(**kw)
        with tempfile.TemporaryDirectory() as d:
            open(os.path.join(d, "i.pcm"), "wb").write(s.pcm)
            subprocess.check_call([exe, "enc", dll, *enc, "0", os.path.join(d, "i.pcm"), os.path.join(d, "o.flac")], stdout=subprocess.DEVNULL)
            ref = os.path.getsize(os.path.join(d, "o.flac"))
        ours = len(pyencemu.encode(s.pcm, kw["ch"], kw["bps"], kw["sr"], bs=kw["bs"], lpc=kw["lpc"], maxpo=kw["maxpo"]))
        assert ours <= ref * 1.01, (ours, ref)

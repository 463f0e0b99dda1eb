import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "corpus")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


def has_gpu() -> bool:
    try:
        from birdnest.audio_b200 import _abi
        return _abi.lib().bnflac_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def streams():
    """Small synthetic streams covering every construct of the path (SURVEY 8c/8d shapes at reduced length)."""
    import pycorpus
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = pycorpus.make(**CASES[name])
        return cache[name]
    return get


# name -> pycorpus.make kwargs.  Shapes follow BASELINE.json configs[0..4] at lengths the oracle finishes in seconds.
CASES = {
    "cfg1_16bit_stereo_lpc8": dict(ch=2, bps=16, sr=44100, seconds=4, bs=4096, lpc=8, maxpo=5),
    "cfg2_24bit_stereo_lpc12": dict(ch=2, bps=24, sr=96000, seconds=2, bs=4096, lpc=12, maxpo=6),
    "cfg3_24bit_8ch_lpc32_rice2_po8": dict(ch=8, bps=24, sr=192000, samples=16384 * 5, bs=16384, lpc=32, minpo=8, maxpo=8, noise=19, search=0),
    "cfg4_clip_mono_fixed": dict(ch=1, bps=16, sr=44100, seconds=0.7, bs=576, lpc=0),
    "cfg4_clip_stereo_var": dict(ch=2, bps=16, sr=44100, seconds=1.1, lpc=8, var=(4096, 1152, 4080, 720, 16, 192, 2304)),
    "cfg5_6ch_special": dict(ch=6, bps=24, sr=48000, seconds=1.5, bs=1152, lpc=8, kind=1, period=1152),
    "mono_special_escape_verbatim": dict(ch=1, bps=16, sr=44100, seconds=2, bs=576, lpc=0, kind=1, period=576, esc=5, verb=7),
    "stereo_escape_lpc": dict(ch=2, bps=16, sr=48000, seconds=1, bs=1024, lpc=6, esc=3),
    "force_left_side": dict(ch=2, bps=16, sr=44100, seconds=0.5, bs=1024, lpc=4, stereo=2),
    "force_side_right": dict(ch=2, bps=24, sr=44100, seconds=0.5, bs=1024, lpc=4, stereo=3),
    "force_mid_side": dict(ch=2, bps=24, sr=44100, seconds=0.5, bs=1024, lpc=4, stereo=4),
    "independent_stereo": dict(ch=2, bps=16, sr=44100, seconds=0.5, bs=1024, lpc=4, stereo=0),
    "bps8_3ch": dict(ch=3, bps=8, sr=8000, seconds=2, bs=256, lpc=4, noise=3),
    "bps12_sihdr_padding": dict(ch=2, bps=12, sr=22050, seconds=2, bs=1000, lpc=6, noise=4, sihdr=1, pad=100),
    "bps20_4ch_odd_bs_zeropart": dict(ch=4, bps=20, sr=37800, seconds=1, bs=777, lpc=10, noise=9, zeropart=1),
    "ch5_24bit": dict(ch=5, bps=24, sr=48000, seconds=0.5, bs=2048, lpc=12, noise=10),
    "ch7_16bit": dict(ch=7, bps=16, sr=48000, seconds=0.5, bs=4608, lpc=8),
    "tiny_blocks": dict(ch=2, bps=16, sr=8000, seconds=0.5, bs=16, lpc=2, maxpo=2),
    "silence_mono": dict(ch=1, bps=16, sr=44100, seconds=30, bs=4096, lpc=8, noise=0, kind=2),
    "tiled_fixed": dict(ch=2, bps=24, sr=96000, seconds=0.5, bs=4096, lpc=12, maxpo=6, tile=3),
    "tiled_variable": dict(ch=2, bps=16, sr=44100, seconds=0.5, bs=1024, lpc=4, tovar=1, tile=2),
    "odd_rate_bs_big": dict(ch=1, bps=24, sr=12345, seconds=3, bs=65535, lpc=16, maxpo=0, noise=8),
    "short_single_frame": dict(ch=2, bps=16, sr=44100, samples=100, bs=4096, lpc=8),
}

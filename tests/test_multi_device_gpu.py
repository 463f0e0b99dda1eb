"""One process, several devices (bnflac_opts.device): frame-range shards decoded on different GPUs through libbnflac and
concatenated (no collective), handles on two devices alive at the same time (function attributes and SM counts are
per device), and the caller's current device left as it was.  Skipped on boxes with one GPU."""
import hashlib
import os

import pytest

from conftest import ROOT, has_gpu

pytestmark = pytest.mark.gpu


def _ndev():
    from birdnest.audio_b200 import _abi
    return _abi.lib().bnflac_device_count()


@pytest.mark.parametrize("name", ["cfg2_24bit_stereo_lpc12", "cfg3_24bit_8ch_lpc32_rice2_po8", "cfg4_clip_stereo_var"])
def test_shards_on_two_devices_concatenate(streams, name):
    if _ndev() < 2:
        pytest.skip("needs two GPUs")
    import torch
    from birdnest.audio_b200 import _abi
    s = streams(name)
    torch.cuda.set_device(0)
    hs = [_abi.open_memory(s.flac, device=d, shard_index=d, shard_count=2) for d in (0, 1)]     # both alive at once
    try:
        parts = []
        for d in (1, 0):                                                                          # device 1 first: k_scan's opt-in shared memory on the second device
            out = bytearray(len(s.pcm) * s.tiles + 4096)
            k = hs[d].decode_all(out)
            parts.append((d, bytes(out[:k])))
            assert torch.cuda.current_device() == 0, "the caller's current device was switched"
        whole = b"".join(p for _, p in sorted(parts))
    finally:
        for h in hs:
            h.close()
    assert whole == s.pcm * s.tiles
    assert hashlib.md5(whole).digest() == s.md5 or s.tiles > 1


def test_device_resident_decode_on_the_second_device(streams):
    if _ndev() < 2:
        pytest.skip("needs two GPUs")
    import torch
    from birdnest.audio_b200 import _abi
    s = streams("cfg1_16bit_stereo_lpc8")
    torch.cuda.set_device(0)
    d_in = torch.zeros(len(s.flac) + 256, dtype=torch.uint8, device="cuda:1")
    d_in[:len(s.flac)] = torch.frombuffer(bytearray(s.flac), dtype=torch.uint8).to("cuda:1")
    d_out = torch.empty(len(s.pcm) + 256, dtype=torch.uint8, device="cuda:1")
    torch.cuda.synchronize(1)
    with _abi.open_device(d_in.data_ptr(), len(s.flac), s.flac[:65536], device=1, keep=d_in) as h:
        for _ in range(3):                                   # the second and third pass are launched without host hand-offs
            _, w = h.decode_device(d_out.data_ptr(), d_out.numel())
            assert bytes(d_out[:w].cpu().numpy()) == s.pcm
        assert torch.cuda.current_device() == 0


def test_first_decode_of_a_process_from_many_threads():
    """Handles are single-threaded, different handles may be used from different threads (include/bnflac.h) -- also when the very first
    decodes of a process start at the same moment: the once-per-device kernel attributes (k_scan's 226 KB of dynamic shared memory) must be
    in place before ANY thread launches.  A fresh process, eight threads, eight handles, first pass of each at once."""
    import subprocess
    import sys
    if not has_gpu():
        pytest.skip("no CUDA device")
    code = r'''
import sys, threading
sys.path[:0] = [%r, %r, %r]
import pycorpus
from birdnest.audio_b200 import _abi
s = pycorpus.make(ch=2, bps=24, sr=96000, seconds=1, bs=4096, lpc=12, maxpo=6)
hs = [_abi.open_memory(s.flac) for _ in range(8)]
out, bar = [None] * 8, threading.Barrier(8)
def work(i):
    bar.wait()
    out[i] = bytes(hs[i].decode_all())
ts = [threading.Thread(target=work, args=(i,)) for i in range(8)]
[t.start() for t in ts]; [t.join() for t in ts]
assert all(o == s.pcm for o in out), [None if o is None else len(o) for o in out]
print("ok")
''' % (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "corpus"))
    for _ in range(3):       # a race: give it a few chances to show
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]

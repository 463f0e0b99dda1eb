"""Balanced decode schedule (kernels_decode.cuh): jobs cut between decode warps and handed over through HBM.  On the benchmark streams
the cut happens once per slot; here BNFLAC_BALANCE_SLOTS (read once per process, hence the subprocesses) shrinks the launch to a few
warps so that every case stream -- all channel counts, sample widths, predictor orders, escape / VERBATIM / CONSTANT subframes,
wasted bits -- is cut many times, at every tile position, and must still decode bit for bit."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT, has_gpu

pytestmark = pytest.mark.gpu

CODE = r'''
import sys
sys.path[:0] = [%r, %r, %r, %r]
import pycorpus
from conftest import CASES
from birdnest.audio_b200 import _abi
bad = []
for name in sorted(CASES):
    s = pycorpus.make(**CASES[name])
    want = s.pcm * s.tiles
    with _abi.open_memory(s.flac) as h:
        got = h.decode_all()
        again = h.decode_all() if len(want) < (64 << 20) else got      # second pass: launched without host hand-offs
    if got != want or again != want:
        n = min(len(got), len(want)); d = next((i for i in range(0, n, 4096) if got[i:i + 4096] != want[i:i + 4096]), n)
        bad.append((name, len(got), len(want), d))
print("bad", bad)
print("ok" if not bad else "FAIL")
'''


@pytest.mark.parametrize("slots", [2, 6, 14])
def test_cut_jobs_decode_bit_exact(slots):
    if not has_gpu():
        pytest.skip("no CUDA device")
    code = CODE % (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "corpus"), os.path.join(ROOT, "tests"))
    env = dict(os.environ, BNFLAC_BALANCE_SLOTS=str(slots), BNFLAC_TRACE="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (r.stdout[-2000:], r.stderr[-2000:])
    assert "balanced over" in r.stderr, "the balanced schedule was never used"

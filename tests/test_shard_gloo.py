"""N>1 host logic on CPU (gloo, world_size 2): frame-range shard planning + the host-side gather.  The per-shard decode is
done by the ORACLE here (the checker), so the partition/ordering logic the GPU path shares is what is under test."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import CASES, ROOT


def _worker(rank, world, port, name, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "corpus"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes as C
    import pycorpus
    import pyoracle
    from birdnest.audio_b200.sharding import shard_ranges, owned_frames
    s = pycorpus.make(**CASES[name])
    si = pyoracle.streaminfo(s.flac)
    lo, hi = shard_ranges(len(s.flac), si.first_frame_offset, world)[rank]
    # frames whose sync lies in [lo, hi) belong to this rank (the rule the CUDA scan kernel applies via own_begin/own_end)
    mine = owned_frames(s.frame_off, lo, hi)
    pcm = b""
    if mine:
        b, e = s.frame_off[mine[0]], s.frame_off[mine[-1] + 1]
        need = sum(s.frame_bs[i] for i in mine) * si.channels * ((si.bits_per_sample + 7) // 8)
        buf = C.create_string_buffer(need + 1)
        nf = C.c_size_t()
        got = pyoracle.lib().fo_decode_range(s.flac, len(s.flac), C.byref(si), b, e, buf, need, C.byref(nf))
        assert got == need and nf.value == len(mine)
        pcm = buf.raw[:need]
    # host-side gather (no collective on the data path; this is the result hand-off only)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(pcm)], dtype=torch.int64))
    mx = int(max(t.item() for t in sizes))
    mine_t = torch.zeros(mx, dtype=torch.uint8)
    mine_t[:len(pcm)] = torch.frombuffer(bytearray(pcm), dtype=torch.uint8) if pcm else mine_t[:0]
    parts = [torch.zeros(mx, dtype=torch.uint8) for _ in range(world)] if rank == 0 else None
    dist.gather(mine_t, parts, dst=0)
    if rank == 0:
        whole = b"".join(bytes(parts[r][:int(sizes[r].item())].numpy()) for r in range(world))
        q.put(whole == s.pcm * s.tiles)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["cfg1_16bit_stereo_lpc8", "cfg4_clip_stereo_var", "tiled_fixed"])
def test_two_rank_frame_range_shards_gather_to_the_whole(name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_shard_ranges_partition_the_frame_bytes():
    """shard_ranges() is the engine's own arithmetic (bnflac_shard_range through the C ABI): one source for planner and engine."""
    from birdnest.audio_b200.sharding import shard_ranges, owned_frames
    from birdnest.audio_b200 import _abi
    assert _abi.shard_range(1_000_003, 4242, 1, 3) == (4242 + (1_000_003 - 4242) // 3, 4242 + (1_000_003 - 4242) * 2 // 3)
    assert _abi.shard_range(10, 4, 0, 0) == (4, 10)                      # count 0 = whole stream
    with pytest.raises(_abi.BnflacError):
        _abi.shard_range(10, 4, 3, 3)
    for n in (1, 2, 3, 4, 8):
        r = shard_ranges(1_000_003, 4242, n)
        assert r[0][0] == 4242 and r[-1][1] == 1_000_003
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
    offs = [100, 250, 251, 900, 1000]
    got = [owned_frames(offs, lo, hi) for lo, hi in shard_ranges(1000, 100, 3)]
    assert sorted(sum(got, [])) == [0, 1, 2, 3] and all(g == sorted(g) for g in got)

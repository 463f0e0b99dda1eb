"""Frame-range sharding plan: FLAC frames are independent and self-positioned, so a stream is split into
`count` byte ranges of its frame data; a shard owns the frames whose sync code lies inside its range, decodes them with
no data-path collective, and the host concatenates the PCM slices in shard order.

The arithmetic is the engine's own: shard_ranges() asks libbnflac (bnflac_shard_range, host-only, no device needed), the
same function bnflac_opts.shard_index / shard_count go through.
"""
from typing import List, Sequence, Tuple

from . import _abi


def shard_ranges(stream_len: int, first_frame_offset: int, count: int) -> List[Tuple[int, int]]:
    """[own_begin, own_end) of every shard, in stream order."""
    return [_abi.shard_range(stream_len, first_frame_offset, i, count) for i in range(count)]


def owned_frames(frame_offsets: Sequence[int], lo: int, hi: int) -> List[int]:
    """Indices of the frames whose first byte is in [lo, hi).  frame_offsets may carry a trailing end-of-stream offset."""
    return [i for i, o in enumerate(frame_offsets) if lo <= o < hi and (i + 1 < len(frame_offsets))]

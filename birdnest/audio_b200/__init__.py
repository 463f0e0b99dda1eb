"""birdnest.audio_b200 -- B200-native FLAC decode engine behind BirdNest.Audio's FLACDecoder Stream surface.

Only what the hot path needs lives here:
  csrc/           hand-written sm_100a CUDA kernels + the C-ABI host runtime (libbnflac.so)
  _abi.py         ctypes binding of include/bnflac.h (what the C# P/Invoke shim binds)
  flac_decoder.py Python mirror of the reference's `FLACDecoder : Stream` (FLACDecoder.cs:14-598) used by the tests
There is no CPU decode path: importing works anywhere, decoding needs the built library and a CUDA device.
"""
from ._abi import (BnflacError, Info, FrameRec, SubframeRec, Timing, lib, lib_path, open_memory, open_device,  # noqa: F401
                   Handle, AL_FORMAT_NAMES, STATE_NAMES)
from .flac_decoder import (FLACDecoder, FLACPacket, FLACPacketQueue, EmptyStubLogger, ALFormat,  # noqa: F401
                           ApplicationException)

__all__ = ["FLACDecoder", "FLACPacket", "FLACPacketQueue", "EmptyStubLogger", "ALFormat", "ApplicationException",
           "BnflacError", "Handle", "open_memory", "open_device", "lib", "lib_path"]

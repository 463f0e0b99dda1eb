"""Python mirror of the reference's `BirdNest.Audio.FLACDecoder : System.IO.Stream`.

Same names, argument meaning and error behaviour as Library/BirdNest.Audio/FLACDecoder.cs, so that the parity
tests read like tests of the reference class; the native codec underneath is libbnflac.so (CUDA, sm_100a)
instead of LibFlac.dll.  The C# replacement a maintainer would drop in is shown in INTEGRATION.md; the C++
host mirror is csrc/flac_decoder.hpp.

  reference member (FLACDecoder.cs)                         here
  ctor(Stream, IFLACPacketQueue, IFLACDecoderLogger[, byte[]])  :23,72   FLACDecoder(stream, queue, logger[, buffer])
  Read(byte[], int, int)                                       :124     Read(buffer, offset, count)
  Format / Channels / SampleRate / BitsPerSample / Duration    :426-430 same-named properties
  Length, CanRead, CanSeek, CanWrite                           :240-265 same
  Flush / Seek / SetLength / Write / Position                  :109-122,235-238,267-274  raise NotImplementedError
  Dispose (closes the inner stream)                            :285-319 Dispose() / close() / context manager
  FLACCheck / ErrorCallback exception texts                    :98-105,590-594  ApplicationException, same text
"""
from __future__ import annotations

import collections
import datetime
import enum
from typing import Optional

from . import _abi


class ApplicationException(Exception):
    """Stands in for System.ApplicationException (FLACDecoder.cs:54,62,103,220,593)."""


class ALFormat(enum.IntEnum):
    """OpenTK.Audio.OpenAL.ALFormat values the reference maps (FLACDecoder.cs:454-465)."""
    Unmapped = 0
    Mono8 = 0x1100
    Mono16 = 0x1101
    Stereo8 = 0x1102
    Stereo16 = 0x1103


class FLACPacket:
    """FLACPacket.cs:3-10."""
    __slots__ = ("SampleRate", "Channels", "BlockSize", "Data", "Offset")

    def __init__(self):
        self.SampleRate = 0
        self.Channels = 0
        self.BlockSize = 0
        self.Data = b""
        self.Offset = 0


class FLACPacketQueue:
    """FLACPacketQueue.cs:5-36 (ConcurrentQueue<FLACPacket>)."""

    def __init__(self):
        self._q = collections.deque()

    def IsEmpty(self) -> bool:
        return not self._q

    def Enqueue(self, packet: FLACPacket) -> None:
        self._q.append(packet)

    def TryPeek(self):
        return (True, self._q[0]) if self._q else (False, None)

    def TryDequeue(self):
        return (True, self._q.popleft()) if self._q else (False, None)


class EmptyStubLogger:
    """EmptyStubLogger.cs:3-13."""

    def Warning(self, message: str) -> None:  # noqa: N802 (reference name)
        pass


class FLACDecoder:
    DEFAULT_MAX_BUFFER_SIZE = 16384   # FLACDecoder.cs:21
    PACKET_BYTES = 1 << 20            # PCM bytes fetched from the engine per queued packet (reference: one frame per packet)

    def __init__(self, stream, queue, logger, buffer: Optional[bytearray] = None, *, device: int = -1, strict_reference: bool = False):
        self.mStream = stream
        self.mPacketQueue = queue
        self.mLogger = logger
        self.mInstreamBuffer = buffer if buffer is not None else bytearray(self.DEFAULT_MAX_BUFFER_SIZE)
        self.mHitEOFYet = False
        self.mIsDisposed = False
        self._strict = strict_reference
        self._handle = None
        self.Format = ALFormat.Unmapped
        self.Channels = 0
        self.SampleRate = 0
        self.BitsPerSample = 0
        self.Duration = datetime.timedelta(0)
        self.mFLACLength = 0
        self._errors_checked = False
        # SetupDecoder + SetupFLACStream (FLACDecoder.cs:49-64): the engine pulls the stream through the same
        # ReadCallback contract (<= len(mInstreamBuffer) bytes per Stream.Read, short read => end of stream)
        try:
            self._handle = _abi.open_callbacks(self._read_callback, device=device, flags=_abi.OPT_LAZY_PULL)   # metadata now, stream bytes on demand
        except _abi.BnflacError as e:
            if e.code in (_abi.ERR_NOT_FLAC, _abi.ERR_TRUNCATED):
                # process_until_end_of_metadata fails in the reference (FLACDecoder.cs:66-70)
                raise ApplicationException("FLAC: Could not Could not process until end of metadata - EndOfStream!") from e
            if e.code == _abi.ERR_ABORTED:
                raise ApplicationException("FLAC: Could not Could not process until end of metadata - Aborted!") from e
            raise
        self._metadata_callback(self._handle.info())

    # ---- callbacks (FLACDecoder.cs:325-363, 431-473) -------------------------------------------------------
    def _read_callback(self, nbytes: int):
        if self.mInstreamBuffer is None:
            return None                                        # ReadStatusAbort
        if nbytes <= 0:
            self.mHitEOFYet = True
            return None
        if self.mHitEOFYet:
            return b""                                         # the short read already reported the end (FLACDecoder.cs:345-350)
        out = bytearray()
        while len(out) < nbytes:                               # the engine asks for 1 MiB; the Stream is read <= buffer-size at a time
            length = min(nbytes - len(out), len(self.mInstreamBuffer))
            chunk = self.mStream.read(length)
            if chunk is None:
                chunk = b""
            self.mInstreamBuffer[:len(chunk)] = chunk
            out += chunk
            if len(chunk) < length:
                self.mHitEOFYet = True
                break
        return bytes(out)

    def _metadata_callback(self, info) -> None:
        self.BitsPerSample = info.bits_per_sample
        self.Channels = info.channels
        self.SampleRate = info.sample_rate
        self.mBlockAlign = info.block_align
        self.mTotalSamples = info.total_samples & 0xFFFFFFFF    # FLACDecoder.cs:449 quirk (low 32 bits only)
        self.mFLACLength = info.length_reference
        self.Duration = datetime.timedelta(seconds=info.duration_seconds)
        self._bytes_per_sample = info.bytes_per_sample
        if self.BitsPerSample == 16:
            self.Format = ALFormat.Stereo16 if self.Channels == 2 else ALFormat.Mono16
        elif self.BitsPerSample == 8:
            self.Format = ALFormat.Stereo8 if self.Channels == 2 else ALFormat.Mono8
        else:
            self.mLogger.Warning("FLAC: Unsupported sample bit size: {0}\n".format(self.BitsPerSample))

    # ---- Stream surface --------------------------------------------------------------------------------
    def Flush(self):
        raise NotImplementedError()

    def Seek(self, offset, origin):
        raise NotImplementedError()

    def SetLength(self, value):
        raise NotImplementedError()

    def Write(self, buffer, offset, count):
        raise NotImplementedError()

    @property
    def Position(self):
        raise NotImplementedError()

    @Position.setter
    def Position(self, value):
        raise NotImplementedError()

    @property
    def CanRead(self) -> bool:
        return bool(self.mStream.readable())

    @property
    def CanSeek(self) -> bool:
        return False

    @property
    def CanWrite(self) -> bool:
        return False

    @property
    def Length(self) -> int:
        return self.mFLACLength

    def Read(self, buffer, offset: int, count: int) -> int:
        """FLACDecoder.cs:124-205: fill buffer[offset:offset+count] from queued packets, decoding more when empty."""
        localOffset = offset
        spaceRemaining = count
        bytesRead = 0
        while spaceRemaining > 0:
            self.RequestAnotherFLACPacket()
            ok, current = self.mPacketQueue.TryPeek()
            if not ok:
                break
            bytesLeft = len(current.Data) - current.Offset
            if bytesLeft > spaceRemaining:
                buffer[localOffset:localOffset + spaceRemaining] = current.Data[current.Offset:current.Offset + spaceRemaining]
                current.Offset += spaceRemaining
                bytesRead += spaceRemaining
                spaceRemaining = 0
            elif bytesLeft > 0:
                buffer[localOffset:localOffset + bytesLeft] = current.Data[current.Offset:current.Offset + bytesLeft]
                localOffset += bytesLeft
                spaceRemaining -= bytesLeft
                bytesRead += bytesLeft
                self.PopTopOffQueue()
            else:
                self.PopTopOffQueue()
        return bytesRead

    def RequestAnotherFLACPacket(self) -> None:
        """FLACDecoder.cs:207-224.  One packet = up to PACKET_BYTES of decoded PCM (the reference queues one frame)."""
        if not self.mPacketQueue.IsEmpty():
            return
        if self._handle is None:
            return
        state = self._handle.state()
        if state < 4:  # EndOfStream
            if self._strict and self.BitsPerSample != 16:
                # WriteCallback rejects anything but 16-bit (FLACDecoder.cs:526-530) -> process_single fails
                self.mLogger.Warning("FLAC: Unsupported bit-rate: {0}".format(self.BitsPerSample))
                raise ApplicationException("FLAC: Could not process single - Aborted!")
            data = bytearray(self.PACKET_BYTES)
            try:
                n = self._handle.read_into(data)
            except _abi.BnflacError as e:
                raise ApplicationException("FLAC: Could not process single - {0}!".format(_abi.STATE_NAMES[self._handle.state()])) from e
            self._raise_frame_errors()
            if n:
                packet = FLACPacket()
                packet.Channels = self.Channels
                packet.SampleRate = self.SampleRate
                packet.BlockSize = n // max(1, self.Channels * self._bytes_per_sample)
                packet.Offset = 0
                packet.Data = bytes(data[:n])
                self.mPacketQueue.Enqueue(packet)
        elif state >= 5:  # OggError and worse
            raise ApplicationException("FLAC: Decoding returned with critical state: {0}".format(_abi.STATE_NAMES[state]))

    def _raise_frame_errors(self) -> None:
        """ErrorCallback (FLACDecoder.cs:590-594) throws on the first decode error the native codec reports."""
        if self._errors_checked:
            return
        errs = self._handle.errors_so_far()        # polled after every packet: nothing is decoded or pulled ahead for it
        if errs:
            self._errors_checked = True
            names = ["LostSync", "BadHeader", "FrameCrcMismatch", "UnparsableStream"]
            raise ApplicationException("FLAC: Could not decode frame: {0} - {1}!".format(names[errs[0]], "ReadFrame" if errs[0] >= 2 else "SearchForFrameSync"))

    def PopTopOffQueue(self) -> None:
        ok, _ = self.mPacketQueue.TryDequeue()
        if not ok:
            raise Exception("FLAC - queue error")

    def CopyTo(self, destination, bufferSize: int = 81920) -> None:
        """System.IO.Stream.CopyTo as OpenALDemo uses it (Program.cs:33): Read(buf,0,81920) until 0."""
        buf = bytearray(bufferSize)
        while True:
            n = self.Read(buf, 0, bufferSize)
            if n == 0:
                break
            destination.write(bytes(buf[:n]))

    # ---- IDisposable ----------------------------------------------------------------------------------
    def Dispose(self) -> None:
        if self.mIsDisposed:
            return
        self.mHitEOFYet = False
        if self._handle is not None:
            self._handle.close()          # finish + delete (FLACDecoder.cs:294-303)
            self._handle = None
        self.mStream.close()              # the decoder owns and closes the input stream (:308)
        self.mInstreamBuffer = None
        self.mIsDisposed = True

    close = Dispose

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.Dispose()

// FLACDecoder.cs -- drop-in replacement for Library/BirdNest.Audio/FLACDecoder.cs: the same public surface
// (class BirdNest.Audio.FLACDecoder : System.IO.Stream, both constructors, Read, Format / Channels / SampleRate /
// BitsPerSample / Duration / Length, CanRead / CanSeek / CanWrite, the throwing members, Dispose closing the inner
// stream, the ApplicationException texts) over libbnflac's C ABI instead of LibFlac.dll.
//
// What changed against the reference file:
//   * SetupDecoder/SetupFLACStream/SetupStreamInfo (:49-70)  -> one bnflac_open_callbacks + bnflac_info
//   * ReadCallback (:325-363) keeps its contract; the engine asks for 1 MiB at a time, the Stream is still read
//     <= mInstreamBuffer.Length bytes per call and a short read still means end of stream
//   * WriteCallback (:520-580), the per-sample managed interleave, is gone: the GPU writes interleaved LE PCM
//   * RequestAnotherFLACPacket (:207-224) queues one packet per bulk bnflac_read instead of one per frame
//   * Seek/Tell/Length/EOF callbacks (:365-423,475-488) are not needed by the engine
// Not compiled here (no C# toolchain in the build image); birdnest/audio_b200/flac_decoder.py and
// csrc/flac_decoder.hpp are the same logic in the two languages the image can run, and are what the tests drive.
using System;
using System.IO;
using System.Runtime.InteropServices;
using OpenTK.Audio.OpenAL;
using LibBnFlacSharp;

namespace BirdNest.Audio
{
	public class FLACDecoder : Stream
	{
		private Stream mStream;
		private IFLACPacketQueue mPacketQueue;
		private IFLACDecoderLogger mLogger;
		private byte[] mInstreamBuffer;
		private const Int64 DEFAULT_MAX_BUFFER_SIZE = 16384;
		private const int PACKET_BYTES = 1 << 20;     // PCM bytes pulled from the engine per queued packet

		private IntPtr mDecoderContext;
		private LibBnFlac.ReadCallback mReadCallback;  // kept alive as long as the handle: with LazyPull the library calls it from bnflac_read
		private bool mHitEOFYet;
		private bool mErrorsChecked;
		private int mBytesPerSample;

		public FLACDecoder (Stream stream, IFLACPacketQueue queue, IFLACDecoderLogger logger)
			: this (stream, queue, logger, new byte[DEFAULT_MAX_BUFFER_SIZE])
		{
		}

		public FLACDecoder (Stream stream, IFLACPacketQueue queue, IFLACDecoderLogger logger, byte[] buffer)
		{
			mStream = stream;
			mPacketQueue = queue;
			mLogger = logger;
			mInstreamBuffer = buffer;
			mHitEOFYet = false;

			mReadCallback = new LibBnFlac.ReadCallback (this.ReadCallback);
			// LazyPull: like the reference, the constructor reads the metadata only; stream bytes are pulled as Read advances
			var opts = new LibBnFlac.Opts { StructSize = (uint) Marshal.SizeOf (typeof (LibBnFlac.Opts)), Device = -1, Flags = (uint) LibBnFlac.OpenFlags.LazyPull };
			int rc = LibBnFlac.bnflac_open_callbacks (mReadCallback, IntPtr.Zero, ref opts, out mDecoderContext);
			if (rc == (int) LibBnFlac.Error.NotFlac || rc == (int) LibBnFlac.Error.Truncated)
				throw new ApplicationException ("FLAC: Could not Could not process until end of metadata - EndOfStream!");
			if (rc == (int) LibBnFlac.Error.Aborted)
				throw new ApplicationException ("FLAC: Could not Could not process until end of metadata - Aborted!");
			if (rc == (int) LibBnFlac.Error.MemoryAllocationError)
				throw new ApplicationException ("FLAC: Could not initialize stream decoder!");
			if (rc != 0)
				throw new ApplicationException ("FLAC: Could not open stream for reading!");

			LibBnFlac.Info info;
			LibBnFlac.bnflac_info (mDecoderContext, out info);
			MetadataCallback (info);
		}

		private void FLACCheck (bool result, string operation)
		{
			if (!result)
			{
				var decoderState = (LibBnFlac.StreamDecoderState) LibBnFlac.bnflac_state (mDecoderContext);
				throw new ApplicationException (string.Format ("FLAC: Could not {0} - {1}!", operation, decoderState));
			}
		}

		#region implemented abstract members of Stream

		public override void Flush () { throw new NotImplementedException (); }
		public override long Seek (long offset, SeekOrigin origin) { throw new NotImplementedException (); }
		public override void SetLength (long value) { throw new NotImplementedException (); }
		public override void Write (byte[] buffer, int offset, int count) { throw new NotImplementedException (); }

		public override int Read (byte[] buffer, int offset, int count)
		{
			int localOffset = offset;
			int spaceRemaining = count;
			int bytesRead = 0;
			while (spaceRemaining > 0)
			{
				RequestAnotherFLACPacket ();
				FLACPacket current;
				if (!mPacketQueue.TryPeek (out current))
					break;
				int bytesLeft = current.Data.Length - current.Offset;
				if (bytesLeft > spaceRemaining)
				{
					Array.Copy (current.Data, current.Offset, buffer, localOffset, spaceRemaining);
					current.Offset += spaceRemaining;
					bytesRead += spaceRemaining;
					spaceRemaining = 0;
				}
				else
				{
					if (bytesLeft > 0)
					{
						Array.Copy (current.Data, current.Offset, buffer, localOffset, bytesLeft);
						localOffset += bytesLeft;
						spaceRemaining -= bytesLeft;
						bytesRead += bytesLeft;
					}
					PopTopOffQueue ();
				}
			}
			return bytesRead;
		}

		private void RequestAnotherFLACPacket ()
		{
			if (!mPacketQueue.IsEmpty ())
				return;
			var state = (LibBnFlac.StreamDecoderState) LibBnFlac.bnflac_state (mDecoderContext);
			if (state < LibBnFlac.StreamDecoderState.EndOfStream)
			{
				var data = new byte[PACKET_BYTES];
				long n;
				var pin = GCHandle.Alloc (data, GCHandleType.Pinned);
				try { n = LibBnFlac.bnflac_read (mDecoderContext, pin.AddrOfPinnedObject (), (UIntPtr) (uint) data.Length); }
				finally { pin.Free (); }
				FLACCheck (n >= 0, "process single");
				RaiseFrameErrors ();
				if (n > 0)
				{
					if (n < data.Length) Array.Resize (ref data, (int) n);
					var packet = new FLACPacket ();
					packet.Channels = this.Channels;
					packet.SampleRate = this.SampleRate;
					packet.BlockSize = (int) (n / Math.Max (1, this.Channels * mBytesPerSample));
					packet.Offset = 0;
					packet.Data = data;
					mPacketQueue.Enqueue (packet);
				}
			}
			else if (state >= LibBnFlac.StreamDecoderState.OggError)
			{
				throw new ApplicationException (string.Format ("FLAC: Decoding returned with critical state: {0}", state));
			}
		}

		// ErrorCallback (reference :590-594) throws on the first decode error the native codec reports
		private void RaiseFrameErrors ()
		{
			if (mErrorsChecked) return;
			IntPtr codes; UIntPtr n;
			// polled after every packet: only what has been decoded so far, nothing is decoded or pulled ahead for it
			if (LibBnFlac.bnflac_errors_so_far (mDecoderContext, out codes, out n) != 0 || n == UIntPtr.Zero) return;
			mErrorsChecked = true;
			var status = (LibBnFlac.DecodeError) Marshal.ReadInt32 (codes);
			var decoderState = status >= LibBnFlac.DecodeError.FrameCrcMismatch ? LibBnFlac.StreamDecoderState.ReadFrame : LibBnFlac.StreamDecoderState.SearchForFrameSync;
			throw new ApplicationException (string.Format ("FLAC: Could not decode frame: {0} - {1}!", status, decoderState));
		}

		private void PopTopOffQueue ()
		{
			FLACPacket top;
			if (!mPacketQueue.TryDequeue (out top))
				throw new Exception ("FLAC - queue error");
		}

		public override bool CanRead { get { return mStream.CanRead; } }
		public override bool CanSeek { get { return false; } }
		public override bool CanWrite { get { return false; } }

		private int mBlockAlign;
		private long mTotalSamples;
		private long mFLACLength;
		public override long Length { get { return mFLACLength; } }

		public override long Position
		{
			get { throw new NotImplementedException (); }
			set { throw new NotImplementedException (); }
		}

		#endregion

		#region IDisposable implementation

		private bool mIsDisposed = false;

		protected override void Dispose (bool disposing)
		{
			if (mIsDisposed)
				return;
			mHitEOFYet = false;
			if (mDecoderContext != IntPtr.Zero)
			{
				LibBnFlac.bnflac_close (mDecoderContext);   // finish + delete
				mDecoderContext = IntPtr.Zero;
			}
			if (disposing)
			{
				mStream.Close ();
				mInstreamBuffer = null;
			}
			mIsDisposed = true;
			base.Dispose (true);
		}

		#endregion

		#region Callbacks

		// Same contract as the reference's ReadCallback: at most mInstreamBuffer.Length bytes per Stream.Read,
		// a short read reports end of stream.  The engine's request (1 MiB) is filled by looping.
		protected int ReadCallback (IntPtr user, IntPtr buffer, ref UIntPtr bytes)
		{
			if (mInstreamBuffer == null)
				return 2;
			long want = (long) bytes.ToUInt64 ();
			if (want <= 0)
			{
				mHitEOFYet = true;
				return 2;
			}
			long done = 0;
			while (done < want)
			{
				int length = (int) Math.Min (want - done, (long) mInstreamBuffer.Length);
				int count = mStream.Read (mInstreamBuffer, 0, length);
				if (count < 0)
				{
					mHitEOFYet = true;
					return 2;
				}
				Marshal.Copy (mInstreamBuffer, 0, new IntPtr (buffer.ToInt64 () + done), count);
				done += count;
				if (count < length)
				{
					mHitEOFYet = true;
					bytes = (UIntPtr) (ulong) done;
					return 1;
				}
			}
			bytes = (UIntPtr) (ulong) done;
			return 0;
		}

		public ALFormat Format { get; private set; }
		public int Channels { get; private set; }
		public int SampleRate { get; private set; }
		public int BitsPerSample { get; private set; }
		public TimeSpan Duration { get; private set; }

		private void MetadataCallback (LibBnFlac.Info info)
		{
			this.BitsPerSample = (int) info.BitsPerSample;
			this.Channels = (int) info.Channels;
			this.SampleRate = (int) info.SampleRate;
			mBytesPerSample = (int) info.BytesPerSample;
			mBlockAlign = (int) info.BlockAlign;
			mTotalSamples = (long) (info.TotalSamples & 0xFFFFFFFFUL);   // the reference keeps the low 32 bits only (:449)
			mFLACLength = (long) info.LengthReference;
			Duration = TimeSpan.FromSeconds (info.DurationSeconds);
			if (this.BitsPerSample == 16)
				this.Format = this.Channels == 2 ? ALFormat.Stereo16 : ALFormat.Mono16;
			else if (this.BitsPerSample == 8)
				this.Format = this.Channels == 2 ? ALFormat.Stereo8 : ALFormat.Mono8;
			else
				mLogger.Warning (string.Format ("FLAC: Unsupported sample bit size: {0}\n", BitsPerSample));
		}

		#endregion
	}
}

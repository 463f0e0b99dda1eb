// LibBnFlac.cs -- P/Invoke declarations for libbnflac (include/bnflac.h), the B200-native codec that replaces
// LibFlac.dll behind BirdNest.Audio.FLACDecoder.  Takes the place of the decoder half of
// Library/LibFLACSharp/LibFLACSharp.cs:42-85,175-212 (16 DllImports + 8 callback delegate types): the new
// boundary is one open call, one info call, a bulk read and a close.
//
// Not compiled in this repository's CI (the build image has no C# toolchain); it is the binding a maintainer adds
// next to LibFLACSharp.cs.  Field order and sizes follow include/bnflac.h exactly (tests/test_abi_cpu.py checks the
// header against the ctypes mirror, which has the same layout).
using System;
using System.Runtime.InteropServices;

namespace LibBnFlacSharp
{
	public static class LibBnFlac
	{
		// resolves to libbnflac.so (Linux: Mono/.NET probe "lib" + name + ".so"), the way DLLName = "LibFlac" did (LibFLACSharp.cs:22)
		private const string DLLName = "bnflac";

		public enum Error : int
		{
			Ok = 0, Arg = -1, NotFlac = -2, Truncated = -3, NoDevice = -4, Cuda = -5, MemoryAllocationError = -6,
			Capacity = -7, Aborted = -8, Unsupported = -9, State = -10
		}

		// same numbering as LibFLAC.StreamDecoderState (LibFLACSharp.cs:24-36)
		public enum StreamDecoderState : int
		{
			SearchForMetadata = 0, ReadMetadata, SearchForFrameSync, ReadFrame, EndOfStream, OggError, SeekError, Aborted,
			MemoryAllocationError, Uninitialized
		}

		// same numbering as LibFLAC.DecodeError (LibFLACSharp.cs:262-268)
		public enum DecodeError : int { LostSync = 0, BadHeader = 1, FrameCrcMismatch = 2, UnparsableStream = 3 }

		[Flags]
		public enum OpenFlags : uint { None = 0, VerifyMd5 = 1, BorrowInput = 2, LazyPull = 4 }   // BNFLAC_OPT_* (include/bnflac.h)

		[StructLayout(LayoutKind.Sequential)]
		public struct Opts
		{
			public uint StructSize;       // = Marshal.SizeOf(typeof(Opts))
			public int Device;            // -1 = current CUDA device
			public IntPtr Stream;         // cudaStream_t or IntPtr.Zero
			public uint ShardIndex, ShardCount;
			public uint Flags;
			public uint ReadChunkFrames;
		}

		[StructLayout(LayoutKind.Sequential)]
		public struct Info
		{
			public uint SampleRate, Channels, BitsPerSample;
			public uint MinBlockSize, MaxBlockSize, MinFrameSize, MaxFrameSize;
			public uint BlockAlign;       // channels * (bps / 8), FLACDecoder.cs:448
			public uint ALFormat;         // 0x1100 Mono8 .. 0x1103 Stereo16, 0 = unmapped (FLACDecoder.cs:454-465)
			public uint BytesPerSample;
			public ulong TotalSamples, PcmBytes, LengthReference;
			public double DurationSeconds;
			[MarshalAs(UnmanagedType.ByValArray, SizeConst = 16)] public byte[] Md5;
			public ulong FirstFrameOffset;
		}

		// shape of FLACDecoder.ReadCallback (FLACDecoder.cs:325-363): 0 continue, 1 end of stream, 2 abort
		[UnmanagedFunctionPointer(CallingConvention.Cdecl)]
		public delegate int ReadCallback(IntPtr user, IntPtr buffer, ref UIntPtr bytes);

		[DllImport(DLLName, CallingConvention = CallingConvention.Cdecl)]
		public static extern int bnflac_open_callbacks(ReadCallback read, IntPtr user, ref Opts opts, out IntPtr handle);

		[DllImport(DLLName, CallingConvention = CallingConvention.Cdecl)]
		public static extern int bnflac_open_memory(byte[] data, UIntPtr len, ref Opts opts, out IntPtr handle);

		[DllImport(DLLName, CallingConvention = CallingConvention.Cdecl)]
		public static extern int bnflac_info(IntPtr handle, out Info info);

		[DllImport(DLLName, CallingConvention = CallingConvention.Cdecl)]
		public static extern int bnflac_state(IntPtr handle);

		// interleaved little-endian PCM into dst; returns bytes written, 0 at end of stream, < 0 = Error
		[DllImport(DLLName, CallingConvention = CallingConvention.Cdecl)]
		public static extern long bnflac_read(IntPtr handle, IntPtr dst, UIntPtr count);

		[DllImport(DLLName, CallingConvention = CallingConvention.Cdecl)]
		public static extern int bnflac_decode_all(IntPtr handle, IntPtr dst, UIntPtr cap, out ulong written);

		[DllImport(DLLName, CallingConvention = CallingConvention.Cdecl)]
		public static extern int bnflac_errors(IntPtr handle, out IntPtr codes, out UIntPtr n);

		[DllImport(DLLName, CallingConvention = CallingConvention.Cdecl)]
		public static extern int bnflac_probe(IntPtr data, UIntPtr len, out Info info);

		[DllImport(DLLName, CallingConvention = CallingConvention.Cdecl)]
		public static extern int bnflac_errors_so_far(IntPtr handle, out IntPtr codes, out UIntPtr n);

		[DllImport(DLLName, CallingConvention = CallingConvention.Cdecl)]
		public static extern void bnflac_close(IntPtr handle);

		[DllImport(DLLName, CallingConvention = CallingConvention.Cdecl)]
		public static extern IntPtr bnflac_strerror(int err);

		[DllImport(DLLName, CallingConvention = CallingConvention.Cdecl)]
		public static extern int bnflac_device_count();

		// ---- encoder (SURVEY 8f-4): replaces the FLAC__stream_encoder_* DllImports of LibFLACSharp.cs:322-387 ----
		[StructLayout(LayoutKind.Sequential)]
		public struct EncOpts
		{
			public uint StructSize;       // = Marshal.SizeOf(typeof(EncOpts))
			public int Device;            // -1 = current CUDA device
			public IntPtr Stream;
			public uint SampleRate, Channels, BitsPerSample;      // set_sample_rate / set_channels / set_bits_per_sample
			public uint BlockSize;        // set_blocksize; 0 = 4096 (or the preset's)
			public uint MaxLpcOrder, QlpPrecision, MinPartitionOrder, MaxPartitionOrder;
			public uint MidSide;          // set_do_mid_side_stereo
			public uint CompressionLevel; // set_compression_level, with Flags |= UseLevel
			public uint Flags;            // 1 NoMd5, 2 InputInt32, 4 UseLevel, 8 FixedOrder
			public ulong FirstFrameNumber; // a stream encoded in several calls: the coded number of this call's first frame
		}

		[StructLayout(LayoutKind.Sequential)]
		public struct EncStats
		{
			public float PlanMs, WriteMs, TotalMs;
			public uint Frames;
			public ulong Bytes;
			public uint MinFrameSize, MaxFrameSize;
			public IntPtr FrameSizes;     // in: IntPtr.Zero or room for FrameSizesCap uints
			public ulong FrameSizesCap;
		}

		[DllImport(DLLName, CallingConvention = CallingConvention.Cdecl)]
		public static extern int bnflac_encode_bound(UIntPtr pcmBytes, ref EncOpts opts, out ulong bound);

		// interleaved little-endian PCM (the layout FLACDecoder.Read returns) -> "fLaC" + STREAMINFO + frames
		[DllImport(DLLName, CallingConvention = CallingConvention.Cdecl)]
		public static extern int bnflac_encode(byte[] pcm, UIntPtr pcmBytes, ref EncOpts opts, byte[] dst, UIntPtr cap, out ulong written, ref EncStats stats);
	}
}

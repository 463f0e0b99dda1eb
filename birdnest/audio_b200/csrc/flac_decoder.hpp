// flac_decoder.hpp -- C++ host mirror of BirdNest.Audio.FLACDecoder (Library/BirdNest.Audio/FLACDecoder.cs) over the
// C ABI of libbnflac (include/bnflac.h).  Header only; same member names, argument meaning and error texts as the C#
// class, so that a C++ caller reads like OpenALDemo/Program.cs:26-38.  The native codec underneath is the CUDA pipeline
// in libbnflac.so; there is no CPU decode path (construction throws when no CUDA device is usable).
//
//   reference member (FLACDecoder.cs)                                here
//   ctor(Stream, IFLACPacketQueue, IFLACDecoderLogger[, byte[]]) :23,72   FLACDecoder(std::istream&, queue*, logger*[, buffer bytes])
//   Read(byte[], int, int) :124                                            Read(uint8_t* buffer, int offset, int count)
//   Format/Channels/SampleRate/BitsPerSample/Duration :426-430            same-named accessors
//   Length :261, CanRead/CanSeek/CanWrite :240-256                        same
//   Flush/Seek/SetLength/Write/Position :109-122,235-238,267-274          throw NotImplementedException
//   Dispose :285-319                                                      Dispose() / destructor
#pragma once
#include <bnflac.h>
#include <algorithm>
#include <cstdint>
#include <deque>
#include <istream>
#include <stdexcept>
#include <string>
#include <vector>

namespace bnflac_host {

struct ApplicationException : std::runtime_error { using std::runtime_error::runtime_error; };   // System.ApplicationException
struct NotImplementedException : std::logic_error { NotImplementedException() : std::logic_error("NotImplementedException") {} };

enum class ALFormat : int { Unmapped = 0, Mono8 = 0x1100, Mono16 = 0x1101, Stereo8 = 0x1102, Stereo16 = 0x1103 };   // FLACDecoder.cs:454-465

struct FLACPacket { int SampleRate = 0, Channels = 0, BlockSize = 0; std::vector<uint8_t> Data; int Offset = 0; };   // FLACPacket.cs:3-10

struct IFLACPacketQueue {                                   // IFLACPacketQueue.cs:3-9
    virtual ~IFLACPacketQueue() {}
    virtual bool IsEmpty() = 0;
    virtual void Enqueue(FLACPacket&& p) = 0;
    virtual bool TryPeek(FLACPacket*& p) = 0;
    virtual bool TryDequeue() = 0;
};
struct FLACPacketQueue : IFLACPacketQueue {                 // FLACPacketQueue.cs:5-36
    std::deque<FLACPacket> q;
    bool IsEmpty() override { return q.empty(); }
    void Enqueue(FLACPacket&& p) override { q.push_back(std::move(p)); }
    bool TryPeek(FLACPacket*& p) override { if (q.empty()) return false; p = &q.front(); return true; }
    bool TryDequeue() override { if (q.empty()) return false; q.pop_front(); return true; }
};
struct IFLACDecoderLogger { virtual ~IFLACDecoderLogger() {} virtual void Warning(const std::string&) {} };   // EmptyStubLogger.cs:3-13

class FLACDecoder {
public:
    static constexpr int DEFAULT_MAX_BUFFER_SIZE = 16384;   // FLACDecoder.cs:21
    static constexpr size_t PACKET_BYTES = 1u << 20;        // PCM bytes per queued packet (the reference queues one frame per packet)

    FLACDecoder(std::istream& stream, IFLACPacketQueue* queue, IFLACDecoderLogger* logger, size_t buffer_bytes = DEFAULT_MAX_BUFFER_SIZE, int device = -1)
        : mStream(stream), mPacketQueue(queue), mLogger(logger), mInstreamBuffer(buffer_bytes) {
        bnflac_opts o{}; o.struct_size = sizeof o; o.device = device;
        o.flags = BNFLAC_OPT_LAZY_PULL;       // like the reference: metadata in the constructor, stream bytes pulled as Read advances
        const int rc = bnflac_open_callbacks(&FLACDecoder::ReadCallback, this, &o, &mHandle);   // SetupDecoder + SetupFLACStream (:49-64)
        if (rc == BNFLAC_ERR_NOT_FLAC || rc == BNFLAC_ERR_TRUNCATED) throw ApplicationException("FLAC: Could not Could not process until end of metadata - EndOfStream!");
        if (rc == BNFLAC_ERR_ABORTED) throw ApplicationException("FLAC: Could not Could not process until end of metadata - Aborted!");
        if (rc) throw ApplicationException(std::string("FLAC: Could not initialize stream decoder - ") + bnflac_strerror(rc) + "!");
        MetadataCallback();
    }
    ~FLACDecoder() { Dispose(); }
    FLACDecoder(const FLACDecoder&) = delete;
    FLACDecoder& operator=(const FLACDecoder&) = delete;

    // ---- Stream surface (FLACDecoder.cs:109-122, 235-274)
    void Flush() { throw NotImplementedException(); }
    long Seek(long, int) { throw NotImplementedException(); }
    void SetLength(long) { throw NotImplementedException(); }
    void Write(const uint8_t*, int, int) { throw NotImplementedException(); }
    long Position() const { throw NotImplementedException(); }
    bool CanRead() const { return mStream.good() || mStream.eof(); }
    bool CanSeek() const { return false; }
    bool CanWrite() const { return false; }
    long long Length() const { return (long long)mInfo.length_reference; }

    // FLACDecoder.cs:124-205.  The reference fills the caller's buffer from queued one-frame packets.  Here the engine already
    // holds decoded PCM in pinned host memory, so an empty queue means: copy straight into the caller's buffer (one copy
    // instead of engine -> packet -> buffer, and no zero-filled megabyte per packet -- measured on the 1 h stream: the packet
    // detour was 4 of the 13 GB of single-threaded memcpy that made up a drained Read loop).  Packets that are already queued
    // (SetPacketMode(true), or a caller that enqueues its own) are served first, exactly as before.
    int Read(uint8_t* buffer, int offset, int count) {
        int localOffset = offset, spaceRemaining = count, bytesRead = 0;
        while (spaceRemaining > 0) {
            if (!mPacketMode && mPacketQueue->IsEmpty()) {
                if (!mHandle) break;
                const int state = bnflac_state(mHandle);
                if (state >= BNFLAC_STATE_OGG_ERROR) throw ApplicationException(std::string("FLAC: Decoding returned with critical state: ") + bnflac_state_name(state));
                if (state >= BNFLAC_STATE_END_OF_STREAM) break;
                const int64_t n = bnflac_read(mHandle, buffer + localOffset, (size_t)spaceRemaining);
                if (n < 0) throw ApplicationException(std::string("FLAC: Could not process single - ") + bnflac_state_name(bnflac_state(mHandle)) + "!");
                RaiseFrameErrors();
                if (n == 0) break;
                localOffset += (int)n; spaceRemaining -= (int)n; bytesRead += (int)n;
                continue;
            }
            RequestAnotherFLACPacket();
            FLACPacket* cur = nullptr;
            if (!mPacketQueue->TryPeek(cur)) break;
            const int bytesLeft = (int)cur->Data.size() - cur->Offset;
            if (bytesLeft > spaceRemaining) {
                std::copy_n(cur->Data.data() + cur->Offset, spaceRemaining, buffer + localOffset);
                cur->Offset += spaceRemaining; bytesRead += spaceRemaining; spaceRemaining = 0;
            } else {
                if (bytesLeft > 0) { std::copy_n(cur->Data.data() + cur->Offset, bytesLeft, buffer + localOffset); localOffset += bytesLeft; spaceRemaining -= bytesLeft; bytesRead += bytesLeft; }
                PopTopOffQueue();
            }
        }
        return bytesRead;
    }
    // System.IO.Stream.CopyTo as OpenALDemo uses it (Program.cs:33)
    void CopyTo(std::vector<uint8_t>& destination, int bufferSize = 81920) {
        std::vector<uint8_t> buf((size_t)bufferSize);
        for (int n; (n = Read(buf.data(), 0, bufferSize)) > 0;) destination.insert(destination.end(), buf.begin(), buf.begin() + n);
    }

    void SetPacketMode(bool on) { mPacketMode = on; }        // true: every Read goes through FLACPacket objects in the queue, like the reference

    ALFormat Format() const { return mFormat; }
    int Channels() const { return (int)mInfo.channels; }
    int SampleRate() const { return (int)mInfo.sample_rate; }
    int BitsPerSample() const { return (int)mInfo.bits_per_sample; }
    double DurationSeconds() const { return mInfo.duration_seconds; }          // TimeSpan Duration (:452)
    const bnflac_info_t& Info() const { return mInfo; }

    void Dispose() {                                                           // FLACDecoder.cs:285-319
        if (mIsDisposed) return;
        mIsDisposed = true;
        if (mHandle) { bnflac_close(mHandle); mHandle = nullptr; }
    }

private:
    // FLACDecoder.cs:325-363: at most mInstreamBuffer bytes per Stream.Read, a short read is the end of the stream
    static int ReadCallback(void* user, uint8_t* buf, size_t* bytes) {
        FLACDecoder* self = static_cast<FLACDecoder*>(user);
        if (self->mInstreamBuffer.empty()) return 2;                           // ReadStatusAbort
        size_t done = 0; const size_t want = *bytes;
        bool eof = self->mHitEOFYet;
        // the managed byte[] + Marshal.Copy of the reference is one copy too many here: each Stream.Read (still at most
        // mInstreamBuffer.size() bytes, FLACDecoder.cs:336) lands directly in the engine's buffer
        while (done < want && !eof) {
            const size_t len = std::min(want - done, self->mInstreamBuffer.size());
            self->mStream.read(reinterpret_cast<char*>(buf + done), (std::streamsize)len);
            const size_t got = (size_t)self->mStream.gcount();
            done += got;
            if (got < len) { eof = true; self->mHitEOFYet = true; }
        }
        *bytes = done;
        return eof ? 1 : 0;
    }
    void MetadataCallback() {                                                  // FLACDecoder.cs:431-473
        bnflac_info(mHandle, &mInfo);
        if (mInfo.bits_per_sample == 16) mFormat = mInfo.channels == 2 ? ALFormat::Stereo16 : ALFormat::Mono16;
        else if (mInfo.bits_per_sample == 8) mFormat = mInfo.channels == 2 ? ALFormat::Stereo8 : ALFormat::Mono8;
        else if (mLogger) mLogger->Warning("FLAC: Unsupported sample bit size: " + std::to_string(mInfo.bits_per_sample) + "\n");
    }
    void RequestAnotherFLACPacket() {                                          // FLACDecoder.cs:207-224
        if (!mPacketQueue->IsEmpty() || !mHandle) return;
        const int state = bnflac_state(mHandle);
        if (state < BNFLAC_STATE_END_OF_STREAM) {
            FLACPacket p; p.Data.resize(PACKET_BYTES);
            const int64_t n = bnflac_read(mHandle, p.Data.data(), p.Data.size());
            if (n < 0) throw ApplicationException(std::string("FLAC: Could not process single - ") + bnflac_state_name(bnflac_state(mHandle)) + "!");
            RaiseFrameErrors();
            if (n > 0) {
                p.Data.resize((size_t)n); p.Channels = Channels(); p.SampleRate = SampleRate();
                p.BlockSize = (int)(n / std::max<int64_t>(1, (int64_t)mInfo.channels * mInfo.bytes_per_sample));
                mPacketQueue->Enqueue(std::move(p));
            }
        } else if (state >= BNFLAC_STATE_OGG_ERROR)
            throw ApplicationException(std::string("FLAC: Decoding returned with critical state: ") + bnflac_state_name(state));
    }
    void RaiseFrameErrors() {                                                  // ErrorCallback, FLACDecoder.cs:590-594
        if (mErrorsChecked) return;
        const uint32_t* codes = nullptr; size_t n = 0;                         // polled after every packet: no decode / pull ahead
        if (bnflac_errors_so_far(mHandle, &codes, &n) != 0 || !n) return;
        mErrorsChecked = true;
        throw ApplicationException(std::string("FLAC: Could not decode frame: ") + bnflac_frame_status_name((int)codes[0] + 1) + " - " + (codes[0] >= 2 ? "ReadFrame" : "SearchForFrameSync") + "!");
    }
    void PopTopOffQueue() { if (!mPacketQueue->TryDequeue()) throw std::runtime_error("FLAC - queue error"); }

    std::istream& mStream;
    IFLACPacketQueue* mPacketQueue;
    IFLACDecoderLogger* mLogger;
    std::vector<uint8_t> mInstreamBuffer;
    bnflac_t* mHandle = nullptr;
    bnflac_info_t mInfo{};
    ALFormat mFormat = ALFormat::Unmapped;
    bool mHitEOFYet = false, mIsDisposed = false, mErrorsChecked = false, mPacketMode = false;
};

} // namespace bnflac_host

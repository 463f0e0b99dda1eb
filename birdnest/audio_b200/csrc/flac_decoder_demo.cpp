// flac_decoder_demo.cpp -- the shape of Library/OpenALDemo/Program.cs:26-38 in C++: open a .flac, construct the decoder,
// read the properties OpenAL is given, drain the stream with CopyTo, report.  (No OpenAL here: playback is out of scope.)
//   flacdecoder_demo file.flac [out.pcm]
//   flacdecoder_demo --bench file.flac [device] [reps]   the drop-in surface end to end, timed inside the process (bench.py's
//       e2e_stream): FLACDecoder over a std::istream whose Read hands out <= 16 KiB per call (FLACDecoder.cs:336), drained with
//       Read(buf, 0, 81920) into pageable memory exactly as Stream.CopyTo(MemoryStream) does (Program.cs:33); prints one line
//       per repetition: first_read_ms (constructor + first Read), total_ms (constructor .. last Read), bytes, reads.
//   flacdecoder_demo --bench-drain ...                   the same Read loop without the MemoryStream (the 81,920-byte buffer is
//       reused): what the decoder and its Stream buffering cost, without the caller's growing 2 GB destination.
#include "flac_decoder.hpp"
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>

static double ms_since(std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }

static int bench(const char* path, int device, int reps, bool keep) {
    for (int r = 0; r < reps; r++) {
        std::ifstream fs(path, std::ios::binary);
        if (!fs) { std::fprintf(stderr, "cannot open %s\n", path); return 2; }
        try {
            bnflac_host::FLACPacketQueue queue;
            bnflac_host::IFLACDecoderLogger logger;
            const auto t0 = std::chrono::steady_clock::now();
            bnflac_host::FLACDecoder reader(fs, &queue, &logger, bnflac_host::FLACDecoder::DEFAULT_MAX_BUFFER_SIZE, device);
            std::vector<uint8_t> ms;                                  // MemoryStream: grows as it is written
            std::vector<uint8_t> buf(81920);                          // Stream.CopyTo's buffer
            double first = -1; size_t reads = 0, bytes = 0;
            unsigned long long sum = 0;                               // the bytes are really there
            for (int n; (n = reader.Read(buf.data(), 0, (int)buf.size())) > 0; reads++) {
                if (first < 0) first = ms_since(t0);
                if (keep) ms.insert(ms.end(), buf.begin(), buf.begin() + n);      // MemoryStream.Write: what CopyTo does with every buffer
                else sum += buf[0] + buf[(size_t)n - 1];                          // (--bench-drain: the Read loop alone, the buffer is reused)
                bytes += (size_t)n;
            }
            const double total = ms_since(t0);
            for (size_t i = 0; i < ms.size(); i += 4099) sum += ms[i];
            std::printf("rep %d first_read_ms %.3f total_ms %.3f bytes %zu reads %zu checksum %llu\n", r, first, total, bytes, reads, sum);
        } catch (const std::exception& e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 3 && (!std::strcmp(argv[1], "--bench") || !std::strcmp(argv[1], "--bench-drain")))
        return bench(argv[2], argc > 3 ? std::atoi(argv[3]) : -1, argc > 4 ? std::atoi(argv[4]) : 3, !std::strcmp(argv[1], "--bench"));
    if (argc < 2) { std::fprintf(stderr, "usage: %s file.flac [out.pcm] | --bench file.flac [device] [reps]\n", argv[0]); return 2; }
    std::ifstream fs(argv[1], std::ios::binary);
    if (!fs) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
    try {
        bnflac_host::FLACPacketQueue queue;
        bnflac_host::IFLACDecoderLogger logger;
        bnflac_host::FLACDecoder reader(fs, &queue, &logger);
        std::printf("SampleRate %d  Channels %d  BitsPerSample %d  Duration %.3f s  Format 0x%04x  Length %lld\n", reader.SampleRate(), reader.Channels(),
                    reader.BitsPerSample(), reader.DurationSeconds(), (int)reader.Format(), reader.Length());
        std::vector<uint8_t> pcm;
        const auto t0 = std::chrono::steady_clock::now();
        reader.CopyTo(pcm);
        std::printf("decoded %zu PCM bytes in %.2f ms\n", pcm.size(), ms_since(t0));
        if (argc > 2) { std::ofstream out(argv[2], std::ios::binary); out.write(reinterpret_cast<const char*>(pcm.data()), (std::streamsize)pcm.size()); }
    } catch (const std::exception& e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
    return 0;
}

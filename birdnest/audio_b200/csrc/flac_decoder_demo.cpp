// flac_decoder_demo.cpp -- the shape of Library/OpenALDemo/Program.cs:26-38 in C++: open a .flac, construct the decoder,
// read the properties OpenAL is given, drain the stream with CopyTo, report.  (No OpenAL here: playback is out of scope.)
#include "flac_decoder.hpp"
#include <chrono>
#include <cstdio>
#include <fstream>

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s file.flac [out.pcm]\n", argv[0]); return 2; }
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
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        std::printf("decoded %zu PCM bytes in %.2f ms\n", pcm.size(), ms);
        if (argc > 2) { std::ofstream out(argv[2], std::ios::binary); out.write(reinterpret_cast<const char*>(pcm.data()), (std::streamsize)pcm.size()); }
    } catch (const std::exception& e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
    return 0;
}

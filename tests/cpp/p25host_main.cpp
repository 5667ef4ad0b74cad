// p25host_main.cpp -- drives the C++ host layer (include/p25cu.hpp) in the reference's two driver shapes and prints
// what it saw, for tests/test_cpp_host.py to compare with the oracle.
//
//   p25host_main replay a.f32 b.f32 ...   src/replay.rs:26-57: f32le 48 kHz recordings -> ReplayReceiver
//   p25host_main sdr a.u8 b.u8 ...        src/sdr.rs:25-33 + src/demod.rs:62-119 + src/recv.rs:140-167:
//                                         32,768-byte u8 IQ chunks -> DemodTask -> MessageReceiver
//   p25host_main consumer a.f32 ...       replay, then every stream's events through p25cu::RecvConsumer: what RecvTask
//                                         and the hub make of them (src/recv.rs:237-342, src/hub.rs:335-443)
//   p25host_main fields KIND:HEX ...      no GPU: one event per argument (KIND = 7 TSBK, 3 LinkControl, 8 VoiceTerm;
//                                         HEX = payload) through RecvConsumer
// Lines:  E <stream> <sample> <kind> <len> <payload hex>     one per MessageEvent, in delivery order
//         T <stream> <sample> <talkgroup> <rx_freq_hz>       add_talkgroup (src/recv.rs:325-342)
//         H <stream> <json>                                  one hub SerdeEvent (src/hub.rs:505-523)
//         P <chunk> <stream> <dBm>                           signal power (every 4th chunk)
//         B <chunk> <n_out>                                  baseband samples per stream produced by the chunk
//         S <family> <words> <errs> <fixed>                  merged stats;   V <n> voice frames handed to the audio sink
#include <cstdio>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

#include "p25cu.hpp"

static void print_event(const p25cu::MessageEvent& e) {
    std::printf("E %u %llu %u %u ", e.stream, (unsigned long long)e.sample, (unsigned)e.kind, e.len);
    for (uint32_t i = 0; i < e.len; i++) std::printf("%02x", e.payload[i]);
    std::printf("\n");
}

static void print_stats(const p25cu::Stats& st) {
    for (int f = 0; f < P25CU_ST_FAMILIES; f++)
        std::printf("S %s %llu %llu %llu\n", p25cu::Stats::family_name(f), (unsigned long long)st.code[f].words,
                    (unsigned long long)st.code[f].errs, (unsigned long long)st.code[f].fixed);
}

int main(int argc, char** argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s replay|sdr FILE...\n", argv[0]);
        return 2;
    }
    const std::string mode = argv[1];
    if (mode == "fields") {
        p25cu::RecvConsumer rc;
        for (int i = 2; i < argc; i++) {
            const std::string a = argv[i];
            p25cu::MessageEvent e{};
            e.stream = 0;
            e.sample = (unsigned long long)(i - 2);
            e.kind = static_cast<p25cu::MessageEvent::Kind>(std::stoi(a.substr(0, a.find(':'))));
            const std::string hex = a.substr(a.find(':') + 1);
            e.len = (uint32_t)(hex.size() / 2);
            for (uint32_t b = 0; b < e.len && b < 60; b++) e.payload[b] = (uint8_t)std::stoi(hex.substr(2 * b, 2), nullptr, 16);
            rc.handle(e);
        }
        for (const auto& t : rc.talkgroups) std::printf("T 0 %llu %u %llu\n", (unsigned long long)t.sample, t.tg, (unsigned long long)t.rx_freq);
        for (const auto& j : rc.hub_json) std::printf("H 0 %s\n", j.c_str());
        return 0;
    }
    std::vector<std::unique_ptr<std::ifstream>> files;
    for (int i = 2; i < argc; i++) {
        files.emplace_back(new std::ifstream(argv[i], std::ios::binary));
        if (!*files.back()) {
            std::fprintf(stderr, "cannot open %s\n", argv[i]);
            return 2;
        }
    }
    const uint32_t S = (uint32_t)files.size();
    try {
        if (mode == "consumer") {
            p25cu::ReplayReceiver rx(S, nullptr);
            std::vector<std::istream*> in;
            for (auto& f : files) in.push_back(f.get());
            rx.replay(in);
            std::vector<p25cu::RecvConsumer> rc(S);          // one RecvTask per stream, like one p25rx process per channel
            for (const auto& e : rx.events()) rc[e.stream].handle(e);
            for (uint32_t s = 0; s < S; s++) {
                for (const auto& t : rc[s].talkgroups) std::printf("T %u %llu %u %llu\n", s, (unsigned long long)t.sample, t.tg, (unsigned long long)t.rx_freq);
                for (const auto& j : rc[s].hub_json) std::printf("H %u %s\n", s, j.c_str());
            }
        } else if (mode == "replay") {
            unsigned long long voice = 0;
            p25cu::ReplayReceiver rx(S, [&](const p25cu::MessageEvent& e) {
                const p25cu::VoiceFrame vf = e.voice_frame();     // what AudioTask::play receives (src/audio.rs:75-87)
                (void)vf;
                voice++;
            });
            std::vector<std::istream*> in;
            for (auto& f : files) in.push_back(f.get());
            rx.replay(in);
            for (const auto& e : rx.events()) print_event(e);
            print_stats(rx.merged_stats());
            std::printf("V %llu\n", voice);
        } else if (mode == "sdr") {
            p25cu::Context ctx(S, p25cu::Format::U8Iq, 5, p25cu::BUF_SAMPLES);
            p25cu::DemodTask demod(ctx);
            p25cu::MessageReceiver recv(ctx);
            p25cu::Stats stats;
            std::vector<unsigned char> chunk(S * p25cu::BUF_BYTES);
            for (unsigned c = 0;; c++) {
                bool full = true;
                for (uint32_t s = 0; s < S; s++) {
                    files[s]->read((char*)chunk.data() + s * p25cu::BUF_BYTES, p25cu::BUF_BYTES);
                    full = full && (std::size_t)files[s]->gcount() == p25cu::BUF_BYTES;
                }
                if (!full) break;                                                  // partial transfers are dropped
                const p25cu::DemodTask::Chunk out = demod.run_chunk(chunk.data(), p25cu::BUF_SAMPLES);
                std::printf("B %u %zu\n", c, out.n_out);
                if (out.power)
                    for (uint32_t s = 0; s < S; s++) std::printf("P %u %u %.4f\n", c, s, (*out.power)[s]);
                for (const auto& e : recv.feed_demodulated()) {
                    if (e.kind == p25cu::MessageEvent::Error) stats.record_err(e);
                    print_event(e);
                }
            }
            for (uint32_t s = 0; s < S; s++) stats.merge(ctx.stats(s, true));
            print_stats(stats);
        } else {
            std::fprintf(stderr, "unknown mode %s\n", mode.c_str());
            return 2;
        }
    } catch (const p25cu::Error& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}

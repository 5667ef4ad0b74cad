// p25cu.hpp -- C++17 host layer over the C ABI of include/p25cu.h.
//
// The reference is Rust; no Rust toolchain exists where this repository is built and tested, so the host side
// above the ABI is C++ (header-only, no dependencies beyond the C ABI).  It mirrors the reference's interface for
// this path, batched over the streams of one GPU context -- same names, argument meaning and error behaviour:
//
//   p25cu::DemodTask        reference src/demod.rs:25-119   (one run-loop iteration per run_chunk call)
//   p25cu::power_dbm        reference src/demod.rs:123-134  (computed inside run_chunk, every 4th chunk, :67, :95-101)
//   p25cu::MessageReceiver  p25::message::receiver::MessageReceiver as used at src/recv.rs:81, :207, :136
//   p25cu::MessageEvent     the variants matched at src/recv.rs:214-233
//   p25cu::Stats            p25::stats::Stats: merge / clear / record_err (src/recv.rs:159, :212, :215; src/hub.rs:557-581)
//   p25cu::ReplayReceiver   reference src/replay.rs:11-57
//
// Where the reference aborts (`.expect()`, `panic = "abort"`, Cargo.toml:50-51) this layer throws p25cu::Error
// carrying the p25cu_status and the library's message; decode failures stay values (MessageEvent::Error).
#ifndef P25CU_HPP
#define P25CU_HPP

#include <array>
#include <cstdint>
#include <cstring>
#include <functional>
#include <istream>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "p25cu.h"

namespace p25cu {

constexpr std::size_t BUF_BYTES = 32768;            // reference src/consts.rs:6
constexpr std::size_t BUF_SAMPLES = BUF_BYTES / 2;  // reference src/consts.rs:8
constexpr unsigned SDR_SAMPLE_RATE = 240000;        // reference src/consts.rs:11
constexpr unsigned BASEBAND_SAMPLE_RATE = 48000;    // reference src/consts.rs:13

struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string& what) : std::runtime_error("p25cu status " + std::to_string(s) + ": " + what), status(s) {}
};

enum class Format : int32_t { U8Iq = P25CU_FMT_U8_IQ, Cf32Iq = P25CU_FMT_CF32_IQ };

// ---------------------------------------------------------------------------------------------- events
struct NetworkId {          // p25::trunking / nid: NetworkId { access code, data unit } (src/policy.rs:96-131)
    uint16_t nac;
    uint8_t data_unit;      // DUID: 0x0 HDU, 0x3 TDU, 0x5 LDU1, 0x7 TSDU, 0xA LDU2, 0xC PDU, 0xF TDULC
};

struct VoiceFrame {         // src/audio.rs:76: chunks u0..u7 and the per-code-word corrected-bit counts
    std::array<uint32_t, 8> chunks;
    std::array<uint32_t, 7> errors;
};

struct MessageEvent {
    enum Kind : uint32_t {
        Error = P25CU_EV_ERROR, PacketNID = P25CU_EV_NID, VoiceHeader = P25CU_EV_VOICE_HEADER,
        LinkControl = P25CU_EV_LINK_CONTROL, CryptoControl = P25CU_EV_CRYPTO_CONTROL,
        LowSpeedDataFragment = P25CU_EV_LSD, VoiceFrameEv = P25CU_EV_VOICE_FRAME,
        TrunkingControl = P25CU_EV_TSBK, VoiceTerm = P25CU_EV_VOICE_TERM
    };
    uint32_t stream;        // which of the context's streams (the reference has one)
    uint64_t sample;        // absolute 48 kHz sample index at which feed() would have returned this event
    Kind kind;
    uint32_t len;
    std::array<uint8_t, 60> payload;

    uint32_t error_code() const { return word(0); }                       // Error(P25Error)
    NetworkId nid() const { return NetworkId{uint16_t(payload[0] | payload[1] << 8), payload[2]}; }
    const uint8_t* bytes() const { return payload.data(); }               // TSBK 12 B, LC 9 B, CC 12 B, HDU 15 B
    uint32_t lsd() const { return word(0); }                              // LowSpeedDataFragment(u32), src/recv.rs:226
    VoiceFrame voice_frame() const {
        VoiceFrame v;
        for (int i = 0; i < 8; i++) v.chunks[i] = word(i);
        for (int i = 0; i < 7; i++) v.errors[i] = word(8 + i);
        return v;
    }

private:
    uint32_t word(int i) const {
        uint32_t w;
        std::memcpy(&w, payload.data() + 4 * i, 4);
        return w;
    }
};

// ---------------------------------------------------------------------------------------------- stats
struct CodeStats {          // src/hub.rs:574-581
    uint64_t words = 0, errs = 0, size = 0, fixed = 0;
};

struct Stats {              // family order of src/hub.rs:557-572
    std::array<CodeStats, P25CU_ST_FAMILIES> code{};
    uint64_t recorded_errors = 0;

    static const char* family_name(int f) {
        static const char* n[P25CU_ST_FAMILIES] = {"bch", "cyclic", "golayStd", "golayExt", "golayShort", "hammingStd",
                                                   "hammingShort", "rsShort", "rsMed", "rsLong", "viterbiDibit", "viterbiTribit"};
        return n[f];
    }
    void merge(const Stats& o) {                                          // Stats::merge, src/recv.rs:212
        for (int f = 0; f < P25CU_ST_FAMILIES; f++) {
            code[f].words += o.code[f].words;
            code[f].errs += o.code[f].errs;
            code[f].fixed += o.code[f].fixed;
            code[f].size = o.code[f].size;
        }
    }
    void record_err(const MessageEvent&) { recorded_errors++; }            // Stats::record_err, src/recv.rs:215
    void clear() { *this = Stats{}; }                                      // Stats::clear, src/recv.rs:159
};

// ---------------------------------------------------------------------------------------------- context
class Context {
public:
    Context(uint32_t n_streams, Format fmt = Format::U8Iq, int decimation = 5, std::size_t max_chunk_samples = BUF_SAMPLES,
            std::size_t max_baseband = 0, int device = 0, uint32_t event_slots = 0)
        : n_streams_(n_streams), decimation_(decimation) {
        p25cu_config cfg{};
        cfg.device = device;
        cfg.n_streams = n_streams;
        cfg.format = static_cast<int32_t>(fmt);
        cfg.decimation = decimation;
        cfg.max_chunk_samples = max_chunk_samples;
        cfg.max_baseband = max_baseband;
        cfg.abi_version = P25CU_ABI_VERSION;
        cfg.event_slots = event_slots;
        const int rc = p25cu_create(&cfg, &ctx_);
        if (rc != P25CU_OK) throw Error(rc, p25cu_last_error(nullptr));
    }
    ~Context() { p25cu_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;

    uint32_t n_streams() const { return n_streams_; }
    int decimation() const { return decimation_; }
    p25cu_ctx* raw() { return ctx_; }

    // Surface 1.  iq: [n_streams][n_in] samples of the context's format (host memory).  Returns outputs per stream.
    std::size_t demod(const void* iq, std::size_t n_in, float* baseband_out, float* power_dbm = nullptr) {
        std::size_t n_out = 0;
        check(p25cu_demod(ctx_, iq, n_in, 0, baseband_out, &n_out, power_dbm));
        return n_out;
    }
    // Surface 2.  baseband: [n_streams][n] float32 at 48 kHz, or nullptr for the device-resident output of demod().
    void decode(const float* baseband, std::size_t n) { check(p25cu_decode(ctx_, baseband, n)); }
    void process(const void* iq, std::size_t n_in) { check(p25cu_process(ctx_, iq, n_in, 0)); }
    void resync(uint32_t stream) { check(p25cu_resync(ctx_, stream)); }     // MessageReceiver::resync, src/recv.rs:136

    // Every queued event, ordered by (stream, sample): per stream the order in which feed() returns them.
    std::vector<MessageEvent> poll() {
        const p25cu_event* ev = nullptr;
        std::size_t n = 0;
        check(p25cu_poll_view(ctx_, &ev, &n));
        std::vector<MessageEvent> out(n);
        for (std::size_t i = 0; i < n; i++) {
            out[i].stream = ev[i].stream;
            out[i].sample = ev[i].sample;
            out[i].kind = static_cast<MessageEvent::Kind>(ev[i].kind);
            out[i].len = ev[i].len;
            std::memcpy(out[i].payload.data(), ev[i].payload, 60);
        }
        return out;
    }

    Stats stats(uint32_t stream, bool clear = false) {
        p25cu_stats raw{};
        check(p25cu_get_stats(ctx_, stream, &raw, clear ? 1 : 0));
        Stats s;
        for (int f = 0; f < P25CU_ST_FAMILIES; f++)
            s.code[f] = CodeStats{raw.code[f].words, raw.code[f].errs, raw.code[f].size, raw.code[f].fixed};
        return s;
    }

private:
    void check(int rc) {
        if (rc != P25CU_OK) throw Error(rc, p25cu_last_error(ctx_));
    }
    p25cu_ctx* ctx_ = nullptr;
    uint32_t n_streams_;
    int decimation_;
};

// ---------------------------------------------------------------------------------------------- DemodTask
// One iteration of DemodTask::run (src/demod.rs:70-117) for every stream: IQ chunk in, 48 kHz baseband chunk out,
// signal power in dBm on every 4th chunk (Throttler::new(4), src/demod.rs:67, :95-101).
class DemodTask {
public:
    explicit DemodTask(Context& ctx) : ctx_(ctx) {}

    struct Chunk {
        std::vector<float> baseband;              // [n_streams][n_out]
        std::size_t n_out = 0;
        std::optional<std::vector<float>> power;  // HubEvent::UpdateSignalPower, one value per stream
    };

    Chunk run_chunk(const void* iq, std::size_t n_in) {
        Chunk c;
        const bool want_power = notifier_ == 0;
        notifier_ = (notifier_ + 1) % 4;
        c.baseband.resize((n_in / ctx_.decimation() + 1) * ctx_.n_streams());
        std::vector<float> pw(want_power ? ctx_.n_streams() : 0);
        c.n_out = ctx_.demod(iq, n_in, c.baseband.data(), want_power ? pw.data() : nullptr);
        // the library writes rows of n_out samples back to back
        c.baseband.resize(c.n_out * ctx_.n_streams());
        if (want_power) c.power = std::move(pw);
        return c;
    }

private:
    Context& ctx_;
    unsigned notifier_ = 0;
};

// ---------------------------------------------------------------------------------------------- MessageReceiver
// feed() takes one chunk of every stream ([n_streams][n]) instead of one sample and returns every event of the
// chunk, ordered by (stream, sample); stats are merged like `stats.merge(&mut self.msg)` (src/recv.rs:212).
class MessageReceiver {
public:
    explicit MessageReceiver(Context& ctx) : ctx_(ctx) {}
    std::vector<MessageEvent> feed(const float* samples, std::size_t n_per_stream) {
        ctx_.decode(samples, n_per_stream);
        return ctx_.poll();
    }
    std::vector<MessageEvent> feed_demodulated() {      // the device-resident output of the preceding DemodTask / demod()
        ctx_.decode(nullptr, 0);
        return ctx_.poll();
    }
    void resync(uint32_t stream) { ctx_.resync(stream); }

private:
    Context& ctx_;
};

// ---------------------------------------------------------------------------------------------- ReplayReceiver
// src/replay.rs:11-57 for one or more f32le / 48 kHz / mono recordings replayed side by side.
class ReplayReceiver {
public:
    static constexpr std::size_t READ_BYTES = 32768;    // src/replay.rs:27

    ReplayReceiver(uint32_t n_streams, std::function<void(const MessageEvent&)> audio, int device = 0)
        : ctx_(n_streams, Format::U8Iq, 5, BUF_SAMPLES, READ_BYTES / 4, device), msg_(ctx_), audio_(std::move(audio)) {}

    // Reads every stream in 32,768-byte blocks until the shortest recording ends.  Unlike src/replay.rs:36 a short
    // final read is fed at its true length (the reference re-feeds the stale tail of its buffer).
    void replay(const std::vector<std::istream*>& streams) {
        const std::size_t S = streams.size();
        std::vector<std::vector<char>> pend(S);
        std::vector<float> chunk;
        for (;;) {
            std::size_t n = READ_BYTES / 4;
            for (std::size_t s = 0; s < S; s++) {
                if (pend[s].size() < READ_BYTES) {
                    const std::size_t have = pend[s].size();
                    pend[s].resize(have + READ_BYTES);
                    streams[s]->read(pend[s].data() + have, READ_BYTES);
                    pend[s].resize(have + static_cast<std::size_t>(streams[s]->gcount()));
                }
                n = std::min(n, pend[s].size() / 4);
            }
            if (n == 0) break;
            chunk.resize(S * n);
            for (std::size_t s = 0; s < S; s++) {
                std::memcpy(chunk.data() + s * n, pend[s].data(), 4 * n);
                pend[s].erase(pend[s].begin(), pend[s].begin() + 4 * n);
            }
            feed(chunk.data(), n);
        }
    }

    void feed(const float* samples, std::size_t n_per_stream) {          // src/replay.rs:40-57
        for (const MessageEvent& e : msg_.feed(samples, n_per_stream)) {
            if (e.kind == MessageEvent::Error) stats_.record_err(e);
            else if (e.kind == MessageEvent::VoiceFrameEv && audio_) audio_(e);
            events_.push_back(e);
        }
    }

    // `self.stats.merge(&mut self.msg)` over all streams (src/replay.rs:49)
    const Stats& merged_stats() {
        for (uint32_t s = 0; s < ctx_.n_streams(); s++) stats_.merge(ctx_.stats(s, true));
        return stats_;
    }
    const std::vector<MessageEvent>& events() const { return events_; }
    Context& context() { return ctx_; }

private:
    Context ctx_;
    MessageReceiver msg_;
    std::function<void(const MessageEvent&)> audio_;
    Stats stats_;
    std::vector<MessageEvent> events_;
};

}  // namespace p25cu

#endif  // P25CU_HPP

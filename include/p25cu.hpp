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

    // Packed, asynchronous drain (p25cu_poll_start / p25cu_poll_packed): start the compaction of everything decoded so
    // far, collect it later -- while the next chunk's kernels run -- as MessageEvents.
    void poll_start() { check(p25cu_poll_start(ctx_)); }
    std::vector<MessageEvent> poll_packed(bool* more = nullptr) {
        const uint32_t* w = nullptr;
        std::size_t nw = 0, ne = 0;
        int m = 0;
        check(p25cu_poll_packed(ctx_, &w, &nw, &ne, &m));
        if (more) *more = m != 0;
        std::vector<MessageEvent> out;
        out.reserve(ne);
        for (std::size_t cur = 0; cur + 3 <= nw;) {
            const uint32_t w2 = w[cur + 2], len = w2 >> 24;
            MessageEvent e{};
            e.stream = w[cur];
            e.sample = uint64_t(w[cur + 1]) | uint64_t(w2 & 0xFFFF) << 32;
            e.kind = static_cast<MessageEvent::Kind>((w2 >> 16) & 0xFF);
            e.len = len;
            std::memcpy(e.payload.data(), w + cur + 3, len);
            out.push_back(e);
            cur += 3 + (len + 3) / 4;
        }
        return out;
    }
    // Page-locked chunk buffers: the reference's buffer pools (src/demod.rs:63, :103; src/sdr.rs:25-33)
    void* host_alloc(std::size_t bytes) {
        void* p = nullptr;
        check(p25cu_host_alloc(ctx_, bytes, &p));
        return p;
    }
    void host_free(void* p) { check(p25cu_host_free(ctx_, p)); }

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


// ============================================================================================== consumers' view
// the fields the reference's consumers read from the raw payloads the hot path delivers --
// RecvTask::handle_tsbk / handle_lc / add_talkgroup (src/recv.rs:237-342) and the hub's status broadcasts
// (src/hub.rs:335-443).  Integer / byte work on the host.  The reference takes these parsers from the p25 crate
// (not vendored): the layouts below are the CAI's [STD], written from memory like the rest of spec/, and are the same
// ones p25rx_b200/consumers.py implements (tests/test_cpp_host.py diffs the two).

inline uint16_t crc_ccitt_p25(const uint8_t* d, std::size_t n) {      // x^16 + x^12 + x^5 + 1, zero start, inverted
    uint32_t crc = 0;
    for (std::size_t i = 0; i < n; i++) {
        crc ^= uint32_t(d[i]) << 8;
        for (int b = 0; b < 8; b++) crc = (crc & 0x8000) ? ((crc << 1) ^ 0x1021) & 0xFFFF : (crc << 1) & 0xFFFF;
    }
    return uint16_t(crc ^ 0xFFFF);
}

struct Channel {                     // ch.id(), ch.number(): src/recv.rs:335-337
    uint16_t bits;
    uint8_t id() const { return uint8_t(bits >> 12); }
    uint16_t number() const { return bits & 0xFFF; }
};

enum class TsbkOpcode : uint8_t {
    GroupVoiceGrant = 0x00, GroupVoiceUpdate = 0x02, GroupVoiceUpdateExplicit = 0x03, UnitVoiceGrant = 0x04,
    LocRegResponse = 0x2B, UnitRegResponse = 0x2C, UnitDeregAck = 0x2F, AltControlChannel = 0x39,
    RfssStatusBroadcast = 0x3A, NetworkStatusBroadcast = 0x3B, AdjacentSite = 0x3C, ChannelParamsUpdate = 0x3D
};
enum class LinkControlOpcode : uint8_t {
    GroupVoiceTraffic = 0x00, GroupVoiceUpdate = 0x02, UnitVoiceTraffic = 0x03, CallTermination = 0x0F,
    SystemServiceBroadcast = 0x20, AltControlChannel = 0x21, AdjacentSite = 0x22, RfssStatusBroadcast = 0x23,
    NetworkStatusBroadcast = 0x24
};

// The 12 bytes of a TrunkingControl event with the accessors RecvTask::handle_tsbk uses (src/recv.rs:238-249).
struct TsbkFields {
    std::array<uint8_t, 12> raw{};
    explicit TsbkFields(const uint8_t* p) { std::memcpy(raw.data(), p, 12); }
    bool is_tail() const { return raw[0] & 0x80; }
    bool protected_() const { return raw[0] & 0x40; }
    uint8_t opcode_bits() const { return raw[0] & 0x3F; }
    std::optional<TsbkOpcode> opcode() const {
        switch (opcode_bits()) {
            case 0x00: case 0x02: case 0x03: case 0x04: case 0x2B: case 0x2C: case 0x2F: case 0x39: case 0x3A: case 0x3B:
            case 0x3C: case 0x3D: return static_cast<TsbkOpcode>(opcode_bits());
            default: return std::nullopt;
        }
    }
    uint8_t mfg() const { return raw[1]; }
    bool crc_valid() const { return crc_ccitt_p25(raw.data(), 10) == uint16_t(raw[10] << 8 | raw[11]); }
    const uint8_t* payload() const { return raw.data() + 2; }     // 8 bytes
};

// The 9 bytes of a LinkControl / VoiceTerm event (src/recv.rs:277-306).
struct LinkControlFields {
    std::array<uint8_t, 9> raw{};
    explicit LinkControlFields(const uint8_t* p) { std::memcpy(raw.data(), p, 9); }
    std::optional<LinkControlOpcode> opcode() const {
        switch (raw[0] & 0x3F) {
            case 0x00: case 0x02: case 0x03: case 0x0F: case 0x20: case 0x21: case 0x22: case 0x23: case 0x24:
                return static_cast<LinkControlOpcode>(raw[0] & 0x3F);
            default: return std::nullopt;
        }
    }
    const uint8_t* payload() const { return raw.data() + 1; }     // 8 bytes
};

namespace fields {
inline uint32_t be(const uint8_t* p, int n) {
    uint32_t v = 0;
    for (int i = 0; i < n; i++) v = v << 8 | p[i];
    return v;
}
struct GroupVoiceGrant {             // tsbk::GroupVoiceGrant (src/recv.rs:256-258)
    const uint8_t* p;
    explicit GroupVoiceGrant(const TsbkFields& t) : p(t.payload()) {}
    uint8_t opts() const { return p[0]; }
    Channel channel() const { return Channel{uint16_t(be(p + 1, 2))}; }
    uint16_t talkgroup() const { return uint16_t(be(p + 3, 2)); }
    uint32_t src_unit() const { return be(p + 5, 3); }
};
struct GroupTrafficUpdate {          // fields::GroupTrafficUpdate::updates (src/recv.rs:260-262, :300-301, :308-312)
    const uint8_t* p;
    explicit GroupTrafficUpdate(const uint8_t* payload) : p(payload) {}
    std::array<std::pair<Channel, uint16_t>, 2> updates() const {
        return {{{Channel{uint16_t(be(p, 2))}, uint16_t(be(p + 2, 2))}, {Channel{uint16_t(be(p + 4, 2))}, uint16_t(be(p + 6, 2))}}};
    }
};
struct ChannelParamsUpdate {         // fields::ChannelParamsUpdate (src/recv.rs:264-266): id 4 | bandwidth 9 | offset 9 | spacing 10 | base 32
    uint8_t id;
    uint32_t bandwidth_hz;
    int64_t tx_offset_hz;
    uint32_t spacing_hz;
    uint64_t base_hz;
    explicit ChannelParamsUpdate(const uint8_t* p) {
        uint64_t v = 0;
        for (int i = 0; i < 8; i++) v = v << 8 | p[i];
        id = uint8_t(v >> 60);
        bandwidth_hz = uint32_t((v >> 51) & 0x1FF) * 125;
        const uint32_t off = uint32_t((v >> 42) & 0x1FF);
        tx_offset_hz = ((off & 0x100) ? 1 : -1) * int64_t(off & 0xFF) * 250000;
        spacing_hz = uint32_t((v >> 32) & 0x3FF) * 125;
        base_hz = (v & 0xFFFFFFFFull) * 5;
    }
    uint64_t rx_freq(uint16_t number) const { return base_hz + uint64_t(spacing_hz) * number; }   // ChannelParams::rx_freq, src/recv.rs:336
};
struct RfssStatusBroadcast {         // src/hub.rs:528-537: area, system, rfss, site
    const uint8_t* p;
    explicit RfssStatusBroadcast(const uint8_t* payload) : p(payload) {}
    uint8_t area() const { return p[0]; }
    uint16_t system() const { return uint16_t((p[1] & 0xF) << 8 | p[2]); }
    uint8_t rfss() const { return p[3]; }
    uint8_t site() const { return p[4]; }
    Channel channel() const { return Channel{uint16_t(be(p + 5, 2))}; }
};
struct NetworkStatusBroadcast {      // src/hub.rs:539-547: area, wacn, system
    const uint8_t* p;
    explicit NetworkStatusBroadcast(const uint8_t* payload) : p(payload) {}
    uint8_t area() const { return p[0]; }
    uint32_t wacn() const { return uint32_t(p[1]) << 12 | uint32_t(p[2]) << 4 | p[3] >> 4; }
    uint16_t system() const { return uint16_t((p[3] & 0xF) << 8 | p[4]); }
    Channel channel() const { return Channel{uint16_t(be(p + 5, 2))}; }
};
struct AdjacentSite {                // src/hub.rs:423-442: area, rfss, system, site, channel
    const uint8_t* p;
    explicit AdjacentSite(const uint8_t* payload) : p(payload) {}
    uint8_t area() const { return p[0]; }
    uint16_t system() const { return uint16_t((p[1] & 0xF) << 8 | p[2]); }
    uint8_t rfss() const { return p[3]; }
    uint8_t site() const { return p[4]; }
    Channel channel() const { return Channel{uint16_t(be(p + 5, 2))}; }
};
struct AltControlChannel {           // src/hub.rs:405-421: rfss, site, two (channel, services) alternatives
    const uint8_t* p;
    explicit AltControlChannel(const uint8_t* payload) : p(payload) {}
    uint8_t rfss() const { return p[0]; }
    uint8_t site() const { return p[1]; }
    std::array<std::pair<Channel, uint8_t>, 2> alts() const {
        return {{{Channel{uint16_t(be(p + 2, 2))}, p[4]}, {Channel{uint16_t(be(p + 5, 2))}, p[7]}}};
    }
};
struct LocRegResponse {              // src/hub.rs:362-371
    const uint8_t* p;
    explicit LocRegResponse(const TsbkFields& t) : p(t.payload()) {}
    uint8_t response() const { return p[0] & 3; }
    uint8_t rfss() const { return p[3]; }
    uint8_t site() const { return p[4]; }
    uint32_t dest_unit() const { return be(p + 5, 3); }
};
struct UnitRegResponse {             // src/hub.rs:372-381
    const uint8_t* p;
    explicit UnitRegResponse(const TsbkFields& t) : p(t.payload()) {}
    uint8_t response() const { return (p[0] >> 4) & 3; }
    uint16_t system() const { return uint16_t((p[0] & 0xF) << 8 | p[1]); }
    uint32_t src_id() const { return be(p + 2, 3); }
    uint32_t src_addr() const { return be(p + 5, 3); }
};
struct UnitDeregAck {                // src/hub.rs:382-390
    const uint8_t* p;
    explicit UnitDeregAck(const TsbkFields& t) : p(t.payload()) {}
    uint32_t wacn() const { return uint32_t(p[1]) << 12 | uint32_t(p[2]) << 4 | p[3] >> 4; }
    uint16_t system() const { return uint16_t((p[3] & 0xF) << 8 | p[4]); }
    uint32_t src_unit() const { return be(p + 5, 3); }
};
struct GroupVoiceTraffic {           // control::GroupVoiceTraffic::src_unit (src/hub.rs:394-396)
    const uint8_t* raw;
    explicit GroupVoiceTraffic(const LinkControlFields& lc) : raw(lc.raw.data()) {}
    uint16_t talkgroup() const { return uint16_t(be(raw + 4, 2)); }
    uint32_t src_unit() const { return be(raw + 6, 3); }
};
}  // namespace fields

// `self.channels` of RecvTask and of the hub's State (src/recv.rs:265, :335-338; src/hub.rs:476-483, :505-506)
class ChannelParamsMap {
public:
    void update(const fields::ChannelParamsUpdate& p) { m_[p.id & 0xF] = p; }
    const fields::ChannelParamsUpdate* lookup(uint8_t id) const { return m_[id & 0xF] ? &*m_[id & 0xF] : nullptr; }

private:
    std::array<std::optional<fields::ChannelParamsUpdate>, 16> m_{};
};

// What RecvTask and the hub make of a stream's events: talkgroups with their traffic-channel frequency
// (add_talkgroup, src/recv.rs:325-342: only TalkGroup::Other, only known channel identifiers) and the JSON objects the
// hub streams to its subscribers ({"event": ..., "payload": ...}, src/hub.rs:335-443, :505-547).
class RecvConsumer {
public:
    struct Talkgroup {
        uint64_t sample;
        uint16_t tg;
        uint64_t rx_freq;
    };
    std::vector<Talkgroup> talkgroups;
    std::vector<std::string> hub_json;      // one serialised SerdeEvent per entry

    void handle(const MessageEvent& e) {                                   // src/recv.rs:214-233
        if (e.kind == MessageEvent::TrunkingControl) handle_tsbk(e, TsbkFields(e.bytes()));
        else if (e.kind == MessageEvent::LinkControl || e.kind == MessageEvent::VoiceTerm) handle_lc(e, LinkControlFields(e.bytes()));
    }
    const ChannelParamsMap& channels() const { return channels_; }

private:
    static bool other(uint16_t tg) { return tg != 0x0000 && tg != 0x0001 && tg != 0xFFFF; }     // TalkGroup::Other(_)
    void add_talkgroup(const MessageEvent& e, uint16_t tg, Channel ch) {                           // src/recv.rs:325-342
        if (!other(tg)) return;
        const fields::ChannelParamsUpdate* p = channels_.lookup(ch.id());
        if (!p) return;
        talkgroups.push_back(Talkgroup{e.sample, tg, p->rx_freq(ch.number())});
    }
    void emit(const char* name, const std::string& payload) { hub_json.push_back(std::string("{\"event\": \"") + name + "\", \"payload\": " + payload + "}"); }
    static std::string kv(std::initializer_list<std::pair<const char*, uint64_t>> f) {
        std::string s = "{";
        bool first = true;
        for (const auto& x : f) {
            if (!first) s += ", ";
            first = false;
            s += std::string("\"") + x.first + "\": " + std::to_string(x.second);
        }
        return s + "}";
    }
    void rfss(const uint8_t* p) { fields::RfssStatusBroadcast f(p); emit("rfssStatus", kv({{"area", f.area()}, {"system", f.system()}, {"rfss", f.rfss()}, {"site", f.site()}})); }
    void net(const uint8_t* p) { fields::NetworkStatusBroadcast f(p); emit("networkStatus", kv({{"area", f.area()}, {"wacn", f.wacn()}, {"system", f.system()}})); }
    void adjacent(const uint8_t* p) {
        fields::AdjacentSite f(p);
        const fields::ChannelParamsUpdate* c = channels_.lookup(f.channel().id());
        if (!c) return;
        emit("adjacentSite", kv({{"area", f.area()}, {"rfss", f.rfss()}, {"system", f.system()}, {"site", f.site()}, {"freq", c->rx_freq(f.channel().number())}}));
    }
    void alt(const uint8_t* p) {
        fields::AltControlChannel f(p);
        for (const auto& a : f.alts()) {
            const fields::ChannelParamsUpdate* c = channels_.lookup(a.first.id());
            if (!c) continue;
            emit("altControl", kv({{"rfss", f.rfss()}, {"site", f.site()}, {"freq", c->rx_freq(a.first.number())}}));
        }
    }
    void handle_tsbk(const MessageEvent& e, const TsbkFields& t) {         // src/recv.rs:237-274 + src/hub.rs:346-392
        if (t.mfg() != 0 || !t.crc_valid()) return;
        const auto op = t.opcode();
        if (!op) return;
        switch (*op) {
            case TsbkOpcode::GroupVoiceGrant: { fields::GroupVoiceGrant g(t); add_talkgroup(e, g.talkgroup(), g.channel()); break; }
            case TsbkOpcode::GroupVoiceUpdate:
                for (const auto& u : fields::GroupTrafficUpdate(t.payload()).updates()) add_talkgroup(e, u.second, u.first);
                break;
            case TsbkOpcode::ChannelParamsUpdate: channels_.update(fields::ChannelParamsUpdate(t.payload())); break;
            case TsbkOpcode::RfssStatusBroadcast: rfss(t.payload()); break;
            case TsbkOpcode::NetworkStatusBroadcast: net(t.payload()); break;
            case TsbkOpcode::AltControlChannel: alt(t.payload()); break;
            case TsbkOpcode::AdjacentSite: adjacent(t.payload()); break;
            case TsbkOpcode::LocRegResponse: { fields::LocRegResponse f(t); emit("locReg", kv({{"response", f.response()}, {"rfss", f.rfss()}, {"site", f.site()}, {"unit", f.dest_unit()}})); break; }
            case TsbkOpcode::UnitRegResponse: { fields::UnitRegResponse f(t); emit("unitReg", kv({{"response", f.response()}, {"system", f.system()}, {"unitId", f.src_id()}, {"unitAddr", f.src_addr()}})); break; }
            case TsbkOpcode::UnitDeregAck: { fields::UnitDeregAck f(t); emit("unitDereg", kv({{"wacn", f.wacn()}, {"system", f.system()}, {"unit", f.src_unit()}})); break; }
            default: break;
        }
    }
    void handle_lc(const MessageEvent& e, const LinkControlFields& lc) {   // src/recv.rs:277-306 + src/hub.rs:393-404
        const auto op = lc.opcode();
        if (!op) return;
        switch (*op) {
            case LinkControlOpcode::GroupVoiceTraffic: emit("srcUnit", std::to_string(fields::GroupVoiceTraffic(lc).src_unit())); break;
            case LinkControlOpcode::GroupVoiceUpdate:
                for (const auto& u : fields::GroupTrafficUpdate(lc.payload()).updates()) add_talkgroup(e, u.second, u.first);
                break;
            case LinkControlOpcode::RfssStatusBroadcast: rfss(lc.payload()); break;
            case LinkControlOpcode::NetworkStatusBroadcast: net(lc.payload()); break;
            case LinkControlOpcode::AdjacentSite: adjacent(lc.payload()); break;
            case LinkControlOpcode::AltControlChannel: alt(lc.payload()); break;
            default: break;
        }
    }
    ChannelParamsMap channels_;
};

}  // namespace p25cu

#endif  // P25CU_HPP

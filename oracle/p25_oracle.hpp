// ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED.
//
// CPU restatement of the p25rx baseband hot path, written
// sample-at-a-time like the reference.  It is the checker for the CUDA library in
// p25rx_b200/csrc and the timed "restated-reference" CPU baseline of bench.py.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load it.  The product library never links or calls anything in oracle/.
//
// PARITY UNPINNED: the reference (kchmck/p25rx) delegates all arithmetic on this path
// to crates that are not vendored in /root/reference and cannot be built here (no Rust
// toolchain, no network):
//     p25 (git a96c564), p25_filts (git 0d34fc2), static_decimate (git e9e00a0),
//     static_fir 0.2.0, demod_fm 1.0.0, moving_avg 0.1.0, rtlsdr_iq 0.1.0,
//     cai_golay 0.1.1, cai_cyclic 0.1.2, binfield_matrix 0.2.0      (Cargo.lock pins)
// and the reference's own tests hold no vector for this path.  What is
// restated here is (a) the chain structure and parameters visible at the reference's
// call sites, cited per class below, (b) the published P25 CAI algorithms (TIA-102.BAAA-A)
// whose constants are cross-checked algebraically in tests/test_spec_selfcheck.py, and
// (c) build-defined choices (filter taps, sync rule) recorded in spec/p25_spec.py.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "p25_tables.h"

namespace p25o {

struct cf32 {
    float re, im;
};

// ---------------------------------------------------------------------------
// Event record: same layout as include/p25cu.h p25cu_event (80 bytes).
// Variants follow p25::message::receiver::MessageEvent as matched at
// reference src/recv.rs:214-233 and src/replay.rs:51-55.
// ---------------------------------------------------------------------------
enum EventKind : uint32_t {
    EV_ERROR = 0,
    EV_NID = 1,
    EV_VOICE_HEADER = 2,
    EV_LINK_CONTROL = 3,
    EV_CRYPTO_CONTROL = 4,
    EV_LSD = 5,
    EV_VOICE_FRAME = 6,
    EV_TSBK = 7,
    EV_VOICE_TERM = 8,
};
enum ErrorCode : uint32_t { ERR_BCH = 1, ERR_RS = 2, ERR_VITERBI_DIBIT = 3, ERR_VITERBI_TRIBIT = 4, ERR_UNKNOWN_NID = 5 };

struct Event {
    uint32_t stream;
    uint32_t kind;
    uint64_t sample;  // index of the baseband sample whose feed() returned this event
    uint32_t len;
    uint8_t payload[60];
};
static_assert(sizeof(Event) == 80, "event layout");

// Stats families in the order of reference src/hub.rs:557-572.
enum StatFamily { ST_BCH, ST_CYCLIC, ST_GOLAY_STD, ST_GOLAY_EXT, ST_GOLAY_SHORT, ST_HAMMING_STD, ST_HAMMING_SHORT,
                  ST_RS_SHORT, ST_RS_MED, ST_RS_LONG, ST_VITERBI_DIBIT, ST_VITERBI_TRIBIT, ST_FAMILIES };
struct CodeStats {  // reference src/hub.rs:574-581
    uint64_t words, errs, size, fixed;
};
struct Stats {
    CodeStats c[ST_FAMILIES];
    Stats() { clear(); }
    void clear() {
        for (int i = 0; i < ST_FAMILIES; i++) c[i] = CodeStats{0, 0, P25_STATS_SIZE[i], 0};
    }
    void ok(int f, unsigned fixed) {
        c[f].words++;
        c[f].fixed += fixed;
    }
    void bad(int f) {
        c[f].words++;
        c[f].errs++;
    }
};

// ---------------------------------------------------------------------------
// FEC primitives (published P25 algorithms; tables from spec/p25_spec.py)
// ---------------------------------------------------------------------------
int gf_mul(int a, int b);
int gf_div(int a, int b);
// binary BCH(63,16,23): word63 bit i = coefficient of x^i.  returns false if > 11 errors.
bool bch_decode(uint64_t word63, uint16_t* data, int* nerr);
// Golay codes: return number of corrected bits, or -1 if unrecoverable.
int golay23_decode(uint32_t word, uint32_t* data12);
int golay24_decode(uint32_t word, uint32_t* data12);
int golay18_decode(uint32_t word, uint32_t* data6);
int hamming15_decode(uint32_t word, uint32_t* data11);
int hamming10_decode(uint32_t word, uint32_t* data6);
int cyclic16_decode(uint32_t word, uint32_t* data8);
// Reed-Solomon over GF(64): sym[0] is the highest-degree symbol.  Corrects in place.
// returns number of corrected symbols or -1.
int rs_decode(uint8_t* sym, int n, int k);
// half-rate trellis: 98 received dibits -> 12 bytes.  returns corrected bit count or -1.
int trellis_half_decode(const uint8_t* dibits98, uint8_t* out12);
// 3/4-rate trellis: 98 received dibits -> 18 bytes.  returns corrected bit count or -1.
int trellis_34_decode(const uint8_t* dibits98, uint8_t* out18);
// IMBE frame: 72 received dibits -> u0..u7 and 7 error counts.
void imbe_decode(const uint8_t* dibits72, uint32_t chunks[8], uint32_t errors[7]);
uint16_t crc_ccitt_p25(const uint8_t* data, int n);

// ---------------------------------------------------------------------------
// Demodulation chain: reference src/demod.rs:82-114
// ---------------------------------------------------------------------------
// static_fir::FirFilter (call sites src/demod.rs:51,:93): direct-form FIR, ring history.
template <int N>
class FirFilter {
public:
    explicit FirFilter(const float* taps) : taps_(taps), pos_(0) { std::memset(hist_, 0, sizeof hist_); }
    void push(cf32 s) {
        pos_ = (pos_ + 1) % N;
        hist_[pos_] = s;
    }
    cf32 eval() const {
        float re = 0.f, im = 0.f;
        int p = pos_;
        for (int k = 0; k < N; k++) {  // k = 0 is the newest sample
            re = std::fmaf(taps_[k], hist_[p].re, re);
            im = std::fmaf(taps_[k], hist_[p].im, im);
            p = (p == 0) ? N - 1 : p - 1;
        }
        return cf32{re, im};
    }
    cf32 feed(cf32 s) {
        push(s);
        return eval();
    }

private:
    const float* taps_;
    cf32 hist_[N];
    int pos_;
};

// static_decimate::Decimator (call sites src/demod.rs:50,:87): every input enters the
// history, the dot product is evaluated on every D-th input; phase persists across calls.
template <int N>
class Decimator {
public:
    Decimator(const float* taps, int factor) : fir_(taps), factor_(factor), phase_(0) {}
    bool feed(cf32 s, cf32* out) {
        fir_.push(s);
        if (++phase_ == factor_) {
            phase_ = 0;
            *out = fir_.eval();
            return true;
        }
        return false;
    }

private:
    FirFilter<N> fir_;
    int factor_, phase_;
};

// demod_fm::FmDemod::new(5000, 48000) (src/demod.rs:54,:109-111)
class FmDemod {
public:
    FmDemod() : prev_{0.f, 0.f} {}
    float feed(cf32 s) {
        float re = s.re * prev_.re + s.im * prev_.im;
        float im = s.im * prev_.re - s.re * prev_.im;
        prev_ = s;
        return std::atan2(im, re) * P25_FM_GAIN;
    }

private:
    cf32 prev_;
};

// moving_avg::MovingAverage::new(10) (src/demod.rs:52,:114).  Restated as a direct
// 10-term sum (oldest first) so that it is a pure function of the window.
class MovingAverage {
public:
    MovingAverage() : pos_(0) { std::memset(h_, 0, sizeof h_); }
    float feed(float s) {
        h_[pos_] = s;
        pos_ = (pos_ + 1) % P25_BOXCAR;
        float acc = 0.f;
        for (int i = 0; i < P25_BOXCAR; i++) acc += h_[(pos_ + i) % P25_BOXCAR];
        return acc / (float)P25_BOXCAR;
    }

private:
    float h_[P25_BOXCAR];
    int pos_;
};

enum IqFormat { FMT_U8 = 0, FMT_CF32 = 1 };

// DemodTask::run loop body (src/demod.rs:70-117) for one stream.
class DemodChain {
public:
    // mode 0: the reference chain (/5); mode 1: prepend the /10 stage for 2.4 MS/s input (declared extension,
    // SURVEY F4); mode 2: input is already at 48 kS/s (one output channel of the wideband channelizer): only the
    // reference's 48 kHz stages run (src/demod.rs:93-114)
    DemodChain(int fmt, int mode)
        : fmt_(fmt), use_front_(mode == 1), skip_decim_(mode == 2), front_(P25_TAPS_FRONT_H, P25_DECIM_FRONT),
          decim_(P25_TAPS_DECIM_H, P25_DECIM_NATIVE), chan_(P25_TAPS_CHAN_H) {}
    // feeds n IQ samples, writes baseband samples, returns their count.
    // power_dbm (nullable) receives power_dbm() of this call's channel-filtered samples.
    size_t feed(const void* iq, size_t n, float* out, float* power_dbm);

private:
    int fmt_;
    bool use_front_, skip_decim_;
    Decimator<P25_TAPS_FRONT> front_;
    Decimator<P25_TAPS_DECIM> decim_;
    FirFilter<P25_TAPS_CHAN> chan_;
    FmDemod fm_;
    MovingAverage avg_;
};

// ---------------------------------------------------------------------------
// p25::message::receiver::MessageReceiver -- call sites src/recv.rs:81,:207,:136 and
// src/replay.rs:21,:44.  feed() takes one 48 kHz baseband sample and returns at most
// one event.
// ---------------------------------------------------------------------------
class MessageReceiver {
public:
    MessageReceiver();
    bool feed(float s, Event* ev);  // true if *ev was filled
    void resync();
    Stats stats;
    uint64_t samples_fed() const { return n_; }
    // introspection for tests
    int state() const { return state_; }

private:
    enum State { ST_SYNC = 0, ST_NID = 1, ST_PAYLOAD = 2, ST_FLUSH = 3 };
    void enter_sync();
    void lock(uint64_t idx);
    float sample_at(int64_t idx) const;
    int decide(float s) const;
    bool on_dibit(int d, uint64_t idx, Event* ev);
    bool on_nid(uint64_t idx, Event* ev);
    bool on_payload(uint64_t idx, Event* ev);
    bool fail(uint32_t code, uint64_t idx, Event* ev);

    float hist_[256];
    uint64_t n_;
    int state_;
    // sync detector
    bool det_have_prev_, det_prev_above_;
    float det_prev_corr_;
    // symbol clock + slicer
    uint64_t next_sym_;
    float pth_, mid_, nth_;
    uint32_t frame_pos_;
    // NID
    uint64_t nid_bits_;
    int nid_cnt_;
    // payload
    int duid_;
    int cnt_;          // data dibits collected for the current unit / block
    int blocks_;       // TSBK blocks decoded in this TSDU
    int part_;         // LDU part index
    int chunks_;       // LC/CC hamming chunks collected
    uint8_t buf_[800];
    uint8_t hex_[36];
};

}  // namespace p25o

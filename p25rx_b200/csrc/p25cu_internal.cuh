// p25cu_internal.cuh -- structures shared by the kernels and the C ABI (not installed).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/p25cu.h"
#include "p25_fec.cuh"

// ---------------------------------------------------------------------------
// Data layout in HBM
//
//   iq chunk      [S][n]            cf32 (8 B) or u8 pairs (2 B), stream-major (caller's layout)
//   iq tail       2 x [S][HT]       cf32, last HT input samples of every stream (ping-pong)
//   baseband      [S][row_stride]   f32; row = 256-sample history | this chunk's samples
//   walker state  [S]               WalkState, one 1024 B record per stream
//   event slots   [S][ev_cap]       p25cu_event (80 B); per-stream fill counts in WalkState
//   event dense   [sum counts]      compacted (stream, sample)-ordered copy for the host
//   stats         [S][12][3]        u32 words / errs / fixed
// ---------------------------------------------------------------------------
#define P25CU_BB_HIST 256          // baseband history kept in front of every row (>= 231 + lock margin)
#define P25CU_WALK_WARPS 2         // streams (warps) per walker CTA: small enough to co-reside with 3 ddc_fm CTAs per SM

enum { WS_SYNC = 0, WS_NID = 1, WS_PAYLOAD = 2, WS_FLUSH = 3 };

// Per-stream receiver state, persisted between chunks.  Field-for-field the state of the
// restated MessageReceiver (oracle/p25_oracle.hpp); the sample history lives in the baseband row.
struct __align__(16) WalkState {
    unsigned long long next_sym;  // absolute index of the next symbol instant
    unsigned long long pos;       // SYNC: absolute index of the next sample the detector has not seen
    int state, duid, cnt, blocks, part, chunks;
    int have_prev, prev_above;
    float prev_corr, pth, mid, nth;
    unsigned frame_pos;
    unsigned n_events;            // events written to this stream's slots since the last poll
    unsigned overflow;
    unsigned resync_req;          // set by p25cu_resync, honoured at the next chunk
    unsigned char hex[40];
    unsigned char scratch[64];    // decoder work area (syndromes, Viterbi survivors); not carried state
    unsigned char buf[840];       // data dibits of the unit being received (one per byte)
};
static_assert(sizeof(WalkState) % 16 == 0, "WalkState must be a multiple of 16 bytes");

struct WalkParams {
    const float* bb;              // baseband rows
    size_t row_stride;            // floats per row
    unsigned long long p0;        // absolute index of the first sample of this chunk
    unsigned n;                   // samples in this chunk (per stream)
    unsigned n_streams;
    WalkState* states;
    p25cu_event* slots;
    unsigned ev_cap;              // slots per stream
    unsigned* stats;              // [S][12][3]
    const P25DevTables* tables;
    float* bb_next;               // rows the NEXT chunk will be decoded from: receives the 256-sample history
};

struct DdcParams {
    const void* iq;               // chunk, [S][n]
    const void* tail_in;          // [S][HT] samples in the stream's own format (uchar2 or float2)
    void* tail_out;               // [S][HT]
    float* bb;                    // baseband rows (outputs start at column P25CU_BB_HIST)
    size_t row_stride;
    float* power_sum;             // [S], atomically accumulated sum |c|^2 of this chunk
    unsigned long long a0;        // absolute input index of the first sample of this chunk
    unsigned long long m0;        // absolute output index of the first output of this chunk
    unsigned n;                   // input samples per stream in this chunk
    unsigned n_out;               // outputs per stream in this chunk
    unsigned n_streams;
    unsigned seg_out;             // outputs per segment (multiple of the per-iteration block)
    unsigned n_seg;               // segments per stream
    unsigned ht;                  // tail length in samples
    int aligned16;                // every chunk row starts 16-byte aligned and holds whole 16-byte groups
    unsigned* work_counter;       // device ticket counter of the /50 kernel's dynamic work split (monotonic)
    unsigned* ticket_base;        // host: tickets consumed by earlier launches
};

// kernels (defined in ddc_fm.cu / decode_walk.cu)
cudaError_t p25cu_launch_ddc(const DdcParams& p, int format, int decimation, cudaStream_t st);
cudaError_t p25cu_ddc_upload_taps();
unsigned p25cu_ddc_tail_len(int decimation);
cudaError_t p25cu_launch_walk(const WalkParams& p, cudaStream_t st, unsigned max_blocks);
cudaError_t p25cu_launch_compact(const WalkState* states, const p25cu_event* slots, unsigned ev_cap, unsigned n_streams,
                                 unsigned* offsets, p25cu_event* dense, cudaStream_t st);
cudaError_t p25cu_launch_fec_selftest(const P25DevTables* tables, int kind, void* words, size_t count, int n, int k,
                                      void* out_data, int32_t* out_nerr, cudaStream_t st);

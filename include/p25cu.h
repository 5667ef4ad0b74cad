/* p25cu.h -- C ABI of the B200-native P25 baseband hot path (libp25cu.so).
 *
 * Drop-in boundary for the two surfaces the reference (kchmck/p25rx) exposes on this
 * path, batched over many independent streams:
 *
 *   Surface 1, demod:  DemodTask::new / DemodTask::run        reference src/demod.rs:44-59, :62-119
 *                      power_dbm                               reference src/demod.rs:123-134
 *   Surface 2, decode: MessageReceiver::new / feed / resync    reference src/recv.rs:81, :207, :136
 *                      (p25 crate, driven by                   reference src/replay.rs:21, :40-57)
 *                      Stats::merge / clear                    reference src/recv.rs:159, :212
 *
 * Plain pointers and sizes only; no CUDA or torch types.  Every function returns 0 on
 * success or a negative p25cu_status; nothing aborts (the reference's `.expect()` sites
 * become error codes).  A context belongs to one host thread at a time and owns one GPU's
 * device memory; contexts are fully independent (one per GPU, no inter-GPU traffic).
 *
 * There is no CPU fallback: p25cu_create fails with P25CU_ERR_CUDA when no sm_100 device
 * is usable.
 */
#ifndef P25CU_H
#define P25CU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P25CU_ABI_VERSION 1   /* additions since round 1 are new entry points only; existing signatures are unchanged */

typedef enum {
    P25CU_OK = 0,
    P25CU_ERR_ARG = -1,      /* bad argument (null pointer, size over the configured maximum, ...) */
    P25CU_ERR_CUDA = -2,     /* CUDA runtime error; text in p25cu_last_error() */
    P25CU_ERR_STATE = -3,    /* call sequence error (e.g. decode of device baseband before any demod) */
    P25CU_ERR_OVERFLOW = -4  /* more events than the per-stream slots hold; extra events were dropped */
} p25cu_status;

/* Input sample formats of p25cu_demod.
 * U8_IQ  : interleaved unsigned bytes I,Q as delivered by librtlsdr   (reference src/sdr.rs:25-33,
 *          src/demod.rs:72-84; mapped through the rtlsdr_iq table)
 * CF32_IQ: interleaved float32 I,Q (declared extension) */
typedef enum { P25CU_FMT_U8_IQ = 0, P25CU_FMT_CF32_IQ = 1 } p25cu_format;

/* Decimation from the input rate to the 48 kHz baseband rate (reference src/consts.rs:11-13).
 * 5   : 240 kS/s input, the reference's own chain               (reference src/demod.rs:50)
 * 50  : 2.4 MS/s input, a /10 front stage ahead of the same chain (BASELINE.json configs[0]); cf32 or u8
 * 400 : wideband channelizer (BASELINE.json configs[2]; declared extension, the reference has one tuner and one
 *       channel, src/sdr.rs:61-68): every input row is a 19.2 MS/s cf32 capture that a polyphase filter bank splits
 *       into 1,536 channels 12.5 kHz apart, each delivered at 48 kS/s and run through the reference's 48 kHz stages
 *       (src/demod.rs:93-114).  n_streams must be a multiple of 1536: stream = capture * 1536 + channel, channel k
 *       is centred k * 12.5 kHz above the capture centre (k >= 768: (k - 1536) * 12.5 kHz).  format must be CF32_IQ;
 *       p25cu_demod takes [n_streams / 1536][n_in] samples. */
typedef struct {
    int32_t device;             /* CUDA device ordinal */
    uint32_t n_streams;         /* independent streams held by this context */
    int32_t format;             /* p25cu_format */
    int32_t decimation;         /* 5 or 50 */
    uint64_t max_chunk_samples; /* largest n_in_per_stream (demod) this context will be given (at most 2^30) */
    uint64_t max_baseband;      /* largest n_per_stream for p25cu_decode of caller-provided baseband;
                                   0 = derive from max_chunk_samples / decimation */
    uint32_t abi_version;       /* P25CU_ABI_VERSION */
    uint32_t event_slots;       /* event slots per stream between two p25cu_poll calls; 0 = sized for one
                                   chunk (max baseband / 200 + 16) */
} p25cu_config;

/* MessageEvent variants as matched at reference src/recv.rs:214-233. */
typedef enum {
    P25CU_EV_ERROR = 0,          /* payload: uint32 p25cu_error_code */
    P25CU_EV_NID = 1,            /* payload: NAC low byte, NAC high nibble, DUID */
    P25CU_EV_VOICE_HEADER = 2,   /* payload: 15 bytes  MI(9) MFID ALGID KID(2) TGID(2) */
    P25CU_EV_LINK_CONTROL = 3,   /* payload: 9 bytes */
    P25CU_EV_CRYPTO_CONTROL = 4, /* payload: 12 bytes  MI(9) ALGID KID(2) */
    P25CU_EV_LSD = 5,            /* payload: uint32, two decoded low-speed-data bytes */
    P25CU_EV_VOICE_FRAME = 6,    /* payload: uint32 chunks[8] (u0..u7), uint32 errors[7]  (src/audio.rs:76) */
    P25CU_EV_TSBK = 7,           /* payload: 12 bytes incl. CRC; CRC is checked by the consumer (src/recv.rs:242) */
    P25CU_EV_VOICE_TERM = 8      /* payload: 9 bytes link control of a TDULC */
} p25cu_event_kind;

typedef enum {
    P25CU_E_BCH = 1, P25CU_E_RS = 2, P25CU_E_VITERBI_DIBIT = 3, P25CU_E_VITERBI_TRIBIT = 4, P25CU_E_UNKNOWN_NID = 5
} p25cu_error_code;

/* Packet data units (DUID 0xC) are received -- header block, then the data blocks it announces, 3/4-rate trellis for
 * confirmed data, 1/2-rate otherwise -- but no MessageEvent variant carries packet data (src/recv.rs:214-233): they
 * only show up in the viterbiDibit / viterbiTribit stats (src/hub.rs:569-570) and as P25CU_E_VITERBI_* errors. */
typedef struct {
    uint32_t stream;     /* stream index within the context */
    uint32_t kind;       /* p25cu_event_kind */
    uint64_t sample;     /* absolute index (since create / stream start) of the 48 kHz baseband sample
                            at which the reference's feed() would have returned this event */
    uint32_t len;        /* valid payload bytes */
    uint8_t payload[60];
} p25cu_event;           /* 80 bytes */

/* p25::stats::Stats as serialised at reference src/hub.rs:557-581, same family order. */
enum { P25CU_ST_BCH, P25CU_ST_CYCLIC, P25CU_ST_GOLAY_STD, P25CU_ST_GOLAY_EXT, P25CU_ST_GOLAY_SHORT,
       P25CU_ST_HAMMING_STD, P25CU_ST_HAMMING_SHORT, P25CU_ST_RS_SHORT, P25CU_ST_RS_MED, P25CU_ST_RS_LONG,
       P25CU_ST_VITERBI_DIBIT, P25CU_ST_VITERBI_TRIBIT, P25CU_ST_FAMILIES };
typedef struct {
    uint64_t words, errs, size, fixed;   /* CodeStats, reference src/hub.rs:574-581 */
} p25cu_code_stats;
typedef struct {
    p25cu_code_stats code[P25CU_ST_FAMILIES];
} p25cu_stats;

typedef struct p25cu_ctx p25cu_ctx;

/* ---- lifecycle: replaces DemodTask::new + MessageReceiver::new per stream
 *      (reference src/demod.rs:44-59, src/recv.rs:81, src/main.rs:249-251) ---- */
int p25cu_create(const p25cu_config* cfg, p25cu_ctx** out);
void p25cu_destroy(p25cu_ctx* ctx);
/* Text of the last failure on this context (or of the last failed p25cu_create when ctx is NULL). */
const char* p25cu_last_error(const p25cu_ctx* ctx);

/* ---- Surface 1: one DemodTask::run iteration for every stream (reference src/demod.rs:70-117).
 * iq: [n_streams][n_in_per_stream] samples of cfg.format, stream-major, contiguous.
 *     iq_on_device == 0: host memory.  The chunk is copied to one of two device staging buffers on a copy stream, so
 *     the copy of chunk k+1 runs beside the kernels of chunk k; the call returns when ITS copy has completed: the
 *     caller may refill the buffer immediately (like a Checkout going back to the reference's pool, src/demod.rs:70).
 *     Pinned memory (p25cu_host_alloc / p25cu_host_register) is copied by DMA at full PCIe rate; pageable memory is
 *     staged by the driver.
 *     iq_on_device != 0: iq is a device pointer on cfg.device (no copy).  The library reads it on its own CUDA
 *     stream: whatever produced the buffer must have completed (or the caller must have synchronised) before the
 *     call, and the buffer must stay valid and unmodified until the work is done (p25cu_sync / p25cu_poll).
 * baseband_out (nullable): host [n_streams][*n_out] float32; NULL keeps the result on the device
 *     for p25cu_decode(ctx, NULL, ...).
 * n_out: receives floor((n_in + phase) / decimation), the same for every stream.
 * power_dbm (nullable): host float[n_streams], power_dbm() of this chunk's channel-filtered samples.
 * Chunk length is arbitrary; decimator phase and all filter histories carry over in ctx. */
int p25cu_demod(p25cu_ctx* ctx, const void* iq, size_t n_in_per_stream, int iq_on_device,
                float* baseband_out, size_t* n_out, float* power_dbm);

/* ---- Surface 2: MessageReceiver::feed over a chunk of every stream (reference src/recv.rs:204-234,
 * src/replay.rs:40-57).
 * baseband: host [n_streams][n_per_stream] float32 at 48 kHz (copied before the call returns), or NULL to consume the
 *     device-resident output of the preceding p25cu_demod (n_per_stream is then ignored).  Only the newest demodulated
 *     chunk stays on the device: once decoding has begun, a p25cu_decode(NULL) that finds more than one undecoded
 *     chunk returns P25CU_ERR_STATE (event sample indices and the sync history assume a gap-free stream).  A context
 *     used for demodulation only (p25cu_demod with baseband_out, never decoded) is fine.
 * Events are queued inside ctx until a poll. */
int p25cu_decode(p25cu_ctx* ctx, const float* baseband, size_t n_per_stream);

/* p25cu_demod followed by p25cu_decode of its device-resident output, one call (the hot path). */
int p25cu_process(p25cu_ctx* ctx, const void* iq, size_t n_in_per_stream, int iq_on_device);

/* Drain queued events, ordered by (stream, sample): per-stream order equals the order in which the
 * reference's feed() returns them.  Waits for queued GPU work.  *n receives the number written;
 * returns P25CU_ERR_OVERFLOW (after writing what fits) if cap was too small or slots overflowed. */
int p25cu_poll(p25cu_ctx* ctx, p25cu_event* out, size_t cap, size_t* n);
/* Zero-copy variant: *events points at ctx-owned pinned host memory holding all queued events in the same
 * order; valid until the next p25cu_poll / p25cu_poll_view / p25cu_destroy on this context. */
int p25cu_poll_view(p25cu_ctx* ctx, const p25cu_event** events, size_t* n);
/* Number of events p25cu_poll would return now (waits for queued GPU work). */
int p25cu_pending(p25cu_ctx* ctx, size_t* n);

/* ---- Packed, asynchronous event drain (the production path for many thousands of channels).
 * Records are variable-length, 32-bit aligned, ordered by (stream, sample):
 *     word 0            stream
 *     word 1            sample, bits 0..31
 *     word 2            sample bits 32..47 | kind << 16 | len << 24
 *     ceil(len / 4) payload words (bytes beyond len are zero)
 * i.e. a PacketNID takes 16 bytes and a TSBK 24 instead of 80.
 * p25cu_poll_start queues the compaction of everything decoded so far behind the work already submitted and returns at
 * once; the compaction kernel writes the records straight into pinned host memory, so no separate copy follows and
 * the next chunk's kernels run beside it.  At most two started polls may be outstanding.
 * p25cu_poll_packed collects the oldest started poll (starting one first if none is outstanding) and waits for it only.
 * *words points at ctx-owned pinned memory, valid until two further polls have been started.  *more (nullable) is set
 * when the host buffer (256 MiB) could not hold every queued event: the remaining streams' events come with the next poll.
 * Do not interleave with p25cu_poll / p25cu_poll_view while a started poll is outstanding (P25CU_ERR_STATE). */
int p25cu_poll_start(p25cu_ctx* ctx);
int p25cu_poll_packed(p25cu_ctx* ctx, const uint32_t** words, size_t* n_words, size_t* n_events, int* more);
/* Expand packed records into p25cu_event records on the host (no GPU work).  *n receives the number written. */
int p25cu_unpack_events(const uint32_t* words, size_t n_words, p25cu_event* out, size_t cap, size_t* n);

/* ---- Pinned host buffers for sample chunks: the counterpart of the reference's buffer pools (src/demod.rs:63, :103;
 * src/sdr.rs:25-33).  p25cu_host_alloc returns page-locked memory; p25cu_host_register page-locks memory the caller
 * already owns (e.g. a Rust Vec<u8>), p25cu_host_unregister releases it. */
int p25cu_host_alloc(p25cu_ctx* ctx, size_t bytes, void** out);
int p25cu_host_free(p25cu_ctx* ctx, void* p);
int p25cu_host_register(p25cu_ctx* ctx, void* p, size_t bytes);
int p25cu_host_unregister(p25cu_ctx* ctx, void* p);

/* MessageReceiver::resync (reference src/recv.rs:136, :179): takes effect at the next chunk. */
int p25cu_resync(p25cu_ctx* ctx, uint32_t stream);

/* Stats::merge / Stats::clear (reference src/recv.rs:159, :212; schema src/hub.rs:557-581). */
int p25cu_get_stats(p25cu_ctx* ctx, uint32_t stream, p25cu_stats* out, int clear);

/* ---- measurement helpers (bench.py): the CUDA stream all work of ctx is queued on (a cudaStream_t
 * as void*), a wait for it, and the number of kernel launches issued so far. */
void* p25cu_cuda_stream(p25cu_ctx* ctx);
int p25cu_sync(p25cu_ctx* ctx);
/* Pipelining of consecutive chunks (default: on for decimation 50, off otherwise): the decode walker of chunk k runs on a second CUDA stream
 * concurrently with the demod kernel of chunk k+1 (double-buffered baseband).  Results do not depend on it.
 * on = 0 serialises both kernels on the stream returned by p25cu_cuda_stream (per-kernel timing). */
int p25cu_set_overlap(p25cu_ctx* ctx, int on);
uint64_t p25cu_launch_count(const p25cu_ctx* ctx);
/* Demod-kernel timing: while enabled, every p25cu_demod / p25cu_process records a CUDA event pair on the launching
 * stream right around its kernel launch(es) (after any wait for a baseband buffer), at most 128 launches.  Each call
 * waits for the stream, returns the average duration and the number of launches recorded since the previous call,
 * clears the record and sets the enable state. */
int p25cu_demod_timing(p25cu_ctx* ctx, int enable, double* avg_ms, unsigned* count);
/* Device pointer/row stride (in floats) of the baseband produced by the last p25cu_demod. */
int p25cu_device_baseband(p25cu_ctx* ctx, const float** ptr, size_t* row_stride, size_t* n_out);

/* Copy the first n baseband samples of one stream of the last p25cu_demod / p25cu_process to the host (parity tests
 * sample a few streams of a large batch this way). */
int p25cu_read_baseband(p25cu_ctx* ctx, uint32_t stream, float* out, size_t n);

/* Channelizer mode only (test hook): while enabled, every p25cu_demod also keeps the channel spectra of its chunk,
 * c_k[m] = channel k at 48 kS/s AFTER the channel-select filter (the filter is folded into the polyphase prototype, so
 * no unfiltered spectrum exists on the device).  p25cu_channelizer_output copies them, [captures][*n_rows][1536]
 * complex float32, to `out` (nullable: only report *n_rows). */
int p25cu_set_keep_spectra(p25cu_ctx* ctx, int on);
int p25cu_channelizer_output(p25cu_ctx* ctx, float* out, size_t* n_rows);

/* ---- FEC unit entry points: run the device decoders on caller-provided code words (one word per
 * GPU thread), used by the parity tests to compare each decoder with the oracle in bulk.
 * kind: 0 bch63 (uint64 words) 1 golay23 2 golay24 3 golay18 4 hamming15 5 hamming10 6 cyclic16
 *       (uint32 words; out_data uint32, out_nerr int32, -1 = unrecoverable)
 *       7 rs (n,k) on `count` blocks of n bytes, corrected in place in `words`
 *       8 half-rate trellis: `count` blocks of 98 dibit bytes -> 12 bytes each in out_data
 *       9 imbe: `count` blocks of 72 dibit bytes -> 15 uint32 each in out_data
 *       10..13: the warp-cooperative forms the decode walker uses (one warp per word) of kinds 7, 9, 0, 8
 *       14 3/4-rate trellis: `count` blocks of 98 dibit bytes -> 18 bytes each in out_data; 15 its warp-cooperative form */
int p25cu_fec_selftest(p25cu_ctx* ctx, int kind, void* words, size_t count, int n, int k,
                       void* out_data, int32_t* out_nerr);

#ifdef __cplusplus
}
#endif
#endif /* P25CU_H */

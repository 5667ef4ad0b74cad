//! Raw FFI of include/p25cu.h (ABI version 1).  UNVERIFIED SOURCE: never compiled in the authoring environment.
//! Each function names the reference call it replaces (paths relative to kchmck/p25rx).
#![allow(non_camel_case_types)]
use libc::{c_char, c_int, c_void, size_t};

pub const P25CU_ABI_VERSION: u32 = 1;
pub const P25CU_FMT_U8_IQ: i32 = 0; // rtlsdr bytes, src/sdr.rs:25-33, src/demod.rs:72-84
pub const P25CU_FMT_CF32_IQ: i32 = 1; // declared extension

pub const P25CU_OK: c_int = 0;
pub const P25CU_ERR_ARG: c_int = -1;
pub const P25CU_ERR_CUDA: c_int = -2;
pub const P25CU_ERR_STATE: c_int = -3;
pub const P25CU_ERR_OVERFLOW: c_int = -4;

#[repr(C)]
pub struct p25cu_config {
    pub device: i32,
    pub n_streams: u32,
    pub format: i32,
    pub decimation: i32, // 5 (src/demod.rs:50), 50 (2.4 MS/s input) or 400 (19.2 MS/s channelizer, 1536 streams per row)
    pub max_chunk_samples: u64,
    pub max_baseband: u64,
    pub abi_version: u32,
    pub event_slots: u32,
}

/// kind: 0 Error, 1 PacketNID, 2 VoiceHeader, 3 LinkControl, 4 CryptoControl, 5 LowSpeedDataFragment,
/// 6 VoiceFrame, 7 TrunkingControl, 8 VoiceTerm -- the MessageEvent variants matched at src/recv.rs:214-233.
#[repr(C)]
#[derive(Copy, Clone)]
pub struct p25cu_event {
    pub stream: u32,
    pub kind: u32,
    pub sample: u64,
    pub len: u32,
    pub payload: [u8; 60],
}

#[repr(C)]
#[derive(Copy, Clone, Default)]
pub struct p25cu_code_stats {
    pub words: u64,
    pub errs: u64,
    pub size: u64,
    pub fixed: u64,
}

/// Family order of src/hub.rs:557-572.
#[repr(C)]
#[derive(Copy, Clone, Default)]
pub struct p25cu_stats {
    pub code: [p25cu_code_stats; 12],
}

pub enum p25cu_ctx {}

extern "C" {
    /// DemodTask::new + MessageReceiver::new for every stream (src/demod.rs:44-59, src/recv.rs:81).
    pub fn p25cu_create(cfg: *const p25cu_config, out: *mut *mut p25cu_ctx) -> c_int;
    pub fn p25cu_destroy(ctx: *mut p25cu_ctx);
    pub fn p25cu_last_error(ctx: *const p25cu_ctx) -> *const c_char;
    /// One DemodTask::run iteration for every stream (src/demod.rs:70-117); power_dbm = src/demod.rs:123-134.
    pub fn p25cu_demod(ctx: *mut p25cu_ctx, iq: *const c_void, n_in_per_stream: size_t, iq_on_device: c_int,
                       baseband_out: *mut f32, n_out: *mut size_t, power_dbm: *mut f32) -> c_int;
    /// `for &s in samples { msg.feed(s) }` (src/recv.rs:148-150, src/replay.rs:43-47).
    pub fn p25cu_decode(ctx: *mut p25cu_ctx, baseband: *const f32, n_per_stream: size_t) -> c_int;
    pub fn p25cu_process(ctx: *mut p25cu_ctx, iq: *const c_void, n_in_per_stream: size_t, iq_on_device: c_int) -> c_int;
    /// The Option<MessageEvent> results of feed(), ordered by (stream, sample) (src/recv.rs:207-233).
    pub fn p25cu_poll(ctx: *mut p25cu_ctx, out: *mut p25cu_event, cap: size_t, n: *mut size_t) -> c_int;
    pub fn p25cu_poll_view(ctx: *mut p25cu_ctx, events: *mut *const p25cu_event, n: *mut size_t) -> c_int;
    pub fn p25cu_pending(ctx: *mut p25cu_ctx, n: *mut size_t) -> c_int;
    /// MessageReceiver::resync (src/recv.rs:136, :179); effective at the next chunk.
    pub fn p25cu_resync(ctx: *mut p25cu_ctx, stream: u32) -> c_int;
    /// Stats::merge / clear (src/recv.rs:159, :212).
    pub fn p25cu_get_stats(ctx: *mut p25cu_ctx, stream: u32, out: *mut p25cu_stats, clear: c_int) -> c_int;
    pub fn p25cu_sync(ctx: *mut p25cu_ctx) -> c_int;
    pub fn p25cu_set_overlap(ctx: *mut p25cu_ctx, on: c_int) -> c_int;
    pub fn p25cu_channelizer_output(ctx: *mut p25cu_ctx, out: *mut f32, n_rows: *mut size_t) -> c_int;
    // round 2 (include/p25cu.h): packed asynchronous drain, pinned chunk buffers
    pub fn p25cu_poll_start(ctx: *mut p25cu_ctx) -> i32;
    pub fn p25cu_poll_packed(ctx: *mut p25cu_ctx, words: *mut *const u32, n_words: *mut usize, n_events: *mut usize,
                             more: *mut i32) -> i32;
    pub fn p25cu_unpack_events(words: *const u32, n_words: usize, out: *mut p25cu_event, cap: usize, n: *mut usize) -> i32;
    pub fn p25cu_host_alloc(ctx: *mut p25cu_ctx, bytes: usize, out: *mut *mut std::ffi::c_void) -> i32;
    pub fn p25cu_host_free(ctx: *mut p25cu_ctx, p: *mut std::ffi::c_void) -> i32;
    pub fn p25cu_host_register(ctx: *mut p25cu_ctx, p: *mut std::ffi::c_void, bytes: usize) -> i32;
    pub fn p25cu_host_unregister(ctx: *mut p25cu_ctx, p: *mut std::ffi::c_void) -> i32;
}

//! Safe adapters over libp25cu.so that keep p25rx's own event types flowing.
//!
//! UNVERIFIED SOURCE: written against the reference's call sites, never compiled (no Rust toolchain in the
//! authoring environment).  The p25 crate types named in comments (`NetworkId`, `TsbkFields`,
//! `LinkControlFields`, `VoiceFrame`, ...) are the ones p25rx imports at src/recv.rs:7-13; constructing them is
//! left to the maintainer because that crate is not vendored with the reference.
//!
//! Usage inside p25rx (replacing the `demod` and `receiver` threads of src/main.rs:270-287 for N tuners):
//!
//! ```ignore
//! let mut gpu = Batch::new(0, n_tuners, Format::U8Iq, 5, BUF_SAMPLES)?;      // src/consts.rs:6-8
//! loop {
//!     let chunks: Vec<Checkout<Vec<u8>>> = readers.iter().map(|r| r.recv().unwrap()).collect();  // src/demod.rs:70
//!     gpu.process_u8(&chunks)?;                                              // src/demod.rs:72-117 + src/recv.rs:148-150
//!     gpu.drain(|stream, ev| recv_tasks[stream as usize].handle_event(ev))?; // src/recv.rs:214-233, unchanged handlers
//! }
//! ```
pub mod sys;

use std::ffi::CStr;
use std::ptr;

#[derive(Debug)]
pub struct Error {
    pub status: i32,
    pub text: String,
}

#[derive(Copy, Clone)]
pub enum Format {
    U8Iq = 0,
    Cf32Iq = 1,
}

/// One decoded event, still in wire form (see include/p25cu.h for the payload layouts).
pub enum RawEvent<'a> {
    Error(u32),                               // MessageEvent::Error        -> stats.record_err, src/recv.rs:215
    PacketNid { nac: u16, duid: u8 },         // MessageEvent::PacketNID    -> handle_nid, src/recv.rs:216-222
    VoiceHeader(&'a [u8]),                    // 15 bytes                    -> src/recv.rs:223
    LinkControl(&'a [u8]),                    // 9 bytes                     -> handle_lc, src/recv.rs:224
    CryptoControl(&'a [u8]),                  // 12 bytes                    -> src/recv.rs:225
    LowSpeedData(u32),                        // src/recv.rs:226
    VoiceFrame { chunks: [u32; 8], errors: [u32; 7] }, // -> AudioEvent::VoiceFrame, src/recv.rs:227-230, src/audio.rs:76
    TrunkingControl(&'a [u8]),                // 12 bytes incl. CRC          -> handle_tsbk, src/recv.rs:231
    VoiceTerm(&'a [u8]),                      // 9 bytes                     -> src/recv.rs:232
}

pub struct Batch {
    ctx: *mut sys::p25cu_ctx,
    n_streams: usize,
    chunk: usize,
    staging: Vec<u8>,
    notifier: u32, // Throttler::new(4), src/demod.rs:67
}

unsafe impl Send for Batch {}

impl Batch {
    pub fn new(device: i32, n_streams: usize, fmt: Format, decimation: i32, max_chunk_samples: usize) -> Result<Batch, Error> {
        let cfg = sys::p25cu_config {
            device,
            n_streams: n_streams as u32,
            format: fmt as i32,
            decimation,
            max_chunk_samples: max_chunk_samples as u64,
            max_baseband: 0,
            abi_version: sys::P25CU_ABI_VERSION,
            event_slots: 0,
        };
        let mut ctx = ptr::null_mut();
        let rc = unsafe { sys::p25cu_create(&cfg, &mut ctx) };
        if rc != sys::P25CU_OK {
            let text = unsafe { CStr::from_ptr(sys::p25cu_last_error(ptr::null())) }.to_string_lossy().into_owned();
            return Err(Error { status: rc, text });
        }
        Ok(Batch { ctx, n_streams, chunk: max_chunk_samples, staging: Vec::new(), notifier: 0 })
    }

    fn check(&self, rc: i32) -> Result<(), Error> {
        if rc == sys::P25CU_OK {
            return Ok(());
        }
        let text = unsafe { CStr::from_ptr(sys::p25cu_last_error(self.ctx)) }.to_string_lossy().into_owned();
        Err(Error { status: rc, text })
    }

    /// One DemodTask::run iteration (src/demod.rs:70-117) plus the receiver's sample loop (src/recv.rs:148-150) for
    /// every tuner.  `chunks[i]` is tuner i's BUF_BYTES-long buffer (src/sdr.rs:25-33).
    pub fn process_u8<B: AsRef<[u8]>>(&mut self, chunks: &[B]) -> Result<(), Error> {
        assert_eq!(chunks.len(), self.n_streams);
        let bytes = chunks[0].as_ref().len();
        assert!(bytes / 2 <= self.chunk);
        self.staging.clear();
        for c in chunks {
            assert_eq!(c.as_ref().len(), bytes);
            self.staging.extend_from_slice(c.as_ref());
        }
        let rc = unsafe { sys::p25cu_process(self.ctx, self.staging.as_ptr() as *const _, bytes / 2, 0) };
        self.check(rc)
    }

    /// Signal power of every tuner in dBm like power_dbm (src/demod.rs:123-134); the reference reports it on
    /// every 4th chunk (src/demod.rs:67, :95-101): returns None on the other three.
    pub fn demod_with_power<B: AsRef<[u8]>>(&mut self, chunks: &[B]) -> Result<Option<Vec<f32>>, Error> {
        let want = self.notifier == 0;
        self.notifier = (self.notifier + 1) % 4;
        let bytes = chunks[0].as_ref().len();
        self.staging.clear();
        for c in chunks {
            self.staging.extend_from_slice(c.as_ref());
        }
        let mut power = vec![0f32; if want { self.n_streams } else { 0 }];
        let mut n_out = 0usize;
        let rc = unsafe {
            sys::p25cu_demod(self.ctx, self.staging.as_ptr() as *const _, bytes / 2, 0, ptr::null_mut(), &mut n_out,
                             if want { power.as_mut_ptr() } else { ptr::null_mut() })
        };
        self.check(rc)?;
        let rc = unsafe { sys::p25cu_decode(self.ctx, ptr::null(), 0) };
        self.check(rc)?;
        Ok(if want { Some(power) } else { None })
    }

    /// Replay shape (src/replay.rs:40-57): 48 kHz f32 baseband, `samples[s * n .. (s + 1) * n]` belongs to stream s.
    pub fn feed_baseband(&mut self, samples: &[f32], n_per_stream: usize) -> Result<(), Error> {
        assert_eq!(samples.len(), n_per_stream * self.n_streams);
        let rc = unsafe { sys::p25cu_decode(self.ctx, samples.as_ptr(), n_per_stream) };
        self.check(rc)
    }

    /// Hands every queued event to `f` in (stream, sample) order: per stream exactly the order in which the
    /// reference's feed() returns them, so handle_nid / handle_tsbk / handle_lc (src/recv.rs:216-232) run unchanged.
    pub fn drain<F: FnMut(u32, u64, RawEvent)>(&mut self, mut f: F) -> Result<(), Error> {
        let mut evs: *const sys::p25cu_event = ptr::null();
        let mut n = 0usize;
        let rc = unsafe { sys::p25cu_poll_view(self.ctx, &mut evs, &mut n) };
        if rc != sys::P25CU_OK && rc != sys::P25CU_ERR_OVERFLOW {
            return self.check(rc);
        }
        let evs = unsafe { std::slice::from_raw_parts(evs, n) };
        for e in evs {
            let p = &e.payload[..e.len as usize];
            let word = |i: usize| u32::from_le_bytes([p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]]);
            let ev = match e.kind {
                0 => RawEvent::Error(word(0)),
                1 => RawEvent::PacketNid { nac: p[0] as u16 | (p[1] as u16) << 8, duid: p[2] },
                2 => RawEvent::VoiceHeader(p),
                3 => RawEvent::LinkControl(p),
                4 => RawEvent::CryptoControl(p),
                5 => RawEvent::LowSpeedData(word(0)),
                6 => {
                    let mut chunks = [0u32; 8];
                    let mut errors = [0u32; 7];
                    for i in 0..8 {
                        chunks[i] = word(i);
                    }
                    for i in 0..7 {
                        errors[i] = word(8 + i);
                    }
                    RawEvent::VoiceFrame { chunks, errors }
                }
                7 => RawEvent::TrunkingControl(p),
                _ => RawEvent::VoiceTerm(p),
            };
            f(e.stream, e.sample, ev);
        }
        self.check(rc)
    }

    pub fn resync(&mut self, stream: u32) -> Result<(), Error> {
        let rc = unsafe { sys::p25cu_resync(self.ctx, stream) };
        self.check(rc)
    }

    pub fn stats(&mut self, stream: u32, clear: bool) -> Result<sys::p25cu_stats, Error> {
        let mut st = sys::p25cu_stats::default();
        let rc = unsafe { sys::p25cu_get_stats(self.ctx, stream, &mut st, clear as i32) };
        self.check(rc)?;
        Ok(st)
    }
}

impl Drop for Batch {
    fn drop(&mut self) {
        unsafe { sys::p25cu_destroy(self.ctx) }
    }
}

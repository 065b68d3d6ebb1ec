//! Rust-side binding of `libfastq_b200.so` (C ABI: include/fastq_b200.h) for the `fastq` crate.
//!
//! NOT COMPILED in the build environment of this repository (no rustc / cargo there); kept thin
//! for that reason.  It is what a maintainer of aseyboldt/fastq-rs would add next to
//! src/lib.rs: the existing `Parser::each` / `parallel_each` keep their signatures and semantics,
//! GPU delimiting replaces `IdxRecord::from_buffer` (src/records.rs:201-247) underneath, and two new
//! methods (`count`, `stats`) never materialise records on the host.
//!
//! build.rs:   println!("cargo:rustc-link-lib=dylib=fastq_b200");

use std::ffi::CStr;
use std::io::{self, Error, ErrorKind, Read};
use std::os::raw::{c_char, c_int};
use std::ptr;
use std::slice;

pub const FQB_ABI_VERSION: u32 = 2;
pub const FQB_F_HIST: u32 = 0x01;
pub const FQB_F_INDEX: u32 = 0x02;
pub const FQB_F_PARTIAL: u32 = 0x40; // one refill of a longer stream: tail_offset instead of a truncation error

#[repr(C)]
pub struct FqbConfig {
    pub abi_version: u32,
    pub device: i32,
    pub max_len: u32,
    pub reserved0: u32,
    pub slot_bytes: u64,
    pub n_slots: u32,
    pub reserved1: u32,
}

#[repr(C)]
#[derive(Default)]
pub struct FqbResult {
    pub status: i32,
    pub finished: i32,
    pub n_records: u64,
    pub n_lines: u64,
    pub err_offset: u64,
    pub tail_offset: u64,
    pub line_phase: u32,
    pub reserved: u32,
}

pub enum FqbCtx {}

extern "C" {
    fn fqb_create(cfg: *const FqbConfig, out: *mut *mut FqbCtx) -> c_int;
    fn fqb_destroy(ctx: *mut FqbCtx);
    fn fqb_strerror(status: c_int) -> *const c_char;
    fn fqb_last_error(ctx: *mut FqbCtx) -> *const c_char;
    fn fqb_stats_words(max_len: u32) -> usize;
    fn fqb_stats_len_hist_off(max_len: u32) -> usize;
    fn fqb_stats_base_hist_off(max_len: u32) -> usize;
    fn fqb_stats_qual_hist_off(max_len: u32) -> usize;
    fn fqb_stream_begin(ctx: *mut FqbCtx, flags: u32) -> c_int;
    fn fqb_stream_acquire(ctx: *mut FqbCtx, pinned: *mut *mut u8, cap: *mut u64) -> c_int;
    fn fqb_stream_submit(ctx: *mut FqbCtx, n_valid: u64) -> c_int;
    fn fqb_stream_finish(ctx: *mut FqbCtx, res: *mut FqbResult, host_stats: *mut u64) -> c_int;
    fn fqb_parse_host(
        ctx: *mut FqbCtx, bytes: *const u8, n: u64, flags: u32, res: *mut FqbResult,
        host_stats: *mut u64, host_index: *mut u32, index_cap: u64, n_index: *mut u64,
    ) -> c_int;
}

/// Owner of one `fqb_ctx` (streams, pinned ring, device ring, accumulators).
pub struct Ctx(*mut FqbCtx);

impl Ctx {
    pub fn new(max_len: u32) -> io::Result<Ctx> {
        let cfg = FqbConfig { abi_version: FQB_ABI_VERSION, device: 0, max_len, reserved0: 0,
                              slot_bytes: 0, n_slots: 0, reserved1: 0 };
        let mut h = ptr::null_mut();
        let rc = unsafe { fqb_create(&cfg, &mut h) };
        if rc != 0 {
            // no CPU fallback behind this ABI: without a CUDA device the fast path does not exist
            return Err(Error::new(ErrorKind::Other, strerror(rc)));
        }
        Ok(Ctx(h))
    }
    fn check(&self, rc: c_int) -> io::Result<()> {
        if rc == 0 { return Ok(()); }
        let detail = unsafe { CStr::from_ptr(fqb_last_error(self.0)) }.to_string_lossy().into_owned();
        Err(Error::new(ErrorKind::Other, format!("{} ({})", strerror(rc), detail)))
    }
}

impl Drop for Ctx {
    fn drop(&mut self) { unsafe { fqb_destroy(self.0) } }
}

fn strerror(status: c_int) -> String {
    unsafe { CStr::from_ptr(fqb_strerror(status)) }.to_string_lossy().into_owned()
}

/// 1..=5 are the reference's grammar errors with the reference's exact messages
/// (src/records.rs:143-146,157-160,233-238; src/lib.rs:278-291).
fn status_to_result(res: &FqbResult) -> io::Result<()> {
    match res.status {
        0 => Ok(()),
        s @ 1..=5 => Err(Error::new(ErrorKind::InvalidData, strerror(s))),
        s => Err(Error::new(ErrorKind::Other, strerror(s))),
    }
}

/// The per-position statistics a closure over `Record::seq()/qual()` would compute.
pub struct Stats {
    pub max_len: u32,
    words: Vec<u64>,
}

impl Stats {
    pub fn n_records(&self) -> u64 { self.words[0] }
    pub fn n_bases(&self) -> u64 { self.words[1] }
    /// `len_hist[min(len, max_len + 1)]`
    pub fn len_hist(&self) -> &[u64] {
        let o = unsafe { fqb_stats_len_hist_off(self.max_len) };
        &self.words[o..o + self.max_len as usize + 2]
    }
    /// `[pos][A, C, G, T, N, other]`
    pub fn base_hist(&self) -> &[u64] {
        let o = unsafe { fqb_stats_base_hist_off(self.max_len) };
        &self.words[o..o + 6 * self.max_len as usize]
    }
    /// `[pos][raw quality byte]`
    pub fn qual_hist(&self) -> &[u64] {
        let o = unsafe { fqb_stats_qual_hist_off(self.max_len) };
        &self.words[o..o + 256 * self.max_len as usize]
    }
}

/// read with the EINTR retry of Buffer::read_into (src/buffer.rs:86-91)
fn read_retry<R: Read>(reader: &mut R, buf: &mut [u8]) -> io::Result<usize> {
    loop {
        match reader.read(buf) {
            Err(ref e) if e.kind() == ErrorKind::Interrupted => continue,
            other => return other,
        }
    }
}

/// Drive the pinned ring from any `Read`: the thread_reader protocol (src/thread_reader.rs:40-50)
/// with `fqb_stream_acquire` as `empty_recv.recv()` and `fqb_stream_submit` as `full_send.send()`.
fn stream_through<R: Read>(ctx: &Ctx, reader: &mut R, flags: u32, want_stats: bool, max_len: u32)
    -> io::Result<(FqbResult, Vec<u64>)> {
    ctx.check(unsafe { fqb_stream_begin(ctx.0, flags) })?;
    loop {
        let (mut p, mut cap) = (ptr::null_mut(), 0u64);
        ctx.check(unsafe { fqb_stream_acquire(ctx.0, &mut p, &mut cap) })?;
        let slot = unsafe { slice::from_raw_parts_mut(p, cap as usize) };
        let n = read_retry(reader, slot)?;
        ctx.check(unsafe { fqb_stream_submit(ctx.0, n as u64) })?;
        if n == 0 { break; }
    }
    let mut res = FqbResult::default();
    let mut words = vec![0u64; if want_stats { unsafe { fqb_stats_words(max_len) } } else { 0 }];
    let wp = if want_stats { words.as_mut_ptr() } else { ptr::null_mut() };
    ctx.check(unsafe { fqb_stream_finish(ctx.0, &mut res, wp) })?;
    Ok((res, words))
}

/// Methods added to `fastq::Parser<R>` (the struct itself lives in src/lib.rs:132-135).
pub trait GpuParser {
    /// examples/fastq-count.rs on the GPU: number of records; `Err` exactly where `each` errs.
    fn count(self) -> io::Result<u64>;
    /// Per-position base / quality histograms; `Err` exactly where `each` errs.
    fn stats(self, max_len: u32) -> io::Result<Stats>;
}

impl<R: Read> GpuParser for crate::Parser<R> {
    fn count(self) -> io::Result<u64> {
        let ctx = Ctx::new(1)?;
        let mut reader = self.into_reader();
        let (res, _) = stream_through(&ctx, &mut reader, 0, false, 1)?;
        status_to_result(&res)?;
        Ok(res.n_records)
    }

    fn stats(self, max_len: u32) -> io::Result<Stats> {
        let ctx = Ctx::new(max_len)?;
        let mut reader = self.into_reader();
        let (res, words) = stream_through(&ctx, &mut reader, FQB_F_HIST, true, max_len)?;
        status_to_result(&res)?;
        Ok(Stats { max_len, words })
    }
}

/// `Parser::each` with GPU delimiting for in-memory input: the closure runs here, over the line-end
/// index the library returns (four u32 per record = the `IdxRecord` of src/records.rs:56-63).
/// All records before the first bad one are delivered, then the error (src/lib.rs:226-237).
pub fn each_in_memory<F>(bytes: &[u8], mut func: F) -> io::Result<bool>
where
    F: FnMut(crate::RefRecord) -> bool,
{
    let ctx = Ctx::new(1)?;
    let mut res = FqbResult::default();
    let mut index = vec![0u32; bytes.len().max(1)];
    let mut n_index = 0u64;
    ctx.check(unsafe {
        fqb_parse_host(ctx.0, bytes.as_ptr(), bytes.len() as u64, FQB_F_INDEX, &mut res, ptr::null_mut(),
                       index.as_mut_ptr(), index.len() as u64, &mut n_index)
    })?;
    let mut start = 0usize;
    let mut hi = 0u64; // offsets are strictly increasing: a decrease of the low 32 bits is a 4 GiB wrap
    let mut prev = 0u32;
    let mut ends = [0usize; 4];
    for r in 0..res.n_records as usize {
        for k in 0..4 {
            let lo = index[4 * r + k];
            if lo < prev { hi += 1 << 32; }
            prev = lo;
            ends[k] = (hi + lo as u64) as usize;
        }
        let rec = crate::RefRecord::from_parts(&bytes[start..=ends[3]], ends[0] - start, ends[1] - start,
                                              ends[2] - start, ends[3] - start);
        if !func(rec) { return Ok(false); }
        start = ends[3] + 1;
    }
    status_to_result(&res)?;
    Ok(true)
}

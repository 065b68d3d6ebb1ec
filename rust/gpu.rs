//! `src/gpu.rs` for the `fastq` crate (aseyboldt/fastq-rs 0.6.0): the B200 path behind the crate's own
//! `Parser::each` / `Parser::parallel_each`, over the C ABI of `libfastq_b200.so` (include/fastq_b200.h).
//!
//! NOT COMPILED in the build environment of this repository (no rustc / cargo there).  It is written as an
//! IN-CRATE module so that it needs nothing the crate does not have; the whole crate-side patch is:
//!
//!   1. `src/lib.rs`:      `mod gpu;`                                  (next to `mod records;`, src/lib.rs:119-121)
//!   2. `src/records.rs`:  a crate-visible constructor for `IdxRecord`, whose four line-end fields are private
//!                         to that module (src/records.rs:57-63):
//!          impl IdxRecord {
//!              pub(crate) fn from_parts(head: usize, seq: usize, sep: usize, qual: usize, data: (usize, usize)) -> IdxRecord {
//!                  IdxRecord { head, seq, sep, qual, data }
//!              }
//!          }
//!   3. `build.rs`:        `println!("cargo:rustc-link-lib=dylib=fastq_b200");`
//!   4. (to make it THE path) the bodies of `Parser::each` (src/lib.rs:221-238) and `Parser::parallel_each`
//!      (src/lib.rs:509-566) become `self.gpu_each(func)` / `self.gpu_parallel_each(n_threads, func)`; their
//!      signatures gain `R: Send` (the reader moves to a reader thread, as in `thread_reader`).
//!
//! Everything else used here exists in the crate as it is: `Parser { reader, buffer }` (private fields, visible to a
//! child module), `RecordSet::from_records` (private fn of the crate root, src/lib.rs:314), `IdxRecord::to_ref_record`
//! (src/records.rs:178), `RefRecord`, `BUFSIZE`.
//!
//! What runs where: record delimiting ('\n' scan, '@' / '+' / length validation, src/records.rs:201-247) on the GPU;
//! the caller's closures here, over `RefRecord`s that borrow the pinned ring (inside `each`) or over `RecordSet`s
//! that own a copy of their 68 KiB (inside `parallel_each`, like the reference's own `RecordSet`).

use std::io::{Error, ErrorKind, Read, Result};
use std::os::raw::c_int;
use std::sync::mpsc::sync_channel;
use std::sync::Arc;
use std::{ptr, slice, thread};

use crate::records::IdxRecord;
use crate::{Parser, RecordSet, RefRecord, BUFSIZE};

pub const FQB_ABI_VERSION: u32 = 3;
const FQB_OK: c_int = 0;
const FQB_E_CANCELLED: c_int = 53;

#[repr(C)]
struct FqbConfig {
    abi_version: u32,
    device: i32,
    max_len: u32,
    reserved0: u32,
    slot_bytes: u64,
    n_slots: u32,
    reserved1: u32,
}

#[repr(C)]
struct FqbResult {
    status: i32,
    finished: i32,
    n_records: u64,
    n_lines: u64,
    err_offset: u64,
    tail_offset: u64,
    line_phase: u32,
    reserved: u32,
}

#[repr(C)]
struct FqbBatch {
    bytes: *const u8,
    n_bytes: u64,
    n_avail: u64,
    stream_offset: u64,
    line_ends: *const u32,
    n_records: u64,
    first_record: u64,
    err_offset: u64,
    token: u64,
    status: i32,
    last: i32,
}

enum FqbCtx {}

extern "C" {
    fn fqb_create(cfg: *const FqbConfig, out: *mut *mut FqbCtx) -> c_int;
    fn fqb_destroy(ctx: *mut FqbCtx);
    fn fqb_stream_acquire(ctx: *mut FqbCtx, pinned: *mut *mut u8, cap: *mut u64) -> c_int;
    fn fqb_stream_submit(ctx: *mut FqbCtx, n_valid: u64) -> c_int;
    fn fqb_batch_begin(ctx: *mut FqbCtx, flags: u32) -> c_int;
    fn fqb_batch_close(ctx: *mut FqbCtx) -> c_int;
    fn fqb_next_batch(ctx: *mut FqbCtx, out: *mut FqbBatch) -> c_int;
    fn fqb_release_batch(ctx: *mut FqbCtx, token: u64) -> c_int;
    fn fqb_batch_cancel(ctx: *mut FqbCtx) -> c_int;
    fn fqb_batch_end(ctx: *mut FqbCtx, res: *mut FqbResult) -> c_int;
}

/// One `fqb_ctx` (device buffers, pinned ring, streams).  The producer and the consumer side of the batch mode are
/// two threads by contract (include/fastq_b200.h), hence `Send + Sync` for the raw handle.
struct Ctx(*mut FqbCtx);
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}
impl Ctx {
    fn new() -> Result<Ctx> {
        let cfg = FqbConfig { abi_version: FQB_ABI_VERSION, device: 0, max_len: 150, reserved0: 0,
                              slot_bytes: 8 << 20, n_slots: 4, reserved1: 0 };
        let mut p: *mut FqbCtx = ptr::null_mut();
        match unsafe { fqb_create(&cfg, &mut p) } {
            FQB_OK => Ok(Ctx(p)),
            rc => Err(Error::new(ErrorKind::Other, format!("fqb_create failed ({}): no CUDA device? there is no CPU fallback", rc))),
        }
    }
}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { fqb_destroy(self.0) }
    }
}

/// status 1..5 -> the crate's own errors, same kind and text
/// (src/records.rs:143-146,157-160,234-237; src/lib.rs:279-282,287-290)
fn grammar_error(status: i32) -> Error {
    let msg = match status {
        1 => "Fastq headers must start with '@'",
        2 => "Sequence and quality not separated by +",
        3 => "Sequence and quality length mismatch",
        4 => "Fastq record is too long",
        5 => "Possibly truncated input file",
        _ => "fastq_b200: device or library failure",
    };
    Error::new(if (1..=5).contains(&status) { ErrorKind::InvalidData } else { ErrorKind::Other }, msg)
}

/// The reader side: `thread_reader`'s loop (src/thread_reader.rs:40-50) with the pinned ring slots as its buffers.
/// One `read` per acquire (short reads top the slot up on the next turn), `Interrupted` retried (src/buffer.rs:86-91).
fn pump<R: Read>(ctx: &Ctx, reader: &mut R) -> Result<()> {
    loop {
        let (mut slot, mut cap) = (ptr::null_mut::<u8>(), 0u64);
        match unsafe { fqb_stream_acquire(ctx.0, &mut slot, &mut cap) } {
            FQB_OK => {}
            FQB_E_CANCELLED => return Ok(()), // the consumer stopped (closure returned false, or an error batch)
            rc => return Err(grammar_error(rc)),
        }
        let dest = unsafe { slice::from_raw_parts_mut(slot, cap as usize) };
        let n = loop {
            match reader.read(dest) {
                Ok(n) => break n,
                Err(ref e) if e.kind() == ErrorKind::Interrupted => continue,
                Err(e) => {
                    unsafe {
                        fqb_stream_submit(ctx.0, 0);
                        fqb_batch_cancel(ctx.0);
                    }
                    return Err(e); // reader errors pass through unchanged
                }
            }
        };
        if unsafe { fqb_stream_submit(ctx.0, n as u64) } != FQB_OK {
            unsafe { fqb_batch_cancel(ctx.0) };
            return Err(grammar_error(100));
        }
        if n == 0 {
            return match unsafe { fqb_batch_close(ctx.0) } {
                FQB_OK => Ok(()),
                rc => Err(grammar_error(rc)),
            };
        }
    }
}

/// the k-th record of a batch, which starts at `bytes[start]`: its four line ends relative to the record start
/// (the fields of the crate's `IdxRecord`, src/records.rs:56-63) and its length (`data.1 - data.0`, :240-246)
#[inline]
fn record_parts(b: &FqbBatch, k: usize, start: usize) -> (usize, usize, usize, usize, usize) {
    let le = unsafe { slice::from_raw_parts(b.line_ends.add(4 * k), 4) };
    let base = b.stream_offset as u32;
    let rel = |x: u32| x.wrapping_sub(base) as usize - start;
    (rel(le[0]), rel(le[1]), rel(le[2]), rel(le[3]), rel(le[3]) + 1)
}

/// `Buffer::replace_buffer` + `read_into` in stream coordinates (src/buffer.rs:30-48,74-100): where the reference's
/// next buffer fill ends, given that the record at `cur` was incomplete in the fill that ended at `fill_end`.
fn next_fill(fill_end: u64, cur: u64) -> u64 {
    let n = fill_end - cur;
    let new_end = (n + 15) & !15;
    let free = BUFSIZE as u64 - new_end;
    fill_end + if free < 4096 { free } else { free - free % 4096 }
}

impl<R: Read + Send> Parser<R> {
    /// `Parser::each` (src/lib.rs:221-238) with GPU delimiting: every record in order, `Ok(true)` at the end of the
    /// input, `Ok(false)` if the closure stopped, `Err` -- after all records in front of the bad one -- on bad input.
    pub fn gpu_each<F>(self, mut func: F) -> Result<bool>
    where
        F: FnMut(RefRecord) -> bool,
    {
        let Parser { mut reader, buffer: _ } = self;
        let ctx = Ctx::new()?;
        if unsafe { fqb_batch_begin(ctx.0, 0) } != FQB_OK {
            return Err(grammar_error(100));
        }
        let outcome = thread::scope(|s| {
            let producer = s.spawn(|| pump(&ctx, &mut reader));
            let mut result: Result<bool> = Ok(true);
            loop {
                let mut b: FqbBatch = unsafe { std::mem::zeroed() };
                match unsafe { fqb_next_batch(ctx.0, &mut b) } {
                    FQB_OK => {}
                    FQB_E_CANCELLED => break, // the reader failed: its error is the result (below)
                    rc => {
                        result = Err(grammar_error(rc));
                        break;
                    }
                }
                let bytes = unsafe { slice::from_raw_parts(b.bytes, b.n_bytes as usize) };
                let mut start = 0usize;
                let mut stopped = false;
                for k in 0..b.n_records as usize {
                    let (head, seq, sep, qual, len) = record_parts(&b, k, start);
                    let rec = IdxRecord::from_parts(head, seq, sep, qual, (start, start + len));
                    start += len;
                    if !func(rec.to_ref_record(bytes)) {
                        stopped = true;
                        break;
                    }
                }
                if b.token != u64::MAX {
                    unsafe { fqb_release_batch(ctx.0, b.token) };
                }
                if stopped {
                    result = Ok(false);
                    break;
                }
                if b.status != 0 {
                    result = Err(grammar_error(b.status));
                    break;
                }
                if b.last != 0 {
                    break;
                }
            }
            unsafe { fqb_batch_cancel(ctx.0) }; // (no-op after a clean end; stops a reader that is still reading)
            match producer.join().expect("reader thread paniced") {
                Err(e) => Err(e), // the reader's own io::Error wins, as in RecordRefIter::advance
                Ok(()) => result,
            }
        });
        unsafe { fqb_batch_end(ctx.0, ptr::null_mut()) };
        outcome
    }

    /// `Parser::parallel_each` (src/lib.rs:509-566): `n_threads` workers `worker-{i}`, one `sync_channel(10)` each,
    /// RecordSets dealt round-robin, a failed send stops the producer, results collected in worker order.  The
    /// RecordSets are cut where the reference's buffer fills would cut them (`next_fill`), own a copy of their bytes
    /// and are dropped -- like the reference's -- when the fill they belong to is the one that meets the error.
    pub fn gpu_parallel_each<O, S, F>(self, n_threads: usize, func: F) -> Result<S>
    where
        S: std::iter::FromIterator<O>,
        O: Send + 'static,
        F: Send + Sync + 'static,
        F: Fn(Box<dyn Iterator<Item = RecordSet>>) -> O,
    {
        let Parser { mut reader, buffer: _ } = self;
        let ctx = Ctx::new()?;
        if unsafe { fqb_batch_begin(ctx.0, 0) } != FQB_OK {
            return Err(grammar_error(100));
        }
        let mut senders = vec![];
        let mut threads: Vec<thread::JoinHandle<_>> = vec![];
        let func = Arc::new(func);
        for i in 0..n_threads {
            let (tx, rx) = sync_channel::<RecordSet>(10);
            let func = func.clone();
            let t = thread::Builder::new().name(format!("worker-{}", i)).spawn(move || func(Box::new(rx.into_iter())))?;
            senders.push(tx);
            threads.push(t);
        }
        let io_error = thread::scope(|s| -> Option<Error> {
            let producer = s.spawn(|| pump(&ctx, &mut reader));
            let mut err = None;
            let mut turn = senders.iter().cycle();
            // the set being built: records complete inside the current fill of the reference's buffer
            let (mut buf, mut recs): (Vec<u8>, Vec<IdxRecord>) = (Vec::with_capacity(BUFSIZE), Vec::new());
            let mut fill_end = BUFSIZE as u64;
            let mut hung_up = false;
            let mut flush = |buf: &mut Vec<u8>, recs: &mut Vec<IdxRecord>| -> bool {
                let set = RecordSet::from_records(std::mem::take(buf).into_boxed_slice(), std::mem::take(recs));
                turn.next().map_or(true, |tx| tx.send(set).is_ok())
            };
            // (the reference's first set is empty: its buffer starts empty, src/lib.rs:381-391)
            if n_threads > 0 && !flush(&mut buf, &mut recs) {
                hung_up = true;
            }
            'batches: while !hung_up {
                let mut b: FqbBatch = unsafe { std::mem::zeroed() };
                match unsafe { fqb_next_batch(ctx.0, &mut b) } {
                    FQB_OK => {}
                    FQB_E_CANCELLED => break,
                    rc => {
                        err = Some(grammar_error(rc));
                        break;
                    }
                }
                let bytes = unsafe { slice::from_raw_parts(b.bytes, b.n_bytes as usize) };
                let mut start = 0usize;
                for k in 0..b.n_records as usize {
                    let (head, seq, sep, qual, len) = record_parts(&b, k, start);
                    let (s_abs, e_abs) = (b.stream_offset + start as u64, b.stream_offset + (start + len) as u64 - 1);
                    while e_abs >= fill_end {
                        // incomplete in this fill: the set goes out, the buffer is refilled (src/lib.rs:393-415)
                        if !flush(&mut buf, &mut recs) {
                            hung_up = true; // a worker quit: stop parsing (src/lib.rs:540-542)
                            break 'batches;
                        }
                        fill_end = next_fill(fill_end, s_abs);
                    }
                    let at = buf.len();
                    buf.extend_from_slice(&bytes[start..start + len]);
                    start += len;
                    recs.push(IdxRecord::from_parts(head, seq, sep, qual, (at, at + len)));   // data = range in the set's own buffer (src/lib.rs:417-418)
                }
                let (status, last) = (b.status, b.last);
                if b.token != u64::MAX {
                    unsafe { fqb_release_batch(ctx.0, b.token) };
                }
                if status != 0 {
                    // records of the fill in which the reference detects the error are dropped with it
                    // (src/lib.rs:375,399-410); `buf` / `recs` are simply not sent.  (The Python mirror of this
                    // shim -- fastq_rs_b200/parser.py: record_sets -- also replays the case of an error the
                    // reference only detects one fill later, and is checked set by set against the oracle.)
                    err = Some(grammar_error(status));
                    break;
                }
                if last != 0 {
                    flush(&mut buf, &mut recs);
                    break;
                }
            }
            unsafe { fqb_batch_cancel(ctx.0) };
            match producer.join().expect("reader thread paniced") {
                Err(e) => Some(e),
                Ok(()) => err,
            }
        });
        ::std::mem::drop(senders); // the workers' iterators end (src/lib.rs:551)
        unsafe { fqb_batch_end(ctx.0, ptr::null_mut()) };
        let results = threads.into_iter().map(|t| t.join());
        if let Some(e) = io_error {
            for r in results {
                r.expect("Panic in worker thread."); // src/lib.rs:556-559
            }
            return Err(e);
        }
        Ok(results.map(|r| r.expect("Panic in worker thread.")).collect())
    }
}

// ---- the fast paths that never materialise records on the host: count() and stats() -------------------------
extern "C" {
    fn fqb_stream_begin(ctx: *mut FqbCtx, flags: u32) -> c_int;
    fn fqb_stream_finish(ctx: *mut FqbCtx, res: *mut FqbResult, host_stats: *mut u64) -> c_int;
    fn fqb_stats_words(max_len: u32) -> usize;
    fn fqb_stats_len_hist_off(max_len: u32) -> usize;
    fn fqb_stats_base_hist_off(max_len: u32) -> usize;
    fn fqb_stats_qual_hist_off(max_len: u32) -> usize;
}
const FQB_F_HIST: u32 = 0x01;

/// The per-position statistics a closure over `Record::seq()` / `qual()` (src/records.rs:82-90) would compute.
pub struct Stats {
    pub max_len: u32,
    words: Vec<u64>,
}

impl Stats {
    pub fn n_records(&self) -> u64 {
        self.words[0]
    }
    pub fn n_bases(&self) -> u64 {
        self.words[1]
    }
    /// `len_hist[min(len, max_len + 1)]`
    pub fn len_hist(&self) -> &[u64] {
        let o = unsafe { fqb_stats_len_hist_off(self.max_len) };
        &self.words[o..o + self.max_len as usize + 2]
    }
    /// `[pos][A, C, G, T, N, other]` (the alphabet of `validate_dnan`, src/records.rs:29-33)
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

impl<R: Read> Parser<R> {
    /// the reader through the pinned ring on the caller's thread (acquire = `empty_recv.recv()`, submit =
    /// `full_send.send()`, src/thread_reader.rs:40-50); H2D copies and kernels run behind it on side streams
    fn gpu_stream(self, flags: u32, max_len: u32) -> Result<(FqbResult, Vec<u64>)> {
        let Parser { mut reader, buffer: _ } = self;
        let cfg = FqbConfig { abi_version: FQB_ABI_VERSION, device: 0, max_len, reserved0: 0, slot_bytes: 0, n_slots: 0, reserved1: 0 };
        let mut p: *mut FqbCtx = ptr::null_mut();
        if unsafe { fqb_create(&cfg, &mut p) } != FQB_OK {
            return Err(grammar_error(100));
        }
        let ctx = Ctx(p);
        if unsafe { fqb_stream_begin(ctx.0, flags) } != FQB_OK {
            return Err(grammar_error(100));
        }
        loop {
            let (mut slot, mut cap) = (ptr::null_mut::<u8>(), 0u64);
            if unsafe { fqb_stream_acquire(ctx.0, &mut slot, &mut cap) } != FQB_OK {
                return Err(grammar_error(100));
            }
            let dest = unsafe { slice::from_raw_parts_mut(slot, cap as usize) };
            let n = loop {
                match reader.read(dest) {
                    Err(ref e) if e.kind() == ErrorKind::Interrupted => continue,
                    other => break other?,
                }
            };
            if unsafe { fqb_stream_submit(ctx.0, n as u64) } != FQB_OK {
                return Err(grammar_error(100));
            }
            if n == 0 {
                break;
            }
        }
        let mut res: FqbResult = unsafe { std::mem::zeroed() };
        let mut words = vec![0u64; if flags & FQB_F_HIST != 0 { unsafe { fqb_stats_words(max_len) } } else { 0 }];
        let wp = if words.is_empty() { ptr::null_mut() } else { words.as_mut_ptr() };
        if unsafe { fqb_stream_finish(ctx.0, &mut res, wp) } != FQB_OK {
            return Err(grammar_error(100));
        }
        if res.status != 0 {
            return Err(grammar_error(res.status));
        }
        Ok((res, words))
    }

    /// examples/fastq-count.rs on the GPU: the number of records; `Err` exactly where `each` errs.
    pub fn gpu_count(self) -> Result<u64> {
        Ok(self.gpu_stream(0, 1)?.0.n_records)
    }

    /// per-position base / quality histograms over `max_len` positions; `Err` exactly where `each` errs
    pub fn gpu_stats(self, max_len: u32) -> Result<Stats> {
        let (_, words) = self.gpu_stream(FQB_F_HIST, max_len)?;
        Ok(Stats { max_len, words })
    }
}

//! `BatchReader` / `Batch` / `Cursor` over the C ABI.  UNCOMPILED: see ../README.md.
//!
//! The device keeps a batch as rows sorted by (key, val, time) plus a hash index; the ABI's
//! cursor calls are batched (`mzgpu_batch_seek_keys`, `mzgpu_batch_key_page`, `mzgpu_batch_rows`).
//! `GpuCursor` answers DD's one-key-at-a-time `Cursor` protocol from a host-side window: a
//! page of key runs and the rows of the current key, refilled by one device call when the
//! cursor leaves the window.  Walks such as `mz_join_core`'s key merge
//! (src/compute/src/render/join/mz_join_core.rs:606-621: seek_key on both cursors) or
//! `walk_cursor` (src/compute/src/render/context.rs:1299-1355) therefore cost one device round
//! trip per page, not per key.  (The join/reduce *operators* of the core do not go through
//! this cursor at all: they probe on the device; this is for operators that stay in Rust.)
use std::rc::Rc;
use differential_dataflow::trace::{Batch, BatchReader, Cursor, Description};
use timely::progress::frontier::{Antichain, AntichainRef};

use super::sys::{self, KeyRun, R32};
use super::worker_ctx;

const PAGE_KEYS: u64 = 4096;

struct Handle(*mut sys::Batch);
impl Drop for Handle { fn drop(&mut self) { unsafe { sys::mzgpu_batch_release(self.0) } } }

#[derive(Clone)]
pub struct GpuBatch { h: Rc<Handle>, desc: Description<u64>, len: usize }

impl GpuBatch {
    /// Takes over the reference the library returned.
    pub unsafe fn from_raw(h: *mut sys::Batch) -> Self {
        let d = sys::mzgpu_batch_desc(h);
        let ac = |t: u64| if t == sys::FRONTIER_EMPTY { Antichain::new() } else { Antichain::from_elem(t) };
        GpuBatch { desc: Description::new(ac(d.lower), ac(d.upper), ac(d.since)), len: sys::mzgpu_batch_len(h) as usize, h: Rc::new(Handle(h)) }
    }
    pub fn raw(&self) -> *mut sys::Batch { self.h.0 }
    pub fn export_rows(&self) -> Vec<R32> {
        let mut v = vec![R32::default(); self.len];
        unsafe { sys::check(worker_ctx(), sys::mzgpu_batch_rows(self.raw(), 0, self.len as u64, v.as_mut_ptr() as *mut _, sys::MEM_HOST)).expect("batch_rows"); }
        v
    }
}

impl BatchReader for GpuBatch {
    type Key<'a> = &'a u64;
    type Val<'a> = &'a u64;
    type Time = u64;
    type TimeGat<'a> = &'a u64;
    type Diff = i64;
    type DiffGat<'a> = &'a i64;
    type Cursor = GpuCursor;
    fn cursor(&self) -> GpuCursor { GpuCursor::new() }
    fn len(&self) -> usize { self.len }
    fn description(&self) -> &Description<u64> { &self.desc }
}

impl Batch for GpuBatch {
    type Merger = GpuMerger;
    fn empty(lower: Antichain<u64>, upper: Antichain<u64>) -> Self {
        use differential_dataflow::trace::Builder;
        super::builder::GpuBuilder::with_capacity(0, 0, 0).done(Description::new(lower, upper, Antichain::from_elem(0)))
    }
}

/// Batch::Merger: the device merges in one step when `done` is called (fuel is accounted by the
/// spine on the library side; a stand-alone merger has nothing to do in `work`).
pub struct GpuMerger { since: u64 }
impl differential_dataflow::trace::Merger<GpuBatch> for GpuMerger {
    fn new(_b1: &GpuBatch, _b2: &GpuBatch, frontier: AntichainRef<u64>) -> Self {
        GpuMerger { since: frontier.iter().next().copied().unwrap_or(sys::FRONTIER_EMPTY) }
    }
    fn work(&mut self, _b1: &GpuBatch, _b2: &GpuBatch, fuel: &mut isize) { *fuel = (*fuel).max(1); }
    fn done(self, b1: &GpuBatch, b2: &GpuBatch) -> GpuBatch {
        let mut out = std::ptr::null_mut();
        unsafe {
            sys::check(worker_ctx(), sys::mzgpu_batch_merge(b1.raw(), b2.raw(), self.since, &mut out)).expect("batch_merge");
            GpuBatch::from_raw(out)
        }
    }
}

/// Host-side window over a batch: `runs` = a page of distinct keys, `rows` = updates of the
/// page's keys (fetched with one `mzgpu_batch_rows` call per page).
pub struct GpuCursor {
    runs: Vec<KeyRun>,
    page_first_ordinal: u64,
    rows: Vec<R32>,
    rows_first: u64,
    key_i: usize,  // index into `runs`
    val_i: usize,  // row index (into `rows`) of the current val's first update
    at_end: bool,
}

impl GpuCursor {
    fn new() -> Self { GpuCursor { runs: vec![], page_first_ordinal: 0, rows: vec![], rows_first: 0, key_i: 0, val_i: 0, at_end: false } }

    fn load_page(&mut self, b: &GpuBatch, first_ordinal: u64) {
        let mut runs = vec![KeyRun::default(); PAGE_KEYS as usize];
        let mut n = 0u64;
        unsafe { sys::check(worker_ctx(), sys::mzgpu_batch_key_page(b.raw(), first_ordinal, PAGE_KEYS, sys::MEM_HOST, runs.as_mut_ptr(), &mut n)).expect("key_page"); }
        runs.truncate(n as usize);
        self.page_first_ordinal = first_ordinal;
        self.at_end = runs.is_empty();
        self.runs = runs;
        self.key_i = 0;
        self.load_rows(b);
    }
    fn load_rows(&mut self, b: &GpuBatch) {
        if let (Some(f), Some(l)) = (self.runs.first(), self.runs.last()) {
            let (first, len) = (f.first, l.first + l.len - f.first);
            self.rows.resize(len as usize, R32::default());
            unsafe { sys::check(worker_ctx(), sys::mzgpu_batch_rows(b.raw(), first, len, self.rows.as_mut_ptr() as *mut _, sys::MEM_HOST)).expect("batch_rows"); }
            self.rows_first = first;
        }
        self.rewind_vals_inner();
    }
    fn run(&self) -> &KeyRun { &self.runs[self.key_i] }
    fn key_rows(&self) -> std::ops::Range<usize> {
        let r = self.run();
        let a = (r.first - self.rows_first) as usize;
        a..a + r.len as usize
    }
    fn rewind_vals_inner(&mut self) { if !self.runs.is_empty() && self.key_i < self.runs.len() { self.val_i = self.key_rows().start; } }
}

impl Cursor for GpuCursor {
    type Key<'a> = &'a u64;
    type Val<'a> = &'a u64;
    type Time = u64;
    type TimeGat<'a> = &'a u64;
    type Diff = i64;
    type DiffGat<'a> = &'a i64;
    type Storage = GpuBatch;

    fn key_valid(&self, _b: &GpuBatch) -> bool { !self.at_end && self.key_i < self.runs.len() }
    fn val_valid(&self, _b: &GpuBatch) -> bool { self.key_valid(_b) && self.val_i < self.key_rows().end }
    fn key<'a>(&self, _b: &'a GpuBatch) -> &'a u64 { unsafe { &*(&self.run().key as *const u64) } }
    fn val<'a>(&self, _b: &'a GpuBatch) -> &'a u64 { unsafe { &*(&self.rows[self.val_i].val as *const u64) } }

    /// (time, diff) of every update of the current (key, val): the rows that share the val.
    fn map_times<L: FnMut(&u64, &i64)>(&mut self, _b: &GpuBatch, mut logic: L) {
        let end = self.key_rows().end;
        let v = self.rows[self.val_i].val;
        let mut i = self.val_i;
        while i < end && self.rows[i].val == v { logic(&self.rows[i].time, &self.rows[i].diff); i += 1; }
    }
    fn step_key(&mut self, b: &GpuBatch) {
        self.key_i += 1;
        if self.key_i >= self.runs.len() && !self.at_end { let next = self.page_first_ordinal + self.runs.len() as u64; self.load_page(b, next); }
        self.rewind_vals_inner();
    }
    /// seek_key: one batched seek (n = 1 here; operators that know their probe keys up front call
    /// `mzgpu_batch_seek_keys` with all of them) positions the window at the first key >= `key`.
    fn seek_key(&mut self, b: &GpuBatch, key: &u64) {
        if self.key_valid(b) {
            // inside the current page?
            if let Some(p) = self.runs[self.key_i..].iter().position(|r| r.key >= *key) { self.key_i += p; self.rewind_vals_inner(); return; }
        }
        let mut run = KeyRun::default();
        unsafe { sys::check(worker_ctx(), sys::mzgpu_batch_seek_keys(b.raw(), key, 1, sys::MEM_HOST, &mut run)).expect("seek_keys"); }
        if run.len == 0 { self.at_end = true; self.runs.clear(); return; }
        // continue paging from the found key: its ordinal is not known, so the window is this run
        // followed by on-demand pages located through `first`
        self.runs = vec![run];
        self.key_i = 0;
        self.at_end = false;
        self.page_first_ordinal = u64::MAX - 1;  // (step_key re-seeks with key + 1, below)
        self.load_rows(b);
    }
    fn step_val(&mut self, _b: &GpuBatch) {
        let end = self.key_rows().end;
        let v = self.rows[self.val_i].val;
        while self.val_i < end && self.rows[self.val_i].val == v { self.val_i += 1; }
    }
    fn seek_val(&mut self, _b: &GpuBatch, val: &u64) {
        let end = self.key_rows().end;
        while self.val_i < end && self.rows[self.val_i].val < *val { self.val_i += 1; }
    }
    fn rewind_keys(&mut self, b: &GpuBatch) { self.load_page(b, 0); }
    fn rewind_vals(&mut self, _b: &GpuBatch) { self.rewind_vals_inner(); }
}

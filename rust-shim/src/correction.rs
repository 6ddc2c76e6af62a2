//! The MV sink's correction buffer on the device.  UNCOMPILED: see ../README.md.
//!
//! Same surface as `CorrectionV2<D>` (src/compute/src/sink/correction_v2.rs:188-390: `insert`,
//! `insert_negated`, `updates_before`, `advance_since`, `consolidate_at_since`) for
//! `D = (u64, u64)`.  The reference keeps chains of sorted chunks and merges them lazily; the device
//! buffer stashes inserts as time-major rows and consolidates everything buffered by (time, data)
//! at a read (include/mzgpu.h, "correction buffer") — the updates a reader sees are the same, in the
//! same order.
use timely::progress::frontier::Antichain;

use super::sys::{self, R32};
use super::worker_ctx;

pub struct GpuCorrection { h: *mut sys::Correction, out: *mut sys::Buf }

impl GpuCorrection {
    pub fn new() -> Self {
        let (mut h, mut out) = (std::ptr::null_mut(), std::ptr::null_mut());
        unsafe {
            sys::check(worker_ctx(), sys::mzgpu_correction_new(worker_ctx(), &mut h)).expect("correction_new");
            sys::check(worker_ctx(), sys::mzgpu_buf_new(worker_ctx(), sys::ROW_R32, &mut out)).expect("buf_new");
        }
        GpuCorrection { h, out }
    }
    pub fn insert(&mut self, updates: &mut Vec<((u64, u64), u64, i64)>) { self.insert_inner(updates, false) }
    pub fn insert_negated(&mut self, updates: &mut Vec<((u64, u64), u64, i64)>) { self.insert_inner(updates, true) }
    fn insert_inner(&mut self, updates: &mut Vec<((u64, u64), u64, i64)>, negate: bool) {
        unsafe {
            sys::check(worker_ctx(), sys::mzgpu_correction_insert(self.h, updates.as_ptr() as *const R32, updates.len() as u64,
                                                                  sys::MEM_HOST, negate as i32)).expect("correction_insert");
        }
        updates.clear();
    }
    /// Consolidated updates at times not beyond `upper` (times advanced to the buffer's since),
    /// ordered by (time, data) — what `updates_before` iterates.
    pub fn updates_before(&mut self, upper: &Antichain<u64>) -> Vec<((u64, u64), u64, i64)> {
        let up = upper.elements().first().copied().unwrap_or(sys::FRONTIER_EMPTY);
        unsafe {
            sys::check(worker_ctx(), sys::mzgpu_buf_clear(self.out)).expect("buf_clear");
            sys::check(worker_ctx(), sys::mzgpu_correction_updates_before(self.h, up, self.out)).expect("updates_before");
            let n = sys::mzgpu_buf_len(self.out) as usize;
            let mut rows = vec![((0u64, 0u64), 0u64, 0i64); n];
            let mut got = 0u64;
            sys::check(worker_ctx(), sys::mzgpu_buf_download(self.out, rows.as_mut_ptr() as *mut _, n as u64, sys::MEM_HOST, &mut got))
                .expect("buf_download");
            rows.truncate(got as usize);
            rows
        }
    }
    pub fn advance_since(&mut self, since: Antichain<u64>) {
        let s = since.elements().first().copied().unwrap_or(sys::FRONTIER_EMPTY);
        unsafe { sys::check(worker_ctx(), sys::mzgpu_correction_advance_since(self.h, s)).expect("advance_since"); }
    }
    pub fn consolidate_at_since(&mut self) {
        unsafe { sys::check(worker_ctx(), sys::mzgpu_correction_consolidate_at_since(self.h)).expect("consolidate_at_since"); }
    }
    pub fn len(&self) -> u64 { unsafe { sys::mzgpu_correction_len(self.h) } }
}
impl Drop for GpuCorrection {
    fn drop(&mut self) { unsafe { sys::mzgpu_correction_free(self.h); sys::mzgpu_buf_free(self.out); } }
}

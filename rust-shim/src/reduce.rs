//! `mz_reduce_abelian` instances over a device arrangement.  UNCOMPILED: see ../README.md.
//!
//! One handle covers the plans of src/compute/src/render/reduce.rs that are reductions of an
//! arrangement with an abelian diff: accumulable COUNT/SUM (`build_accumulable`, :1261-1471: explode
//! into the `Accum` semigroup, arrange, `reduce_abelian`, `finalize_accum` :1575-1584,1671-1700),
//! DISTINCT (`build_distinct`, :264-334), threshold (src/compute/src/render/threshold.rs:33-77),
//! MIN / MAX (`build_bucketed_negated_output`, :1050-1135) and TopK (`BasicTopKPlan`,
//! src/compute/src/render/top_k.rs:215-248).  An activation takes the input updates that arrived
//! (device buffer or host slice), seals the operator's own input arrangement at `upper` and appends
//! the output corrections (`sys::Rout`) — the `(key, aggregates) +-1` rows `reduce_abelian` emits.
use super::sys::{self, R32};
use super::worker_ctx;

pub enum ReduceKind {
    CountSumI64, CountSumF64, Distinct, Threshold, Min, Max,
    /// `limit < 0`: LIMIT NULL
    TopK { limit: i64, offset: u64, descending: bool },
}

pub struct GpuReduce { h: *mut sys::Reduce }

impl GpuReduce {
    pub fn new(kind: ReduceKind) -> Result<Self, (i32, String)> {
        let mut h = std::ptr::null_mut();
        let st = unsafe {
            match kind {
                ReduceKind::TopK { limit, offset, descending } =>
                    sys::mzgpu_topk_new(worker_ctx(), limit, offset, descending as i32, &mut h),
                k => sys::mzgpu_reduce_new(worker_ctx(), match k {
                    ReduceKind::CountSumI64 => sys::AGG_COUNT_SUM_I64,
                    ReduceKind::CountSumF64 => sys::AGG_COUNT_SUM_F64,
                    ReduceKind::Distinct => sys::AGG_DISTINCT,
                    ReduceKind::Threshold => sys::AGG_THRESHOLD,
                    ReduceKind::Min => sys::AGG_MIN,
                    _ => sys::AGG_MAX,
                }, &mut h),
            }
        };
        // plans the device subset cannot hold (a TopK window wider than 32 on a wide group, ...)
        // come back as E_UNSUPPORTED: the caller renders the Rust operator instead
        unsafe { sys::check(worker_ctx(), st)?; }
        Ok(GpuReduce { h })
    }
    /// One activation over host updates `((key, val), time, diff)`.
    pub fn step_host(&mut self, updates: &[((u64, u64), u64, i64)], upper: u64, out: *mut sys::Buf) -> Result<(), (i32, String)> {
        let st = unsafe {
            sys::mzgpu_reduce_accumulable(self.h, updates.as_ptr() as *const R32, updates.len() as u64, sys::MEM_HOST, upper, out)
        };
        unsafe { sys::check(worker_ctx(), st) }
    }
    /// One activation over a device buffer (the result collection of a delta / linear join): the
    /// row count stays on the device, nothing is read back.
    pub fn step(&mut self, rows: *mut sys::Buf, upper: u64, out: *mut sys::Buf) -> Result<(), (i32, String)> {
        unsafe { sys::check(worker_ctx(), sys::mzgpu_reduce_accumulable_buf(self.h, rows, upper, out)) }
    }
    /// The operator's input arrangement (for compaction: `TraceManager::maintenance`,
    /// src/compute/src/arrangement/manager.rs:55-73).
    pub fn input_trace(&self) -> *mut sys::Spine { unsafe { sys::mzgpu_reduce_input_trace(self.h) } }
}
impl Drop for GpuReduce {
    fn drop(&mut self) { unsafe { sys::mzgpu_reduce_free(self.h) } }
}

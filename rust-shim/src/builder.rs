//! `Builder` over the C ABI.  UNCOMPILED: see ../README.md.
//!
//! `Builder::{with_capacity, push, done, seal}` as `arrange_core` and the merge batcher use
//! them (src/timely-util/src/operator.rs:647-677).  When the chain comes from a `GpuBatcher`
//! it is the sealed batch already (`seal` just returns it); `push`/`done` serve callers that
//! hold plain update vectors (`mzgpu_builder_*`).
use differential_dataflow::trace::{Builder, Description};

use super::batch::GpuBatch;
use super::batcher::SealedChunk;
use super::sys;
use super::worker_ctx;

pub struct GpuBuilder { h: *mut sys::Builder }

impl Builder for GpuBuilder {
    type Input = SealedChunk;
    type Time = u64;
    type Output = GpuBatch;

    fn with_capacity(_keys: usize, _vals: usize, upds: usize) -> Self {
        let mut h = std::ptr::null_mut();
        unsafe { sys::check(worker_ctx(), sys::mzgpu_builder_new(worker_ctx(), sys::ROW_R32, upds as u64, &mut h)).expect("builder_new"); }
        GpuBuilder { h }
    }

    /// A chunk that is already a device batch is appended by its rows (device to device).
    fn push(&mut self, chunk: &mut Self::Input) {
        let rows = chunk.0.export_rows();
        unsafe { sys::check(worker_ctx(), sys::mzgpu_builder_push(self.h, rows.as_ptr() as *const _, rows.len() as u64, sys::MEM_HOST)).expect("builder_push"); }
    }

    fn done(self, description: Description<u64>) -> GpuBatch {
        let d = sys::Desc {
            lower: description.lower().elements().first().copied().unwrap_or(sys::FRONTIER_EMPTY),
            upper: description.upper().elements().first().copied().unwrap_or(sys::FRONTIER_EMPTY),
            since: description.since().elements().first().copied().unwrap_or(sys::FRONTIER_EMPTY),
        };
        let mut out = std::ptr::null_mut();
        unsafe { sys::check(worker_ctx(), sys::mzgpu_builder_done(self.h, d, &mut out)).expect("builder_done"); }
        unsafe { GpuBatch::from_raw(out) }
    }

    /// The GPU batcher's chain is one finished batch: nothing to build.
    fn seal(chain: &mut Vec<Self::Input>, description: Description<u64>) -> GpuBatch {
        match chain.len() {
            1 => chain.pop().unwrap().0,
            _ => {
                let mut b = Self::with_capacity(0, 0, 0);
                for mut c in chain.drain(..) { b.push(&mut c); }
                b.done(description)
            }
        }
    }
}

impl Drop for GpuBuilder {
    fn drop(&mut self) { unsafe { sys::mzgpu_builder_free(self.h) } }
}

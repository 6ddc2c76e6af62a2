//! `Batcher` over the C ABI.  UNCOMPILED: see ../README.md.
//!
//! Replaces `MergeBatcher<Vec<..>, ColumnationChunker<..>, ColMerger<..>>` /
//! `ColumnMerger` (src/timely-util/src/columnar/batcher.rs:617-802) for
//! `((u64, u64), u64, i64)` updates.  `arrange_core` (src/timely-util/src/operator.rs:572-633)
//! calls `push_container` per input container, `seal::<Bu>(upper)` when the frontier advances and
//! `frontier()` afterwards to downgrade its capabilities: exactly the three ABI calls below.
use std::marker::PhantomData;
use differential_dataflow::logging::Logger;
use differential_dataflow::trace::{Batcher, Builder, Description};
use timely::progress::frontier::{Antichain, AntichainRef};

use super::batch::GpuBatch;
use super::sys::{self, R32};
use super::worker_ctx;

pub struct GpuBatcher {
    h: *mut sys::Batcher,
    lower: Antichain<u64>,
    frontier: Antichain<u64>,
}

/// The chain a GPU batcher hands to a builder: the sealed batch itself (already sorted,
/// consolidated and indexed on the device), wrapped so that `Builder::seal` can take it.
pub struct SealedChunk(pub GpuBatch);

impl Batcher for GpuBatcher {
    type Input = Vec<((u64, u64), u64, i64)>;
    type Output = SealedChunk;
    type Time = u64;

    fn new(_logger: Option<Logger>, _operator_id: usize) -> Self {
        let mut h = std::ptr::null_mut();
        unsafe { sys::check(worker_ctx(), sys::mzgpu_batcher_new(worker_ctx(), sys::ROW_R32, &mut h)).expect("batcher_new"); }
        GpuBatcher { h, lower: Antichain::from_elem(0), frontier: Antichain::new() }
    }

    /// Chunker::push_into: here the container is only stashed (one H2D copy); sorting happens
    /// once, at seal, over everything buffered (include/mzgpu.h, a2).
    fn push_container(&mut self, container: &mut Self::Input) {
        // ((k, v), t, d) has the layout of mzgpu_r32 (four 8-byte words, no padding)
        let rows = container.as_ptr() as *const R32;
        unsafe {
            sys::check(worker_ctx(), sys::mzgpu_batcher_push(self.h, rows as *const _, container.len() as u64, sys::MEM_HOST))
                .expect("batcher_push");
        }
        container.clear();
    }

    /// MergeBatcher::seal: ship `time < upper`, keep the rest, description
    /// `[previous upper, upper)` with `since = [0]`.
    fn seal<B: Builder<Input = Self::Output, Time = u64>>(&mut self, upper: Antichain<u64>) -> B::Output {
        let up = upper.elements().first().copied().unwrap_or(sys::FRONTIER_EMPTY);
        let (mut batch, mut new_lower) = (std::ptr::null_mut(), 0u64);
        unsafe { sys::check(worker_ctx(), sys::mzgpu_batcher_seal(self.h, up, &mut batch, &mut new_lower)).expect("batcher_seal"); }
        self.frontier = if new_lower == sys::FRONTIER_EMPTY { Antichain::new() } else { Antichain::from_elem(new_lower) };
        let desc = Description::new(self.lower.clone(), upper.clone(), Antichain::from_elem(0));
        self.lower = upper;
        let mut chain = vec![SealedChunk(unsafe { GpuBatch::from_raw(batch) })];
        B::seal(&mut chain, desc)
    }

    fn frontier(&mut self) -> AntichainRef<'_, u64> { self.frontier.borrow() }
}

impl Drop for GpuBatcher {
    fn drop(&mut self) { unsafe { sys::mzgpu_batcher_free(self.h) } }
}

#[allow(dead_code)]
struct _Marker(PhantomData<R32>);

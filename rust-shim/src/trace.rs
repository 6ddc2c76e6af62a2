//! `TraceReader` / `Trace` over the C ABI.  UNCOMPILED: see ../README.md.
//!
//! Stands where `RowRowSpine` / `Spine<Rc<OrdValBatch<..>>>` stand in
//! src/compute/src/typedefs.rs:46-126.  The fueled merge schedule (levels, fuel, roll-up,
//! ExertionLogic) runs inside the library (`mzgpu_spine_*`), so this type only forwards.
use differential_dataflow::trace::cursor::CursorList;
use differential_dataflow::trace::{ExertionLogic, Trace, TraceReader};
use timely::dataflow::operators::generic::OperatorInfo;
use timely::progress::frontier::{Antichain, AntichainRef};

use super::batch::{GpuBatch, GpuCursor};
use super::sys;
use super::worker_ctx;

pub struct GpuSpine {
    h: *mut sys::Spine,
    logical: Antichain<u64>,
    physical: Antichain<u64>,
    proportionality: u32,
}

fn t(a: AntichainRef<u64>) -> u64 { a.iter().next().copied().unwrap_or(sys::FRONTIER_EMPTY) }

impl TraceReader for GpuSpine {
    type Key<'a> = &'a u64;
    type Val<'a> = &'a u64;
    type Time = u64;
    type TimeGat<'a> = &'a u64;
    type Diff = i64;
    type DiffGat<'a> = &'a i64;
    type Batch = GpuBatch;
    type Storage = Vec<GpuBatch>;
    type Cursor = CursorList<GpuCursor>;

    /// cursor_through(upper): the batches whose upper <= `upper` (mz_join_core.rs:243-246).
    fn cursor_through(&mut self, upper: AntichainRef<u64>) -> Option<(Self::Cursor, Self::Storage)> {
        let mut raw = vec![std::ptr::null_mut(); 128];
        let mut n = 0u32;
        let st = unsafe { sys::mzgpu_spine_batches_through(self.h, t(upper), raw.as_mut_ptr(), raw.len() as u32, &mut n) };
        if st == sys::E_FRONTIER { return None; }
        unsafe { sys::check(worker_ctx(), st).expect("batches_through"); }
        let storage: Vec<GpuBatch> = raw[..n as usize].iter().map(|&b| unsafe { sys::mzgpu_batch_retain(b); GpuBatch::from_raw(b) }).collect();
        let cursors = storage.iter().map(|b| { use differential_dataflow::trace::BatchReader; b.cursor() }).collect::<Vec<_>>();
        Some((CursorList::new(cursors, &storage), storage))
    }
    fn set_logical_compaction(&mut self, frontier: AntichainRef<u64>) {
        self.logical = frontier.to_owned();
        unsafe { sys::check(worker_ctx(), sys::mzgpu_spine_set_logical_compaction(self.h, t(frontier))).expect("logical"); }
    }
    fn get_logical_compaction(&mut self) -> AntichainRef<'_, u64> { self.logical.borrow() }
    fn set_physical_compaction(&mut self, frontier: AntichainRef<u64>) {
        self.physical = frontier.to_owned();
        unsafe { sys::check(worker_ctx(), sys::mzgpu_spine_set_physical_compaction(self.h, t(frontier))).expect("physical"); }
    }
    fn get_physical_compaction(&mut self) -> AntichainRef<'_, u64> { self.physical.borrow() }
    fn map_batches<F: FnMut(&GpuBatch)>(&self, mut f: F) {
        let mut raw = vec![std::ptr::null_mut(); 128];
        let mut n = 0u32;
        unsafe { sys::check(worker_ctx(), sys::mzgpu_spine_batches_through(self.h, sys::FRONTIER_EMPTY, raw.as_mut_ptr(), raw.len() as u32, &mut n)).expect("batches"); }
        for &b in &raw[..n as usize] { unsafe { sys::mzgpu_batch_retain(b); f(&GpuBatch::from_raw(b)); } }
    }
    fn read_upper(&mut self, target: &mut Antichain<u64>) {
        target.clear();
        let u = unsafe { sys::mzgpu_spine_read_upper(self.h) };
        if u != sys::FRONTIER_EMPTY { target.insert(u); }
    }
}

impl Trace for GpuSpine {
    fn new(_info: OperatorInfo, _logging: Option<differential_dataflow::logging::Logger>, _activator: Option<timely::scheduling::activate::Activator>) -> Self {
        let mut h = std::ptr::null_mut();
        unsafe { sys::check(worker_ctx(), sys::mzgpu_spine_new(worker_ctx(), sys::ROW_R32, 1, &mut h)).expect("spine_new"); }
        GpuSpine { h, logical: Antichain::from_elem(0), physical: Antichain::from_elem(0), proportionality: 16 }
    }
    /// Trace::exert with Materialize's ExertionLogic (src/cluster/src/client.rs:227-254)
    /// evaluated inside the library.
    fn exert(&mut self) {
        let e = unsafe { sys::mzgpu_spine_exert_logic(self.h, self.proportionality) };
        if e != 0 { unsafe { sys::check(worker_ctx(), sys::mzgpu_spine_exert(self.h, e, std::ptr::null_mut())).expect("exert"); } }
    }
    fn set_exert_logic(&mut self, _logic: ExertionLogic) { /* the library carries the reference's rule; `proportionality` is its one parameter */ }
    fn insert(&mut self, batch: GpuBatch) {
        unsafe { sys::check(worker_ctx(), sys::mzgpu_spine_insert(self.h, batch.raw())).expect("spine_insert"); }
    }
    fn close(&mut self) {
        use differential_dataflow::trace::Batch;
        let u = unsafe { sys::mzgpu_spine_read_upper(self.h) };
        if u != sys::FRONTIER_EMPTY { self.insert(GpuBatch::empty(Antichain::from_elem(u), Antichain::new())); }
    }
}

impl Drop for GpuSpine {
    fn drop(&mut self) { unsafe { sys::mzgpu_spine_free(self.h) } }
}

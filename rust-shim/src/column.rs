//! `Column<C>` in and out of the device.  UNCOMPILED: see ../README.md.
//!
//! `Column` (src/timely-util/src/columnar.rs:54-222) is the container the reference's dataflow edges
//! carry; serialized (`Column::Bytes` / `Column::Align`) it is `columnar::bytes::indexed`.  The merge
//! batcher's input is `Column<((K, V), T, R)>` (`Col2ValBatcher`, columnar.rs:41-45): with this module a
//! GPU batcher takes those containers AS BYTES -- no `into_index_iter()` row loop on the host, the
//! transposition to packed rows is one kernel (materialize_b200/csrc/column.cu) -- and the read side
//! of an arrangement (`walk_cursor`, src/compute/src/render/context.rs:1299-1355) hands containers
//! back the same way.
use columnar::bytes::indexed;
use columnar::Borrow;
use mz_timely_util::columnar::Column;

use super::sys;
use super::worker_ctx;

/// `((u64, u64), u64, i64)` updates: the fixed-width subset (`mzgpu_r32`).
pub type U64Update = ((u64, u64), u64, i64);

/// A device row buffer (`mzgpu_buf`) owned by this handle.
pub struct DeviceRows { pub(crate) h: *mut sys::Buf }

impl DeviceRows {
    pub fn new() -> Self {
        let mut h = std::ptr::null_mut();
        unsafe { sys::check(worker_ctx(), sys::mzgpu_buf_new(worker_ctx(), sys::ROW_R32, &mut h)).expect("buf_new"); }
        DeviceRows { h }
    }

    /// `Column::borrow()` + drain, on the device.  `Typed` containers are serialized first (the same
    /// `indexed::encode` `ColumnBuilder` runs when it mints an `Align`, builder.rs:56-70); `Bytes` and
    /// `Align` go down as they are.
    pub fn extend_from_column(&mut self, column: &Column<U64Update>) {
        let owned;
        let words: &[u64] = match column {
            Column::Typed(t) => {
                let mut alloc = Vec::with_capacity(indexed::length_in_words(&t.borrow()));
                indexed::encode(&mut alloc, &t.borrow());
                owned = alloc;
                &owned
            }
            // `from_bytes` (columnar.rs:179-195) only keeps `Bytes` when the slice is u64 aligned
            Column::Bytes(b) => bytemuck::cast_slice(b),
            Column::Align(a) => a,
        };
        unsafe {
            sys::check(worker_ctx(),
                       sys::mzgpu_column_decode(worker_ctx(), sys::COLUMN_U64X4, words.as_ptr(), words.len() as u64,
                                                sys::MEM_HOST, self.h)).expect("column_decode");
        }
    }

    /// `ColumnBuilder` over the whole buffer (builder.rs:28-111): the `Column::Align` containers
    /// `push_into` would have minted for these rows, then the remainder `finish` hands out.
    pub fn into_columns(&self) -> Vec<Column<U64Update>> {
        let (mut n_words, mut n_chunks) = (0u64, 0u32);
        unsafe {
            // size query: MZGPU_E_CAPACITY reports the totals
            let _ = sys::mzgpu_column_build(self.h, sys::COLUMN_U64X4, std::ptr::null_mut(), 0, sys::MEM_HOST, &mut n_words,
                                            std::ptr::null_mut(), 0, &mut n_chunks);
        }
        let mut words = vec![0u64; n_words as usize];
        let mut sizes = vec![0u64; n_chunks as usize];
        unsafe {
            sys::check(worker_ctx(),
                       sys::mzgpu_column_build(self.h, sys::COLUMN_U64X4, words.as_mut_ptr(), n_words, sys::MEM_HOST,
                                               &mut n_words, sizes.as_mut_ptr(), n_chunks, &mut n_chunks)).expect("column_build");
        }
        let mut out = Vec::with_capacity(sizes.len());
        let mut at = 0usize;
        for s in sizes {
            out.push(Column::Align(words[at..at + s as usize].to_vec()));
            at += s as usize;
        }
        out
    }
}

impl Drop for DeviceRows {
    fn drop(&mut self) { unsafe { sys::mzgpu_buf_free(self.h) } }
}

/// `Batcher::push_container` for `Input = Column<U64Update>`: decode on the device, push the buffer.
pub fn push_column(batcher: *mut sys::Batcher, column: &mut Column<U64Update>) {
    let mut rows = DeviceRows::new();
    rows.extend_from_column(column);
    unsafe { sys::check(worker_ctx(), sys::mzgpu_batcher_push_buf(batcher, rows.h)).expect("batcher_push_buf"); }
    *column = Default::default();
}

/// `walk_cursor` over one batch, fuel rows at a time (context.rs:1299-1355): every call yields the next
/// container; `None` once the batch (or the seeked key's run) is exhausted.
pub struct ColumnWalk { batch: *mut sys::Batch, key: Option<u64>, next: u64, done: bool }

impl ColumnWalk {
    pub fn new(batch: *mut sys::Batch, key: Option<u64>) -> Self { ColumnWalk { batch, key, next: 0, done: false } }

    pub fn step(&mut self, fuel: usize) -> Option<Column<U64Update>> {
        if self.done { return None; }
        let cap = unsafe { sys::mzgpu_column_length_in_words(sys::COLUMN_U64X4, fuel as u64, 0, 0) };
        let mut words = vec![0u64; cap as usize];
        let (mut n_words, mut n_rows) = (0u64, 0u64);
        let key_ptr = self.key.as_ref().map_or(std::ptr::null(), |k| k as *const u64);
        unsafe {
            sys::check(worker_ctx(),
                       sys::mzgpu_batch_walk_column(self.batch, key_ptr, self.next, fuel as u64, sys::COLUMN_U64X4,
                                                    words.as_mut_ptr(), cap, sys::MEM_HOST, &mut n_words, &mut n_rows))
                .expect("batch_walk_column");
        }
        self.next += n_rows;
        self.done = (n_rows as usize) < fuel;
        if n_rows == 0 { return None; }
        words.truncate(n_words as usize);
        Some(Column::Align(words))
    }
}

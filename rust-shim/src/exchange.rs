//! The `Exchange(key.hashed())` pact between workers that own GPUs.  UNCOMPILED: see ../README.md.
//!
//! `mz_arrange` routes every update by `(update.0).0.hashed()` before it reaches the batcher
//! (src/compute/src/extensions/arrange.rs:116; columnar containers: src/timely-util/src/columnar.rs:227-237).
//! When the arrangement lives on the device the rows should not come back to the host to be
//! routed: the worker hands the device buffers of one *round* (the exchange points of operators
//! that run side by side) to the library, which partitions by `mzgpu_route(key, peers)` and
//! delivers over NVLink peer memory (update batches: one scatter + one gather kernel, no host
//! wait) or with a grouped NCCL all-to-all (bulk hydration chunks).  All workers call the rounds
//! in the same order — true for timely dataflows by construction.
use super::sys;
use super::worker_ctx;

/// How the workers of this process group are connected (include/mzgpu.h, "exchange").
pub enum Mesh {
    /// `mzgpu_comm_unique_id` on worker 0, the 128 bytes distributed over the cluster's own
    /// bootstrap channel (src/cluster/src/communication.rs), `mzgpu_comm_init` everywhere.
    Nccl,
    /// Landing zones exported with CUDA IPC and mapped by every peer (`mzgpu_comm_p2p_export` /
    /// `_import`): `landing_rows` = what one worker may send one peer per buffer and round.
    PeerMemory { landing_rows: u64 },
}

pub struct GpuExchange { peer_memory: bool }

impl GpuExchange {
    /// Worker 0 mints the NCCL id; `broadcast` is the cluster bootstrap (host side, as timely's own).
    pub fn connect_nccl(worker: usize, mut broadcast: impl FnMut(&mut [u8; sys::COMM_ID_BYTES])) -> Self {
        let mut id = [0u8; sys::COMM_ID_BYTES];
        unsafe {
            if worker == 0 { assert_eq!(sys::mzgpu_comm_unique_id(id.as_mut_ptr()), sys::OK); }
            broadcast(&mut id);
            sys::check(worker_ctx(), sys::mzgpu_comm_init(worker_ctx(), id.as_ptr())).expect("comm_init");
        }
        GpuExchange { peer_memory: false }
    }
    /// Export this worker's landing zone, all-gather the handles (`allgather`: host bootstrap),
    /// map every peer's zone.  Update-batch rounds go over peer memory afterwards.
    pub fn connect_peer_memory(&mut self, landing_rows: u64, peers: usize,
                               mut allgather: impl FnMut(&[u8; sys::P2P_HANDLE_BYTES]) -> Vec<u8>) {
        let mut h = [0u8; sys::P2P_HANDLE_BYTES];
        unsafe {
            sys::check(worker_ctx(), sys::mzgpu_comm_p2p_export(worker_ctx(), landing_rows, sys::ROW_R32, h.as_mut_ptr()))
                .expect("p2p_export");
            let all = allgather(&h);
            assert_eq!(all.len(), peers * sys::P2P_HANDLE_BYTES);
            sys::check(worker_ctx(), sys::mzgpu_comm_p2p_import(worker_ctx(), all.as_ptr())).expect("p2p_import");
        }
        self.peer_memory = true;
    }
    /// One round: `ins[i]` is partitioned by key, `outs[i]` receives this worker's share of every
    /// peer's `ins[i]`.  `recv_ub[i]`: what this worker can receive at most, if the dataflow knows
    /// (the global batch size for arrangement inputs); `None` = the landing capacity.
    pub fn round(&self, ins: &mut [*mut sys::Buf], outs: &mut [*mut sys::Buf], recv_ub: Option<&[u64]>, bulk: bool) {
        assert_eq!(ins.len(), outs.len());
        let k = ins.len() as u32;
        let st = unsafe {
            if self.peer_memory && !bulk {
                sys::mzgpu_exchange_p2p(worker_ctx(), k, ins.as_mut_ptr(), outs.as_mut_ptr(),
                                        recv_ub.map_or(std::ptr::null(), |u| u.as_ptr()))
            } else {
                sys::mzgpu_exchange_many(worker_ctx(), k, ins.as_mut_ptr(), outs.as_mut_ptr())
            }
        };
        unsafe { sys::check(worker_ctx(), st).expect("exchange round"); }
    }
    /// The routing function itself (FNV-1a of the key word, modulo the peers): what a host-side
    /// `Exchange` pact must use for rows that take the Rust path into the same arrangement.
    pub fn route(key: u64, peers: u32) -> u32 { unsafe { sys::mzgpu_route(key, peers) } }
}

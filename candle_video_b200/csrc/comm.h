// Peer-memory communicator for the multi-GPU paths (one process per GPU, NVLink 5 / NVSwitch P2P).
//
// No NCCL on the data path: every rank owns one "symmetric heap" allocation whose CUDA IPC handle is exchanged once
// (through the host harness, e.g. torch.distributed all_gather of 64 bytes); buffers are carved out of it in the same
// order on every rank, so a peer's copy of a buffer is `peer_base + offset`.  Producer kernels then store straight into
// the consumer rank's buffer (Ulysses head scatter from the q/k-norm+RoPE kernel, attention epilogue writing each
// query block to the rank that owns those tokens, VAE halo rows written by the pixel-norm kernel) and a tiny
// signal/wait kernel on system-scope flags orders producer and consumer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "model_common.h"

namespace ltxv {

constexpr int kMaxRanks = 8;

class PeerComm {
public:
    PeerComm(int nranks, int rank, int device, size_t heap_bytes);
    ~PeerComm();
    int nranks() const { return nranks_; }
    int rank() const { return rank_; }
    void get_handle(void* out64) const;           // cudaIpcMemHandle_t of the local heap
    void open_peers(const void* all_handles);     // nranks * 64 bytes, indexed by rank
    bool ready() const { return opened_; }

    // bump allocator over the symmetric heap (256-byte aligned); identical call sequences on all ranks
    size_t alloc(size_t bytes);
    void reset_allocator() { top_ = kReserved; }
    void* local(size_t off) const { return static_cast<char*>(heap_) + off; }
    void* peer(int r, size_t off) const { return static_cast<char*>(peer_base_[r]) + off; }

    // barrier across ranks [first, first+count) on `s`: everything enqueued before it on those ranks (including
    // stores into peer memory) is visible to everything enqueued after it on those ranks.  `domain` selects an
    // independent flag set / epoch counter: 0 = all ranks, 1 = sub-groups (every rank must use a domain with the
    // same sequence of partners).
    void barrier(cudaStream_t s, int domain = 0, int first = 0, int count = -1);
    uint64_t launches() const { return launches_; }
    uint64_t id() const { return id_; }  // unique per communicator instance (cache key for heap carve-outs)

private:
    static constexpr size_t kReserved = 4096;  // flags live at the start of the heap
    int nranks_, rank_, device_;
    void* heap_ = nullptr;
    size_t heap_bytes_ = 0, top_ = kReserved;
    void* peer_base_[kMaxRanks] = {nullptr};
    bool opened_ = false;
    uint32_t epoch_[2][kMaxRanks] = {{0}, {0}};  // [domain][partner]
    uint64_t launches_ = 0;
    uint64_t id_ = 0;
};

uint64_t comm_launch_count();

}  // namespace ltxv

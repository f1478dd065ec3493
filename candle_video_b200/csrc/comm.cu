#include "comm.h"

#include "profile.h"

#include <atomic>
#include <string.h>

namespace ltxv {

namespace {
std::atomic<uint64_t> g_comm_launches{0};
std::atomic<uint64_t> g_comm_ids{0};

__device__ __forceinline__ unsigned long long globaltimer_ns_host_safe() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct PeerFlags {
    uint32_t* p[kMaxRanks];
    uint32_t epoch[kMaxRanks];  // per-partner epoch: ranks a and b count only the barriers that include both of them
};

// One block, one thread per peer.  Thread r publishes `epoch` into rank r's flag slot for this rank, then waits
// until rank r has published `epoch` (or later) into ours.  System-scope release/acquire orders the peer stores
// issued by earlier kernels of this stream before the flag, and the flag before later kernels' loads.
__global__ void barrier_kernel(PeerFlags flags, uint32_t* local_flags, int first, int count, int rank) {
    const int r = first + threadIdx.x;
    if (static_cast<int>(threadIdx.x) >= count) return;
    __threadfence_system();
    if (r != rank) {
        const uint32_t epoch = flags.epoch[r];
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags.p[r] + rank), "r"(epoch) : "memory");
        const unsigned long long t0 = globaltimer_ns_host_safe();
        uint32_t v;
        unsigned spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(local_flags + r) : "memory");
            if (static_cast<int32_t>(v - epoch) >= 0) break;
            if ((++spins & 0xfff) == 0 && globaltimer_ns_host_safe() - t0 > 8000000000ull) __trap();  // 8 s: a peer died
        } while (true);
    }
    __threadfence_system();
}
}  // namespace

uint64_t comm_launch_count() { return g_comm_launches.load(); }

PeerComm::PeerComm(int nranks, int rank, int device, size_t heap_bytes)
    : nranks_(nranks), rank_(rank), device_(device), heap_bytes_(heap_bytes) {
    if (nranks < 1 || nranks > kMaxRanks) fail("communicator size %d out of range (1..%d)", nranks, kMaxRanks);
    if (rank < 0 || rank >= nranks) fail("rank %d out of range for %d ranks", rank, nranks);
    require_cuda_device(device);
    id_ = g_comm_ids.fetch_add(1) + 1;
    if (heap_bytes < kReserved * 2) fail("symmetric heap too small");
    LTXV_CUDA(cudaMalloc(&heap_, heap_bytes));
    LTXV_CUDA(cudaMemset(heap_, 0, kReserved));
    LTXV_CUDA(cudaDeviceSynchronize());
    peer_base_[rank] = heap_;
    if (nranks == 1) opened_ = true;
}

PeerComm::~PeerComm() {
    for (int r = 0; r < nranks_; ++r)
        if (r != rank_ && peer_base_[r] != nullptr) cudaIpcCloseMemHandle(peer_base_[r]);
    if (heap_) cudaFree(heap_);
}

void PeerComm::get_handle(void* out64) const {
    cudaIpcMemHandle_t h;
    LTXV_CUDA(cudaIpcGetMemHandle(&h, heap_));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(out64, &h, 64);
}

void PeerComm::open_peers(const void* all_handles) {
    LTXV_CUDA(cudaSetDevice(device_));
    for (int r = 0; r < nranks_; ++r) {
        if (r == rank_) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(all_handles) + 64 * r, 64);
        LTXV_CUDA(cudaIpcOpenMemHandle(&peer_base_[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    opened_ = true;
}

size_t PeerComm::alloc(size_t bytes) {
    const size_t off = (top_ + 255) & ~size_t(255);
    if (off + bytes > heap_bytes_)
        fail("symmetric heap exhausted: need %zu bytes at offset %zu of %zu", bytes, off, heap_bytes_);
    top_ = off + bytes;
    return off;
}

void PeerComm::barrier(cudaStream_t s, int domain, int first, int count) {
    if (count < 0) count = nranks_;
    if (count <= 1) return;
    if (domain < 0 || domain > 1 || first < 0 || first + count > nranks_ || rank_ < first || rank_ >= first + count)
        fail("invalid barrier group [%d,%d) / domain %d for rank %d", first, first + count, domain, rank_);
    if (!opened_) fail("communicator peers have not been opened");
    PeerFlags f;
    // flag word for (domain, source rank) lives at heap offset (domain * kMaxRanks + src) * 4 on every rank.  The
    // epoch is kept per partner: two ranks may have taken part in different numbers of sub-group barriers with
    // third parties (e.g. the CFG branch groups run different numbers of forwards), but the sequence of barriers that
    // contain BOTH of them is the same on both sides.
    for (int r = 0; r < kMaxRanks; ++r) {
        f.p[r] = r < nranks_ ? static_cast<uint32_t*>(peer_base_[r]) + domain * kMaxRanks : nullptr;
        f.epoch[r] = 0;
    }
    for (int r = first; r < first + count; ++r)
        if (r != rank_) f.epoch[r] = ++epoch_[domain][r];
    {
        // class 7 of ltxv_profile_*: the barrier kernel's duration = flag round trip over NVLink + waiting for the slowest
        // partner (load imbalance); "work" counts the partners
        ProfScope prof(PROF_OTHER, static_cast<double>(count - 1), s);
        barrier_kernel<<<1, 32, 0, s>>>(f, static_cast<uint32_t*>(heap_) + domain * kMaxRanks, first, count, rank_);
    }
    ++launches_;
    g_comm_launches.fetch_add(1, std::memory_order_relaxed);
    LTXV_CUDA(cudaGetLastError());
}

}  // namespace ltxv

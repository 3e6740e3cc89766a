// p2p.cuh — exchanges over peer memory (CUDA IPC mappings, NVLink 5 / NVSwitch) that run INSIDE
// the compute kernels, one process per GPU:
//
//  * scalar all-reduce: the reduction finaliser of the kernel that produced the local value does
//    the exchange itself, so a dot product + all-reduce is ONE launch (the reference: device
//    reduction, exec->synchronize(), MPI_Allreduce on the host, core/distributed/vector.cpp:317-407);
//  * halo exchange of the distributed SpMV: the first CTAs of the SpMV kernel store the entries
//    the neighbours need straight into the neighbours' receive windows and publish an epoch flag;
//    the CTAs that own boundary rows run last and wait for the neighbours' flags before they add
//    the non-local entries (the reference: row_gather, exec->synchronize(), host staging,
//    MPI_Ialltoallv, a second SpMV launch — core/distributed/matrix.cpp:263-335).
//
// Every wait is bounded by a wall-clock timeout (globaltimer); a rank that gives up raises a
// sticky error word that the host reads at every poll.
//
// Scalar block (4 KB per rank), written in the style of NCCL's LL protocol — every 8-byte word
// carries 4 bytes of payload and the 32-bit epoch tag, so a word is either old or complete and
// NO fence / flag round trip is needed (one NVLink one-way latency per all-reduce instead of a
// store, a system fence round trip and a flag store):
//   slots  u64[2][kP2pMaxRanks][8]   value c of rank r = words 2c (low half | tag << 32) and
//                                    2c+1 (high half | tag << 32), written BY r INTO everybody's block
//   epoch  u64                       number of all-reduces done (local)
//   error  int                       set when a peer did not show up in time
// Double-buffered by epoch parity: a rank can be at most one all-reduce ahead of a peer.
#pragma once
#include <cstdint>

#include <cuda_runtime.h>

namespace gkob200 {

constexpr int kP2pMaxRanks = 16;
constexpr size_t kP2pSlotsOff = 0, kP2pEpochOff = 2 * kP2pMaxRanks * 8 * sizeof(unsigned long long),
                 kP2pErrorOff = kP2pEpochOff + 8, kP2pBlockBytes = 4096;
static_assert(kP2pErrorOff + 4 <= kP2pBlockBytes, "scalar block layout");

struct P2pDev {
    int rank, size;
    unsigned long long timeout_ns;        // per wait; GKOB200_P2P_TIMEOUT_MS (default 30 s)
    unsigned char* block[kP2pMaxRanks];   // block[rank] is local, the others are IPC mappings
};

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Spin until *flag >= target (system-scope acquire).  false: timed out.
__device__ __forceinline__ bool wait_flag_ge(const unsigned long long* flag, unsigned long long target,
                                             unsigned long long timeout_ns)
{
    if (ld_acquire_sys(flag) >= target) return true;
    const unsigned long long t0 = global_timer_ns();
    unsigned spins = 0;
    while (ld_acquire_sys(flag) < target) {
        if ((++spins & 0x3ffu) == 0 && global_timer_ns() - t0 > timeout_ns) return false;
    }
    return true;
}

__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// ONE thread: store my `count` (<= 4) values, tagged with the epoch, into every rank's block;
// wait until every rank's tagged words for this epoch stand in my own block; sum in rank order
// (the same order on every rank: identical bits everywhere, run-to-run reproducible).
// Returns false (and raises the block's error word) when a peer did not show up in time; the
// caller must then stop the solve on this rank (the values in `buf` are a partial sum).
template <typename V>
__device__ __forceinline__ bool peer_allreduce(const P2pDev& pr, V* buf, int count)
{
    unsigned char* mine = pr.block[pr.rank];
    unsigned long long* epoch = reinterpret_cast<unsigned long long*>(mine + kP2pEpochOff);
    const unsigned long long e = *epoch + 1;
    const int parity = static_cast<int>(e & 1);
    const unsigned long long tag = (e & 0xffffffffull) << 32;
    unsigned long long w[8];
    for (int c = 0; c < 4; ++c) {
        const unsigned long long bits =
            c < count ? static_cast<unsigned long long>(__double_as_longlong(static_cast<double>(buf[c]))) : 0ull;
        w[2 * c] = (bits & 0xffffffffull) | tag;
        w[2 * c + 1] = (bits >> 32) | tag;
    }
    for (int r = 0; r < pr.size; ++r) {
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(pr.block[r] + kP2pSlotsOff) +
                                  (parity * kP2pMaxRanks + pr.rank) * 8;
        for (int i = 0; i < 2 * count; ++i) st_relaxed_sys(dst + i, w[i]);
    }
    double tot[4] = {0.0, 0.0, 0.0, 0.0};
    bool ok = true;
    const unsigned long long t0 = global_timer_ns();
    for (int r = 0; r < pr.size; ++r) {
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(mine + kP2pSlotsOff) +
                                        (parity * kP2pMaxRanks + r) * 8;
        for (int c = 0; c < count; ++c) {
            unsigned long long lo = ld_relaxed_sys(src + 2 * c), hi = ld_relaxed_sys(src + 2 * c + 1);
            unsigned spins = 0;
            while (ok && ((lo ^ tag) >> 32 || (hi ^ tag) >> 32)) {
                if ((++spins & 0x3ffu) == 0 && global_timer_ns() - t0 > pr.timeout_ns) {
                    ok = false;
                    *reinterpret_cast<volatile int*>(mine + kP2pErrorOff) = 1;
                    break;
                }
                lo = ld_relaxed_sys(src + 2 * c);
                hi = ld_relaxed_sys(src + 2 * c + 1);
            }
            tot[c] += __longlong_as_double(static_cast<long long>((lo & 0xffffffffull) | (hi << 32)));
        }
    }
    for (int c = 0; c < count; ++c) buf[c] = static_cast<V>(tot[c]);
    *epoch = e;
    return ok;
}

// ------------------------------- halo window -----------------------------------
// One window per distributed matrix and rank (cudaMalloc'ed, mapped by every peer):
//   [0,128)      started[r]  u64  latest apply (epoch) rank r has entered   (written by r)
//   [128,256)    arrived[r]  u64  latest epoch whose entries from r are complete here (written by r)
//   [256,264)    epoch       u64  number of applies entered on this rank (bumped by the kernel
//                                 that precedes the SpMV: halo_epoch_bump / dist_cg_direction)
//   [264,268)    push ticket u32
//   [268,272)    error       int
//   [1024, ...)  receive buffer 0, receive buffer 1 (epoch parity), recv_stride values each
// Flow control: entries of epoch e go to buffer e&1; a sender waits until the receiver has
// ENTERED epoch e-1 (hence finished reading epoch e-2, stream order) before it overwrites.
constexpr size_t kHaloStartedOff = 0, kHaloArrivedOff = 128, kHaloEpochOff = 256, kHaloTicketOff = 264,
                 kHaloErrorOff = 268, kHaloDataOff = 1024;

struct HaloDev {
    int rank, size;
    unsigned long long timeout_ns;
    unsigned char* window;                    // local window
    int n_push_ctas;                          // CTAs at the front of the SpMV grid that push
    int n_interior;                           // row-block slots without non-local rows (run first)
    // ---- push side ----
    long long send_total;
    const int* gather;                        // local row of send entry t
    int n_send_peers;
    int send_peer[kP2pMaxRanks];
    long long send_begin[kP2pMaxRanks + 1];   // entries [send_begin[i], send_begin[i+1]) go to send_peer[i]
    unsigned char* dst_data[kP2pMaxRanks];    // peer window + kHaloDataOff + (my offset in its buffer) * sizeof(V)
    long long dst_stride[kP2pMaxRanks];       // bytes between the peer's two receive buffers
    unsigned long long* dst_arrived[kP2pMaxRanks];   // &peer.arrived[rank]
    // ---- neighbours in either direction: told when this rank enters an epoch ----
    int n_neighbours;
    unsigned long long* nb_started[kP2pMaxRanks];    // &peer.started[rank]
    // ---- receive side ----
    int n_recv_peers;
    int recv_peer[kP2pMaxRanks];
    long long recv_stride;                    // values between my two receive buffers
    // ---- non-local block (row-compressed) and the CTA order ----
    const int* order;                         // slot -> row block; boundary blocks last
    const int2* nl_thread_range;              // boundary slot j, thread t: entry range [x, y) of row
                                              // (block*128 + t) in the non-local block  [128 per slot]
    const int* nl_row_ptrs;
    const int* nl_cols;
    const void* nl_vals;
};

}  // namespace gkob200

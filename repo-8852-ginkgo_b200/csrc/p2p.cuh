// p2p.cuh — scalar all-reduce over peer memory (CUDA IPC mappings of one 4 KB block per rank,
// NVLink / NVSwitch), callable from inside a kernel: the reduction finaliser of the kernel that
// produced the local value does the exchange itself, so a dot product + all-reduce is ONE launch
// (the reference: device reduction, exec->synchronize(), MPI_Allreduce on the host,
// core/distributed/vector.cpp:317-407).
//   slots  double[2][kP2pMaxRanks][4]   values written BY rank r INTO everybody's block
//   flags  u64   [2][kP2pMaxRanks]      epoch of the last complete write of rank r
//   epoch  u64                          number of all-reduces done (local)
//   error  int                          set when a peer did not show up (bounded spin)
// Double-buffered by epoch parity: a rank can be at most one all-reduce ahead of a peer.
#pragma once
#include <cstdint>

namespace gkob200 {

constexpr int kP2pMaxRanks = 16;
constexpr size_t kP2pSlotsOff = 0, kP2pFlagsOff = 2 * kP2pMaxRanks * 4 * sizeof(double),
                 kP2pEpochOff = kP2pFlagsOff + 2 * kP2pMaxRanks * sizeof(unsigned long long),
                 kP2pErrorOff = kP2pEpochOff + 8, kP2pBlockBytes = 4096;

struct P2pDev {
    int rank, size;
    unsigned char* block[kP2pMaxRanks];   // block[rank] is local, the others are IPC mappings
};

// ONE thread: push my `count` (<= 4) values into every rank's block, publish the epoch with a
// system-scope release, wait for every rank's epoch in my own block, sum in rank order (the
// same order on every rank: identical bits everywhere, run-to-run reproducible).
template <typename V>
__device__ __forceinline__ void peer_allreduce(const P2pDev& pr, V* buf, int count)
{
    unsigned char* mine = pr.block[pr.rank];
    unsigned long long* epoch = reinterpret_cast<unsigned long long*>(mine + kP2pEpochOff);
    const unsigned long long e = *epoch + 1;
    const int parity = static_cast<int>(e & 1);
    double v[4];
    for (int c = 0; c < 4; ++c) v[c] = c < count ? static_cast<double>(buf[c]) : 0.0;
    for (int r = 0; r < pr.size; ++r) {
        volatile double* dst = reinterpret_cast<volatile double*>(pr.block[r] + kP2pSlotsOff) +
                               (parity * kP2pMaxRanks + pr.rank) * 4;
        for (int c = 0; c < count; ++c) dst[c] = v[c];
    }
    // ONE system-scope fence orders all the value stores before all the flag stores (a release
    // store per flag would wait for an NVLink round trip eight times: 25 us instead of 5)
    __threadfence_system();
    for (int r = 0; r < pr.size; ++r) {
        unsigned long long* f = reinterpret_cast<unsigned long long*>(pr.block[r] + kP2pFlagsOff) +
                                parity * kP2pMaxRanks + pr.rank;
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(f), "l"(e) : "memory");
    }
    double tot[4] = {0.0, 0.0, 0.0, 0.0};
    for (int r = 0; r < pr.size; ++r) {
        const unsigned long long* f = reinterpret_cast<const unsigned long long*>(mine + kP2pFlagsOff) +
                                      parity * kP2pMaxRanks + r;
        unsigned long long seen = 0;
        long long spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(f) : "memory");
        } while (seen < e && ++spins < (1ll << 24));
        if (seen < e) *reinterpret_cast<volatile int*>(mine + kP2pErrorOff) = 1;
        const volatile double* src = reinterpret_cast<const volatile double*>(mine + kP2pSlotsOff) +
                                     (parity * kP2pMaxRanks + r) * 4;
        for (int c = 0; c < count; ++c) tot[c] += src[c];
    }
    for (int c = 0; c < count; ++c) buf[c] = static_cast<V>(tot[c]);
    *epoch = e;
}

}  // namespace gkob200

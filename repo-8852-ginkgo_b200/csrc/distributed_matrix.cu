// distributed_matrix.cu — row-partitioned distributed::Matrix apply and distributed CG
// over the GPUs of one NVLink/NVSwitch box, one process per GPU.
//
// Replaces experimental::distributed::Matrix::{communicate, apply_impl}
// (reference core/distributed/matrix.cpp:263-369) and the all_reduce steps of
// experimental::distributed::Vector (core/distributed/vector.cpp:317-407).  The reference
// packs with row_gather, calls exec->synchronize() (a device-wide sync), stages through
// HOST buffers unless MPI is GPU-aware, posts MPI_Ialltoallv and launches a second SpMV.
//
// Fused path (every rank on its own GPU, peers mappable, one right-hand side, local block on
// the CSR row-block kernel): ONE launch per apply.  The first CTAs of the SpMV grid store the
// entries the neighbours need straight into the neighbours' receive windows over peer memory
// and publish an epoch flag; the row blocks that own non-local entries are scheduled last,
// wait for the neighbours' flags and continue their row sums with the non-local entries
// (csr_spmv.cu, p2p.cuh).  No pack kernel, no NCCL call, no second launch, no events; the
// distributed CG iteration is four launches, captured in a CUDA graph.
//
// Fallback path (GKOB200_P2P=0, ranks sharing a GPU, many right-hand sides, other formats):
//   * the pack kernel runs on the compute stream, an event hands the send buffer to a
//     dedicated communication stream, the halo moves GPU->GPU as grouped ncclSend/ncclRecv,
//   * the local SpMV overlaps the exchange on the compute stream,
//   * a second event gates the non-local SpMV  x += A_nl * ghost,
//   * dot products are reduced with ncclAllReduce (or the peer-memory all-reduce).
// No device-wide synchronisation anywhere on either path.
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "p2p.cuh"
#include "solver_common.cuh"

#define GKOB200_NCCL(call)                                        \
    do {                                                          \
        ncclResult_t r__ = (call);                                \
        if (r__ != ncclSuccess) return 1000 + static_cast<int>(r__); \
    } while (0)

using gkob200::HaloDev;
using gkob200::P2pDev;
using gkob200::kP2pMaxRanks;
using gkob200::kP2pBlockBytes;
using gkob200::kP2pErrorOff;

struct gkob200_dist_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, size = 1;
    cudaStream_t comm_stream = nullptr;
    bool p2p = false;
    unsigned long long timeout_ns = 30ull * 1000 * 1000 * 1000;
    P2pDev p2p_dev{};
    P2pDev* p2p_dev_ptr = nullptr;   // device copy (for kernels that get it through a pointer)
};

struct gkob200_dist_matrix {
    gkob200_dist_comm* comm = nullptr;
    gkob200_matrix local{}, non_local{};
    const int32_t* gather_idxs = nullptr;  // device, send_total entries (local row indices)
    std::vector<int64_t> send_sizes, send_offsets, recv_sizes, recv_offsets;
    int64_t send_total = 0, recv_total = 0;
    gkob200::DevBuf send_buf, recv_buf, consts;
    int64_t buf_nrhs = 0;
    cudaEvent_t packed = nullptr, received = nullptr;
    int64_t launches = 0;
    bool exchanged = false;   // the last apply all-reduced its fused dot itself
    // ---- fused halo (peer-memory window) ----
    bool fused = false;
    unsigned char* window = nullptr;                  // local window (cudaMalloc)
    unsigned char* peer_window[kP2pMaxRanks] = {};    // IPC mappings (peer_window[rank] == window)
    gkob200::DevBuf halo_dev, order, nl_thread_range;
    int n_push = 0, n_interior = 0, n_runs = 0;
    int run_slot[gkob200::kHaloRuns + 1] = {}, run_block[gkob200::kHaloRuns] = {};
};

namespace gkob200 {
namespace {

template <typename V>
ncclDataType_t nccl_type();
template <>
ncclDataType_t nccl_type<double>() { return ncclDouble; }
template <>
ncclDataType_t nccl_type<float>() { return ncclFloat; }

// stand-alone all-reduce (ranks whose producing kernel cannot do the exchange itself)
template <typename V>
__global__ void p2p_allreduce_kernel(P2pDev pr, V* buf, int count, const int* skip, int* on_fail)
{
    // `skip`: the solver's stopped flag — identical on every rank (it derives from all-reduced
    // values), and the fused exchanges inside the solver kernels honour it too
    if (skip && *skip) return;
    if (!peer_allreduce(pr, buf, count) && on_fail) *on_fail = 1;
}

// the kernel in front of a fused-halo SpMV enters the next epoch (p2p.cuh)
__global__ void halo_epoch_bump(unsigned char* window, const int* skip)
{
    if (skip && *skip) return;
    ++*reinterpret_cast<unsigned long long*>(window + kHaloEpochOff);
}

// Collective over the communicator: every rank allocates `bytes` of zeroed device memory and
// maps the allocations of all the others (CUDA IPC).  Each rank contributes `extra` (<= 160
// bytes) that every rank receives in `extras` (size * 160 bytes).  *ok = 1 only if EVERY rank
// could map every peer, all ranks sit on distinct GPUs (two spinning ranks time-slicing one GPU
// would wait for each other) and every rank passed want != 0; otherwise nothing stays mapped.
constexpr size_t kExtraBytes = 160;
int p2p_shared_alloc(gkob200_dist_comm* c, size_t bytes, int want, const void* extra, size_t extra_bytes,
                     unsigned char** ptrs, std::vector<unsigned char>* extras, int* ok_out)
{
    struct Record {
        cudaIpcMemHandle_t handle;
        char uuid[16];
        int want;
        char pad[12];
        unsigned char extra[kExtraBytes];
    };
    static_assert(sizeof(cudaIpcMemHandle_t) == 64 && sizeof(Record) == 256, "record size");
    *ok_out = 0;
    if (extra_bytes > kExtraBytes) return GKOB200_EINVAL;
    unsigned char* local = nullptr;
    GKOB200_CUDA(cudaMalloc(&local, bytes));
    GKOB200_CUDA(cudaMemset(local, 0, bytes));
    Record mine{};
    int dev = 0;
    GKOB200_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    GKOB200_CUDA(cudaGetDeviceProperties(&prop, dev));
    memcpy(mine.uuid, &prop.uuid, 16);
    mine.want = want;
    if (extra && extra_bytes) memcpy(mine.extra, extra, extra_bytes);
    int ok = cudaIpcGetMemHandle(&mine.handle, local) == cudaSuccess ? 1 : 0;
    cudaGetLastError();
    unsigned char* d_rec = nullptr;
    GKOB200_CUDA(cudaMalloc(&d_rec, sizeof(Record) * (c->size + 1)));
    GKOB200_CUDA(cudaMemcpy(d_rec, &mine, sizeof(Record), cudaMemcpyHostToDevice));
    GKOB200_NCCL(ncclAllGather(d_rec, d_rec + sizeof(Record), sizeof(Record), ncclChar, c->comm, c->comm_stream));
    GKOB200_CUDA(cudaStreamSynchronize(c->comm_stream));
    std::vector<Record> all(c->size);
    GKOB200_CUDA(cudaMemcpy(all.data(), d_rec + sizeof(Record), sizeof(Record) * c->size, cudaMemcpyDeviceToHost));
    for (int r = 0; r < c->size; ++r) ptrs[r] = nullptr;
    for (int r = 0; r < c->size; ++r) {
        if (!all[r].want) ok = 0;
        for (int q = 0; q < r; ++q)
            if (memcmp(all[r].uuid, all[q].uuid, 16) == 0) ok = 0;   // two ranks on one GPU
    }
    for (int r = 0; r < c->size; ++r) {
        if (r == c->rank) {
            ptrs[r] = local;
        } else if (ok) {
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                ok = 0;
            }
            ptrs[r] = static_cast<unsigned char*>(ptr);
        }
    }
    // all or nobody
    int* d_ok = reinterpret_cast<int*>(d_rec);
    GKOB200_CUDA(cudaMemcpy(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice));
    GKOB200_NCCL(ncclAllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, c->comm, c->comm_stream));
    GKOB200_CUDA(cudaStreamSynchronize(c->comm_stream));
    GKOB200_CUDA(cudaMemcpy(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(d_rec);
    if (extras) {
        extras->resize(static_cast<size_t>(c->size) * kExtraBytes);
        for (int r = 0; r < c->size; ++r) memcpy(extras->data() + r * kExtraBytes, all[r].extra, kExtraBytes);
    }
    if (!ok) {
        for (int r = 0; r < c->size; ++r) {
            if (r != c->rank && ptrs[r]) cudaIpcCloseMemHandle(ptrs[r]);
            ptrs[r] = nullptr;
        }
        cudaGetLastError();
        cudaFree(local);
    }
    *ok_out = ok;
    return 0;
}

void p2p_shared_free(gkob200_dist_comm* c, unsigned char** ptrs)
{
    for (int r = 0; r < c->size; ++r) {
        if (!ptrs[r]) continue;
        if (r == c->rank)
            cudaFree(ptrs[r]);
        else
            cudaIpcCloseMemHandle(ptrs[r]);
        ptrs[r] = nullptr;
    }
    cudaGetLastError();
}

// Maps one 4 KB scalar block per rank (p2p.cuh) for the in-kernel all-reduces.
int p2p_setup(gkob200_dist_comm* c)
{
    const char* env = getenv("GKOB200_P2P");
    if (const char* t = getenv("GKOB200_P2P_TIMEOUT_MS")) {
        const long long ms = atoll(t);
        if (ms > 0) c->timeout_ns = static_cast<unsigned long long>(ms) * 1000000ull;
    }
    const int want = !((env && env[0] == '0') || c->size > kP2pMaxRanks);
    int ok = 0;
    unsigned char* ptrs[kP2pMaxRanks] = {};
    if (c->size > kP2pMaxRanks) return 0;   // (cannot even take part in the exchange)
    int rc = p2p_shared_alloc(c, kP2pBlockBytes, want, nullptr, 0, ptrs, nullptr, &ok);
    if (rc) return rc;
    c->p2p = ok != 0;
    if (c->p2p) {
        c->p2p_dev.rank = c->rank;
        c->p2p_dev.size = c->size;
        c->p2p_dev.timeout_ns = c->timeout_ns;
        for (int r = 0; r < c->size; ++r) c->p2p_dev.block[r] = ptrs[r];
        GKOB200_CUDA(cudaMalloc(&c->p2p_dev_ptr, sizeof(P2pDev)));
        GKOB200_CUDA(cudaMemcpy(c->p2p_dev_ptr, &c->p2p_dev, sizeof(P2pDev), cudaMemcpyHostToDevice));
    }
    return 0;
}

// in-place sum of `count` scalars over all ranks, on stream s
template <typename V>
int comm_allreduce(gkob200_dist_comm* c, cudaStream_t s, V* buf, size_t count, const int* skip = nullptr,
                   int* on_fail = nullptr)
{
    if (!c || c->size == 1 || count == 0) return 0;
    if (c->p2p && count <= 4) {
        p2p_allreduce_kernel<V><<<1, 1, 0, s>>>(c->p2p_dev, buf, static_cast<int>(count), skip, on_fail);
        GKOB200_CHECK_LAUNCH();
        return 0;
    }
    GKOB200_NCCL(ncclAllReduce(buf, buf, count, nccl_type<V>(), ncclSum, c->comm, s));
    return 0;
}

// Measurement-only switches (tools/dist_ab.py; results are WRONG with any bit set):
//   bit 0: do not enter a new halo epoch (no rank ever waits for a neighbour)
//   bit 1: skip the scalar all-reduces of the distributed CG
//   bit 2: the push CTAs of the fused SpMV do nothing       bit 3: no non-local tail
inline int dist_debug()
{
    const char* e = getenv("GKOB200_DIST_DEBUG");
    return e ? atoi(e) : 0;
}

// ---- fused halo: plan ------------------------------------------------------------------
// true when matrix_apply runs this descriptor on the bulk-async CSR row-block kernel with
// 32-bit indices (the only kernel that carries the halo exchange)
bool local_block_takes_halo(const gkob200_matrix& A)
{
    if (A.format != GKOB200_FMT_CSR || A.index_type != GKOB200_I32) return false;
    int strategy = A.csr_strategy;
    if (strategy == GKOB200_CSR_AUTO)
        strategy = A.csr_max_block_nnz > 0 ? gkob200_csr_pick_strategy(A.n_rows, A.nnz, -1, A.csr_max_block_nnz)
                                           : GKOB200_CSR_MERGE_PATH;
    if (strategy != GKOB200_CSR_CLASSICAL) return false;
    if (reinterpret_cast<uintptr_t>(A.values) % 16 || reinterpret_cast<uintptr_t>(A.col_idxs) % 16) return false;
    if (const char* e = getenv("GKOB200_CSR_ROWBLOCK"))
        if (e[0] == 'p') return false;
    return A.n_rows > 0;
}

// Collective (called by every rank from gkob200_dist_matrix_create).  Allocates and maps the
// receive windows, exchanges the offsets at which every rank's entries land in its peers'
// buffers, builds the CTA order (row blocks with non-local rows last) and uploads the plan.
int halo_setup(gkob200_dist_matrix* m)
{
    gkob200_dist_comm* c = m->comm;
    if (!c || c->size == 1 || !c->p2p) return 0;
    const char* env = getenv("GKOB200_FUSED_HALO");
    const size_t vbytes = m->local.value_type == GKOB200_F64 ? 8 : 4;
    const bool nl_ok = m->non_local.nnz == 0 ||
                       (m->non_local.format == GKOB200_FMT_CSR_ROWS && m->non_local.index_type == GKOB200_I32 &&
                        m->non_local.value_type == m->local.value_type);
    const int want = !(env && env[0] == '0') && local_block_takes_halo(m->local) && nl_ok &&
                     m->recv_total < (int64_t(1) << 31) && m->send_total < (int64_t(1) << 31);
    const int64_t recv_stride = (m->recv_total + 31) / 32 * 32 + 32;
    struct Extra {
        int64_t recv_offsets[kP2pMaxRanks];
        int64_t recv_stride;
        int64_t value_bytes;
    } mine{};
    static_assert(sizeof(Extra) <= kExtraBytes, "extra");
    for (int p = 0; p < c->size; ++p) mine.recv_offsets[p] = m->recv_offsets[p];
    mine.recv_stride = recv_stride;
    mine.value_bytes = static_cast<int64_t>(vbytes);
    std::vector<unsigned char> extras;
    int ok = 0;
    const size_t win_bytes = kHaloDataOff + 2 * static_cast<size_t>(recv_stride) * vbytes;
    int rc = p2p_shared_alloc(c, win_bytes, want, &mine, sizeof(mine), m->peer_window, &extras, &ok);
    if (rc) return rc;
    if (!ok) return 0;
    m->window = m->peer_window[c->rank];
    auto extra_of = [&](int r) { return reinterpret_cast<const Extra*>(extras.data() + r * kExtraBytes); };

    HaloDev H{};
    H.rank = c->rank;
    H.size = c->size;
    H.timeout_ns = c->timeout_ns;
    H.window = m->window;
    H.send_total = m->send_total;
    H.gather = m->gather_idxs;
    H.recv_stride = recv_stride;
    bool neighbour[kP2pMaxRanks] = {};
    for (int p = 0; p < c->size; ++p) {
        if (p == c->rank) continue;
        if (m->send_sizes[p] > 0) {
            const int i = H.n_send_peers++;
            H.send_peer[i] = p;
            H.send_begin[i] = m->send_offsets[p];
            H.send_begin[i + 1] = m->send_offsets[p + 1];
            const Extra* ex = extra_of(p);
            H.dst_data[i] = m->peer_window[p] + kHaloDataOff + static_cast<size_t>(ex->recv_offsets[c->rank]) * vbytes;
            H.dst_stride[i] = ex->recv_stride * static_cast<int64_t>(vbytes);
            H.dst_arrived[i] = reinterpret_cast<unsigned long long*>(m->peer_window[p] + kHaloArrivedOff) + c->rank;
            neighbour[p] = true;
        }
        if (m->recv_sizes[p] > 0) {
            H.recv_peer[H.n_recv_peers++] = p;
            neighbour[p] = true;
        }
    }
    for (int p = 0; p < c->size; ++p)
        if (neighbour[p])
            H.nb_started[H.n_neighbours++] =
                reinterpret_cast<unsigned long long*>(m->peer_window[p] + kHaloStartedOff) + c->rank;
    // (send lists are grouped by peer in rank order: send_begin[i+1] == the next peer's begin,
    //  peers without entries are skipped — their ranges are empty)
    m->n_push = H.n_neighbours == 0 ? 0
                                    : static_cast<int>(std::min<int64_t>(
                                          sm_count(), std::max<int64_t>(1, ceildiv(m->send_total, 4096))));
    H.n_push_ctas = m->n_push;

    // CTA order: row blocks (128 rows) without non-local rows first, the others last
    const int64_t n_blocks = ceildiv(m->local.n_rows, 128);
    std::vector<int32_t> row_list(static_cast<size_t>(m->non_local.nnz > 0 ? m->non_local.n_listed : 0));
    if (!row_list.empty())
        GKOB200_CUDA(cudaMemcpy(row_list.data(), m->non_local.row_list, row_list.size() * sizeof(int32_t),
                                cudaMemcpyDeviceToHost));
    std::vector<int32_t> order, boundary, nl_ptrs(row_list.size() + 1, 0);
    std::vector<int2> thread_range;
    if (!row_list.empty())
        GKOB200_CUDA(cudaMemcpy(nl_ptrs.data(), m->non_local.row_ptrs, nl_ptrs.size() * sizeof(int32_t),
                                cudaMemcpyDeviceToHost));
    std::vector<char> is_boundary(static_cast<size_t>(n_blocks), 0);
    for (size_t j = 0; j < row_list.size(); ++j) {
        const int32_t blk = row_list[j] / 128;
        if (blk < 0 || blk >= n_blocks || (j > 0 && row_list[j] <= row_list[j - 1])) return GKOB200_EINVAL;
        if (boundary.empty() || boundary.back() != blk) {
            boundary.push_back(blk);
            thread_range.resize(boundary.size() * 128, make_int2(0, 0));
            is_boundary[blk] = 1;
        }
        thread_range[(boundary.size() - 1) * 128 + static_cast<size_t>(row_list[j] % 128)] =
            make_int2(nl_ptrs[j], nl_ptrs[j + 1]);
    }
    order.reserve(static_cast<size_t>(n_blocks));
    for (int64_t b = 0; b < n_blocks; ++b)
        if (!is_boundary[b]) order.push_back(static_cast<int32_t>(b));
    H.n_interior = static_cast<int>(order.size());
    m->n_interior = H.n_interior;
    // interior slots as runs of consecutive blocks (slab partitions: one run; <= kHaloRuns runs
    // travel in the kernel parameters, otherwise the kernel looks the slot up in `order`)
    {
        std::vector<int> rs, rb;
        for (size_t sl = 0; sl < order.size(); ++sl)
            if (sl == 0 || order[sl] != order[sl - 1] + 1) {
                rs.push_back(static_cast<int>(sl));
                rb.push_back(order[sl]);
            }
        if (!rs.empty() && rs.size() <= static_cast<size_t>(kHaloRuns)) {
            m->n_runs = static_cast<int>(rs.size());
            for (int i = 0; i < m->n_runs; ++i) {
                m->run_slot[i] = rs[i];
                m->run_block[i] = rb[i];
            }
            m->run_slot[m->n_runs] = H.n_interior;
        }
    }
    order.insert(order.end(), boundary.begin(), boundary.end());
    if ((rc = m->order.alloc(order.size() * sizeof(int32_t) + 16))) return rc;
    if ((rc = m->nl_thread_range.alloc(thread_range.size() * sizeof(int2) + 16))) return rc;
    GKOB200_CUDA(cudaMemcpy(m->order.p, order.data(), order.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (!thread_range.empty())
        GKOB200_CUDA(cudaMemcpy(m->nl_thread_range.p, thread_range.data(), thread_range.size() * sizeof(int2),
                                cudaMemcpyHostToDevice));
    H.order = m->order.as<int32_t>();
    H.nl_thread_range = m->nl_thread_range.as<int2>();
    H.nl_row_ptrs = static_cast<const int32_t*>(m->non_local.row_ptrs);
    H.nl_cols = static_cast<const int32_t*>(m->non_local.col_idxs);
    H.nl_vals = m->non_local.values;
    if ((rc = m->halo_dev.alloc(sizeof(HaloDev)))) return rc;
    GKOB200_CUDA(cudaMemcpy(m->halo_dev.p, &H, sizeof(HaloDev), cudaMemcpyHostToDevice));
    m->fused = true;
    return 0;
}

template <typename V>
__global__ void __launch_bounds__(256)
    pack_halo(int64_t n_send, int64_t k, const int32_t* __restrict__ gather, const V* __restrict__ b, int64_t bs,
              V* __restrict__ out, const int* skip)
{
    if (skip && *skip) return;
    const int64_t total = n_send * k;
    for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t i = t / k, j = t % k;
        out[t] = b[static_cast<int64_t>(gather[i]) * bs + j];
    }
}

template <typename V>
int ensure_buffers(gkob200_dist_matrix* m, int64_t nrhs)
{
    if (m->buf_nrhs >= nrhs && m->consts.p) return 0;
    int rc;
    if ((rc = m->send_buf.alloc(static_cast<size_t>(m->send_total * nrhs + 1) * sizeof(V)))) return rc;
    if ((rc = m->recv_buf.alloc(static_cast<size_t>(m->recv_total * nrhs + 1) * sizeof(V)))) return rc;
    if (!m->consts.p) {
        if ((rc = m->consts.alloc(2 * sizeof(V)))) return rc;
        const V one = V(1);
        GKOB200_CUDA(cudaMemcpy(m->consts.p, &one, sizeof(V), cudaMemcpyHostToDevice));
    }
    m->buf_nrhs = nrhs;
    return 0;
}

// x = A b  or  x = alpha A b + beta x  on the local rows.  `fusion` (nrhs == 1): skip flag
// and the dot w.(A b) are attached to the LAST SpMV of the sequence.  `epoch_bumped`: the
// caller's previous kernel already entered the next halo epoch (distributed CG).
template <typename V>
int dist_apply(gkob200_dist_matrix* m, cudaStream_t s, const V* b, int64_t bs, int64_t nrhs, const V* alpha,
               const V* beta, V* x, int64_t xs, const SpmvFusion<V>* fusion, bool epoch_bumped = false)
{
    int rc;
    gkob200_dist_comm* c = m->comm;
    if (m->fused && nrhs == 1) {
        // ---- fused path: one launch (+ the deferred-reduction finish when a dot is attached)
        SpmvFusion<V> fu;
        if (fusion) fu = *fusion;
        if (!epoch_bumped && !(dist_debug() & 1)) {
            halo_epoch_bump<<<1, 1, 0, s>>>(m->window, fu.skip);
            GKOB200_CHECK_LAUNCH();
            ++m->launches;
        }
        fu.halo = m->halo_dev.as<HaloDev>();
        fu.halo_push_ctas = m->n_push;
        fu.halo_n_interior = m->n_interior;
        fu.halo_debug = dist_debug();
        fu.halo_runs = m->n_runs;
        for (int i = 0; i < kHaloRuns; ++i) {
            fu.halo_run_slot[i] = m->run_slot[i];
            fu.halo_run_block[i] = m->run_block[i];
        }
        fu.halo_run_slot[kHaloRuns] = m->run_slot[kHaloRuns];
        m->exchanged = fu.out != nullptr && fu.p2p != nullptr;
        if ((rc = matrix_apply<V>(s, m->local, b, bs, 1, alpha, beta, x, xs, &fu))) return rc;
        m->launches += fu.out ? 2 : 1;
        return 0;
    }
    if ((rc = ensure_buffers<V>(m, nrhs))) return rc;
    const bool has_halo = (m->send_total > 0 || m->recv_total > 0) && c && c->size > 1;
    V* send = m->send_buf.as<V>();
    V* recv = m->recv_buf.as<V>();
    const int* skip = fusion ? fusion->skip : nullptr;
    if (has_halo) {
        if (m->send_total > 0) {
            pack_halo<V><<<grid_for(m->send_total * nrhs, 256, 4), 256, 0, s>>>(m->send_total, nrhs, m->gather_idxs, b, bs,
                                                                               send, skip);
            GKOB200_CHECK_LAUNCH();
            ++m->launches;
        }
        GKOB200_CUDA(cudaEventRecord(m->packed, s));
        GKOB200_CUDA(cudaStreamWaitEvent(c->comm_stream, m->packed, 0));
        GKOB200_NCCL(ncclGroupStart());
        for (int p = 0; p < c->size; ++p) {
            if (p == c->rank) continue;
            if (m->send_sizes[p] > 0)
                GKOB200_NCCL(ncclSend(send + m->send_offsets[p] * nrhs, static_cast<size_t>(m->send_sizes[p] * nrhs),
                                      nccl_type<V>(), p, c->comm, c->comm_stream));
            if (m->recv_sizes[p] > 0)
                GKOB200_NCCL(ncclRecv(recv + m->recv_offsets[p] * nrhs, static_cast<size_t>(m->recv_sizes[p] * nrhs),
                                      nccl_type<V>(), p, c->comm, c->comm_stream));
        }
        GKOB200_NCCL(ncclGroupEnd());
        GKOB200_CUDA(cudaEventRecord(m->received, c->comm_stream));
        ++m->launches;
    }
    const bool nl = has_halo && m->recv_total > 0 && m->non_local.nnz > 0;
    // local block (overlaps the halo exchange)
    // With a row-compressed non-local block the dot w.(A b) is split: the local SpMV
    // reduces w.(A_loc b) into out[0], the non-local kernel its own contribution into out[1]
    // (the caller adds the two after the all-reduce).  With a full-height non-local CSR the
    // dot is taken once, by the non-local SpMV, on the final result, and out[1] is zeroed.
    const bool split = nl && m->non_local.format == GKOB200_FMT_CSR_ROWS;
    SpmvFusion<V> fl, fn;
    if (fusion) {
        fl = *fusion;
        fn = *fusion;
        if (nl && !split) {
            fl.w = nullptr;
            fl.out = nullptr;
        }
        if (split && fn.out) fn.out = fn.out + 1;
        fl.p2p = nullptr;   // only the last kernel of the apply may exchange
        if (!(split && fn.out)) fn.p2p = nullptr;
        // out[1] is only written by the row-compressed non-local kernel: every other case must
        // not hand the previous iteration's (all-reduced) value to the caller's sum again
        if (fusion->out && nrhs == 1 && !split) GKOB200_CUDA(cudaMemsetAsync(fusion->out + 1, 0, sizeof(V), s));
    }
    m->exchanged = fusion && nl && fn.p2p != nullptr;
    if ((rc = matrix_apply<V>(s, m->local, b, bs, nrhs, alpha, beta, x, xs, fusion ? &fl : nullptr))) return rc;
    ++m->launches;
    if (has_halo) GKOB200_CUDA(cudaStreamWaitEvent(s, m->received, 0));
    if (nl) {
        // x += alpha * A_nl * ghost   (reference: non_local_mtx_->apply(alpha|one, recv, one, x))
        const V* one = m->consts.as<V>();
        if ((rc = matrix_apply<V>(s, m->non_local, recv, nrhs, nrhs, alpha ? alpha : one, one, x, xs,
                                  fusion ? &fn : nullptr)))
            return rc;
        ++m->launches;
    }
    return 0;
}

// ------------------------------- distributed CG --------------------------------
// Reference loop: core/solver/cg.cpp:157-193 on distributed::Vector / distributed::Matrix.
// Fused path, per iteration (graph-captured, `chunk` iterations per graph):
//   dist_cg_direction   p = z + (rho/prev_rho) p ; enters the next halo epoch
//   SpMV                q = A p : halo push + local rows + non-local rows + per-CTA partials of p.q
//   finish_partials     p.q summed, all-reduced over peer memory
//   dist_cg_update      x += t p, r -= t q, z = D^-1 r, (r.z, r.r) reduced, all-reduced over peer
//                       memory, rho bookkeeping, stopping criterion
enum { D_RHO = 0, D_PREV_RHO, D_BETA, D_BETA2, D_TAU, D_ORIG_TAU, D_RED0, D_RED1, D_COUNT };

template <typename V>
struct DistCgParams {
    int64_t n;
    V *x, *r, *z, *p, *q;
    const V* inv_diag;
    V* sc;
    SolverState* st;
    uint8_t* stop_status;
    V* hist;
    V factor;
    int64_t max_iters;
    void* ws;
    const P2pDev* p2p;   // non-null: the reduction finaliser all-reduces over peer memory itself
    unsigned long long* halo_epoch;   // non-null: dist_cg_direction enters the next halo epoch
};

template <typename V>
struct Pair;
template <>
struct Pair<double> { using type = double2; };
template <>
struct Pair<float> { using type = float2; };

// x += t p ; r -= t q ; z = M^-1 r ; partial sums (r.z, r.r) -> sc[D_RED0..1]
// Wide: all vectors 16-byte aligned, n even: two rows per thread through 128-bit accesses
template <typename V, int Mode, bool First, bool Wide>
__global__ void __launch_bounds__(256) dist_cg_update(DistCgParams<V> P)
{
    if (P.st->stopped) return;
    V t = V(0);
    bool upd = false;
    if (!First) {
        const V beta = P.sc[D_BETA] + P.sc[D_BETA2];  // local + non-local share of p.q
        upd = beta != V(0);
        if (upd) t = div_rn(P.sc[D_RHO], beta);
    }
    V acc[2] = {V(0), V(0)};
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t tid0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (Wide) {
        using V2 = typename Pair<V>::type;
        V2* __restrict__ x2 = reinterpret_cast<V2*>(P.x);
        V2* __restrict__ r2 = reinterpret_cast<V2*>(P.r);
        V2* __restrict__ z2 = reinterpret_cast<V2*>(P.z);
        const V2* __restrict__ p2 = reinterpret_cast<const V2*>(P.p);
        const V2* __restrict__ q2 = reinterpret_cast<const V2*>(P.q);
        const V2* __restrict__ d2 = reinterpret_cast<const V2*>(P.inv_diag);
        for (int64_t i = tid0; i < P.n / 2; i += step) {
            V2 rv = r2[i];
            if (upd) {
                V2 xv = x2[i];
                const V2 pv = p2[i], qv = q2[i];
                xv.x = add_rn(xv.x, mul_rn(t, pv.x));
                xv.y = add_rn(xv.y, mul_rn(t, pv.y));
                rv.x = sub_rn(rv.x, mul_rn(t, qv.x));
                rv.y = sub_rn(rv.y, mul_rn(t, qv.y));
                x2[i] = xv;
                r2[i] = rv;
            }
            V2 zv = rv;
            if (Mode == 1) {
                const V2 dv = d2[i];
                zv.x = mul_rn(rv.x, dv.x);
                zv.y = mul_rn(rv.y, dv.y);
                z2[i] = zv;
            }
            acc[0] += rv.x * zv.x + rv.y * zv.y;
            acc[1] += rv.x * rv.x + rv.y * rv.y;
        }
    } else {
        for (int64_t i = tid0; i < P.n; i += step) {
            V ri = P.r[i];
            if (upd) {
                P.x[i] = add_rn(P.x[i], mul_rn(t, P.p[i]));
                ri = sub_rn(ri, mul_rn(t, P.q[i]));
                P.r[i] = ri;
            }
            V zi = ri;
            if (Mode == 1) {
                zi = mul_rn(ri, P.inv_diag[i]);
                P.z[i] = zi;
            }
            acc[0] += ri * zi;
            acc[1] += ri * ri;
        }
    }
    DistCgParams<V> Q = P;
    grid_reduce<2>(acc, ws_partials<V>(P.ws), ws_ticket(P.ws), [Q](V(&tot)[2]) {
        Q.sc[D_RED0] = tot[0];
        Q.sc[D_RED1] = tot[1];
        if (Q.p2p) {
            // all-reduce over peer memory + rho bookkeeping + criterion, all in this finaliser
            // (otherwise: ncclAllReduce + dist_cg_scalars, two more launches)
            if (!peer_allreduce(*Q.p2p, Q.sc + D_RED0, 2)) {
                Q.st->stopped = 1;   // a peer never showed up: the host finds the error word
                return;
            }
            if (!First) Q.sc[D_PREV_RHO] = Q.sc[D_RHO];
            Q.sc[D_RHO] = Q.sc[D_RED0];
            Q.sc[D_TAU] = sqrt_rn(Q.sc[D_RED1]);
            criterion_check(Q.st, 1, Q.sc + D_TAU, Q.sc + D_ORIG_TAU, Q.factor, Q.max_iters, true, Q.stop_status,
                            Q.hist, true);
        }
    });
}

// after the all-reduce: rho bookkeeping + criterion on the GLOBAL residual norm
template <typename V, bool First>
__global__ void dist_cg_scalars(DistCgParams<V> P)
{
    if (P.st->stopped) return;
    if (!First) P.sc[D_PREV_RHO] = P.sc[D_RHO];
    P.sc[D_RHO] = P.sc[D_RED0];
    P.sc[D_TAU] = sqrt_rn(P.sc[D_RED1]);
    criterion_check(P.st, 1, P.sc + D_TAU, P.sc + D_ORIG_TAU, P.factor, P.max_iters, true, P.stop_status, P.hist, true);
}

template <typename V, bool ZisR, bool Wide>
__global__ void __launch_bounds__(256) dist_cg_direction(DistCgParams<V> P)
{
    if (P.st->stopped) return;
    if (P.halo_epoch && blockIdx.x == 0 && threadIdx.x == 0) ++*P.halo_epoch;
    const V prev = P.sc[D_PREV_RHO];
    const bool zero_prev = prev == V(0);
    const V t = zero_prev ? V(0) : div_rn(P.sc[D_RHO], prev);
    const V* __restrict__ z = ZisR ? P.r : P.z;
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t tid0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (Wide) {
        using V2 = typename Pair<V>::type;
        const V2* __restrict__ z2 = reinterpret_cast<const V2*>(z);
        V2* __restrict__ p2 = reinterpret_cast<V2*>(P.p);
        for (int64_t i = tid0; i < P.n / 2; i += step) {
            const V2 zv = z2[i];
            if (zero_prev) {
                p2[i] = zv;
            } else {
                const V2 pv = p2[i];
                V2 o;
                o.x = add_rn(zv.x, mul_rn(t, pv.x));
                o.y = add_rn(zv.y, mul_rn(t, pv.y));
                p2[i] = o;
            }
        }
        return;
    }
    for (int64_t i = tid0; i < P.n; i += step) P.p[i] = zero_prev ? z[i] : add_rn(z[i], mul_rn(t, P.p[i]));
}

template <typename V>
__global__ void __launch_bounds__(256) sq_norm_partial(int64_t n, const V* __restrict__ a, V* out, void* ws)
{
    V acc[1] = {V(0)};
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += step) acc[0] += a[i] * a[i];
    grid_reduce<1>(acc, ws_partials<V>(ws), ws_ticket(ws), [out](V(&tot)[1]) { out[0] = tot[0]; });
}
template <typename V>
__global__ void sqrt_to(V* dst, const V* src) { dst[0] = sqrt_rn(src[0]); }
template <typename V>
__global__ void set_one(V* dst) { dst[0] = V(1); }

template <typename V>
struct DistCgSolver : SolverBase<V> {
    using B = SolverBase<V>;
    using B::M; using B::k; using B::n; using B::stop; using B::launch_count;
    gkob200_dist_matrix* dm = nullptr;
    DevBuf vecs, scal, bigws;
    int64_t ws_blocks = 0;
    // fused path: `chunk` iterations per CUDA graph
    cudaGraphExec_t graph = nullptr;
    cudaStream_t cap_stream = nullptr;
    void* graph_x = nullptr;
    int64_t launches_per_chunk = 0;
    cudaEvent_t ev[2] = {nullptr, nullptr};

    ~DistCgSolver() override
    {
        if (graph) cudaGraphExecDestroy(graph);
        if (cap_stream) cudaStreamDestroy(cap_stream);
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
    }

    int init()
    {
        this->A = dm->local;  // n_rows == local rows; square check in init_base is on the local block
        int rc = this->init_base();
        if (rc) return rc;
        if (k != 1) return GKOB200_EUNSUPPORTED;
        // the p.q dot is fused into the last SpMV of the distributed apply
        if (!matrix_apply_fuses_dot(dm->local, 1) || (dm->non_local.nnz > 0 && !matrix_apply_fuses_dot(dm->non_local, 1)))
            return GKOB200_EUNSUPPORTED;
        if (M.kind != GKOB200_PRECOND_NONE && M.kind != GKOB200_PRECOND_JACOBI_SCALAR) return GKOB200_EUNSUPPORTED;
        if ((rc = vecs.alloc(static_cast<size_t>(n) * 4 * sizeof(V)))) return rc;
        if ((rc = scal.alloc(D_COUNT * sizeof(V)))) return rc;
        ws_blocks = ceildiv(n, 128) + sm_count() + 1;   // row blocks + halo push CTAs
        if (ws_blocks < kReduceMaxBlocks) ws_blocks = kReduceMaxBlocks;
        if ((rc = bigws.alloc(reduce_ws_bytes(ws_blocks)))) return rc;
        if (dm->fused) {
            for (auto& e : ev) GKOB200_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            GKOB200_CUDA(cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking));
        }
        return 0;
    }

    DistCgParams<V> params(V* x)
    {
        DistCgParams<V> P;
        P.n = n;
        P.x = x;
        P.r = vecs.as<V>();
        P.z = P.r + n;
        P.p = P.z + n;
        P.q = P.p + n;
        P.inv_diag = M.kind == GKOB200_PRECOND_JACOBI_SCALAR ? static_cast<const V*>(M.inv_diag) : nullptr;
        P.sc = scal.as<V>();
        P.st = this->st();
        P.stop_status = this->stat();
        P.hist = this->hist.template as<V>();
        P.factor = static_cast<V>(stop.reduction_factor);
        P.max_iters = stop.max_iters;
        P.ws = bigws.p;
        P.p2p = (dm->comm && dm->comm->p2p) ? dm->comm->p2p_dev_ptr : nullptr;
        P.halo_epoch = dm->fused ? reinterpret_cast<unsigned long long*>(dm->window + kHaloEpochOff) : nullptr;
        if (dist_debug() & 1) P.halo_epoch = nullptr;
        if (dist_debug() & 2) P.p2p = nullptr;
        return P;
    }

    // 128-bit accesses: every vector on 16 bytes and an even (float: multiple of 4) row count
    bool wide_ok(const V* x) const
    {
        auto al = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
        return n >= 2 && (static_cast<size_t>(n) * sizeof(V)) % 16 == 0 && al(x) && al(vecs.p) &&
               (M.kind != GKOB200_PRECOND_JACOBI_SCALAR || al(M.inv_diag));
    }

    int allreduce(cudaStream_t s, V* buf, size_t count)
    {
        gkob200_dist_comm* c = dm->comm;
        if (!c || c->size == 1 || (dist_debug() & 2)) return 0;
        ++launch_count;
        return comm_allreduce<V>(c, s, buf, count, &this->st()->stopped, &this->st()->stopped);
    }

    template <bool First>
    int update(cudaStream_t s, V* x)
    {
        DistCgParams<V> P = params(x);
        const bool wide = wide_ok(x);
        const int grid = grid_for(wide ? n / 2 : n, 256, 6);
        if (M.kind == GKOB200_PRECOND_NONE) {
            if (wide) dist_cg_update<V, 0, First, true><<<grid, 256, 0, s>>>(P);
            else dist_cg_update<V, 0, First, false><<<grid, 256, 0, s>>>(P);
        } else {
            if (wide) dist_cg_update<V, 1, First, true><<<grid, 256, 0, s>>>(P);
            else dist_cg_update<V, 1, First, false><<<grid, 256, 0, s>>>(P);
        }
        GKOB200_CHECK_LAUNCH();
        ++launch_count;
        if (P.p2p) return 0;   // exchange + scalars happened in the kernel's finaliser
        int rc = allreduce(s, P.sc + D_RED0, 2);
        if (rc) return rc;
        dist_cg_scalars<V, First><<<1, 1, 0, s>>>(P);
        ++launch_count;
        GKOB200_CHECK_LAUNCH();
        return 0;
    }

    // one iteration: direction, distributed SpMV with the fused dot, update
    int enqueue_iteration(cudaStream_t s, V* x)
    {
        DistCgParams<V> P = params(x);
        const bool wide = wide_ok(x);
        const int grid = grid_for(wide ? n / 2 : n, 256, 6);
        if (M.kind == GKOB200_PRECOND_NONE) {
            if (wide) dist_cg_direction<V, true, true><<<grid, 256, 0, s>>>(P);
            else dist_cg_direction<V, true, false><<<grid, 256, 0, s>>>(P);
        } else {
            if (wide) dist_cg_direction<V, false, true><<<grid, 256, 0, s>>>(P);
            else dist_cg_direction<V, false, false><<<grid, 256, 0, s>>>(P);
        }
        ++launch_count;
        GKOB200_CHECK_LAUNCH();
        SpmvFusion<V> fu;
        fu.skip = &this->st()->stopped;
        fu.on_fail = &this->st()->stopped;
        fu.w = P.p;
        fu.out = P.sc + D_BETA;
        fu.ws = bigws.p;
        fu.ws_blocks = ws_blocks;
        if (P.p2p) {
            fu.p2p = P.p2p;
            fu.p2p_buf = P.sc + D_BETA;
            fu.p2p_count = dm->fused ? 1 : 2;
        }
        const bool no_allreduce = (dist_debug() & 2) != 0;
        int rc;
        if ((rc = dist_apply<V>(dm, s, P.p, 1, 1, nullptr, nullptr, P.q, 1, &fu, dm->fused))) return rc;
        // (a rank whose last SpMV could not exchange all-reduces with the stand-alone kernel)
        if (!dm->exchanged && !no_allreduce && (rc = allreduce(s, P.sc + D_BETA, dm->fused ? 1 : 2))) return rc;
        return update<false>(s, x);
    }

    int build_graph(V* x)
    {
        if (graph && graph_x == x) return 0;
        if (graph) {
            cudaGraphExecDestroy(graph);
            graph = nullptr;
        }
        const int64_t saved = launch_count, saved_m = dm->launches;
        cudaGraph_t g = nullptr;
        GKOB200_CUDA(cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal));
        int rc = 0;
        for (int i = 0; i < this->chunk && rc == 0; ++i) rc = enqueue_iteration(cap_stream, x);
        cudaError_t e = cudaStreamEndCapture(cap_stream, &g);
        launches_per_chunk = (launch_count - saved) + (dm->launches - saved_m);
        launch_count = saved;
        dm->launches = saved_m;
        if (rc) {
            if (g) cudaGraphDestroy(g);
            return rc;
        }
        if (e != cudaSuccess) return static_cast<int>(e);
        e = cudaGraphInstantiate(&graph, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return static_cast<int>(e);
        graph_x = x;
        return 0;
    }

    // sticky error words of the peer-memory exchanges (p2p.cuh)
    int p2p_error()
    {
        gkob200_dist_comm* c = dm->comm;
        if (!c || !c->p2p) return 0;
        int err = 0, herr = 0;
        GKOB200_CUDA(cudaMemcpy(&err, c->p2p_dev.block[c->rank] + kP2pErrorOff, sizeof(int), cudaMemcpyDeviceToHost));
        if (dm->fused) GKOB200_CUDA(cudaMemcpy(&herr, dm->window + kHaloErrorOff, sizeof(int), cudaMemcpyDeviceToHost));
        return err ? 2000 : herr ? 2001 : 0;
    }

    int apply(cudaStream_t s, const void* b_, int64_t bs, void* x_, int64_t xs) override
    {
        const V* b = static_cast<const V*>(b_);
        V* x = static_cast<V*>(x_);
        launch_count = 0;
        this->num_iterations = 0;
        dm->launches = 0;
        if (bs != 1 || xs != 1) return GKOB200_EUNSUPPORTED;
        DistCgParams<V> P = params(x);
        int rc;
        if ((rc = this->reset_state(s))) return rc;
        // stopping_status starts cleared on every apply (cg::initialize does it in the reference,
        // reference/solver/cg_kernels.cpp:60-63): a converged status left by an earlier apply on
        // this solver object would end the new solve at iteration 0
        GKOB200_CUDA(cudaMemsetAsync(this->stat(), 0, static_cast<size_t>(k), s));
        // z = p = q = 0 (cg::initialize): the first direction update is p = z + rho * p_old and
        // must not see the previous solve's p
        // r = b ; r = -A x + r ; baseline norm (sum of local squared norms, then sqrt:
        // core/distributed/vector.cpp:394-407)
        if ((rc = typed::dense_copy(B::tag(), s, n, int64_t(1), b, int64_t(1), P.r, int64_t(1)))) return rc;
        GKOB200_CUDA(cudaMemsetAsync(P.z, 0, static_cast<size_t>(n) * 3 * sizeof(V), s));
        GKOB200_CUDA(cudaMemsetAsync(P.sc, 0, D_COUNT * sizeof(V), s));
        if ((rc = typed::dense_fill(B::tag(), s, int64_t(1), int64_t(1), P.sc + D_PREV_RHO, int64_t(1), V(1)))) return rc;
        if ((rc = dist_apply<V>(dm, s, x, 1, 1, this->neg_one(), this->one(), P.r, 1, nullptr))) return rc;
        launch_count += 3;
        if (stop.baseline == GKOB200_STOP_ABSOLUTE) {
            set_one<V><<<1, 1, 0, s>>>(P.sc + D_ORIG_TAU);
        } else {
            const V* src = stop.baseline == GKOB200_STOP_RHS_NORM ? b : P.r;
            sq_norm_partial<V><<<grid_for(n, 256, 4), 256, 0, s>>>(n, src, P.sc + D_RED0, bigws.p);
            if ((rc = allreduce(s, P.sc + D_RED0, 1))) return rc;
            sqrt_to<V><<<1, 1, 0, s>>>(P.sc + D_ORIG_TAU, P.sc + D_RED0);
            launch_count += 2;
        }
        GKOB200_CHECK_LAUNCH();
        if ((rc = update<true>(s, x))) return rc;
        if (dm->fused) {
            // every rank launches the same number of graphs: the stop flag derives from
            // all-reduced values, so the pinned copies the hosts read are identical
            if ((rc = build_graph(x))) return rc;
            int64_t g = 0;
            bool done = false;
            const int64_t max_chunks = ceildiv(stop.max_iters, this->chunk);
            while (!done) {
                if (g >= max_chunks) break;
                GKOB200_CUDA(cudaGraphLaunch(graph, s));
                launch_count += launches_per_chunk;
                const int slot = static_cast<int>(g & 1);
                GKOB200_CUDA(cudaMemcpyAsync(&this->h_state[slot], this->st(), sizeof(SolverState),
                                             cudaMemcpyDeviceToHost, s));
                GKOB200_CUDA(cudaEventRecord(ev[slot], s));
                if (g >= 1) {
                    GKOB200_CUDA(cudaEventSynchronize(ev[slot ^ 1]));
                    if (this->h_state[slot ^ 1].stopped) done = true;
                }
                ++g;
            }
            GKOB200_CUDA(cudaStreamSynchronize(s));
        } else {
            int64_t it = 0;
            bool stopped = false;
            while (true) {
                if (it % this->chunk == 0 || it >= stop.max_iters) {
                    if ((rc = this->poll(s, &stopped))) return rc;
                    if (stopped) break;
                }
                if ((rc = enqueue_iteration(s, x))) return rc;
                ++it;
            }
        }
        launch_count += dm->launches;
        if ((rc = this->finish(s))) return rc;
        return p2p_error();
    }
};

}  // namespace
}  // namespace gkob200

using namespace gkob200;

extern "C" {

int gkob200_nccl_unique_id(void* out128)
{
    if (!out128) return GKOB200_EINVAL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    GKOB200_NCCL(ncclGetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return 0;
}

int gkob200_dist_comm_create(const void* id128, int rank, int size, gkob200_dist_comm** out)
{
    if (!out || rank < 0 || size < 1 || rank >= size) return GKOB200_EINVAL;
    auto* c = new gkob200_dist_comm();
    c->rank = rank;
    c->size = size;
    if (size > 1) {
        if (!id128) {
            delete c;
            return GKOB200_EINVAL;
        }
        ncclUniqueId id;
        memcpy(&id, id128, sizeof(id));
        ncclResult_t r = ncclCommInitRank(&c->comm, size, id, rank);
        if (r != ncclSuccess) {
            delete c;
            return 1000 + static_cast<int>(r);
        }
    }
    cudaError_t e = cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete c;
        return static_cast<int>(e);
    }
    if (size > 1) {
        int rc = p2p_setup(c);
        if (rc) {
            delete c;
            return rc;
        }
    }
    *out = c;
    return 0;
}

/* 1 when scalar all-reduces run over peer memory (CUDA IPC) instead of NCCL */
int gkob200_dist_comm_uses_p2p(const gkob200_dist_comm* c) { return c && c->p2p ? 1 : 0; }

int gkob200_dist_comm_destroy(gkob200_dist_comm* c)
{
    if (!c) return 0;
    if (c->p2p) {
        cudaDeviceSynchronize();
        p2p_shared_free(c, c->p2p_dev.block);
        cudaFree(c->p2p_dev_ptr);
    }
    if (c->comm) ncclCommDestroy(c->comm);
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
    delete c;
    return 0;
}

int gkob200_dist_comm_rank(const gkob200_dist_comm* c) { return c ? c->rank : -1; }
int gkob200_dist_comm_size(const gkob200_dist_comm* c) { return c ? c->size : -1; }

/* in-place sum all-reduce of `count` values on `stream` (distributed::Vector reductions) */
int gkob200_dist_allreduce_sum_f64(gkob200_dist_comm* c, void* stream, double* buf, int64_t count)
{
    if (!c || count < 0) return GKOB200_EINVAL;
    return comm_allreduce<double>(c, as_stream(stream), buf, static_cast<size_t>(count));
}
int gkob200_dist_allreduce_sum_f32(gkob200_dist_comm* c, void* stream, float* buf, int64_t count)
{
    if (!c || count < 0) return GKOB200_EINVAL;
    return comm_allreduce<float>(c, as_stream(stream), buf, static_cast<size_t>(count));
}
/* all-to-all of `count` int64 per peer (device buffers): setup exchange of halo sizes */
int gkob200_dist_alltoall_i64(gkob200_dist_comm* c, void* stream, const int64_t* send, int64_t* recv, int64_t count)
{
    if (!c || !send || !recv || count < 0) return GKOB200_EINVAL;
    cudaStream_t s = as_stream(stream);
    if (c->size == 1) {
        GKOB200_CUDA(cudaMemcpyAsync(recv, send, sizeof(int64_t) * count, cudaMemcpyDeviceToDevice, s));
        return 0;
    }
    GKOB200_NCCL(ncclGroupStart());
    for (int p = 0; p < c->size; ++p) {
        GKOB200_NCCL(ncclSend(send + p * count, static_cast<size_t>(count), ncclInt64, p, c->comm, s));
        GKOB200_NCCL(ncclRecv(recv + p * count, static_cast<size_t>(count), ncclInt64, p, c->comm, s));
    }
    GKOB200_NCCL(ncclGroupEnd());
    return 0;
}
/* all-to-all-v of int32 (device buffers): gather indices travel from receivers to senders
 * [ref: core/distributed/matrix.cpp:218-221] */
int gkob200_dist_alltoallv_i32(gkob200_dist_comm* c, void* stream, const int32_t* send, const int64_t* send_sizes_host,
                               const int64_t* send_offsets_host, int32_t* recv, const int64_t* recv_sizes_host,
                               const int64_t* recv_offsets_host)
{
    if (!c) return GKOB200_EINVAL;
    cudaStream_t s = as_stream(stream);
    if (c->size == 1) {
        if (send_sizes_host[0] > 0)
            GKOB200_CUDA(cudaMemcpyAsync(recv + recv_offsets_host[0], send + send_offsets_host[0],
                                         sizeof(int32_t) * send_sizes_host[0], cudaMemcpyDeviceToDevice, s));
        return 0;
    }
    GKOB200_NCCL(ncclGroupStart());
    for (int p = 0; p < c->size; ++p) {
        if (send_sizes_host[p] > 0)
            GKOB200_NCCL(ncclSend(send + send_offsets_host[p], static_cast<size_t>(send_sizes_host[p]), ncclInt32, p, c->comm, s));
        if (recv_sizes_host[p] > 0)
            GKOB200_NCCL(ncclRecv(recv + recv_offsets_host[p], static_cast<size_t>(recv_sizes_host[p]), ncclInt32, p, c->comm, s));
    }
    GKOB200_NCCL(ncclGroupEnd());
    return 0;
}

int gkob200_dist_matrix_create(gkob200_dist_comm* comm, const gkob200_matrix* local, const gkob200_matrix* non_local,
                               const int32_t* gather_idxs, const int64_t* send_sizes_host,
                               const int64_t* recv_sizes_host, gkob200_dist_matrix** out)
{
    if (!comm || !local || !out || !send_sizes_host || !recv_sizes_host) return GKOB200_EINVAL;
    auto* m = new gkob200_dist_matrix();
    m->comm = comm;
    m->local = *local;
    if (non_local) m->non_local = *non_local;
    m->gather_idxs = gather_idxs;
    const int P = comm->size;
    m->send_sizes.assign(send_sizes_host, send_sizes_host + P);
    m->recv_sizes.assign(recv_sizes_host, recv_sizes_host + P);
    m->send_offsets.assign(P + 1, 0);
    m->recv_offsets.assign(P + 1, 0);
    for (int p = 0; p < P; ++p) {
        m->send_offsets[p + 1] = m->send_offsets[p] + m->send_sizes[p];
        m->recv_offsets[p + 1] = m->recv_offsets[p] + m->recv_sizes[p];
    }
    m->send_total = m->send_offsets[P];
    m->recv_total = m->recv_offsets[P];
    if (cudaEventCreateWithFlags(&m->packed, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->received, cudaEventDisableTiming) != cudaSuccess) {
        delete m;
        return static_cast<int>(cudaGetLastError());
    }
    // collective: every rank of the communicator creates its part of the matrix here
    const int rc = halo_setup(m);
    if (rc) {
        gkob200_dist_matrix_destroy(m);
        return rc;
    }
    *out = m;
    return 0;
}

/* 1 when apply() runs the halo exchange inside the SpMV launch over peer memory (see the file
 * header); 0: pack + ncclSend/ncclRecv + separate non-local SpMV.  GKOB200_FUSED_HALO=0 in the
 * environment forces the latter. */
int gkob200_dist_matrix_uses_fused_halo(const gkob200_dist_matrix* m) { return m && m->fused ? 1 : 0; }

int gkob200_dist_matrix_destroy(gkob200_dist_matrix* m)
{
    if (!m) return 0;
    if (m->window) {
        cudaDeviceSynchronize();
        p2p_shared_free(m->comm, m->peer_window);
        m->window = nullptr;
    }
    if (m->packed) cudaEventDestroy(m->packed);
    if (m->received) cudaEventDestroy(m->received);
    delete m;
    return 0;
}

int gkob200_dist_matrix_apply(gkob200_dist_matrix* m, void* stream, const void* b, int64_t b_stride, int64_t nrhs,
                              const void* alpha, const void* beta, void* x, int64_t x_stride)
{
    if (!m || (alpha == nullptr) != (beta == nullptr)) return GKOB200_EINVAL;
    if (m->local.value_type == GKOB200_F64)
        return dist_apply<double>(m, as_stream(stream), static_cast<const double*>(b), b_stride, nrhs,
                                  static_cast<const double*>(alpha), static_cast<const double*>(beta),
                                  static_cast<double*>(x), x_stride, nullptr);
    return dist_apply<float>(m, as_stream(stream), static_cast<const float*>(b), b_stride, nrhs,
                             static_cast<const float*>(alpha), static_cast<const float*>(beta), static_cast<float*>(x),
                             x_stride, nullptr);
}

int gkob200_dist_solver_create(int kind, gkob200_dist_matrix* A, const gkob200_precond* M, const gkob200_stop* stop,
                               int64_t nrhs, gkob200_solver** out)
{
    if (!A || !stop || !out) return GKOB200_EINVAL;
    *out = nullptr;
    if (kind != GKOB200_SOLVER_CG) return GKOB200_EUNSUPPORTED;
    gkob200_stop clamped = *stop;
    if (clamped.max_iters < 0) clamped.max_iters = 0;
    if (clamped.max_iters > (int64_t(1) << 40)) clamped.max_iters = int64_t(1) << 40;
    stop = &clamped;
    int rc;
    if (A->local.value_type == GKOB200_F64) {
        auto* s = new DistCgSolver<double>();
        s->dm = A;
        if (M) s->M = *M; else { s->M = gkob200_precond{}; s->M.kind = GKOB200_PRECOND_NONE; }
        s->stop = *stop;
        s->k = nrhs;
        s->nrhs = nrhs;
        if ((rc = s->init())) { delete s; return rc; }
        *out = s;
    } else {
        auto* s = new DistCgSolver<float>();
        s->dm = A;
        if (M) s->M = *M; else { s->M = gkob200_precond{}; s->M.kind = GKOB200_PRECOND_NONE; }
        s->stop = *stop;
        s->k = nrhs;
        s->nrhs = nrhs;
        if ((rc = s->init())) { delete s; return rc; }
        *out = s;
    }
    return 0;
}

}  // extern "C"

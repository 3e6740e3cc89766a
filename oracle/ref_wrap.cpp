// ref_wrap.cpp — TEST INFRASTRUCTURE.  A thin C-ABI over the UNMODIFIED Ginkgo 1.5.0
// reference / OpenMP executors built by oracle/Makefile.ref into oracle/_ref/lib.
// It lets tests pull golden outputs from the real reference (to pin the plain-C
// oracle and to check the CUDA path) and lets bench.py time the reference's own CPU
// implementation (`cpu_baseline.kind == "reference"`).  Our code; it only calls
// Ginkgo's public API.  Never loaded by the product path.
#include <ginkgo/ginkgo.hpp>
#include <ginkgo/core/distributed/partition.hpp>

#include "core/distributed/matrix_kernels.hpp"

#include <omp.h>

#include <chrono>
#include <fstream>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

namespace {

std::shared_ptr<gko::Executor> make_exec(int kind)
{
    // 0: ReferenceExecutor (sequential oracle), 1: OmpExecutor (CPU baseline)
    if (kind == 1) return gko::OmpExecutor::create();
    return gko::ReferenceExecutor::create();
}

template <typename T>
gko::array<T> view(std::shared_ptr<const gko::Executor> exec, int64_t n, const T* p)
{
    return gko::array<T>::view(exec, static_cast<gko::size_type>(n), const_cast<T*>(p));
}

template <typename V>
std::unique_ptr<gko::matrix::Dense<V>> dense_view(std::shared_ptr<const gko::Executor> exec, int64_t n, int64_t k,
                                                  const V* p, int64_t stride)
{
    return gko::matrix::Dense<V>::create(exec, gko::dim<2>(n, k), view<V>(exec, n * stride, p),
                                         static_cast<gko::size_type>(stride));
}

template <typename V, typename I>
std::unique_ptr<gko::matrix::Csr<V, I>> csr_view(std::shared_ptr<const gko::Executor> exec, int64_t n_rows,
                                                 int64_t n_cols, int64_t nnz, const I* rp, const I* ci, const V* va)
{
    using Csr = gko::matrix::Csr<V, I>;
    return Csr::create(exec, gko::dim<2>(n_rows, n_cols), view<V>(exec, nnz, va), view<I>(exec, nnz, ci),
                       view<I>(exec, n_rows + 1, rp), std::make_shared<typename Csr::classical>());
}

// Records ||r|| at every iteration_complete event, the same quantity
// stop::ResidualNorm computes from the residual (core/stop/residual_norm.cpp:196-203).
template <typename V>
struct HistoryLogger : gko::log::Logger {
    mutable std::vector<double> hist;
    mutable int64_t iters = 0;
    void on_iteration_complete(const gko::LinOp*, const gko::size_type& it, const gko::LinOp* r, const gko::LinOp*,
                               const gko::LinOp* tau) const override
    {
        iters = static_cast<int64_t>(it);
        if (auto t = dynamic_cast<const gko::matrix::Dense<V>*>(tau)) {
            // GMRES hands the criterion its implicit residual norm (core/solver/gmres.cpp:236-243)
            hist.push_back(static_cast<double>(t->get_executor()->copy_val_to_host(t->get_const_values())));
        } else if (auto d = dynamic_cast<const gko::matrix::Dense<V>*>(r)) {
            auto nrm = gko::matrix::Dense<V>::create(d->get_executor(), gko::dim<2>(1, d->get_size()[1]));
            d->compute_norm2(nrm.get());
            auto h = nrm->get_executor()->get_master();
            hist.push_back(static_cast<double>(h->copy_val_to_host(nrm->get_const_values())));
        }
    }
    HistoryLogger(std::shared_ptr<const gko::Executor> exec)
        : gko::log::Logger(exec, gko::log::Logger::iteration_complete_mask)
    {}
};

template <typename V, typename I>
std::shared_ptr<gko::LinOp> to_format(std::shared_ptr<gko::matrix::Csr<V, I>> csr, int format, int64_t hybrid_limit)
{
    auto exec = csr->get_executor();
    switch (format) {
    case 0:
        return csr;
    case 1: {
        auto m = gko::share(gko::matrix::Ell<V, I>::create(exec));
        csr->convert_to(m.get());
        return m;
    }
    case 2: {
        auto m = gko::share(gko::matrix::Sellp<V, I>::create(exec));
        csr->convert_to(m.get());
        return m;
    }
    case 3: {
        auto m = gko::share(gko::matrix::Coo<V, I>::create(exec));
        csr->convert_to(m.get());
        return m;
    }
    case 4: {
        using Hyb = gko::matrix::Hybrid<V, I>;
        std::shared_ptr<typename Hyb::strategy_type> strat;
        if (hybrid_limit >= 0)
            strat = std::make_shared<typename Hyb::column_limit>(static_cast<gko::size_type>(hybrid_limit));
        else
            strat = std::make_shared<typename Hyb::automatic>();
        auto m = gko::share(Hyb::create(exec, strat));
        csr->convert_to(m.get());
        return m;
    }
    default:
        return nullptr;
    }
}

template <typename V, typename I>
int spmv_impl(int exec_kind, int format, int64_t hybrid_limit, int64_t n_rows, int64_t n_cols, int64_t nnz,
              const I* rp, const I* ci, const V* va, const V* b, int64_t bs, int64_t nrhs, const V* alpha,
              const V* beta, V* c, int64_t cs, int reps, double* seconds)
{
    try {
        auto exec = make_exec(exec_kind);
        auto csr = gko::share(csr_view<V, I>(exec, n_rows, n_cols, nnz, rp, ci, va));
        auto A = to_format<V, I>(csr, format, hybrid_limit);
        if (!A) return -2;
        auto db = dense_view<V>(exec, n_cols, nrhs, b, bs);
        auto dc = dense_view<V>(exec, n_rows, nrhs, c, cs);
        std::unique_ptr<gko::matrix::Dense<V>> da, dbeta;
        if (alpha) {
            da = gko::initialize<gko::matrix::Dense<V>>({*alpha}, exec);
            dbeta = gko::initialize<gko::matrix::Dense<V>>({*beta}, exec);
        }
        double best = 1e300, total = 0;
        for (int r = 0; r < (reps > 0 ? reps : 1); ++r) {
            exec->synchronize();
            auto t0 = std::chrono::steady_clock::now();
            if (alpha)
                A->apply(da.get(), db.get(), dbeta.get(), dc.get());
            else
                A->apply(db.get(), dc.get());
            exec->synchronize();
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            best = dt < best ? dt : best;
            total += dt;
        }
        if (seconds) {
            seconds[0] = total / (reps > 0 ? reps : 1);
            seconds[1] = best;
        }
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_wrap: %s\n", e.what());
        return -1;
    }
}

// solver_kind 0 CG, 1 BiCGSTAB, 2 GMRES, 3 FCG, 4 CGS; precond_block: 0 none, 1 scalar Jacobi, >1 block Jacobi
template <typename V, typename I>
int64_t solve_impl(int exec_kind, int solver_kind, int format, int64_t hybrid_limit, int64_t n, int64_t nnz,
                   const I* rp, const I* ci, const V* va, int precond_block, int64_t max_iters, double factor,
                   int baseline, int64_t krylov_dim, int64_t nrhs, const V* b, V* x, double* hist, int64_t hist_cap,
                   int64_t* hist_len, double* seconds)
{
    try {
        auto exec = make_exec(exec_kind);
        auto csr = gko::share(csr_view<V, I>(exec, n, n, nnz, rp, ci, va));
        auto A = to_format<V, I>(csr, format, hybrid_limit);
        if (!A) return -2;
        auto db = dense_view<V>(exec, n, nrhs, b, nrhs);
        auto dx = dense_view<V>(exec, n, nrhs, x, nrhs);
        std::vector<std::shared_ptr<const gko::stop::CriterionFactory>> crit;
        crit.push_back(gko::stop::Iteration::build().with_max_iters(static_cast<gko::size_type>(max_iters)).on(exec));
        if (factor > 0) {
            auto mode = baseline == 0   ? gko::stop::mode::rhs_norm
                        : baseline == 1 ? gko::stop::mode::initial_resnorm
                                        : gko::stop::mode::absolute;
            crit.push_back(gko::stop::ResidualNorm<V>::build()
                               .with_reduction_factor(static_cast<gko::remove_complex<V>>(factor))
                               .with_baseline(mode)
                               .on(exec));
        }
        std::shared_ptr<const gko::LinOp> M;
        if (precond_block > 0) {
            M = gko::preconditioner::Jacobi<V, I>::build()
                    .with_max_block_size(static_cast<gko::uint32>(precond_block))
                    .on(exec)
                    ->generate(gko::as<gko::LinOp>(csr));
        }
        std::unique_ptr<gko::LinOp> solver;
        if (solver_kind == 0) {
            auto f = gko::solver::Cg<V>::build().with_criteria(crit);
            if (M) f.with_generated_preconditioner(M);
            solver = f.on(exec)->generate(A);
        } else if (solver_kind == 1) {
            auto f = gko::solver::Bicgstab<V>::build().with_criteria(crit);
            if (M) f.with_generated_preconditioner(M);
            solver = f.on(exec)->generate(A);
        } else if (solver_kind == 2) {
            auto f = gko::solver::Gmres<V>::build().with_criteria(crit).with_krylov_dim(
                static_cast<gko::size_type>(krylov_dim));
            if (M) f.with_generated_preconditioner(M);
            solver = f.on(exec)->generate(A);
        } else if (solver_kind == 3) {
            auto f = gko::solver::Fcg<V>::build().with_criteria(crit);
            if (M) f.with_generated_preconditioner(M);
            solver = f.on(exec)->generate(A);
        } else if (solver_kind == 4) {
            auto f = gko::solver::Cgs<V>::build().with_criteria(crit);
            if (M) f.with_generated_preconditioner(M);
            solver = f.on(exec)->generate(A);
        } else {
            return -2;
        }
        auto logger = std::make_shared<HistoryLogger<V>>(exec);
        solver->add_logger(logger);
        exec->synchronize();
        auto t0 = std::chrono::steady_clock::now();
        solver->apply(db.get(), dx.get());
        exec->synchronize();
        if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        // the last iteration_complete event carries the final iteration count
        // (core/solver/cg.cpp:163-164: logged right before the criterion check)
        const int64_t iters = logger->iters;
        if (hist) {
            int64_t m = static_cast<int64_t>(logger->hist.size());
            if (m > hist_cap) m = hist_cap;
            for (int64_t i = 0; i < m; ++i) hist[i] = logger->hist[i];
            if (hist_len) *hist_len = m;
        }
        return iters;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_wrap: %s\n", e.what());
        return -1;
    }
}

// ---- format conversions through the real reference executor ----------------------
template <typename V, typename I>
int64_t convert_impl(int format, int hyb_kind, int64_t hyb_param, double percent, double ratio, int64_t slice_size,
                     int64_t stride_factor, int64_t n_rows, int64_t n_cols, int64_t nnz, const I* rp, const I* ci,
                     const V* va, int64_t* meta, I* out_idx_a, I* out_idx_b, V* out_vals, uint64_t* out_u64_a,
                     uint64_t* out_u64_b, I* coo_rows, I* coo_cols, V* coo_vals, int64_t cap)
{
    try {
        auto exec = gko::ReferenceExecutor::create();
        auto csr = csr_view<V, I>(exec, n_rows, n_cols, nnz, rp, ci, va);
        if (format == 1) {
            auto m = gko::matrix::Ell<V, I>::create(exec);
            csr->convert_to(m.get());
            const int64_t total = m->get_num_stored_elements();
            meta[0] = m->get_num_stored_elements_per_row();
            meta[1] = m->get_stride();
            if (total > cap) return -3;
            std::copy_n(m->get_const_col_idxs(), total, out_idx_a);
            std::copy_n(m->get_const_values(), total, out_vals);
            return total;
        } else if (format == 2) {
            auto m = gko::matrix::Sellp<V, I>::create(exec, gko::dim<2>{}, slice_size, stride_factor, 0);
            csr->convert_to(m.get());
            const int64_t ns = gko::ceildiv(n_rows, slice_size);
            const int64_t total = m->get_num_stored_elements();
            meta[0] = m->get_total_cols();
            meta[1] = ns;
            if (total > cap) return -3;
            std::copy_n(m->get_const_slice_sets(), ns + 1, out_u64_a);
            std::copy_n(m->get_const_slice_lengths(), ns, out_u64_b);
            std::copy_n(m->get_const_col_idxs(), total, out_idx_a);
            std::copy_n(m->get_const_values(), total, out_vals);
            return total;
        } else if (format == 3) {
            auto m = gko::matrix::Coo<V, I>::create(exec);
            csr->convert_to(m.get());
            if (nnz > cap) return -3;
            std::copy_n(m->get_const_row_idxs(), nnz, out_idx_a);
            std::copy_n(m->get_const_col_idxs(), nnz, out_idx_b);
            std::copy_n(m->get_const_values(), nnz, out_vals);
            return nnz;
        } else if (format == 4) {
            using Hyb = gko::matrix::Hybrid<V, I>;
            std::shared_ptr<typename Hyb::strategy_type> strat;
            switch (hyb_kind) {
            case 0: strat = std::make_shared<typename Hyb::column_limit>(static_cast<gko::size_type>(hyb_param)); break;
            case 1: strat = std::make_shared<typename Hyb::imbalance_limit>(percent); break;
            case 2: strat = std::make_shared<typename Hyb::imbalance_bounded_limit>(percent, ratio); break;
            case 3: strat = std::make_shared<typename Hyb::minimal_storage_limit>(); break;
            default: strat = std::make_shared<typename Hyb::automatic>(); break;
            }
            auto m = Hyb::create(exec, strat);
            csr->convert_to(m.get());
            const int64_t ell_total = m->get_ell_num_stored_elements();
            const int64_t coo_nnz = m->get_coo_num_stored_elements();
            meta[0] = m->get_ell_num_stored_elements_per_row();
            meta[1] = m->get_ell_stride();
            meta[2] = coo_nnz;
            if (ell_total > cap || coo_nnz > cap) return -3;
            std::copy_n(m->get_const_ell_col_idxs(), ell_total, out_idx_a);
            std::copy_n(m->get_const_ell_values(), ell_total, out_vals);
            std::copy_n(m->get_const_coo_row_idxs(), coo_nnz, coo_rows);
            std::copy_n(m->get_const_coo_col_idxs(), coo_nnz, coo_cols);
            std::copy_n(m->get_const_coo_values(), coo_nnz, coo_vals);
            return ell_total;
        }
        return -2;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_wrap: %s\n", e.what());
        return -1;
    }
}

// ---- block-Jacobi generate through the real reference executor -------------------
template <typename V, typename I>
int64_t jacobi_impl(int64_t n, int64_t nnz, const I* rp, const I* ci, const V* va, int max_block_size, int64_t* meta,
                    I* block_ptrs, V* blocks, int64_t cap)
{
    try {
        auto exec = gko::ReferenceExecutor::create();
        auto csr = gko::share(csr_view<V, I>(exec, n, n, nnz, rp, ci, va));
        auto jac = gko::preconditioner::Jacobi<V, I>::build()
                       .with_max_block_size(static_cast<gko::uint32>(max_block_size))
                       .on(exec)
                       ->generate(gko::as<gko::LinOp>(csr));
        const int64_t nb = jac->get_num_blocks();
        const auto& sch = jac->get_storage_scheme();
        meta[0] = nb;
        meta[1] = sch.block_offset;
        meta[2] = sch.group_offset;
        meta[3] = sch.group_power;
        const int64_t stored = jac->get_num_stored_elements();
        if (stored > cap) return -3;
        if (max_block_size > 1) {
            std::copy_n(jac->get_parameters().block_pointers.get_const_data(), nb + 1, block_ptrs);
        }
        std::copy_n(jac->get_blocks(), stored, blocks);
        return stored;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_wrap: %s\n", e.what());
        return -1;
    }
}

// ---- distributed index maps through the real reference kernels (no MPI needed: the
// reference's own tests loop over local_part in one process,
// reference/test/distributed/matrix_kernels.cpp:86-152) ------------------------------
using Part = gko::experimental::distributed::Partition<int32_t, int64_t>;

std::unique_ptr<Part> make_partition(std::shared_ptr<const gko::Executor> exec, int mode, int num_parts,
                                     int64_t global_size, const int64_t* ranges, const int32_t* mapping)
{
    if (mode == 0) return Part::build_from_global_size_uniform(exec, num_parts, global_size);
    if (mode == 1) return Part::build_from_contiguous(exec, view<int64_t>(exec, num_parts + 1, ranges));
    return Part::build_from_mapping(exec, view<int32_t>(exec, global_size, mapping), num_parts);
}

template <typename V>
int dist_build_impl(int mode, int num_parts, int64_t global_size, const int64_t* ranges, const int32_t* mapping,
                    int64_t nnz, const int64_t* rows, const int64_t* cols, const V* vals, int local_part,
                    int64_t* counts, int32_t* lrow, int32_t* lcol, V* lval, int32_t* nrow, int32_t* ncol, V* nval,
                    int32_t* gather, int32_t* recv_sizes, int64_t* nl_to_global, int64_t* part_meta)
{
    try {
        auto exec = gko::ReferenceExecutor::create();
        auto part = make_partition(exec, mode, num_parts, global_size, ranges, mapping);
        gko::device_matrix_data<V, int64_t> input{exec, gko::dim<2>(part->get_size(), part->get_size()),
                                                  view<int64_t>(exec, nnz, rows), view<int64_t>(exec, nnz, cols),
                                                  view<V>(exec, nnz, vals)};
        gko::array<int32_t> a_lrow{exec}, a_lcol{exec}, a_nrow{exec}, a_ncol{exec}, a_gather{exec};
        gko::array<V> a_lval{exec}, a_nval{exec};
        gko::array<int32_t> a_recv{exec, static_cast<gko::size_type>(num_parts)};
        gko::array<int64_t> a_nl2g{exec};
        gko::kernels::reference::distributed_matrix::build_local_nonlocal(
            exec, input, part.get(), part.get(), local_part, a_lrow, a_lcol, a_lval, a_nrow, a_ncol, a_nval, a_gather,
            a_recv, a_nl2g);
        counts[0] = a_lrow.get_num_elems();
        counts[1] = a_nrow.get_num_elems();
        counts[2] = a_gather.get_num_elems();
        std::copy_n(a_lrow.get_const_data(), counts[0], lrow);
        std::copy_n(a_lcol.get_const_data(), counts[0], lcol);
        std::copy_n(a_lval.get_const_data(), counts[0], lval);
        std::copy_n(a_nrow.get_const_data(), counts[1], nrow);
        std::copy_n(a_ncol.get_const_data(), counts[1], ncol);
        std::copy_n(a_nval.get_const_data(), counts[1], nval);
        std::copy_n(a_gather.get_const_data(), counts[2], gather);
        std::copy_n(a_nl2g.get_const_data(), counts[2], nl_to_global);
        std::copy_n(a_recv.get_const_data(), num_parts, recv_sizes);
        // partition description: num_ranges, then range_bounds[num_ranges+1], part_ids, starting idxs, part sizes
        if (part_meta) {
            const int64_t nr = part->get_num_ranges();
            int64_t* o = part_meta;
            *o++ = nr;
            for (int64_t i = 0; i <= nr; ++i) *o++ = part->get_range_bounds()[i];
            for (int64_t i = 0; i < nr; ++i) *o++ = part->get_part_ids()[i];
            for (int64_t i = 0; i < nr; ++i) *o++ = part->get_range_starting_indices()[i];
            for (int i = 0; i < num_parts; ++i) *o++ = part->get_part_sizes()[i];
        }
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_wrap: %s\n", e.what());
        return -1;
    }
}

// device_matrix_data assembly on the reference executor: sort_row_major -> sum_duplicates ->
// remove_zeros (each step optional), then the Csr the reference would read from it.
template <typename V, typename I>
int64_t assemble_impl(int steps, int64_t n_rows, int64_t n_cols, int64_t nnz, I* rows, I* cols, V* vals, I* out_rp)
{
    auto exec = gko::ReferenceExecutor::create();
    gko::device_matrix_data<V, I> data{exec, gko::dim<2>(n_rows, n_cols), static_cast<gko::size_type>(nnz)};
    std::memcpy(data.get_row_idxs(), rows, nnz * sizeof(I));
    std::memcpy(data.get_col_idxs(), cols, nnz * sizeof(I));
    std::memcpy(data.get_values(), vals, nnz * sizeof(V));
    if (steps & 1) data.sort_row_major();
    if (steps & 2) data.sum_duplicates();
    if (steps & 4) data.remove_zeros();
    const int64_t m = static_cast<int64_t>(data.get_num_elems());
    std::memcpy(rows, data.get_const_row_idxs(), m * sizeof(I));
    std::memcpy(cols, data.get_const_col_idxs(), m * sizeof(I));
    std::memcpy(vals, data.get_const_values(), m * sizeof(V));
    if (out_rp) {
        auto csr = gko::matrix::Csr<V, I>::create(exec);
        csr->read(data);
        std::memcpy(out_rp, csr->get_const_row_ptrs(), (n_rows + 1) * sizeof(I));
    }
    return m;
}

// op 0: transpose (out arrays of the transposed matrix), op 1: sort_by_column_index (in place:
// out_ci / out_va receive the sorted arrays, out_rp a copy of rp)
template <typename V, typename I>
int csr_op_impl(int op, int64_t n_rows, int64_t n_cols, int64_t nnz, const I* rp, const I* ci, const V* va, I* out_rp,
                I* out_ci, V* out_va)
{
    auto exec = gko::ReferenceExecutor::create();
    auto A = csr_view<V, I>(exec, n_rows, n_cols, nnz, rp, ci, va);
    if (op == 0) {
        auto T = gko::as<gko::matrix::Csr<V, I>>(A->transpose());
        std::memcpy(out_rp, T->get_const_row_ptrs(), (n_cols + 1) * sizeof(I));
        std::memcpy(out_ci, T->get_const_col_idxs(), nnz * sizeof(I));
        std::memcpy(out_va, T->get_const_values(), nnz * sizeof(V));
    } else {
        auto B = A->clone();
        B->sort_by_column_index();
        std::memcpy(out_rp, B->get_const_row_ptrs(), (n_rows + 1) * sizeof(I));
        std::memcpy(out_ci, B->get_const_col_idxs(), nnz * sizeof(I));
        std::memcpy(out_va, B->get_const_values(), nnz * sizeof(V));
    }
    return 0;
}

// gko::read_generic_raw / write_raw / write_binary_raw on a file
int64_t mtx_read_impl(const char* path, int64_t* dims, int64_t cap, int64_t* rows, int64_t* cols, double* vals)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) return -2;
    try {
        auto data = gko::read_generic_raw<double, gko::int64>(in);
        dims[0] = data.size[0];
        dims[1] = data.size[1];
        const int64_t n = static_cast<int64_t>(data.nonzeros.size());
        if (n > cap) return -3;
        for (int64_t i = 0; i < n; ++i) {
            rows[i] = data.nonzeros[i].row;
            cols[i] = data.nonzeros[i].column;
            vals[i] = data.nonzeros[i].value;
        }
        return n;
    } catch (const std::exception&) {
        return -1;
    }
}

int mtx_write_impl(const char* path, int format, int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* rows,
                   const int64_t* cols, const double* vals, int index32, int value32)
{
    std::ofstream out(path, std::ios::binary);
    if (!out) return -2;
    auto fill = [&](auto& data) {
        for (int64_t i = 0; i < nnz; ++i) data.nonzeros.emplace_back(rows[i], cols[i], vals[i]);
    };
    if (format == 0 || format == 2) {
        gko::matrix_data<double, gko::int64> d(gko::dim<2>(n_rows, n_cols));
        fill(d);
        gko::write_raw(out, d, format == 0 ? gko::layout_type::coordinate : gko::layout_type::array);
    } else if (index32 && value32) {
        gko::matrix_data<float, gko::int32> d(gko::dim<2>(n_rows, n_cols));
        fill(d);
        gko::write_binary_raw(out, d);
    } else if (index32) {
        gko::matrix_data<double, gko::int32> d(gko::dim<2>(n_rows, n_cols));
        fill(d);
        gko::write_binary_raw(out, d);
    } else if (value32) {
        gko::matrix_data<float, gko::int64> d(gko::dim<2>(n_rows, n_cols));
        fill(d);
        gko::write_binary_raw(out, d);
    } else {
        gko::matrix_data<double, gko::int64> d(gko::dim<2>(n_rows, n_cols));
        fill(d);
        gko::write_binary_raw(out, d);
    }
    return 0;
}

}  // namespace

extern "C" {

int64_t ref_mtx_read(const char* path, int64_t* dims, int64_t cap, int64_t* rows, int64_t* cols, double* vals)
{
    return mtx_read_impl(path, dims, cap, rows, cols, vals);
}
int ref_mtx_write(const char* path, int format, int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* rows,
                  const int64_t* cols, const double* vals, int index32, int value32)
{
    return mtx_write_impl(path, format, n_rows, n_cols, nnz, rows, cols, vals, index32, value32);
}

int ref_num_threads(void) { return omp_get_max_threads(); }
void ref_set_num_threads(int n) { omp_set_num_threads(n); }

#define REF_SPMV(V, VT, I, IT)                                                                                   \
    int ref_spmv_##V##_##I(int exec_kind, int format, int64_t hybrid_limit, int64_t n_rows, int64_t n_cols,      \
                           int64_t nnz, const IT* rp, const IT* ci, const VT* va, const VT* b, int64_t bs,       \
                           int64_t nrhs, const VT* alpha, const VT* beta, VT* c, int64_t cs, int reps,           \
                           double* seconds)                                                                      \
    {                                                                                                            \
        return spmv_impl<VT, IT>(exec_kind, format, hybrid_limit, n_rows, n_cols, nnz, rp, ci, va, b, bs, nrhs,  \
                                 alpha, beta, c, cs, reps, seconds);                                             \
    }                                                                                                            \
    int64_t ref_solve_##V##_##I(int exec_kind, int solver_kind, int format, int64_t hybrid_limit, int64_t n,     \
                                int64_t nnz, const IT* rp, const IT* ci, const VT* va, int precond_block,        \
                                int64_t max_iters, double factor, int baseline, int64_t krylov_dim,              \
                                int64_t nrhs, const VT* b, VT* x, double* hist, int64_t hist_cap,                \
                                int64_t* hist_len, double* seconds)                                              \
    {                                                                                                            \
        return solve_impl<VT, IT>(exec_kind, solver_kind, format, hybrid_limit, n, nnz, rp, ci, va,              \
                                  precond_block, max_iters, factor, baseline, krylov_dim, nrhs, b, x, hist,      \
                                  hist_cap, hist_len, seconds);                                                  \
    }
#define REF_CONVERT(V, VT, I, IT)                                                                                \
    int64_t ref_convert_##V##_##I(int format, int hyb_kind, int64_t hyb_param, double percent, double ratio,     \
                                  int64_t slice_size, int64_t stride_factor, int64_t n_rows, int64_t n_cols,     \
                                  int64_t nnz, const IT* rp, const IT* ci, const VT* va, int64_t* meta,          \
                                  IT* out_idx_a, IT* out_idx_b, VT* out_vals, uint64_t* out_u64_a,               \
                                  uint64_t* out_u64_b, IT* coo_rows, IT* coo_cols, VT* coo_vals, int64_t cap)    \
    {                                                                                                            \
        return convert_impl<VT, IT>(format, hyb_kind, hyb_param, percent, ratio, slice_size, stride_factor,      \
                                    n_rows, n_cols, nnz, rp, ci, va, meta, out_idx_a, out_idx_b, out_vals,       \
                                    out_u64_a, out_u64_b, coo_rows, coo_cols, coo_vals, cap);                    \
    }
int ref_dist_build_f64(int mode, int num_parts, int64_t global_size, const int64_t* ranges, const int32_t* mapping,
                       int64_t nnz, const int64_t* rows, const int64_t* cols, const double* vals, int local_part,
                       int64_t* counts, int32_t* lrow, int32_t* lcol, double* lval, int32_t* nrow, int32_t* ncol,
                       double* nval, int32_t* gather, int32_t* recv_sizes, int64_t* nl_to_global, int64_t* part_meta)
{
    return dist_build_impl<double>(mode, num_parts, global_size, ranges, mapping, nnz, rows, cols, vals, local_part,
                                   counts, lrow, lcol, lval, nrow, ncol, nval, gather, recv_sizes, nl_to_global,
                                   part_meta);
}
int64_t ref_jacobi_generate_f64_i32(int64_t n, int64_t nnz, const int32_t* rp, const int32_t* ci, const double* va,
                                    int max_block_size, int64_t* meta, int32_t* block_ptrs, double* blocks,
                                    int64_t cap)
{
    return jacobi_impl<double, int32_t>(n, nnz, rp, ci, va, max_block_size, meta, block_ptrs, blocks, cap);
}
int64_t ref_jacobi_generate_f32_i32(int64_t n, int64_t nnz, const int32_t* rp, const int32_t* ci, const float* va,
                                    int max_block_size, int64_t* meta, int32_t* block_ptrs, float* blocks, int64_t cap)
{
    return jacobi_impl<float, int32_t>(n, nnz, rp, ci, va, max_block_size, meta, block_ptrs, blocks, cap);
}
#define REF_SETUP(V, VT, I, IT)                                                                                  \
    int64_t ref_assemble_##V##_##I(int steps, int64_t n_rows, int64_t n_cols, int64_t nnz, IT* rows, IT* cols,   \
                                   VT* vals, IT* out_rp)                                                         \
    { return assemble_impl<VT, IT>(steps, n_rows, n_cols, nnz, rows, cols, vals, out_rp); }                      \
    int ref_csr_op_##V##_##I(int op, int64_t n_rows, int64_t n_cols, int64_t nnz, const IT* rp, const IT* ci,    \
                             const VT* va, IT* out_rp, IT* out_ci, VT* out_va)                                   \
    { return csr_op_impl<VT, IT>(op, n_rows, n_cols, nnz, rp, ci, va, out_rp, out_ci, out_va); }
REF_SETUP(f64, double, i32, int32_t)
REF_SETUP(f32, float, i32, int32_t)
REF_SETUP(f64, double, i64, int64_t)
REF_CONVERT(f64, double, i32, int32_t)
REF_CONVERT(f32, float, i32, int32_t)
REF_CONVERT(f64, double, i64, int64_t)
REF_SPMV(f64, double, i32, int32_t)
REF_SPMV(f32, float, i32, int32_t)
REF_SPMV(f64, double, i64, int64_t)

}  // extern "C"

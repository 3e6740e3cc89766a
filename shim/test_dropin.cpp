// shim/test_dropin.cpp — the drop-in, end to end: an ordinary Ginkgo program (public API
// only) run once on gko::ReferenceExecutor and once on gko::CudaExecutor, where
// libginkgo_cuda.so is the B200 shim over libgko_b200.so.  Nothing below knows about
// gko_b200: matrix::{Csr,Ell,Sellp,Coo,Hybrid}::apply, solver::{Cg,Bicgstab,Gmres} with
// preconditioner::Jacobi and stop::{Iteration,ResidualNorm} keep their signatures.
#include <ginkgo/ginkgo.hpp>

#include <cmath>
#include <cstdio>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

using V = double;
using I = gko::int32;
using Dense = gko::matrix::Dense<V>;
using Csr = gko::matrix::Csr<V, I>;

static int failures = 0;
#define EXPECT(cond, what)                                            \
    do {                                                              \
        const bool ok__ = (cond);                                     \
        std::printf("%-58s %s\n", what, ok__ ? "ok" : "FAILED");     \
        if (!ok__) ++failures;                                        \
    } while (0)

gko::matrix_data<V, I> stencil27(int nx, int ny, int nz)
{
    gko::matrix_data<V, I> d{gko::dim<2>(nx * ny * nz)};
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x)
                for (int dz = -1; dz <= 1; ++dz)
                    for (int dy = -1; dy <= 1; ++dy)
                        for (int dx = -1; dx <= 1; ++dx) {
                            const int X = x + dx, Y = y + dy, Z = z + dz;
                            if (X < 0 || X >= nx || Y < 0 || Y >= ny || Z < 0 || Z >= nz) continue;
                            const bool diag = !dx && !dy && !dz;
                            d.nonzeros.emplace_back(x + nx * (y + ny * z), X + nx * (Y + ny * Z), diag ? 26.0 : -1.0);
                        }
    return d;
}

double rel_diff(const Dense* a_dev, const Dense* b_ref)
{
    auto a = gko::clone(b_ref->get_executor(), a_dev);
    double num = 0, den = 0;
    for (gko::size_type i = 0; i < a->get_size()[0]; ++i)
        for (gko::size_type j = 0; j < a->get_size()[1]; ++j) {
            const double d = a->at(i, j) - b_ref->at(i, j);
            num += d * d;
            den += b_ref->at(i, j) * b_ref->at(i, j);
        }
    return std::sqrt(num / (den > 0 ? den : 1));
}

template <typename Mtx, typename... Args>
void check_format(const char* name, std::shared_ptr<gko::Executor> ref, std::shared_ptr<gko::Executor> cuda,
                  const gko::matrix_data<V, I>& data, const Dense* b_ref, int nrhs, Args&&... args)
{
    auto A_ref = Mtx::create(ref, std::forward<Args>(args)...);
    A_ref->read(data);
    auto A_dev = gko::clone(cuda, A_ref);
    const auto n = data.size[0];
    auto x_ref = Dense::create(ref, gko::dim<2>(n, nrhs));
    auto x_dev = Dense::create(cuda, gko::dim<2>(n, nrhs));
    auto b_dev = gko::clone(cuda, b_ref);
    A_ref->apply(b_ref, x_ref.get());
    A_dev->apply(b_dev.get(), x_dev.get());
    EXPECT(rel_diff(x_dev.get(), x_ref.get()) <= 1e-14, (std::string(name) + "::apply(b, x)").c_str());
    auto alpha = gko::initialize<Dense>({-0.5}, ref), beta = gko::initialize<Dense>({2.0}, ref);
    auto alpha_d = gko::clone(cuda, alpha), beta_d = gko::clone(cuda, beta);
    A_ref->apply(alpha.get(), b_ref, beta.get(), x_ref.get());
    A_dev->apply(alpha_d.get(), b_dev.get(), beta_d.get(), x_dev.get());
    EXPECT(rel_diff(x_dev.get(), x_ref.get()) <= 1e-14, (std::string(name) + "::apply(alpha, b, beta, x)").c_str());
}

template <typename Solver>
void check_solver(const char* name, std::shared_ptr<gko::Executor> ref, std::shared_ptr<gko::Executor> cuda,
                  std::shared_ptr<Csr> A_ref, const Dense* b_ref, unsigned block_size)
{
    int iters[2];
    std::unique_ptr<Dense> xs[2];
    std::shared_ptr<gko::Executor> execs[2] = {ref, cuda};
    for (int e = 0; e < 2; ++e) {
        auto exec = execs[e];
        auto A = gko::share(gko::clone(exec, A_ref));
        auto b = gko::clone(exec, b_ref);
        auto x = Dense::create(exec, b_ref->get_size());
        x->fill(0.0);
        auto it_crit = gko::share(gko::stop::Iteration::build().with_max_iters(500u).on(exec));
        auto res_crit = gko::share(gko::stop::ResidualNorm<V>::build().with_reduction_factor(1e-10).on(exec));
        auto logger = gko::share(gko::log::Convergence<V>::create(exec));
        it_crit->add_logger(logger);
        res_crit->add_logger(logger);
        auto factory = Solver::build().with_criteria(it_crit, res_crit);
        std::shared_ptr<gko::LinOp> solver;
        if (block_size > 0) {
            auto M = gko::share(gko::preconditioner::Jacobi<V, I>::build()
                                    .with_max_block_size(block_size)
                                    .with_skip_sorting(true)
                                    .on(exec)
                                    ->generate(A));
            solver = factory.with_generated_preconditioner(M).on(exec)->generate(A);
        } else {
            solver = factory.on(exec)->generate(A);
        }
        solver->apply(b.get(), x.get());
        iters[e] = static_cast<int>(logger->get_num_iterations());
        xs[e] = std::move(x);
    }
    char what[160];
    std::snprintf(what, sizeof(what), "%s + Jacobi(%u): iterations ref %d / cuda %d", name, block_size, iters[0], iters[1]);
    // CG / GMRES: +-2 (BASELINE.md par. 5).  BiCGSTAB's convergence is not monotone and reacts to the
    // rounding of its four dot products (sequential sums on the reference executor, pairwise
    // tree sums here): its count is held to +-12 % instead.
    const bool erratic = std::string(name) == "Bicgstab" || std::string(name) == "Cgs";
    const int tol = erratic ? std::max(2, iters[0] * 12 / 100) : 2;
    EXPECT(std::abs(iters[0] - iters[1]) <= tol && iters[0] < 500, what);
    std::snprintf(what, sizeof(what), "%s + Jacobi(%u): solution", name, block_size);
    EXPECT(rel_diff(xs[1].get(), xs[0].get()) <= 1e-8, what);
}

int main()
{
    auto ref = gko::ReferenceExecutor::create();
    if (gko::CudaExecutor::get_num_devices() == 0) {
        std::printf("no CUDA device\n");
        return 2;
    }
    auto cuda = gko::CudaExecutor::create(0, ref);
    std::cout << "cuda module: " << gko::version_info::get().cuda_version << std::endl;
    const auto data = stencil27(24, 23, 22);
    const auto n = data.size[0];
    for (int nrhs : {1, 3}) {
        auto b = Dense::create(ref, gko::dim<2>(n, nrhs));
        for (gko::size_type i = 0; i < n; ++i)
            for (int j = 0; j < nrhs; ++j) b->at(i, j) = std::sin(0.01 * i + j);
        std::printf("--- SpMV, %d right-hand side(s)\n", nrhs);
        check_format<Csr>("Csr(classical)", ref, cuda, data, b.get(), nrhs, std::make_shared<Csr::classical>());
        check_format<Csr>("Csr(merge_path)", ref, cuda, data, b.get(), nrhs, std::make_shared<Csr::merge_path>());
        check_format<Csr>("Csr(load_balance)", ref, cuda, data, b.get(), nrhs, std::make_shared<Csr::load_balance>(132));
        check_format<gko::matrix::Ell<V, I>>("Ell", ref, cuda, data, b.get(), nrhs);
        check_format<gko::matrix::Sellp<V, I>>("Sellp", ref, cuda, data, b.get(), nrhs);
        check_format<gko::matrix::Coo<V, I>>("Coo", ref, cuda, data, b.get(), nrhs);
        using Hyb = gko::matrix::Hybrid<V, I>;
        check_format<Hyb>("Hybrid(column_limit 16)", ref, cuda, data, b.get(), nrhs,
                          std::make_shared<Hyb::column_limit>(16));
    }
    std::printf("--- solvers\n");
    auto A = gko::share(Csr::create(ref, std::make_shared<Csr::classical>()));
    A->read(data);
    auto b = Dense::create(ref, gko::dim<2>(n, 1));
    for (gko::size_type i = 0; i < n; ++i) b->at(i, 0) = std::sin(0.01 * i);
    for (unsigned bs : {0u, 1u, 8u}) {
        check_solver<gko::solver::Cg<V>>("Cg", ref, cuda, A, b.get(), bs);
        check_solver<gko::solver::Bicgstab<V>>("Bicgstab", ref, cuda, A, b.get(), bs);
        check_solver<gko::solver::Gmres<V>>("Gmres", ref, cuda, A, b.get(), bs);
        check_solver<gko::solver::Fcg<V>>("Fcg", ref, cuda, A, b.get(), bs);
        check_solver<gko::solver::Cgs<V>>("Cgs", ref, cuda, A, b.get(), bs);
    }
    std::printf("--- setup kernels\n");
    {
        auto A_dev = gko::clone(cuda, A);
        auto T_ref = gko::as<Csr>(A->transpose());
        auto T_dev = gko::clone(ref, gko::as<Csr>(A_dev->transpose()));
        bool same = T_ref->get_num_stored_elements() == T_dev->get_num_stored_elements();
        for (gko::size_type i = 0; same && i <= T_ref->get_size()[0]; ++i)
            same = T_ref->get_const_row_ptrs()[i] == T_dev->get_const_row_ptrs()[i];
        for (gko::size_type k = 0; same && k < T_ref->get_num_stored_elements(); ++k)
            same = T_ref->get_const_col_idxs()[k] == T_dev->get_const_col_idxs()[k] &&
                   T_ref->get_const_values()[k] == T_dev->get_const_values()[k];
        EXPECT(same, "Csr::transpose identical to the reference executor");
        auto S_dev = gko::clone(cuda, T_ref);
        S_dev->sort_by_column_index();
        auto S_back = gko::clone(ref, S_dev);
        same = true;
        for (gko::size_type k = 0; same && k < T_ref->get_num_stored_elements(); ++k)
            same = T_ref->get_const_col_idxs()[k] == S_back->get_const_col_idxs()[k] &&
                   T_ref->get_const_values()[k] == S_back->get_const_values()[k];
        EXPECT(same, "Csr::sort_by_column_index keeps a sorted matrix");
    }
    std::printf(failures ? "DROPIN_FAILED (%d)\n" : "DROPIN_OK\n", failures);
    return failures ? 1 : 0;
}

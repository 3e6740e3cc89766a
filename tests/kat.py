"""Known-answer tests lifted (as literals) from the reference's own unit tests.
Each entry cites the reference test it restates.  Used twice: against the CPU oracle
(tests/test_oracle_kat.py, no GPU) and against the CUDA path (tests/test_gpu_*.py)."""
import numpy as np
import scipy.sparse as sp


def dense_to_csr(rows, vdtype=np.float64, idtype=np.int32, keep_zeros=False):
    a = np.array(rows, dtype=vdtype)
    m = sp.csr_matrix(a)
    m.sort_indices()
    return m.indptr.astype(idtype), m.indices.astype(idtype), m.data.astype(vdtype), a.shape


# reference/test/matrix/csr_kernels.cpp:98-118 (fixture) — {{1,3,2},{0,5,0}}
CSR_MTX = dict(row_ptrs=[0, 3, 4], col_idxs=[0, 1, 2, 1], values=[1.0, 3.0, 2.0, 5.0], shape=(2, 3))

CSR_APPLY_KATS = [
    # (name, b, alpha, beta, c_in, expected) — csr_kernels.cpp:358-452
    ("AppliesToDenseVector", [[2.0], [1.0], [4.0]], None, None, None, [[13.0], [5.0]]),
    ("AppliesToDenseMatrix", [[2.0, 3.0], [1.0, -1.5], [4.0, 2.5]], None, None, None, [[13.0, 3.5], [5.0, -7.5]]),
    ("AppliesLinearCombinationToDenseVector", [[2.0], [1.0], [4.0]], -1.0, 2.0, [[1.0], [2.0]], [[-11.0], [-1.0]]),
    ("AppliesLinearCombinationToDenseMatrix", [[2.0, 3.0], [1.0, -1.5], [4.0, 2.5]], -1.0, 2.0,
     [[1.0, 0.5], [2.0, -1.5]], [[-11.0, -2.5], [-1.0, 4.5]]),
]

# reference/test/solver/cg_kernels.cpp:153-253 — step kernels on 2x2 vectors, column 1 stopped
CG_STEP1 = dict(p=3.0, z=-2.0, rho=[2.0, 3.0], prev_rho=[8.0, 3.0], stop=[0, 0x41],
                expect_p=[[-1.25, 3.0], [-1.25, 3.0]])
CG_STEP1_DIV0 = dict(p=3.0, z=-2.0, rho=[1.0, 1.0], prev_rho=[0.0, 0.0], stop=[0, 0],
                     expect_p=[[-2.0, -2.0], [-2.0, -2.0]])
CG_STEP2 = dict(x=-2.0, p=3.0, r=4.0, q=-5.0, rho=[2.0, 3.0], beta=[8.0, 3.0], stop=[0, 0x41],
                expect_x=[[-1.25, -2.0], [-1.25, -2.0]], expect_r=[[5.25, 4.0], [5.25, 4.0]])
CG_STEP2_DIV0 = dict(x=-2.0, p=3.0, r=4.0, q=-5.0, rho=[1.0, 1.0], beta=[0.0, 0.0], stop=[0, 0],
                     expect_x=[[-2.0, -2.0], [-2.0, -2.0]], expect_r=[[4.0, 4.0], [4.0, 4.0]])

# cg_kernels.cpp:62-64,255-264 SolvesStencilSystem; :325-340 SolvesMultipleStencilSystems
STENCIL3 = [[2.0, -1.0, 0.0], [-1.0, 2.0, -1.0], [0.0, -1.0, 2.0]]
CG_SOLVE_KATS = [
    ("SolvesStencilSystem", STENCIL3, [[-1.0], [3.0], [1.0]], [[1.0], [3.0], [2.0]], 400, 1.0),
    ("SolvesMultipleStencilSystems", STENCIL3, [[-1.0, 1.0], [3.0, 0.0], [1.0, 1.0]],
     [[1.0, 1.0], [3.0, 1.0], [2.0, 1.0]], 400, 1.0),
]
# cg_kernels.cpp:79-86 mtx_big, :447-494 SolvesBigDenseSystem1/2 (tolerance r*1e2)
BIG6 = [[8828.0, 2673.0, 4150.0, -3139.5, 3829.5, 5856.0],
        [2673.0, 10765.5, 1805.0, 73.0, 1966.0, 3919.5],
        [4150.0, 1805.0, 6472.5, 2656.0, 2409.5, 3836.5],
        [-3139.5, 73.0, 2656.0, 6048.0, 665.0, -132.0],
        [3829.5, 1966.0, 2409.5, 665.0, 4240.5, 4373.5],
        [5856.0, 3919.5, 3836.5, -132.0, 4373.5, 5678.0]]
CG_SOLVE_KATS += [
    ("SolvesBigDenseSystem1", BIG6, [[1300083.0], [1018120.5], [906410.0], [-42679.5], [846779.5], [1176858.5]],
     [[81.0], [55.0], [45.0], [5.0], [85.0], [-10.0]], 100, 1e2),
    ("SolvesBigDenseSystem2", BIG6, [[886630.5], [-172578.0], [684522.0], [-65310.5], [455487.5], [607436.0]],
     [[33.0], [-56.0], [81.0], [-30.0], [21.0], [40.0]], 100, 1e2),
]

# fcg_kernels.cpp: the same fixtures as CG (mtx :62-64, mtx_big :77-84); SolvesStencilSystem :268-280,
# SolvesMultipleStencilSystems :340-356, SolvesBigDenseSystem1/2 :460-475 / :494-509 (tolerance r*1e3)
FCG_SOLVE_KATS = [(n, A, b, x, it, 1e3 if "Big" in n else t) for (n, A, b, x, it, t) in CG_SOLVE_KATS]
# cgs_kernels.cpp: mtx :63-64, SolvesDenseSystem :319-331 (40 iterations, r), SolvesMultipleDenseSystem
# :391-408, mtx_big :76-82, SolvesBigDenseSystem1 :513-527 (r*1e3), SolvesBigDenseSystem2 :545-559 (r*1e2)
CGS_DENSE3 = [[1.0, -3.0, 0.0], [-4.0, 1.0, -3.0], [2.0, -1.0, 2.0]]
CGS_BIG6 = [[-99.0, 87.0, -67.0, -62.0, -68.0, -19.0],
            [-30.0, -17.0, -1.0, 9.0, 23.0, 77.0],
            [80.0, 89.0, 36.0, 94.0, 55.0, 34.0],
            [-31.0, 21.0, 96.0, -26.0, 24.0, -57.0],
            [60.0, 45.0, -16.0, -4.0, 96.0, 24.0],
            [69.0, 32.0, -68.0, 57.0, -30.0, -51.0]]
CGS_SOLVE_KATS = [
    ("SolvesDenseSystem", CGS_DENSE3, [[-1.0], [3.0], [1.0]], [[-4.0], [-1.0], [4.0]], 40, 1.0),
    # (tolerance of this one: half_tol = sqrt(r), encoded as 0)
    ("SolvesMultipleDenseSystem", CGS_DENSE3, [[-1.0, -5.0], [3.0, 1.0], [1.0, -2.0]],
     [[-4.0, 1.0], [-1.0, 2.0], [4.0, -1.0]], 40, 0.0),
    ("SolvesBigDenseSystem1", CGS_BIG6, [[764.0], [-4032.0], [-11855.0], [7111.0], [-12765.0], [-4589.0]],
     [[-13.0], [-49.0], [69.0], [-33.0], [-82.0], [-39.0]], 100, 1e3),
    ("SolvesBigDenseSystem2", CGS_BIG6, [[17356.0], [5466.0], [748.0], [-456.0], [3434.0], [-7020.0]],
     [[-58.0], [98.0], [-16.0], [-58.0], [2.0], [76.0]], 100, 1e2),
]


def rtol(dtype):
    """r<T>::value of the reference tests: 10 * eps (core/test/utils.hpp:213-219)."""
    return 10 * np.finfo(dtype).eps


def rel_frobenius(a, b):
    """GKO_ASSERT_MTX_NEAR's metric: ||a-b||_F / ||b||_F (core/test/utils/assertions.hpp:206-230)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)

# ---- block-Jacobi KATs: reference/test/preconditioner/jacobi_kernels.cpp ----------------
# (row_ptrs, col_idxs, max_block_size, expected block pointers)
JACOBI_FIND_BLOCKS_KATS = [
    ("FindsNaturalBlocks:160-187", [0, 2, 4, 6, 8], [0, 1, 0, 1, 0, 2, 0, 2], 3, [0, 2, 4]),
    ("ExecutesSupervariableAgglomeration:190-219", [0, 2, 4, 6, 8, 9], [0, 1, 0, 1, 2, 3, 2, 3, 4], 3, [0, 2, 5]),
    ("AdheresToBlockSizeBound:222-254", [0, 1, 2, 3, 4, 5, 6, 7], [0, 1, 2, 3, 4, 5, 6], 3, [0, 3, 6, 7]),
    # fixture matrix (:92-104) with unknown block sizes (:257-272)
    ("CanBeGeneratedWithUnknownBlockSizes:257-272", [0, 3, 5, 7, 10, 13], [0, 1, 4, 0, 1, 2, 3, 2, 3, 4, 0, 3, 4], 3,
     [0, 3, 5]),
]
# fixture matrix of the suite (:92-104) and InvertsDiagonalBlocks (:275-299) with block pointers {0,2,5}
JACOBI_MTX = dict(row_ptrs=[0, 3, 5, 7, 10, 13], col_idxs=[0, 1, 4, 0, 1, 2, 3, 2, 3, 4, 0, 3, 4],
                  values=[4.0, -2.0, -2.0, -1.0, 4.0, 4.0, -2.0, -1.0, 4.0, -2.0, -1.0, -1.0, 4.0])
JACOBI_BLOCK_PTRS = [0, 2, 5]
JACOBI_INV_B1 = [[4.0 / 14, 2.0 / 14], [1.0 / 14, 4.0 / 14]]
JACOBI_INV_B2 = [[14.0 / 48, 8.0 / 48, 4.0 / 48], [4.0 / 48, 16.0 / 48, 8.0 / 48], [1.0 / 48, 4.0 / 48, 14.0 / 48]]

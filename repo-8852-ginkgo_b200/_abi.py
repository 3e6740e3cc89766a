"""ctypes declarations of the C-ABI in include/gko_b200.h / gko_b200_solver.h."""
import ctypes as C

vp, i64, i32, u8, u64, f64, f32, sz = (
    C.c_void_p, C.c_int64, C.c_int32, C.c_uint8, C.c_uint64, C.c_double, C.c_float, C.c_size_t,
)

VT = {"f64": f64, "f32": f32}

# error codes
EINVAL, EUNSUPPORTED, EWORKSPACE = -1, -2, -3
REDUCE_WS_BYTES = 1024 + (148 * 16 + 256) * 8 * 8

CSR_CLASSICAL, CSR_MERGE_PATH, CSR_AUTO, CSR_MERGE_PATH_PLANNED = 0, 1, 2, 3
F64, F32 = 0, 1
I32, I64 = 0, 1
FMT_CSR, FMT_ELL, FMT_SELLP, FMT_COO, FMT_HYBRID, FMT_CSR_ROWS = range(6)
PRECOND_NONE, PRECOND_JACOBI_SCALAR, PRECOND_JACOBI_BLOCK = range(3)
STOP_RHS_NORM, STOP_INITIAL_RESNORM, STOP_ABSOLUTE = range(3)
SOLVER_CG, SOLVER_BICGSTAB, SOLVER_GMRES, SOLVER_FCG, SOLVER_CGS = range(5)


class Matrix(C.Structure):
    _fields_ = [
        ("format", i32), ("value_type", i32), ("index_type", i32), ("csr_strategy", i32),
        ("n_rows", i64), ("n_cols", i64), ("nnz", i64),
        ("row_ptrs", vp), ("col_idxs", vp), ("values", vp),
        ("csr_max_block_nnz", i64),
        ("ell_stride", i64), ("ell_width", i64), ("ell_col_idxs", vp), ("ell_values", vp),
        ("slice_size", i64), ("stride_factor", i64), ("n_slices", i64),
        ("slice_sets", vp), ("slice_lengths", vp),
        ("coo_nnz", i64), ("coo_row_idxs", vp), ("coo_col_idxs", vp), ("coo_values", vp),
        ("workspace", vp), ("workspace_bytes", sz),
        ("row_list", vp), ("n_listed", i64),
        ("sellp_max_slice_len", i64), ("sellp_total_cols", i64),
    ]


class Precond(C.Structure):
    _fields_ = [
        ("kind", i32), ("value_type", i32), ("inv_diag", vp),
        ("num_blocks", i64), ("block_pointers", vp), ("blocks", vp),
        ("block_offset", i64), ("group_offset", i64), ("group_power", i32), ("max_block_size", i32),
    ]


class Stop(C.Structure):
    _fields_ = [("max_iters", i64), ("reduction_factor", f64), ("baseline", i32), ("check_every", i32)]


def _decl(lib, name, argtypes, restype=C.c_int):
    fn = getattr(lib, name)
    fn.argtypes = argtypes
    fn.restype = restype
    return fn


def declare(lib):
    d = lambda name, args, res=C.c_int: _decl(lib, name, args, res)  # noqa: E731
    d("gkob200_version", [])
    d("gkob200_sm_count", [])
    d("gkob200_reduce_ws_init", [vp, vp])
    d("gkob200_fill_array", [vp, vp, i64, C.c_int, vp])
    for I in ("i32", "i64"):
        d(f"gkob200_csr_row_stats_{I}", [vp, i64, vp, vp])
    d("gkob200_csr_pick_strategy", [i64, i64, i64, i64])
    for I in ("i32", "i64"):
        d(f"gkob200_csr_merge_plan_{I}", [vp, i64, i64, vp, vp, sz])
    d("gkob200_csr_spmv_workspace_bytes", [i64, i64, i64, C.c_int], sz)
    for V, T in VT.items():
        for I in ("i32", "i64"):
            d(f"gkob200_csr_spmv_{V}_{I}",
              [vp, i64, i64, i64, vp, vp, vp, vp, i64, i64, vp, vp, vp, i64, C.c_int, i64, vp, sz])
        d(f"gkob200_dense_fill_{V}", [vp, i64, i64, vp, i64, T])
        d(f"gkob200_dense_copy_{V}", [vp, i64, i64, vp, i64, vp, i64])
        d(f"gkob200_dense_scale_{V}", [vp, i64, i64, vp, i64, vp, i64])
        d(f"gkob200_dense_inv_scale_{V}", [vp, i64, i64, vp, i64, vp, i64])
        d(f"gkob200_dense_add_scaled_{V}", [vp, i64, i64, vp, i64, vp, i64, vp, i64])
        d(f"gkob200_dense_sub_scaled_{V}", [vp, i64, i64, vp, i64, vp, i64, vp, i64])
        d(f"gkob200_dense_compute_dot_{V}", [vp, i64, i64, vp, i64, vp, i64, vp, vp])
        d(f"gkob200_dense_compute_norm2_{V}", [vp, i64, i64, vp, i64, vp, vp])
        d(f"gkob200_dense_compute_squared_norm2_{V}", [vp, i64, i64, vp, i64, vp, vp])
        d(f"gkob200_dense_compute_norm1_{V}", [vp, i64, i64, vp, i64, vp, vp])
        d(f"gkob200_dense_compute_sqrt_{V}", [vp, i64, vp])
        for I in ("i32", "i64"):
            d(f"gkob200_dense_row_gather_{V}_{I}", [vp, i64, i64, vp, vp, i64, vp, i64])
        d(f"gkob200_cg_initialize_{V}", [vp, i64, i64, vp, i64, vp, vp, vp, vp, i64, vp, vp, vp])
        d(f"gkob200_cg_step_1_{V}", [vp, i64, i64, vp, vp, i64, vp, vp, vp])
        d(f"gkob200_cg_step_2_{V}", [vp, i64, i64, vp, i64, vp, vp, vp, i64, vp, vp, vp])
        d(f"gkob200_residual_norm_{V}", [vp, i64, vp, vp, T, u8, C.c_int, vp, vp])
        d(f"gkob200_implicit_residual_norm_{V}", [vp, i64, vp, vp, T, u8, C.c_int, vp, vp])
        d(f"gkob200_csr_extract_diagonal_{V}_i32", [vp, i64, i64, vp, vp, vp, vp])
        d(f"gkob200_jacobi_invert_diagonal_{V}", [vp, i64, vp, vp])
        d(f"gkob200_jacobi_simple_scalar_apply_{V}", [vp, i64, i64, vp, vp, i64, vp, i64])
        d(f"gkob200_jacobi_scalar_apply_{V}", [vp, i64, i64, vp, vp, vp, i64, vp, vp, i64])
    d("gkob200_set_all_statuses", [vp, i64, u8, C.c_int, vp])
    # formats
    d("gkob200_coo_spmv_workspace_bytes", [i64, C.c_int], sz)
    for V in VT:
        for I in ("i32", "i64"):
            d(f"gkob200_ell_spmv_{V}_{I}", [vp, i64, i64, i64, i64, vp, vp, vp, i64, i64, vp, vp, vp, i64])
            d(f"gkob200_sellp_spmv_{V}_{I}", [vp, i64, i64, i64, vp, vp, vp, vp, vp, i64, i64, vp, vp, vp, i64])
            d(f"gkob200_coo_spmv_{V}_{I}", [vp, i64, i64, i64, vp, vp, vp, vp, i64, i64, vp, vp, vp, i64, vp, sz])
            d(f"gkob200_coo_spmv2_{V}_{I}", [vp, i64, i64, i64, vp, vp, vp, vp, i64, i64, vp, vp, i64, vp, sz])
            d(f"gkob200_csr_convert_to_ell_{V}_{I}", [vp, i64, vp, vp, vp, i64, i64, vp, vp])
            d(f"gkob200_csr_convert_to_sellp_{V}_{I}", [vp, i64, vp, vp, vp, i64, vp, vp, vp])
            d(f"gkob200_csr_convert_to_hybrid_{V}_{I}", [vp, i64, vp, vp, vp, vp, i64, i64, vp, vp, vp, vp, vp])
    # integer kernels
    d("gkob200_prefix_sum_workspace_bytes", [i64], sz)
    for T in ("i32", "i64", "u64"):
        d(f"gkob200_prefix_sum_{T}", [vp, vp, i64, vp, sz])
    d("gkob200_hybrid_compute_coo_row_ptrs", [vp, vp, i64, u64, vp, vp, sz])
    for I in ("i32", "i64"):
        d(f"gkob200_convert_ptrs_to_idxs_{I}", [vp, vp, i64, vp])
        d(f"gkob200_convert_idxs_to_ptrs_{I}", [vp, vp, i64, i64, vp])
        d(f"gkob200_convert_ptrs_to_sizes_{I}", [vp, vp, i64, vp])
        d(f"gkob200_compute_max_row_nnz_{I}", [vp, vp, i64, vp])
        d(f"gkob200_sellp_compute_slice_sets_{I}", [vp, vp, i64, i64, i64, vp, vp, vp, sz])
        d(f"gkob200_row_len_histogram_{I}", [vp, vp, i64, u64, u64, C.c_int, vp])
    # Krylov step kernels
    for V, T in VT.items():
        d(f"gkob200_bicg_initialize_{V}", [vp, i64, i64, vp, i64] + [vp] * 8 + [i64] + [vp] * 3)
        d(f"gkob200_bicg_step_1_{V}", [vp, i64, i64, vp, vp, vp, vp, i64, vp, vp, vp])
        d(f"gkob200_bicg_step_2_{V}", [vp, i64, i64, vp, i64, vp, vp, vp, vp, vp, i64, vp, vp, vp])
        d(f"gkob200_jacobi_block_transpose_{V}", [vp, i64, vp, vp, i64, i64, C.c_int, vp])
        d(f"gkob200_fcg_initialize_{V}", [vp, i64, i64, vp, i64] + [vp] * 5 + [i64] + [vp] * 4)
        d(f"gkob200_fcg_step_1_{V}", [vp, i64, i64, vp, vp, i64, vp, vp, vp])
        d(f"gkob200_fcg_step_2_{V}", [vp, i64, i64, vp, i64, vp, vp, vp, vp, i64, vp, vp, vp])
        d(f"gkob200_cgs_initialize_{V}", [vp, i64, i64, vp, i64] + [vp] * 8 + [i64] + [vp] * 6)
        d(f"gkob200_cgs_step_1_{V}", [vp, i64, i64, vp, vp, vp, vp, i64, vp, vp, vp, vp])
        d(f"gkob200_cgs_step_2_{V}", [vp, i64, i64, vp, vp, vp, vp, i64, vp, vp, vp, vp])
        d(f"gkob200_cgs_step_3_{V}", [vp, i64, i64, vp, vp, vp, i64, vp, i64, vp, vp])
        d(f"gkob200_bicgstab_initialize_{V}", [vp, i64, i64, vp, i64] + [vp] * 8 + [i64] + [vp] * 7)
        d(f"gkob200_bicgstab_step_1_{V}", [vp, i64, i64, vp, vp, vp, i64, vp, vp, vp, vp, vp])
        d(f"gkob200_bicgstab_step_2_{V}", [vp, i64, i64, vp, vp, vp, i64, vp, vp, vp, vp])
        d(f"gkob200_bicgstab_step_3_{V}", [vp, i64, i64, vp, i64, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp])
        d(f"gkob200_bicgstab_finalize_{V}", [vp, i64, i64, vp, i64, vp, i64, vp, vp])
        d(f"gkob200_gmres_initialize_{V}", [vp, i64, i64, i64, vp, i64, vp, i64, vp, vp, vp])
        d(f"gkob200_gmres_restart_{V}", [vp, i64, i64, vp, i64, vp, vp, vp, vp])
        d(f"gkob200_gmres_multi_axpy_{V}", [vp, i64, i64, vp, vp, vp, i64, vp, vp])
        d(f"gkob200_gmres_hessenberg_qr_{V}", [vp, i64, vp, vp, vp, vp, vp, i64, i64, vp, vp])
        d(f"gkob200_gmres_solve_krylov_{V}", [vp, i64, vp, vp, i64, vp, vp, vp])
        d(f"gkob200_jacobi_block_generate_{V}", [vp, i64, vp, vp, vp, i64, vp, i64, i64, C.c_int, vp])
        d(f"gkob200_jacobi_block_simple_apply_{V}", [vp, i64, vp, vp, i64, i64, C.c_int, i64, i64, vp, i64, vp, i64])
        d(f"gkob200_jacobi_block_apply_{V}", [vp, i64, vp, vp, i64, i64, C.c_int, i64, i64, vp, vp, i64, vp, vp, i64])
    d("gkob200_jacobi_find_blocks_workspace_bytes", [i64], sz)
    d("gkob200_jacobi_find_blocks_i32", [vp, i64, vp, vp, i32, vp, vp, vp, sz])
    # matrix assembly (setup path)
    d("gkob200_setup_sort_workspace_bytes", [i64, C.c_int, C.c_int], sz)
    d("gkob200_setup_compact_workspace_bytes", [i64], sz)
    for V in VT:
        for I in ("i32", "i64"):
            d(f"gkob200_coo_sort_row_major_{V}_{I}", [vp, i64, i64, i64, vp, vp, vp, vp, sz])
            d(f"gkob200_coo_sum_duplicates_{V}_{I}", [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, sz])
            d(f"gkob200_coo_remove_zeros_{V}_{I}", [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, sz])
            d(f"gkob200_csr_transpose_{V}_{I}", [vp, i64, i64, i64, vp, vp, vp, vp, vp, vp, vp, sz])
            d(f"gkob200_csr_sort_by_column_index_{V}_{I}", [vp, i64, i64, i64, vp, vp, vp, vp, sz])
    # MatrixMarket / GINKGO binary files (host)
    d("gkob200_mtx_last_error", [], C.c_char_p)
    d("gkob200_mtx_read_open", [C.c_char_p, C.POINTER(vp), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)])
    d("gkob200_mtx_read_close", [vp])
    for V in VT:
        for I in ("i32", "i64"):
            d(f"gkob200_mtx_read_copy_{V}_{I}", [vp, vp, vp, vp])
            d(f"gkob200_mtx_write_{V}_{I}", [C.c_char_p, C.c_int, C.c_int, i64, i64, i64, vp, vp, vp])
    # generators (host)
    d("gkob200_gen_stencil_nnz", [C.c_int, i64, i64, i64, i64, i64], i64)
    for V in VT:
        for P in ("i32", "i64"):
            d(f"gkob200_gen_stencil_csr_{V}_{P}_{P}", [C.c_int, i64, i64, i64, i64, i64, vp, vp, vp])
    d("gkob200_gen_powerlaw_row_ptrs_i64", [i64, u64, f64, f64, i64, vp], i64)
    d("gkob200_gen_powerlaw_fill_f64_i32", [i64, u64, vp, vp, vp, vp])
    # operator descriptor + solver objects
    MP, PP, SP = C.POINTER(Matrix), C.POINTER(Precond), C.POINTER(Stop)
    d("gkob200_matrix_apply", [vp, MP, vp, i64, i64, vp, vp, vp, i64])
    d("gkob200_solver_create", [C.c_int, MP, PP, SP, i64, i64, C.POINTER(vp)])
    d("gkob200_solver_destroy", [vp])
    d("gkob200_solver_apply", [vp, vp, vp, i64, vp, i64])
    d("gkob200_solver_apply_host", [vp, vp, vp, vp])
    d("gkob200_solver_num_iterations", [vp], i64)
    d("gkob200_solver_stop_status", [vp, vp])
    d("gkob200_solver_residual_history", [vp, vp, i64], i64)
    d("gkob200_solver_launch_count", [vp], i64)
    # distributed
    d("gkob200_nccl_unique_id", [vp])
    d("gkob200_dist_comm_create", [vp, C.c_int, C.c_int, C.POINTER(vp)])
    d("gkob200_dist_comm_destroy", [vp])
    d("gkob200_dist_comm_rank", [vp])
    d("gkob200_dist_comm_size", [vp])
    d("gkob200_dist_comm_uses_p2p", [vp])
    d("gkob200_dist_allreduce_sum_f64", [vp, vp, vp, i64])
    d("gkob200_dist_allreduce_sum_f32", [vp, vp, vp, i64])
    d("gkob200_dist_alltoall_i64", [vp, vp, vp, vp, i64])
    d("gkob200_dist_alltoallv_i32", [vp, vp, vp, vp, vp, vp, vp, vp])
    d("gkob200_partition_build_ranges_from_global_size_i64", [vp, i32, i64, vp])
    d("gkob200_partition_build_from_contiguous_i64", [vp, i32, vp, vp, vp])
    d("gkob200_partition_build_from_mapping_i64", [vp, i64, vp, vp, vp, vp, vp, sz])
    d("gkob200_partition_build_starting_indices_i32_i64", [vp, vp, vp, i64, i32, vp, vp, vp])
    for V in VT:
        d(f"gkob200_dist_build_local_nonlocal_{V}",
          [vp, i64, vp, vp, vp, i64, vp, vp, vp, i64, vp, vp, vp, i64, i32, i32] + [vp] * 10)
        d(f"gkob200_dist_vector_build_local_{V}", [vp, i64, vp, vp, vp, i64, vp, vp, vp, i32, vp, i64])
    d("gkob200_dist_matrix_create", [vp, MP, MP, vp, vp, vp, C.POINTER(vp)])
    d("gkob200_dist_matrix_destroy", [vp])
    d("gkob200_dist_matrix_uses_fused_halo", [vp])
    d("gkob200_dist_matrix_apply", [vp, vp, vp, i64, i64, vp, vp, vp, i64])
    d("gkob200_dist_solver_create", [C.c_int, vp, PP, SP, i64, C.POINTER(vp)])

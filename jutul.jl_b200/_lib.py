"""ctypes binding of libjutul_b200.so (include/jutul_b200.h).

The library is the product; this module only marshals arguments. It fails loudly
when the shared object is missing or when no CUDA device is available — there is
no CPU fallback anywhere in this package.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "csrc", "libjutul_b200.so")

JB_OK, JB_NOT_CONVERGED, JB_BREAKDOWN, JB_BAD_PIVOT, JB_NONFINITE, JB_BAD_SOLVE = 0, 1, 2, 3, 4, 5

P = C.c_void_p
PP = C.POINTER(C.c_void_p)
I32, I64, F64 = C.c_int32, C.c_int64, C.c_double
PI64, PF64, PI32 = C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_int32)

# name -> (restype, argtypes). Mirrors include/jutul_b200.h one to one.
PROTOTYPES = {
    "jb_version": (I32, []),
    "jb_ctx_create": (I32, [I32, PP]),
    "jb_ctx_destroy": (I32, [P]),
    "jb_sync": (I32, [P]),
    "jb_last_error": (C.c_char_p, [P]),
    "jb_launch_count": (I64, [P]),
    "jb_stream": (P, [P]),
    "jb_timer_start": (I32, [P]),
    "jb_timer_stop": (I32, [P, PF64]),
    "jb_prof_enable": (I32, [P, I32]),
    "jb_prof_collect": (I32, [P, PF64, PI64]),
    "jb_malloc": (I32, [P, I64, PP]),
    "jb_free": (I32, [P, P]),
    "jb_h2d": (I32, [P, P, P, I64]),
    "jb_d2h": (I32, [P, P, P, I64]),
    "jb_d2d": (I32, [P, P, P, I64]),
    "jb_pinned_alloc": (I32, [I64, PP]),
    "jb_pinned_free": (I32, [P]),
    "jb_mesh_create": (I32, [P, I64, I64, PI64, PP]),
    "jb_mesh_destroy": (I32, [P]),
    "jb_mesh_halfface": (I32, [P, PI64, PI64, PI64, PI64]),
    "jb_csr_create_from_coo": (I32, [P, PI64, PI64, I64, I64, I32, PP]),
    "jb_csr_create_tpfa": (I32, [P, I32, PP]),
    "jb_csr_destroy": (I32, [P]),
    "jb_csr_nnz": (I64, [P]),
    "jb_csr_nrows": (I64, [P]),
    "jb_csr_get": (I32, [P, PI64, PI64]),
    "jb_csr_values_get": (I32, [P, PF64]),
    "jb_csr_values_set": (I32, [P, PF64]),
    "jb_csr_values_ptr": (P, [P]),
    "jb_csr_values_modified": (I32, [P]),
    "jb_csr_create_transpose": (I32, [P, PP]),
    "jb_csr_transpose_update": (I32, [P]),
    "jb_tpfa_create": (I32, [P, P, PP]),
    "jb_tpfa_destroy": (I32, [P]),
    "jb_tpfa_positions": (I32, [P, PI64, PI64]),
    "jb_twophase_create": (I32, [P, PF64, PF64, PF64, PF64, PP]),
    "jb_twophase_destroy": (I32, [P]),
    "jb_twophase_set_sources": (I32, [P, I64, PI64, PF64]),
    "jb_twophase_update_state": (I32, [P, P, P]),
    "jb_twophase_mass": (I32, [P, P, P, P]),
    "jb_twophase_assemble": (I32, [P, P, P, P, F64, P]),
    "jb_twophase_assemble_faces": (I32, [P, P, P, P, F64, P]),
    "jb_twophase_residual": (I32, [P, P, P, P, F64, P]),
    "jb_heat_pattern": (I32, [P, I64, I64, PP]),
    "jb_heat_assemble": (I32, [P, I64, I64, F64, F64, F64, P, P, P]),
    "jb_poisson_assemble": (I32, [P, P, P, P, I32, F64, I64, PI64, PF64, P]),
    "jb_generic_create": (I32, [P, I32, I32, I64, PI64, PI64, PI64, PP]),
    "jb_generic_destroy": (I32, [P]),
    "jb_generic_fill": (I32, [P, PF64, P, I64]),
    "jb_nfvm_create": (I32, [P, I64, I64, I32, PI64, PI64, PF64, PF64, PI64, PI64, PF64, PF64, PF64, PI64, PI64, PF64, PP]),
    "jb_nfvm_destroy": (I32, [P]),
    "jb_nfvm_evaluate_flux": (I32, [P, P, I64, I64, P]),
    "jb_nfvm_stencil": (I32, [P, PI64, PI64, I64, PI64]),
    "jb_nfvm_pattern": (I32, [P, PP]),
    "jb_nfvm_align": (I32, [P, P]),
    "jb_nfvm_positions": (I32, [P, PI64, PI64]),
    "jb_nfvm_assemble": (I32, [P, P, I64, I64, P, P, P, P]),
    "jb_table_create_1d": (I32, [P, I64, PF64, PF64, I32, PP]),
    "jb_table_create_2d": (I32, [P, I64, I64, PF64, PF64, PF64, I32, I32, PP]),
    "jb_table_destroy": (I32, [P]),
    "jb_table_info": (I32, [P, PI64]),
    "jb_table_eval": (I32, [P, I64, P, P, P, P, P]),
    "jb_varprog_create": (I32, [P, I64, I32, P, I32, PP, PP]),
    "jb_varprog_destroy": (I32, [P]),
    "jb_varprog_order": (I32, [P, PI64, PI64]),
    "jb_varprog_evaluate": (I32, [P, PP, PP]),
    "jb_twophase_assemble_props": (I32, [P, P, PP, P, F64, P]),
    "jb_schur_create": (I32, [P, I32, PI64, PI64, PI64, PI64, PI64, PI64, PI64, PI64, PI64, PI64, PP]),
    "jb_schur_destroy": (I32, [P]),
    "jb_schur_size": (I64, [P]),
    "jb_schur_update": (I32, [P, PF64, PF64, PF64]),
    "jb_schur_prepare": (I32, [P, P, PF64]),
    "jb_schur_mul": (I32, [P, F64, P, F64, P]),
    "jb_krylov_set_schur": (I32, [P, P]),
    "jb_schur_dx_update": (I32, [P, P, PF64]),
    "jb_unit_diagonalize_ghosts": (I32, [P, P, I64]),
    "jb_spmv": (I32, [P, F64, P, F64, P]),
    "jb_ilu0_create": (I32, [P, PI64, PP]),
    "jb_ilu0_destroy": (I32, [P]),
    "jb_ilu0_update": (I32, [P]),
    "jb_ilu0_apply": (I32, [P, P, P]),
    "jb_ilu0_info": (I32, [P, PI64]),
    "jb_ilu0_get": (I32, [P, PI64, PI64, PF64, PI64, PI64, PF64, PF64]),
    "jb_diag_precond_create": (I32, [P, I32, F64, PP]),
    "jb_krylov_create": (I32, [P, P, I32, PP]),
    "jb_krylov_destroy": (I32, [P]),
    "jb_krylov_solve": (I32, [P, P, P, F64, F64, I32, I32, I32, PI32, PF64, I32]),
    "jb_krylov_info": (I32, [P, PI64]),
    "jb_krylov_phase_times": (I32, [P, PF64]),
    "jb_krylov_set_gmres": (I32, [P, I32, I32, I32]),
    "jb_scale_system": (I32, [P, P, I32, F64]),
    "jb_update_scalar": (I32, [P, P, P, I64, I64, F64, F64, F64, F64, F64, F64]),
    "jb_update_fraction_pair": (I32, [P, P, P, I64, I64, F64, F64, F64, F64]),
    "jb_copy_strided": (I32, [P, P, I64, P, I64, I64]),
    "jb_update_fractions": (I32, [P, P, P, I64, I32, I64, F64, F64, F64, F64, I32]),
    "jb_increment_norm": (I32, [P, P, I64, I64, PF64, PF64]),
    "jb_maxabs_rows": (I32, [P, P, I32, I64, PF64]),
    "jb_partition_metis": (I32, [I64, I64, PI64, PF64, I64, PI64]),
    "jb_partition_linear": (I32, [I64, I64, PI64]),
    "jb_process_partition": (I32, [I64, I64, PI64, PI64, PF64, PI64]),
    "jb_order_multicolor": (I32, [I64, I64, PI64, PI64, PI64]),
    "jb_order_multicolor_boundary_last": (I32, [I64, I64, PI64, PI64, PI64, PI64]),
    "jb_perm_create": (I32, [P, PI64, I64, PP]),
    "jb_perm_destroy": (I32, [P]),
    "jb_perm_apply": (I32, [P, P, P, I32, I32]),
    "jb_twophase_set_permutation": (I32, [P, P]),
    "jb_nccl_unique_id": (I32, [C.c_char_p]),
    "jb_comm_create": (I32, [P, I32, I32, C.c_char_p, PP]),
    "jb_comm_destroy": (I32, [P]),
    "jb_dist_create": (I32, [P, I64, I64, I32, PI32, PI64, PI64, PI64, PP]),
    "jb_dist_destroy": (I32, [P]),
    "jb_dist_p2p_export": (I32, [P, C.c_char_p]),
    "jb_dist_p2p_open": (I32, [P, C.c_char_p, PI64, PI64, PI64, PI64]),
    "jb_dist_p2p_status": (I32, [P]),
    "jb_dist_p2p_disable": (I32, [P]),
    "jb_dist_halo_exchange": (I32, [P, P, I32]),
    "jb_dist_allreduce": (I32, [P, PF64, I32, I32]),
    "jb_krylov_set_dist": (I32, [P, P]),
    "jb_twophase_set_owned": (I32, [P, I64]),
    "jb_twophase_update_after_step": (I32, [P]),
    "jb_twophase_perform_step_host": (I32, [P, P, P, PF64, PF64, PF64, F64, F64, F64, F64, I32, F64, F64, PF64, PI32, PI32]),
}


class JutulB200Error(RuntimeError):
    pass


_lib = None


def load():
    """Load libjutul_b200.so. Raises if it has not been built (run `python -c 'import
    __graft_entry__ as g; g.build()'` or `make -C jutul.jl_b200/csrc`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise JutulB200Error(
                f"{SO_PATH} not found: the CUDA extension is not built and there is no CPU fallback")
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in PROTOTYPES.items():
            f = getattr(lib, name)
            f.restype = res
            f.argtypes = args
        _lib = lib
    return _lib


def last_error(ctx=None):
    msg = load().jb_last_error(ctx)
    return msg.decode() if msg else ""


def check(rc, ctx=None, what=""):
    """Hard errors (<0) raise; numerical conditions (>0) are returned to the caller."""
    if rc < 0:
        raise JutulB200Error(f"{what} failed with code {rc}: {last_error(ctx) or last_error(None)}")
    return rc

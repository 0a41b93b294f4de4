"""ctypes binding of libsgnn_b200.so (include/sgnn_b200.h).

The product path has NO fallback: if the shared library is missing or a symbol is absent this
module raises at import time, and every non-zero return code becomes a RuntimeError
(scn raises Python exceptions from C++ asserts; the reference caller catches them at
test_scene.py:83).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libsgnn_b200.so')

SGNN_F32, SGNN_BF16 = 0, 1
CONV_TILE = 128


class SgnnGrid(C.Structure):
    _fields_ = [('nb', C.c_int32), ('d0', C.c_int32), ('d1', C.c_int32), ('d2', C.c_int32),
                ('wx', C.c_int32), ('reserved', C.c_int32), ('n_words', C.c_int64),
                ('mask', C.c_void_p), ('prefix', C.c_void_p), ('row_of_rank', C.c_void_p)]


class SgnnEpilogue(C.Structure):
    _fields_ = [('out', C.c_void_p), ('ld', C.c_int32), ('relu', C.c_int32),
                ('scale', C.c_void_p), ('shift', C.c_void_p)]


class SgnnConvArgs(C.Structure):
    _fields_ = [('in_', C.c_void_p), ('ld_in', C.c_int32), ('dtype', C.c_int32),
                ('nbr', C.c_void_p), ('nbr_stride', C.c_int64), ('K', C.c_int32),
                ('child_mode', C.c_int32), ('weight', C.c_void_p), ('cin', C.c_int32),
                ('cout', C.c_int32), ('n_out', C.c_int64), ('residual', C.c_void_p),
                ('ld_res', C.c_int32), ('n_in', C.c_int32),
                ('a', SgnnEpilogue), ('b', SgnnEpilogue), ('flags', C.c_int32), ('reserved', C.c_int32)]


class SgnnBnFold(C.Structure):
    _fields_ = [('scale', C.c_void_p), ('shift', C.c_void_p)]


class SgnnResBlockW(C.Structure):
    _fields_ = [('bn0', SgnnBnFold), ('w0', C.c_void_p), ('bn1', SgnnBnFold), ('w1', C.c_void_p)]


class SgnnEncLevelW(C.Structure):
    _fields_ = [('cin', C.c_int32), ('c', C.c_int32), ('w_in', C.c_void_p), ('res', SgnnResBlockW),
                ('bn_out', SgnnBnFold), ('w_down', C.c_void_p), ('bn_down', SgnnBnFold)]


class SgnnFcnW(C.Structure):
    _fields_ = [('c', C.c_int32), ('reserved', C.c_int32), ('blk', SgnnResBlockW * 3), ('bn_down', SgnnBnFold * 2),
                ('w_down', C.c_void_p * 2), ('bn_join', SgnnBnFold)]


class SgnnDenseLayerW(C.Structure):
    _fields_ = [('w', C.c_void_p), ('bn', SgnnBnFold), ('cout', C.c_int32), ('ksize', C.c_int32),
                ('stride', C.c_int32), ('pad', C.c_int32), ('transposed', C.c_int32), ('cat_with', C.c_int32)]


class SgnnRefineW(C.Structure):
    _fields_ = [('cin', C.c_int32), ('c', C.c_int32), ('w_in', C.c_void_p), ('fcn', SgnnFcnW),
                ('w_up', C.c_void_p), ('bn_up', SgnnBnFold), ('w_occ', C.c_void_p), ('b_occ', C.c_void_p),
                ('w_sdf', C.c_void_p), ('b_sdf', C.c_void_p)]


class SgnnSurfaceW(C.Structure):
    _fields_ = [('cin', C.c_int32), ('c', C.c_int32), ('w_in', C.c_void_p), ('fcn', SgnnFcnW),
                ('w_lin', C.c_void_p), ('b_lin', C.c_void_p)]


class SgnnGeneratorW(C.Structure):
    _fields_ = [('enc', SgnnEncLevelW * 3), ('dense', SgnnDenseLayerW * 6), ('w_heads', C.c_void_p),
                ('nf_coarse', C.c_int32), ('reserved', C.c_int32), ('ref', SgnnRefineW * 3), ('surf', SgnnSurfaceW),
                ('prepared', C.c_void_p), ('prepared_bytes', C.c_size_t), ('tc32_min_rows', C.c_int64),
                ('ur_min_rows', C.c_int64), ('overlap_max_rows', C.c_int64)]


class SgnnGeneratorOut(C.Structure):
    _fields_ = [('n_out', C.c_int64), ('out_locs', C.c_void_p), ('out_sdf', C.c_void_p),
                ('n_cand', C.c_int64 * 4), ('cand_locs', C.c_void_p * 4), ('cand', C.c_void_p * 4),
                ('rows', C.c_int64 * 16), ('arena_used', C.c_size_t), ('arena_needed', C.c_size_t),
                ('conv_ms', C.c_double), ('n_conv', C.c_int64)]


class SgnnExportSeg(C.Structure):
    _fields_ = [('src', C.c_void_p), ('dst', C.c_void_p), ('n', C.c_int64), ('kind', C.c_int32), ('to_i64', C.c_int32),
                ('aux', C.c_int32 * 4)]


EXPORT_COPY32, EXPORT_COORDS, EXPORT_CHILDREN, EXPORT_DENSE_CELLS = 0, 1, 2, 3
GEN_CAND_LOCS = 1
GEN_CAND_PARENTS = 32
GEN_PROFILE = 2
GEN_TC32 = 4
GEN_DENSE_RULES = 8
GEN_PHASES = 16

_P, _I, _L, _Z = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
_G, _E = C.POINTER(SgnnGrid), C.POINTER(SgnnEpilogue)

# name -> (restype, argtypes); mirrors include/sgnn_b200.h one to one
SIGNATURES = {
    'sgnn_scan_scratch_bytes': (_Z, [_L]),
    'sgnn_compact_scratch_bytes': (_Z, [_L]),
    'sgnn_grid_build': (_I, [_G, _P, _I, _L, _P, _P, _P, _Z, _P]),
    'sgnn_grid_coarsen': (_I, [_G, _G, _P, _Z, _P]),
    'sgnn_grid_enumerate': (_I, [_G, _P, _P]),
    'sgnn_grid_lookup': (_I, [_G, _P, _L, _I, _P, _P]),
    'sgnn_rulebook_submanifold': (_I, [_G, _P, _L, _P, _P]),
    'sgnn_rulebook_strided': (_I, [_G, _P, _L, _P, _P, _L, _P]),
    'sgnn_grid_coarse_build': (_I, [_G, _G, _L, _L, _P, _P, _P, _P]),
    'sgnn_conv_forward': (_I, [C.POINTER(SgnnConvArgs), _P]),
    'sgnn_conv_forward_compact': (_I, [C.POINTER(SgnnConvArgs), _P, _P, _P]),
    'sgnn_rulebook_submanifold_plan': (_I, [_G, _P, _L, _P, _P, _Z, _P]),
    'sgnn_rulebook_submanifold_compact': (_I, [_G, _P, _L, _P, _P, _P]),
    'sgnn_conv_tc32_workspace_bytes': (_Z, [_I, _I, _I]),
    'sgnn_conv_forward_tc32': (_I, [C.POINTER(SgnnConvArgs), _P, _Z, _P]),
    'sgnn_conv_tc32_prepare': (_I, [_P, _I, _I, _I, _I, _P, _Z, _P]),
    'sgnn_generator_prepared_bytes': (_Z, [C.POINTER(SgnnGeneratorW)]),
    'sgnn_generator_prepare': (_I, [C.POINTER(SgnnGeneratorW), _P]),
    'sgnn_tile_plan_bytes': (_Z, [_L]),
    'sgnn_tile_plan_build': (_I, [_P, _L, _L, _P, _Z, _P]),
    'sgnn_conv_forward_tc32_ur': (_I, [C.POINTER(SgnnConvArgs), _P, _P, _Z, _P]),
    'sgnn_conv_urc_prepare': (_I, [_P, _I, _P, _Z, _P]),
    'sgnn_conv_forward_tc32_urc': (_I, [C.POINTER(SgnnConvArgs), _P, _P, _Z, _P]),
    'sgnn_deconv_forward': (_I, [_P, _I, _I, _P, _P, _I, _I, _L, _E, _P]),
    'sgnn_unpool': (_I, [_P, _I, _P, _I, _L, _E, _P]),
    'sgnn_affine_relu': (_I, [_P, _I, _P, _I, _L, _I, _P, _P, _I, _P]),
    'sgnn_add_rows': (_I, [_P, _I, _P, _I, _P, _I, _L, _I, _P]),
    'sgnn_copy_cols': (_I, [_P, _I, _P, _I, _L, _I, _P]),
    'sgnn_linear': (_I, [_P, _I, _P, _P, _P, _I, _L, _I, _I, _P]),
    'sgnn_sparse_to_dense': (_I, [_P, _I, _P, _L, _I, _P, _I, _I, _I, _I, _P]),
    'sgnn_dense_conv3d': (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _P, _P, _I, _P, _P]),
    'sgnn_dense_convT3d': (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _P, _P, _I, _P, _P]),
    'sgnn_dense_to_sparse': (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _P, _Z, _P]),
    'sgnn_heads_compact': (_I, [_P, _I, _I, _P, _P, _P, _P, _P, _L, _P, _P, _P, _I, _P, _P, _Z, _P]),
    'sgnn_heads_flags': (_I, [_P, _I, _I, _P, _P, _P, _P, _L, _P, _P, _P, _P, _Z, _P]),
    'sgnn_heads_write': (_I, [_P, _I, _I, _P, _P, _L, _P, _P, _P, _P, _I, _P]),
    'sgnn_heads_write_join': (_I, [_P, _I, _I, _P, _P, _L, _P, _P, _P, _P, _I, _G, _P, _I, _I, _P]),
    'sgnn_dense_flags': (_I, [_P, _I, _L, _P, _P, _P, _P, _Z, _P]),
    'sgnn_dense_write': (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P]),
    'sgnn_generator_forward': (_I, [C.POINTER(SgnnGeneratorW), _P, _I, _P, _L, _I, C.POINTER(C.c_int32), _P, _Z, _I,
                                    C.POINTER(SgnnGeneratorOut), _P]),
    'sgnn_generator_profile_entry': (_I, [_I, C.POINTER(C.c_int64), C.POINTER(C.c_float)]),
    'sgnn_generator_phase_entry': (_I, [_I, C.c_char_p, C.POINTER(C.c_float)]),
    'sgnn_mc_scratch_bytes': (_Z, [_I, _I, _I]),
    'sgnn_mc_count': (_I, [_P, _I, _I, _I, C.c_float, C.c_float, C.c_float, _P, _P, _Z, _P]),
    'sgnn_mc_emit': (_I, [_P, _I, _I, _I, C.c_float, C.c_float, C.c_float, _P, _P, _P]),
    'sgnn_mc_merge_host': (_I, [_P, _L, _P, _P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    'sgnn_mc_merge_host_src': (_I, [_P, _L, _P, _P, _P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    'sgnn_mc_tri_cells': (_I, [_P, _L, _P, _P]),
    'sgnn_mc_table': (_I, [_P]),
    'sgnn_children_coords': (_I, [_P, _L, _P, _P]),
    'sgnn_export': (_I, [C.POINTER(SgnnExportSeg), _I, _P]),
    'sgnn_concat_skip': (_I, [_G, _P, _I, _I, _P, _L, _P, _I, _I, _P]),
    'sgnn_coords_to_i64': (_I, [_P, _L, _P, _P]),
    'sgnn_debug_ur_diag': (_I, [_P]),
    'sgnn_debug_ffma_peak': (_I, [_I, C.POINTER(C.c_double), _P]),
    'sgnn_launch_count': (_L, []),
    'sgnn_version': (_I, []),
    'sgnn_error_string': (C.c_char_p, [_I]),
    'sgnn_last_cuda_error': (_I, []),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            'sgnn_b200: %s not found -- run `python -c "import __graft_entry__ as g; g.build()"` '
            '(there is no CPU or PyTorch fallback for the hot path)' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class SgnnError(RuntimeError):
    pass


def check(rc, what=''):
    if rc != 0:
        msg = lib.sgnn_error_string(rc).decode()
        if rc == -2:
            msg += ' (cudaError %d)' % lib.sgnn_last_cuda_error()
        raise SgnnError('%s failed: %s [%d]' % (what or 'sgnn call', msg, rc))

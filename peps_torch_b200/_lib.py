"""ctypes binding of libctmb.so (include/ctmb.h).  There is NO fallback: if the CUDA library
is missing or fails to load, importing this module raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libctmb.so')

if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} not found: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                      "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
                      "peps_torch_b200 has no CPU fallback.")
lib = C.CDLL(LIB_PATH)

F64, C128 = 0, 1
UP, LEFT, DOWN, RIGHT = 0, 1, 2, 3
LU, RU, RD, LD = 0, 1, 2, 3


class Options(C.Structure):
    _fields_ = [('svd_reltol', C.c_double), ('eps_multiplet', C.c_double), ('multiplet_abstol', C.c_double),
                ('rsvd_rank_factor', C.c_double), ('rsvd_niter', C.c_int), ('jacobi_max_sweeps', C.c_int),
                ('norm_type', C.c_int), ('rsvd_max_rounds', C.c_int), ('seed', C.c_ulonglong), ('rsvd_tol', C.c_double),
                ('projector_method', C.c_int), ('rsvd_stateless', C.c_int)]


class Site(C.Structure):
    _fields_ = [('a', C.c_void_p), ('dims', C.c_int * 5), ('pad', C.c_int),
                ('C', C.c_void_p * 4), ('T', C.c_void_p * 4)]


_vp, _i, _sz = C.c_void_p, C.c_int, C.c_size_t
_PS, _PO = C.POINTER(Site), C.POINTER(Options)
_PI, _PVP = C.POINTER(C.c_int), C.POINTER(C.c_void_p)

ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)

# every symbol include/ctmb.h declares, with its signature
SIGNATURES = {
    'ctmb_set_group': (C.c_int, [_vp, _i, _i, ALLGATHER_FN, _vp]),
    'ctmb_version': (C.c_int, []),
    'ctmb_last_error': (C.c_char_p, []),
    'ctmb_create': (C.c_int, [C.POINTER(_vp), _i]),
    'ctmb_destroy': (C.c_int, [_vp]),
    'ctmb_default_options': (None, [_PO]),
    'ctmb_get_rsvd_status': (C.c_int, [_vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_double),
                             C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), _i]),
    'ctmb_get_counters': (C.c_int, [_vp, C.POINTER(C.c_longlong), C.POINTER(C.c_double)]),
    'ctmb_reset_counters': (C.c_int, [_vp]),
    'ctmb_profile_enable': (C.c_int, [_vp, _i]),
    'ctmb_profile_get': (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                   C.POINTER(C.c_longlong)]),
    'ctmb_einsum2': (C.c_int, [_vp, _i, C.c_char_p, _vp, C.POINTER(C.c_longlong), _i, _vp,
                               C.POINTER(C.c_longlong), _i, _vp, _vp]),
    'ctmb_c2x2': (C.c_int, [_vp, _i, _i, _i, _PS, _vp, _vp, _sz, _vp]),
    'ctmb_c2x2_workspace': (_sz, [_vp, _i, _i, _i, _PS]),
    'ctmb_halves': (C.c_int, [_vp, _i, _i, _i, C.POINTER(_PS), _vp, _vp, _vp, _sz, _vp]),
    'ctmb_halves_workspace': (_sz, [_vp, _i, _i, _i, C.POINTER(_PS)]),
    'ctmb_projectors': (C.c_int, [_vp, _i, _vp, _vp, _i, _i, _i, _PO, _vp, _vp, _vp, _vp, _sz, _vp]),
    'ctmb_projectors_workspace': (_sz, [_vp, _i, _i, _i, _i, _PO]),
    'ctmb_truncated_svd': (C.c_int, [_vp, _i, _vp, _i, _i, _i, _PO, _vp, _vp, _vp, _vp, _sz, _vp]),
    'ctmb_truncated_svd_workspace': (_sz, [_vp, _i, _i, _i, _i, _PO]),
    'ctmb_truncated_eig_sym': (C.c_int, [_vp, _i, _vp, _i, _i, _PO, _vp, _vp, _vp, _sz, _vp]),
    'ctmb_truncated_eig_sym_workspace': (_sz, [_vp, _i, _i, _i, _PO]),
    'ctmb_qr': (C.c_int, [_vp, _i, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    'ctmb_qr_workspace': (_sz, [_vp, _i, _i, _i]),
    'ctmb_move_generic': (C.c_int, [_vp, _i, _i, _i, _i, _PS, _PI, _PI, _PO, _PVP, _PVP, _PVP, _vp, _sz, _vp]),
    'ctmb_move_generic_workspace': (_sz, [_vp, _i, _i, _i, _i, _PS, _PI, _PI, _PO]),
    'ctmb_move_generic_projectors': (C.c_int, [_vp, _i, _i, _i, _i, _PS, _PI, _i, _PI, _PO, _PVP, _PVP, _vp, _sz, _vp]),
    'ctmb_move_generic_absorb': (C.c_int, [_vp, _i, _i, _i, _i, _PS, _PI, _i, _PI, _PO, _PVP, _PVP, _PVP, _PVP, _PVP,
                                           _vp, _sz, _vp]),
    'ctmb_move_c4v': (C.c_int, [_vp, _i, _vp, C.POINTER(C.c_int), _vp, _vp, _i, _PO, _vp, _vp, _vp, _vp, _sz, _vp]),
    'ctmb_move_c4v_workspace': (_sz, [_vp, _i, C.POINTER(C.c_int), _i, _PO]),
    'ctmb_rdm2x2': (C.c_int, [_vp, _i, _i, C.POINTER(_PS), _i, _vp, _vp, _sz, _vp]),
    'ctmb_rdm2x2_workspace': (_sz, [_vp, _i, _i, C.POINTER(_PS), _i]),
    'ctmb_rdm_small': (C.c_int, [_vp, _i, _i, _i, C.POINTER(_PS), _vp, _vp, _sz, _vp]),
    'ctmb_rdm_small_workspace': (_sz, [_vp, _i, _i, _i, C.POINTER(_PS)]),
    'ctmb_sym_pos_def': (C.c_int, [_vp, _i, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    'ctmb_sym_pos_def_workspace': (_sz, [_vp, _i, _i, _i]),
}
for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)          # AttributeError if the library does not export it
    _f.restype = _res
    _f.argtypes = _args


class CtmbError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise CtmbError(lib.ctmb_last_error().decode())


def default_options():
    o = Options()
    lib.ctmb_default_options(C.byref(o))
    return o

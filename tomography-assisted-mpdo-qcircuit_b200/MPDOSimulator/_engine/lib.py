"""ctypes binding of libmpdo_b200.so (C ABI declared in include/mpdo_b200.h).

There is no CPU fallback: if the shared library is missing, or it is asked to run without a CUDA
device, the caller gets an exception (north star: "no Triton and no CPU fallback").
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.normpath(os.path.join(_HERE, '..', '..', 'lib', 'libmpdo_b200.so'))

MPDO_C64, MPDO_C128 = 0, 1


class IdxMap(C.Structure):
    _fields_ = [('d0', C.c_int32), ('d1', C.c_int32), ('s0', C.c_int64), ('s1', C.c_int64), ('s2', C.c_int64)]


class ContractDesc(C.Structure):
    _fields_ = [
        ('M', C.c_int32), ('N', C.c_int32), ('K', C.c_int32), ('batch', C.c_int32),
        ('dtypeA', C.c_int32), ('dtypeB', C.c_int32), ('dtypeC', C.c_int32),
        ('conjA', C.c_int32), ('conjB', C.c_int32), ('acc64', C.c_int32),
        ('a_kfast', C.c_int32), ('b_jfast', C.c_int32), ('ksplit', C.c_int32), ('hermitian', C.c_int32),
        ('alpha', C.c_double), ('beta', C.c_double),
        ('Ab', IdxMap), ('Ai', IdxMap), ('Ak', IdxMap),
        ('Bb', IdxMap), ('Bk', IdxMap), ('Bj', IdxMap),
        ('Cb', IdxMap), ('Ci', IdxMap), ('Cj', IdxMap),
    ]


# name -> (restype, argtypes); every symbol declared in include/mpdo_b200.h
SYMBOLS = {
    'mpdo_contract': (C.c_int, [C.POINTER(ContractDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'mpdo_absorb_1q': (C.c_int, [C.c_int] * 6 + [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    'mpdo_jacobi_rows': (C.c_int, [C.c_int] * 5 + [C.c_int64, C.c_void_p, C.c_double, C.c_int, C.c_void_p,
                                                  C.c_void_p]),
    'mpdo_rows_finalize': (C.c_int, [C.c_int] * 5 + [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                    C.c_int, C.c_double, C.c_void_p]),
    'mpdo_decompose_rows': (C.c_int, [C.c_int] * 3 + [C.c_void_p] * 6 + [C.c_int, C.c_double, C.c_double, C.c_int,
                                                                    C.c_void_p]),
    'mpdo_eigh_psd_scratch_bytes': (C.c_int64, [C.c_int, C.c_int]),
    'mpdo_eigh_psd': (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                C.c_double, C.c_int, C.c_void_p]),
    'mpdo_chol_psd_scratch_bytes': (C.c_int64, [C.c_int, C.c_int]),
    'mpdo_chol_psd': (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_double, C.c_void_p]),
    'mpdo_rowscale': (C.c_int, [C.c_int] * 4 + [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int,
                                               C.c_int, C.c_void_p, C.c_void_p]),
    'mpdo_rank_rule': (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                 C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    'mpdo_qr_step': (C.c_int, [C.c_int] * 6 + [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p]),
    'mpdo_bond_svd_step': (C.c_int, [C.c_int] * 6 + [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_double,
                                                    C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p,
                                                    C.c_void_p]),
    'mpdo_env_sweep': (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]),
    'mpdo_bond_env_step': (C.c_int, [C.c_int] * 5 + [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'mpdo_kappa_truncate': (C.c_int, [C.c_int] * 5 + [C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_int), C.c_void_p,
                                                     C.c_void_p, C.c_void_p]),
    'mpdo_split_2q': (C.c_int, [C.c_int] * 6 + [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                               C.c_double, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                               C.c_void_p]),
    'mpdo_cast': (C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    'mpdo_tc_enable': (C.c_int, [C.c_int]),
    'mpdo_trim_pools': (C.c_int, []),
    'mpdo_pool_stats': (C.c_int, [C.POINTER(C.c_int64)] * 2),
    'mpdo_timing_enable': (C.c_int, [C.c_int]),
    'mpdo_timing_summary': (C.c_int, [C.c_int, C.c_double] + [C.POINTER(C.c_double)] * 3 + [C.POINTER(C.c_int64)] +
                            [C.POINTER(C.c_double)] * 2),
    'mpdo_debug_barrier_timeout': (C.c_int, [C.c_void_p]),
    'mpdo_version': (C.c_int, []),
    'mpdo_last_error': (C.c_char_p, []),
    'mpdo_device_info': (C.c_int, [C.POINTER(C.c_int)] * 4),
    'mpdo_launch_count': (C.c_int64, []),
}

_lib = None


class MpdoLibraryError(RuntimeError):
    pass


def load():
    """Load libmpdo_b200.so and bind every exported symbol. Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MpdoLibraryError(
            f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            f'(nvcc, sm_100a). There is no CPU fallback for the MPDO update path.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().mpdo_last_error()
        raise RuntimeError(f'{what} failed (rc={rc}): {msg.decode() if msg else ""}')

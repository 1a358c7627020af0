"""ctypes binding of ``libsober_b200.so`` (the C ABI declared in ``include/sober_b200.h``).

There is NO fallback: if the shared library is missing or does not export a declared symbol, importing the
ops raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C sober_b200/csrc``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SOBER_B200_LIB: alternative build of the same ABI (kernel-variant experiments, see tools/k1_variants.sh)
LIB_PATH = os.environ.get("SOBER_B200_LIB") or os.path.join(_HERE, "libsober_b200.so")

OK = 0
ABI_VERSION = 2   # SOBER_B200_ABI_VERSION of include/sober_b200.h
_STATUS = {1: "invalid argument", 2: "CUDA error", 3: "unsupported shape/family", 4: "workspace too small"}

RBF, MATERN12, MATERN32, MATERN52, TANIMOTO, TANIMOTO_BITS, HAMMING_LUT = range(7)
# constant folded into the lengthscale so that the kernels see  RBF = exp(-d2), Matern = f(r), r = sqrt(d2)
FAMILY_SCALE = {RBF: 0.5 ** 0.5, MATERN12: 1.0, MATERN32: 3.0 ** 0.5, MATERN52: 5.0 ** 0.5, TANIMOTO: 1.0}
BITS_MAX_D = 2048  # the popcount Tanimoto kernel covers fingerprints of up to 2048 bits
RECORD_MAX_D = 8   # the register kernel of K1 (record layout) covers d <= 8


class GroupArgs(C.Structure):
    """``sober_group_args`` of include/sober_b200.h -- field order and types must match."""
    _fields_ = [
        ("X", C.c_void_p), ("ldx", C.c_int64),
        ("xn", C.c_void_p), ("xn_stride", C.c_int64),
        ("idx", C.c_void_p), ("mu", C.c_void_p),
        ("n_local", C.c_int64), ("pos0", C.c_int64), ("n_global", C.c_int64), ("ES", C.c_int64),
        ("S", C.c_int32), ("L", C.c_int32), ("d", C.c_int32), ("family", C.c_int32),
        ("outputscale", C.c_double),
        ("Zt", C.c_void_p), ("zn", C.c_void_p),
        ("At", C.c_void_p), ("totw", C.c_void_p),
        ("variant", C.c_int32), ("unit_weights", C.c_int32),
        ("rec", C.c_void_p), ("ldr", C.c_int64),
        ("lut", C.c_void_p),
        ("kx", C.c_void_p), ("ldkx", C.c_int64), ("aw", C.c_void_p), ("n_obs", C.c_int32), ("transform", C.c_int32),
    ]


_P, _I64, _I32, _D = C.c_void_p, C.c_int64, C.c_int32, C.c_double

# name -> (restype, argtypes); the single source of truth for tests/test_abi.py as well
PROTOTYPES = {
    "sober_abi_version": (C.c_int, []),
    "sober_last_cuda_error": (C.c_char_p, []),
    "sober_sm_count": (C.c_int, [C.POINTER(C.c_int)]),
    "sober_prepare_points": (C.c_int, [_P, _I64, _I64, _I32, _P, _P, _P, _I64, _P]),
    "sober_make_records": (C.c_int, [_P, _I64, _I32, _P, _P, _P, _P, _I64, _P, _I64, _P]),
    "sober_pack_bits": (C.c_int, [_P, _I64, _I64, _I32, _P, _I32, _P, _P, _P]),
    "sober_row_sqnorm": (C.c_int, [_P, _I64, _I64, _I32, _P, _P]),
    "sober_compact_workspace": (_I64, [_I64]),
    "sober_compact_nonzero": (C.c_int, [_P, _I64, _P, _P, _P, _P, _I64, _P]),
    "sober_group_accumulate_workspace": (_I64, [C.POINTER(GroupArgs)]),
    "sober_group_accumulate": (C.c_int, [C.POINTER(GroupArgs), _P, _I64, _P]),
    "sober_group_accumulate_gram": (C.c_int, [_P, _I64, _I32, _I64, _P, _I64, _I64, _I32, _P, _P, _P]),
    "sober_car_workspace": (_I64, [_I32]),
    "sober_car_eliminate": (C.c_int, [_P, _I32, _I32, _P, _I32, _P, _P, _P, _I64, _P]),
    "sober_car_cluster_fits": (C.c_int, [_I32, _I32, _I32]),
    "sober_car_cluster": (C.c_int, [_P, _P, _I32, _I32, _P, _I32, _P, _P]),
    "sober_car_cluster_profiled": (C.c_int, [_P, _P, _I32, _I32, _P, _I32, _P, _P, _P]),
    "sober_car_cluster_cols_fits": (C.c_int, [_I32, _I32]),
    "sober_car_cluster_cols": (C.c_int, [_P, _I32, _I32, _P, _I32, _P, _P]),
    "sober_car_cluster_cols_profiled": (C.c_int, [_P, _I32, _I32, _P, _I32, _P, _P, _P]),
    "sober_car_panel_fits": (C.c_int, [_I32, _I32]),
    "sober_car_panel_workspace": (_I64, [_I32, _I32]),
    "sober_car_panel": (C.c_int, [_P, _I32, _I32, _P, _I32, _P, _P, _I64, _P]),
    "sober_car_panel_profiled": (C.c_int, [_P, _I32, _I32, _P, _I32, _P, _P, _I64, _P, _P]),
    "sober_car_prepare": (C.c_int, [_P, _I64, _P, _I32, _I32, _P, _I64, _P]),
    "sober_car_summary": (C.c_int, [_P, _I32, _P, _I64, _D, _P, _P, _P]),
    "sober_apply_tail": (C.c_int, [_P, _P, _I32, _P, _P, _I32, _P, _P]),
    "sober_update_compact": (C.c_int, [_P, _P, _I64, _I64, _I64, _I32, _P, _P, _P, _I32, _I32, _I64, _P, _P,
                                       _P, _P, _I64, _I32, _P]),
    "sober_update_compact_dev": (C.c_int, [_P, _P, _I64, _I64, _I64, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32,
                                           _P]),
    "sober_scatter_result": (C.c_int, [_P, _I64, _P, _P, _I64, _P]),
    "sober_project_design": (C.c_int, [_P, _I64, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _P, _I64, _P, _P]),
    "sober_trsm_right_upper": (C.c_int, [_P, _I64, _P, _I64, _I32, _I32, _P, _I64, _P]),
    "sober_fp64_probe": (C.c_int, [_I32, _I64, _P, _P]),
    "sober_cholesky_upper_fits": (C.c_int, [_I32]),
    "sober_cholesky_upper": (C.c_int, [_P, _I64, _I32, _P, _I64, _P, _P]),
    "sober_gp_rows": (C.c_int, [_P, _I64, _P, _I64, _P, _I64, _I32, _D, _P, _D, _D, _D, _D, _P, _P, _P, _P]),
    "sober_kmeans_assign": (C.c_int, [_P, _I64, _I64, _I32, _P, _I32, _P, _P]),
    "sober_partition_stream": (C.c_int, [_I32, C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]),
    "sober_dmma_probe": (C.c_int, [_I32, _I64, _P, _P]),
    "sober_popc_probe": (C.c_int, [_I32, _I64, _P, _P]),
}

_lib = None


class SoberB200Error(RuntimeError):
    pass


def load():
    """Load the shared library once; raise loudly when it is absent or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SoberB200Error(
            "sober_b200: %s not found -- the CUDA extension is not built and there is no CPU fallback. "
            "Run `python -c \"import __graft_entry__ as g; g.build()\"` (or `make -C sober_b200/csrc`)." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise SoberB200Error("sober_b200: %s does not export %s" % (LIB_PATH, name)) from exc
        fn.restype = res
        fn.argtypes = args
    if lib.sober_abi_version() != ABI_VERSION:
        raise SoberB200Error("sober_b200: ABI version mismatch")
    _lib = lib
    return lib


def check(status, what):
    if status != OK:
        lib = load()
        detail = lib.sober_last_cuda_error().decode() if status == 2 else ""
        raise SoberB200Error("sober_b200: %s failed: %s %s" % (what, _STATUS.get(status, status), detail))

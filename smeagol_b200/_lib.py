"""ctypes binding of libsmeagol_b200.so (the C-ABI declared in include/smeagol_b200.h).

There is NO CPU fallback: if the shared library is missing or a CUDA call fails, the product path raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SMEAGOL_B200_LIB", os.path.join(_HERE, "libsmeagol_b200.so"))  # env: tuning builds

SMG_OK = 0
CTR_HITS_APPENDED, CTR_NEAR, CTR_DEFINITE, CTR_ADJ_ACCEPTED, CTR_LAUNCHES, CTR_CANDIDATES, CTR_NEAR_TC, CTR_KMER_ENTRIES = 0, 1, 2, 3, 4, 5, 6, 7
N_COUNTERS = 8
MAX_REGTAB_WIDTH = 32
ROW_DTYPE_FIELDS = [("word_off", "<i8"), ("g_off", "<i8"), ("g_len", "<i8"), ("n_avail", "<i4"), ("n_own", "<i4")]
NEAR_BYTES = 24

# every symbol include/smeagol_b200.h declares (tests/test_cabi.py checks the export list against the header)
EXPORTS = [
    "smg_version", "smg_last_error", "smg_launch_count", "smg_pwmset_create", "smg_pwmset_create_split", "smg_pwmset_create_split3",
    "smg_kmer_index_size", "smg_kmer_index_build", "smg_scan_kmer", "smg_seed_prepare", "smg_seed_index_build", "smg_scan_seed", "smg_pwmset_destroy", "smg_pwmset_info",
    "smg_pack", "smg_scan", "smg_scan_tc", "smg_scan_tc5", "smg_adjudicate", "smg_sort_hits", "smg_sort_hits_bucketed", "smg_predict_dense", "smg_variant_sites",
    "smg_variant_windows", "smg_variant_windows_mode", "smg_pairwise_ncorrs", "smg_window_counts", "smg_fasta_index", "smg_fasta_compact", "smg_kmer_shuffle", "smg_bin_hits",
]


class SmeagolB200Error(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (built by __graft_entry__.build() / smeagol_b200/csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SmeagolB200Error(
            "CUDA library %s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(smeagol_b200 has no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, u64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64
    lib.smg_version.restype = ctypes.c_int
    lib.smg_last_error.restype = ctypes.c_char_p
    lib.smg_launch_count.restype = u64
    lib.smg_pwmset_create.argtypes = [vp, vp, i32, ctypes.POINTER(vp)]
    lib.smg_pwmset_create_split.argtypes = [vp, vp, i32, i32, ctypes.POINTER(vp)]
    lib.smg_pwmset_create_split3.argtypes = [vp, vp, i32, i32, i32, ctypes.POINTER(vp)]
    lib.smg_kmer_index_size.argtypes = [vp, ctypes.c_int, ctypes.POINTER(i64)]
    lib.smg_kmer_index_build.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, i64, vp, u64, vp, vp]
    lib.smg_scan_kmer.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, i32, vp, vp, i32, i32, ctypes.c_int, vp, vp, vp, ctypes.c_int,
                                  ctypes.c_int, vp, vp, u64, vp, vp, u64, vp, vp, vp]
    lib.smg_pwmset_destroy.argtypes = [vp]
    lib.smg_pwmset_info.argtypes = [vp, vp]
    lib.smg_pack.argtypes = [vp, ctypes.c_int, vp, vp, i32, i64, vp, vp, vp, vp, vp]
    lib.smg_scan.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, i32, i32, ctypes.c_int, vp, vp, ctypes.c_int, ctypes.c_int,
                             vp, vp, u64, vp, u64, vp, vp, vp]
    lib.smg_scan_tc.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, i32, i32, ctypes.c_int, vp, vp, vp, vp, ctypes.c_int,
                                ctypes.c_int, ctypes.c_int, vp, u64, vp, vp, u64, vp, vp, vp]
    lib.smg_scan_tc5.argtypes = lib.smg_scan_tc.argtypes
    lib.smg_seed_prepare.argtypes = [vp, ctypes.c_int, vp]
    lib.smg_seed_index_build.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, i64, vp, u64, vp, vp]
    lib.smg_scan_seed.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, i32, vp, vp, i32, i32, ctypes.c_int, vp, vp, vp, ctypes.c_int,
                                  ctypes.c_int, ctypes.c_int, vp, vp, vp, u64, vp, vp, u64, vp, vp, vp]
    lib.smg_pairwise_ncorrs.argtypes = [vp, vp, vp, i32, vp, vp]
    lib.smg_window_counts.argtypes = [vp, u64, ctypes.c_int, ctypes.c_int, vp, vp, vp, i64, i64, vp, vp]
    lib.smg_bin_hits.argtypes = [vp, vp, u64, ctypes.c_int, ctypes.c_int, vp, vp, i32, i32, vp, vp]
    lib.smg_kmer_shuffle.argtypes = [vp, i64, vp, vp, i32, i32, u64, vp, vp, i64, vp]
    lib.smg_fasta_index.argtypes = [vp, i64, vp, u64, vp, vp]
    lib.smg_fasta_compact.argtypes = [vp, i64, vp, i64, vp, vp, vp, vp, ctypes.c_int, vp, ctypes.POINTER(u64), vp]
    lib.smg_adjudicate.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, u64, ctypes.c_int, ctypes.c_int, vp, vp, u64, vp, vp, vp]
    lib.smg_sort_hits.argtypes = [vp, vp, vp, vp, u64, ctypes.c_int, vp, ctypes.POINTER(u64), vp]
    lib.smg_sort_hits_bucketed.argtypes = [vp, vp, vp, vp, u64, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, ctypes.POINTER(u64), vp, vp]
    lib.smg_predict_dense.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp, vp]
    lib.smg_variant_sites.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp]
    lib.smg_variant_windows.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp]
    lib.smg_variant_windows_mode.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, i32, i32, vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("smg_version", "smg_last_error", "smg_launch_count"):
            fn.restype = ctypes.c_int
    _lib = lib
    return lib


def check(rc, what):
    if rc != SMG_OK:
        msg = load().smg_last_error()
        raise SmeagolB200Error("%s failed (status %d): %s" % (what, rc, msg.decode() if msg else "?"))


def launch_count():
    return int(load().smg_launch_count())

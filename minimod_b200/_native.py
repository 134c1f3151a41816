"""ctypes bindings of the two native libraries.

  lib/libminimod_cuda.so   the C ABI of include/minimod_cuda.h (CUDA kernels, sm_100a)
  lib/libminimod_host.so   host-side BAM/FASTA/packing/formatting (minimod_b200/host/capi.cpp)

There is deliberately no Python or CPU implementation of the hot path behind these bindings:
if libminimod_cuda.so is missing or no CUDA device is usable, loading / mmc_create() fails loudly.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIBDIR = os.path.join(HERE, "lib")

MMC_VIEW, MMC_FREQ = 0, 1
MMC_OK, MMC_EINVAL, MMC_ECUDA, MMC_ENOMEM, MMC_EREAD, MMC_ESTATE, MMC_EORDER = 0, -1, -2, -3, -4, -5, -6
MMC_MAX_CODE_LEN, MMC_MAX_CONTEXT, MMC_MAX_MODS = 8, 32, 64


class MmcMod(C.Structure):
    _fields_ = [("code", C.c_char * (MMC_MAX_CODE_LEN + 1)),
                ("context", C.c_char * (MMC_MAX_CONTEXT + 1)),
                ("call_lut", C.c_uint8 * 256)]


class MmcOpts(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("subtool", C.c_int32), ("n_mods", C.c_int32),
                ("mods", C.POINTER(MmcMod)), ("insertions", C.c_int32), ("haplotypes", C.c_int32),
                ("device", C.c_int32), ("n_slots", C.c_int32), ("max_reads", C.c_uint64),
                ("max_bytes", C.c_uint64), ("sparse_capacity", C.c_uint64), ("dense_haps", C.c_int32),
                ("dense_codes", C.c_int32), ("view_capacity", C.c_uint64),
                ("cap_cigar_words", C.c_uint64), ("cap_seq_bytes", C.c_uint64), ("cap_mm_bytes", C.c_uint64),
                ("cap_ml_bytes", C.c_uint64), ("seq_packing", C.c_int32), ("cigar_packing", C.c_int32)]


class MmcBatch(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("max_reads", C.c_uint32),
                ("tid", C.POINTER(C.c_int32)), ("pos", C.POINTER(C.c_int32)),
                ("l_seq", C.POINTER(C.c_uint32)), ("n_cigar", C.POINTER(C.c_uint32)),
                ("mm_len", C.POINTER(C.c_uint32)), ("ml_len", C.POINTER(C.c_uint32)),
                ("cigar_off", C.POINTER(C.c_uint64)), ("seq_off", C.POINTER(C.c_uint64)),
                ("mm_off", C.POINTER(C.c_uint64)), ("ml_off", C.POINTER(C.c_uint64)),
                ("flag", C.POINTER(C.c_uint16)), ("hp", C.POINTER(C.c_uint8)),
                ("cigar", C.POINTER(C.c_uint32)), ("cigar_cap", C.c_uint64), ("cigar_used", C.c_uint64),
                ("seq4", C.POINTER(C.c_uint8)), ("seq_cap", C.c_uint64), ("seq_used", C.c_uint64),
                ("mm", C.POINTER(C.c_char)), ("mm_cap", C.c_uint64), ("mm_used", C.c_uint64),
                ("ml", C.POINTER(C.c_uint8)), ("ml_cap", C.c_uint64), ("ml_used", C.c_uint64),
                ("priv", C.c_void_p),
                ("seq2", C.POINTER(C.c_uint8)), ("seq_exc", C.POINTER(C.c_uint64)), ("seq_exc_cap", C.c_uint64),
                ("seq_exc_used", C.c_uint64), ("seq_packing", C.c_uint32), ("cigar_packing", C.c_uint32),
                ("cig8", C.POINTER(C.c_uint8)), ("cig8_cap", C.c_uint64), ("cig8_used", C.c_uint64),
                ("cig8_off", C.POINTER(C.c_uint64))]


class MmcFreqRec(C.Structure):
    _fields_ = [("tid", C.c_int32), ("pos", C.c_int32), ("n_called", C.c_uint32), ("n_mod", C.c_uint32),
                ("ins_offset", C.c_uint16), ("hap", C.c_int16), ("strand", C.c_uint8), ("code", C.c_uint8),
                ("reserved", C.c_uint16)]


class MmcViewRec(C.Structure):
    _fields_ = [("read", C.c_uint32), ("ref_pos", C.c_int32), ("read_pos", C.c_int32), ("ins_offset", C.c_uint32),
                ("code", C.c_uint8), ("mod_prob", C.c_uint8), ("strand", C.c_uint8), ("hp", C.c_uint8)]


class MmcTimers(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("decode_ms", C.c_double), ("finalize_ms", C.c_double), ("d2h_ms", C.c_double),
                ("batches", C.c_uint64), ("reads", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("deferred_reads", C.c_uint64),
                ("flat_deferred_reads", C.c_uint64)]


class MmhSynthStats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("n_reads", "bases", "ml_entries", "cigar_ops", "mm_bytes", "ref_span", "seq_bytes")]


class MmhStats(C.Structure):
    _fields_ = [("total_reads", C.c_int32), ("n_recs", C.c_int32), ("total_bytes", C.c_int64),
                ("processed_bytes", C.c_int64), ("ml_entries", C.c_int64), ("bases", C.c_int64)]


FREQ_DTYPE = [("tid", "<i4"), ("pos", "<i4"), ("n_called", "<u4"), ("n_mod", "<u4"), ("ins_offset", "<u2"),
              ("hap", "<i2"), ("strand", "u1"), ("code", "u1"), ("reserved", "<u2")]
VIEW_DTYPE = [("read", "<u4"), ("ref_pos", "<i4"), ("read_pos", "<i4"), ("ins_offset", "<u4"),
              ("code", "u1"), ("mod_prob", "u1"), ("strand", "u1"), ("hp", "u1")]

# every symbol include/minimod_cuda.h declares
CUDA_SYMBOLS = [
    "mmc_create", "mmc_destroy", "mmc_strerror", "mmc_abi_version", "mmc_ref_add", "mmc_ref_commit",
    "mmc_batch_acquire", "mmc_batch_submit", "mmc_batch_wait", "mmc_batch_release", "mmc_batch_upload",
    "mmc_batch_launch", "mmc_sync", "mmc_freq_finalize", "mmc_freq_reset", "mmc_code_name", "mmc_view_fetch",
    "mmc_freq_drain", "mmc_freq_undrain",
    "mmc_dense_slice", "mmc_dense_touch", "mmc_touched_range", "mmc_get_timers", "mmc_reset_timers", "mmc_last_decode_ms", "mmc_describe", "mmc_region_reduce",
]


def _declare_cuda(lib):
    vp, i32, u32, u64 = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64
    P = C.POINTER
    sig = {
        "mmc_create": (C.c_int, [P(vp), P(MmcOpts), i32, P(C.c_char_p), P(u32)]),
        "mmc_destroy": (None, [vp]),
        "mmc_strerror": (C.c_char_p, [vp]),
        "mmc_abi_version": (C.c_int, []),
        "mmc_ref_add": (C.c_int, [vp, i32, C.c_char_p, u32]),
        "mmc_ref_commit": (C.c_int, [vp]),
        "mmc_batch_acquire": (C.c_int, [vp, P(P(MmcBatch))]),
        "mmc_batch_submit": (C.c_int, [vp, P(MmcBatch)]),
        "mmc_batch_wait": (C.c_int, [vp, P(MmcBatch)]),
        "mmc_batch_release": (C.c_int, [vp, P(MmcBatch)]),
        "mmc_batch_upload": (C.c_int, [vp, P(MmcBatch)]),
        "mmc_batch_launch": (C.c_int, [vp, P(MmcBatch)]),
        "mmc_sync": (C.c_int, [vp]),
        "mmc_freq_finalize": (C.c_int, [vp, P(P(MmcFreqRec)), P(u64)]),
        "mmc_freq_reset": (C.c_int, [vp]),
        "mmc_freq_drain": (C.c_int, [vp, i32, u32, P(P(MmcFreqRec)), P(u64)]),
        "mmc_freq_undrain": (C.c_int, [vp]),
        "mmc_code_name": (C.c_char_p, [vp, i32]),
        "mmc_view_fetch": (C.c_int, [vp, P(MmcBatch), P(P(MmcViewRec)), P(u64)]),
        "mmc_dense_slice": (C.c_int, [vp, i32, u32, u32, P(vp), P(u64)]),
        "mmc_dense_touch": (C.c_int, [vp, i32, u32, u32]),
        "mmc_touched_range": (C.c_int, [vp, i32, P(u32), P(u32)]),
        "mmc_get_timers": (C.c_int, [vp, P(MmcTimers)]),
        "mmc_reset_timers": (C.c_int, [vp]),
        "mmc_last_decode_ms": (C.c_int, [vp, P(MmcBatch), P(C.c_double)]),
        "mmc_describe": (C.c_char_p, [vp]),
        "mmc_region_reduce": (C.c_int, [P(vp), i32, i32, P(C.c_double), P(u64)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


def _declare_host(lib):
    vp, P = C.c_void_p, C.POINTER
    sig = {
        "mmh_bam_open": (vp, [C.c_char_p, C.c_char_p, C.c_int]),
        "mmh_bam_open_t": (vp, [C.c_char_p, C.c_int, C.c_char_p, C.c_int]),
        "mmh_bam_scan": (C.c_longlong, [vp, P(C.c_ulonglong)]),
        "mmh_bam_close": (None, [vp]),
        "mmh_bam_n_targets": (C.c_int, [vp]),
        "mmh_bam_target_name": (C.c_char_p, [vp, C.c_int]),
        "mmh_bam_target_len": (C.c_uint32, [vp, C.c_int]),
        "mmh_loader_new": (vp, [vp, C.c_int32, C.c_int64, C.c_int, C.c_int, C.c_int]),
        "mmh_loader_free": (None, [vp]),
        "mmh_loader_fill": (C.c_int, [vp, P(MmcBatch), P(MmhStats), C.c_char_p, C.c_int]),
        "mmh_loader_qname": (C.c_char_p, [vp, C.c_uint32]),
        "mmh_fasta_load": (vp, [C.c_char_p, C.c_char_p, C.c_int]),
        "mmh_fasta_free": (None, [vp]),
        "mmh_fasta_n": (C.c_int, [vp]),
        "mmh_fasta_name": (C.c_char_p, [vp, C.c_int]),
        "mmh_fasta_seq": (vp, [vp, C.c_int]),
        "mmh_fasta_len": (C.c_uint64, [vp, C.c_int]),
        "mmh_parse_mods": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, P(MmcMod), C.c_int, C.c_char_p, C.c_int]),
        "mmh_fastfmt_selftest": (C.c_longlong, [C.c_uint, C.c_ulonglong, C.c_char_p, C.c_int]),
        "mmh_write_freq": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, P(C.c_char_p), P(MmcFreqRec),
                                      C.c_uint64, C.c_int, P(C.c_char_p)]),
        "mmh_synth_new": (vp, [C.c_int, C.c_uint64, C.c_int, P(C.c_char_p), P(C.c_uint32), C.c_double]),
        "mmh_synth_new2": (vp, [C.c_int, C.c_uint64, C.c_int, P(C.c_char_p), P(C.c_uint32), P(C.c_uint32), C.c_double]),
        "mmh_synth_free": (None, [vp]),
        "mmh_synth_n_reads": (C.c_uint64, [vp]),
        "mmh_synth_ref": (vp, [vp, C.c_int, P(C.c_uint64)]),
        "mmh_synth_write_fasta": (C.c_int, [vp, C.c_char_p]),
        "mmh_synth_fill": (C.c_int64, [vp, P(MmcBatch), C.c_uint64, C.c_uint64, C.c_int, P(MmhSynthStats)]),
        "mmh_synth_write_bam": (C.c_int, [vp, C.c_char_p, C.c_uint64, C.c_uint64, C.c_int, P(MmhSynthStats)]),
        "mmh_write_view": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, P(C.c_char_p), P(MmcBatch), vp,
                                      P(MmcViewRec), C.c_uint64, C.c_int, P(C.c_char_p)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


_cuda = None
_host = None


def cuda_lib_path():
    return os.path.join(LIBDIR, "libminimod_cuda.so")


def load_cuda(path=None):
    """Load libminimod_cuda.so (the product) -- or, for the CPU-only kernel tests, an explicitly
    named SIMT-emulation build of the same sources.  Never falls back silently."""
    global _cuda
    if path is None:
        if _cuda is not None:
            return _cuda
        path = cuda_lib_path()
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `make lib` (nvcc, sm_100a). "
                               "minimod_b200 has no CPU fallback.")
        _cuda = _declare_cuda(C.CDLL(path))
        return _cuda
    return _declare_cuda(C.CDLL(path))


def load_host():
    global _host
    if _host is None:
        path = os.path.join(LIBDIR, "libminimod_host.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `make host`.")
        _host = _declare_host(C.CDLL(path))
    return _host

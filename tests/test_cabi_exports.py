"""The nvcc-built C-ABI library loads without a GPU and exports every symbol that
include/bchfft_cuda.h declares (no compute calls here)."""
import ctypes
import os
import re

from bch_fft_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "bchfft_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bchfft_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.EXPORTS)


def test_cuda_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.CUDA_SO), "build it first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_lib.CUDA_SO)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    bound = _lib.bind(_lib.CUDA_SO)
    assert bound.bchfft_is_emulated() == 0


def test_host_library_exports_reference_symbols():
    """the drop-in host library carries the reference's hot-path entry points"""
    path = os.path.join(ROOT, "host", "libbchfft_host.so")
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in ("evaluate_polynomial_additive_FFT", "additive_DFT_opt", "additive_DFT", "additive_DFT_ref",
                 "bch_eval_syndrome", "bch_compute_elp", "bch_locate_errors", "bch_decode",
                 "bch_alloc_buffers", "bch_free_buffers", "bch_free_code", "bch_gen_code", "bch_encode",
                 "multiplicative_FFT", "evaluate_polynomial_multiplicative_FFT",
                 "create_binary_finite_field", "free_finite_field", "multiply", "multiply_by_log",
                 "square", "inverse", "get_bit", "set_bit", "add_bit", "degree_to_packed_size",
                 "additive_fft_batch", "bch_decode_batch", "bch_encode_batch", "multiplicative_fft_batch",
                 # host-only symbols the reference's objects export (SURVEY 8b "complete the libs")
                 "evaluate_sparse_polynomial", "evaluate_sparse_polynomial_ref", "evaluate_sparse_polynomial_log",
                 "multiply_polynomial_by_monomial", "euclidean_division_sparse_reductor",
                 "euclidean_division_sparse_reductor_ref", "euclidean_division_packed_inplace_gf2",
                 "word_size", "wordsize_to_max_degree", "degree_packed", "degree_packed_u32",
                 "degree_to_packed_size_u32", "degree_to_packed_pos", "degree_to_packed_pos_u32",
                 "bch_build_polynomial", "bch_explore_cycles", "bch_reduce_syndrome", "bch_find_best_offset",
                 "numFactors", "factors", "factorOffset", "getFactorSum", "DFT_direct",
                 "multiplicative_FFT_compute_step", "multiplicative_FFT_reorder_step",
                 "multiplicative_FFT_step_chaining", "check_fft_polynomials",
                 "absolute_time", "init_time", "urand", "compare_results"):
        assert hasattr(lib, name), name


def test_product_fails_loudly_without_a_device():
    """no CPU fallback: with no CUDA device the C-ABI reports an error instead of computing"""
    lib = _lib.bind(_lib.CUDA_SO)
    if lib.bchfft_device_count() > 0:
        return   # on a GPU box this property is not observable
    import numpy as np
    exp = np.zeros(16, np.uint32)
    h = ctypes.c_void_p()
    rc = lib.bchfft_field_create(0, 4, exp.ctypes.data, exp.ctypes.data, ctypes.byref(h))
    assert rc != 0 and not h.value

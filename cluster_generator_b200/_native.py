"""
ctypes binding of ``libcgb200.so`` (C ABI declared in ``include/cgb200.h``).

There is no CPU fallback: importing this module without the built library, or
calling into it without a CUDA device, raises.  PyTorch is used for device
buffers and streams only; every pointer crossing the ABI is ``tensor.data_ptr()``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CGB200_LIB") or os.path.join(_HERE, "lib", "libcgb200.so")

CGB_OK = 0
CGB_E_INVALID = -1
CGB_E_CUDA = -2
CGB_E_WORKSPACE = -3
CGB_E_NOTMONO = -4
CGB_E_IO = -5
CGB_DTYPE_F64 = 0
CGB_DTYPE_F32 = 1
CGB_NMOMENTS = 16
CGB_FIT_FITPACK, CGB_FIT_PCR = 0, 1
CGB_FIT_REUSE_FACTOR = 0x100
MOMENT_NAMES = (
    "particles", "sum_trials", "max_trials", "sum_v2", "sum_psi", "sum_x", "sum_y", "sum_z",
    "sum_vx", "sum_vy", "sum_vz", "nonpositive_envelopes", "cooperative_finishes", "failures",
    "sum_r", "exact_evaluations",
)
STREAM_RADIUS, STREAM_POS_DIR, STREAM_VEL_TRIAL, STREAM_VEL_DIR = 0, 1, 2, 3


class SamplerDesc(Structure):
    _fields_ = [
        ("d_cdf_P", c_void_p), ("d_cdf_r", c_void_p), ("n_cdf", c_int32), ("n_psi", c_int32),
        ("d_psi_t", c_void_p), ("d_psi_c", c_void_p), ("d_f_t", c_void_p), ("d_f_c", c_void_p),
        ("n_f", c_int32), ("n_buckets", c_int32), ("n_aux", c_int32), ("n_aux_coef", c_int32),
        ("d_aux_t", c_void_p * 2), ("d_aux_c", c_void_p * 2),
    ]


class HostUniforms(Structure):
    _fields_ = [
        ("d_u_r", c_void_p), ("d_u_ct", c_void_p), ("d_u_ph", c_void_p), ("d_u_ctv", c_void_p),
        ("d_u_phv", c_void_p), ("d_u_trial", c_void_p), ("t_cap", c_int64),
    ]


class SampleArgs(Structure):
    _fields_ = [
        ("d_workspace", c_void_p), ("id_begin", c_int64), ("id_end", c_int64),
        ("particle_mass", c_double), ("seed", c_uint64), ("out_dtype", c_int32), ("flags", c_int32),
        ("d_pos", c_void_p), ("d_vel", c_void_p), ("ld", c_int64), ("d_mass", c_void_p),
        ("d_pot", c_void_p), ("d_radius", c_void_p), ("d_speed", c_void_p), ("d_trials", c_void_p),
        ("d_moments", c_void_p), ("hu", HostUniforms), ("radius_period", c_int64),
        ("n_peers", c_int32), ("pad_peers", c_int32), ("d_peer_moments", c_void_p * 8),
    ]


class HaloTables(Structure):
    _fields_ = [
        ("d_dens_pp", c_void_p), ("d_ener_pp", c_void_p), ("n", c_int32), ("pad", c_int32),
        ("center", c_double * 3), ("velocity", c_double * 3), ("radius_cut", c_double),
    ]


# every symbol include/cgb200.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("cgb_version", c_int, []),
    ("cgb_last_error", c_char_p, []),
    ("cgb_device_sm_count", c_int, [POINTER(c_int)]),
    ("cgb_measure_dfma_tflops", c_int, [c_int, POINTER(c_double), c_void_p]),
    ("cgb_measure_ffma_tflops", c_int, [c_int, POINTER(c_double), c_void_p]),
    ("cgb_spline_fit_workspace_bytes", c_size_t, [c_int]),
    ("cgb_spline_fit", c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    ("cgb_test_givens_cycles", c_int, [c_void_p]),
    ("cgb_test_fpgivs", c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    ("cgb_test_fpgivs_redo", c_int, [c_int, POINTER(c_int)]),
    ("cgb_bounds_check_enabled", c_int, []),
    ("cgb_test_bounds_check", c_int, [ctypes.c_longlong, ctypes.c_longlong]),
    ("cgb_test_div_by_rcp", c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    ("cgb_spline_eval", c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    ("cgb_eddington_workspace_bytes", c_size_t, [c_int]),
    ("cgb_eddington_df", c_int, [c_void_p, c_void_p, c_int, c_double, c_double, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    ("cgb_check_virial", c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p]),
    ("cgb_integrate_spline", c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_double,
                                     c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    ("cgb_qags_replay_cache_doubles", c_size_t, [c_int]),
    ("cgb_qags_replay", c_int, [c_void_p, c_void_p, c_int, c_double, c_double, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    ("cgb_qags_trace", c_int, [c_void_p, c_void_p, c_int, c_double, c_double, c_double, c_double,
                               c_double, c_int, c_void_p, c_void_p]),
    ("cgb_sampler_workspace_bytes", c_size_t, [POINTER(SamplerDesc)]),
    ("cgb_sampler_prepare", c_int, [POINTER(SamplerDesc), c_void_p, c_size_t, c_void_p]),
    ("cgb_sample_scratch_bytes", c_size_t, []),
    ("cgb_sample_particles", c_int, [POINTER(SampleArgs), c_void_p, c_size_t, c_void_p]),
    ("cgb_assemble_particles", c_int, [POINTER(SampleArgs), c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    ("cgb_generate_velocities_workspace_bytes", c_size_t, [c_int]),
    ("cgb_generate_velocities", c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                        c_int, c_void_p, c_int64, c_uint64, c_int64, c_void_p,
                                        c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    ("cgb_sample_gas", c_int, [POINTER(SampleArgs), c_void_p, c_void_p, c_void_p]),
    ("cgb_sampler_gas_flags", c_int, [c_void_p, POINTER(c_int32), c_void_p]),
    ("cgb_screen_audit", c_int, [c_void_p, c_int64, c_int64, c_uint64, c_int64, c_void_p, c_void_p]),
    ("cgb_ppform_workspace_bytes", c_size_t, [c_int]),
    ("cgb_ppform_prepare", c_int, [c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    ("cgb_map_field", c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    ("cgb_resample_gas", c_int, [POINTER(HaloTables), c_int, c_void_p, c_int64, c_int64, c_int, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    ("cgb_resample_scalars", c_int, [POINTER(HaloTables), c_int, POINTER(c_void_p), c_int, c_void_p, c_int64,
                                     c_int64, c_int, POINTER(c_void_p), c_int, c_void_p]),
    ("cgb_particle_histograms", c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int64, c_int, c_double,
                                        c_double, c_double, c_double, c_int, c_void_p, c_void_p]),
    ("cgb_peer_alloc", c_int, [c_size_t, POINTER(c_void_p), c_char_p]),
    ("cgb_peer_open", c_int, [c_char_p, POINTER(c_void_p)]),
    ("cgb_peer_close", c_int, [c_void_p]),
    ("cgb_peer_free", c_int, [c_void_p]),
    ("cgb_h5_create", c_int, [c_char_p, c_int, POINTER(c_void_p)]),
    ("cgb_h5_create_group", c_int, [c_void_p, c_char_p]),
    ("cgb_h5_write_dataset", c_int, [c_void_p, c_char_p, c_int, c_int, POINTER(c_int64), c_void_p]),
    ("cgb_h5_write_dataset_device", c_int, [c_void_p, c_char_p, c_int, c_void_p, c_int, c_int64, c_int, c_int64,
                                            c_double, c_void_p]),
    ("cgb_h5_set_attr", c_int, [c_void_p, c_char_p, c_char_p, c_int, c_int64, c_void_p]),
    ("cgb_h5_close", c_int, [c_void_p]),
    ("cgb_h5_open", c_int, [c_char_p, POINTER(c_void_p)]),
    ("cgb_h5_num_children", c_int, [c_void_p, c_char_p, POINTER(c_int)]),
    ("cgb_h5_child_name", c_int, [c_void_p, c_char_p, c_int, c_char_p, c_int, POINTER(c_int)]),
    ("cgb_h5_dataset_info", c_int, [c_void_p, c_char_p, POINTER(c_int), POINTER(c_int), POINTER(c_int64),
                                    POINTER(c_int64)]),
    ("cgb_h5_read_dataset", c_int, [c_void_p, c_char_p, c_void_p, c_size_t]),
    ("cgb_h5_read_dataset_device", c_int, [c_void_p, c_char_p, c_void_p, c_int, c_int64, c_double, c_void_p]),
    ("cgb_h5_num_attrs", c_int, [c_void_p, c_char_p, POINTER(c_int)]),
    ("cgb_h5_attr_name", c_int, [c_void_p, c_char_p, c_int, c_char_p, c_int]),
    ("cgb_h5_attr_info", c_int, [c_void_p, c_char_p, c_char_p, POINTER(c_int), POINTER(c_int64), POINTER(c_size_t)]),
    ("cgb_h5_read_attr", c_int, [c_void_p, c_char_p, c_char_p, c_void_p, c_size_t]),
]

_LIB = None


def load_library(path: str = LIB_PATH) -> ctypes.CDLL:
    """dlopen the library and bind every declared symbol (no compute, no GPU needed)."""
    global _LIB
    if _LIB is not None and path == LIB_PATH:
        return _LIB
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: the CUDA extension is not built. Run "
            f"`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)."
        )
    lib = ctypes.CDLL(path)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.cgb_version() != 100:
        raise RuntimeError(f"libcgb200 version mismatch: {lib.cgb_version()}")
    if path == LIB_PATH:
        _LIB = lib
    return lib


def lib() -> ctypes.CDLL:
    return load_library()


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "cluster_generator_b200 needs a CUDA device (sm_100a); there is no CPU fallback."
        )
    return torch.device("cuda", torch.cuda.current_device())


def check(rc: int) -> None:
    """Turn a C-ABI status into the exception type the reference would raise."""
    if rc == CGB_OK:
        return
    msg = lib().cgb_last_error().decode("utf-8", "replace")
    if rc in (CGB_E_INVALID, CGB_E_NOTMONO):
        raise ValueError(msg)
    if rc == CGB_E_IO:
        raise IOError(msg)
    raise RuntimeError(f"libcgb200 error {rc}: {msg}")


def stream_ptr() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor, or NULL for None."""
    if t is None:
        return c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "C-ABI buffers must be contiguous CUDA tensors"
    return c_void_p(t.data_ptr())


def dev_f64(a, device=None) -> torch.Tensor:
    """Host array -> contiguous float64 CUDA tensor."""
    import numpy as np

    device = device or require_cuda()
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float64).contiguous()
    arr = np.ascontiguousarray(a, dtype=np.float64)
    if not arr.flags.writeable:
        arr = arr.copy()
    return torch.from_numpy(arr).to(device)


def workspace(nbytes: int, device=None) -> torch.Tensor:
    """256-byte aligned scratch buffer (torch's caching allocator aligns to 512 B)."""
    device = device or require_cuda()
    t = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
    assert t.data_ptr() % 256 == 0
    return t

"""ctypes binding of the C ABI in include/fem_helmholtz.h (libfemhelmholtz_b200.so).

There is deliberately NO fallback: if the shared library is missing or no CUDA device is
visible, every compute entry point raises.  PyTorch is used only for device memory, streams
and (elsewhere) torch.distributed.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# (FH_LIB: development switch -- load another build of the same library, e.g. an instrumented one)
LIB_PATH = os.environ.get("FH_LIB") or os.path.join(_HERE, "lib", "libfemhelmholtz_b200.so")

FH_BC_NEUMANN, FH_BC_SOMMERFELD, FH_BC_DIRICHLET = 0, 1, 2
FH_H_SCALAR, FH_H_NODAL = 0, 1
BC_CODES = {"neumann": FH_BC_NEUMANN, "sommerfeld": FH_BC_SOMMERFELD, "dirichlet": FH_BC_DIRICHLET}


class FHError(RuntimeError):
    """A call into libfemhelmholtz_b200 failed (the message comes from fh_last_error())."""


class Problem1d(C.Structure):
    _fields_ = [
        ("n_elem", C.c_int64), ("h", C.c_double), ("x", C.c_void_p), ("h_mode", C.c_int),
        ("bc_left", C.c_int), ("bc_right", C.c_int),
        ("g_left_re", C.c_double), ("g_left_im", C.c_double),
        ("g_right_re", C.c_double), ("g_right_im", C.c_double),
        ("cos_theta", C.c_double), ("gauss_xi", C.c_double), ("gauss_w", C.c_double),
    ]


_P = C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "fh_last_error": (C.c_char_p, []),
    "fh_abi_version": (C.c_int, []),
    "fh_device_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                 C.POINTER(C.c_size_t)]),
    "fh_shape_functions_1d": (C.c_int, [C.c_double, _P, _P, _P]),
    "fh_element_matrices_1d": (C.c_int, [C.c_double, C.c_double, C.c_double, _P, _P, _P]),
    "fh_boundary_matrices_1d": (C.c_int, [C.c_double, C.c_double, _P, _P, _P]),
    "fh_assemble_1d_c128": (C.c_int, [C.POINTER(Problem1d), _P, _P, C.c_int64, _P, _P, _P, _P, _P]),
    "fh_assemble_1d_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int, C.c_int64]),
    "fh_assemble_1d_ws_c128": (C.c_int, [C.POINTER(Problem1d), _P, _P, C.c_int64, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "fh_assemble_1d_slab_c128": (C.c_int, [C.POINTER(Problem1d), C.c_double, C.c_double, _P, C.c_int64, C.c_int64, _P, _P,
                                           _P, _P, _P]),
    "fh_helmholtz1d_residual_workspace_bytes": (C.c_size_t, []),
    "fh_helmholtz1d_residual_c128": (C.c_int, [C.POINTER(Problem1d), C.c_double, C.c_double, _P, C.c_int64, C.c_int64, _P,
                                               _P, _P, _P, _P, C.c_size_t, _P]),
    "fh_tridiag_batched_max_n": (C.c_int64, []),
    "fh_tridiag_solve_batched_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "fh_tridiag_solve_batched_c128": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64, _P, _P,
                                                C.c_size_t, _P]),
    "fh_helmholtz1d_sweep_fused_c128": (C.c_int, [C.POINTER(Problem1d), _P, _P, C.c_int64, _P, _P, _P,
                                                  C.c_size_t, _P]),
    "fh_helmholtz1d_sweep_assemble_solve_c128": (C.c_int, [C.POINTER(Problem1d), _P, _P, C.c_int64, _P, _P, _P, _P, _P,
                                                           _P, _P, C.c_size_t, _P]),
    "fh_tridiag_solve_large_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "fh_tridiag_solve_large_c128": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, _P, _P, C.c_size_t, _P]),
    "fh_helmholtz1d_solve_large_fused_c128": (C.c_int, [C.POINTER(Problem1d), C.c_double, C.c_double, _P, _P, _P,
                                                        C.c_size_t, _P]),
    "fh_helmholtz1d_solve_large_fused_kdev_c128": (C.c_int, [C.POINTER(Problem1d), _P, _P, _P, _P, C.c_size_t, _P]),
    "fh_helmholtz1d_assemble_solve_large_c128": (C.c_int, [C.POINTER(Problem1d), C.c_double, C.c_double, _P, _P, _P, _P,
                                                           _P, _P, _P, _P, C.c_size_t, _P]),
    "fh_large_uniform_plan": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, _P, C.c_int]),
    "fh_spike_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "fh_spike_reduce_c128": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int, C.c_int, _P, C.c_size_t, _P, _P]),
    "fh_helmholtz1d_spike_reduce_c128": (C.c_int, [C.POINTER(Problem1d), C.c_double, C.c_double, C.c_int64,
                                                   C.c_int64, _P, C.c_size_t, _P, _P]),
    "fh_spike_solve_interface_c128": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "fh_spike_backsub_c128": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int, C.c_int, _P, C.c_size_t, _P, _P, _P,
                                        _P]),
    "fh_helmholtz1d_spike_backsub_c128": (C.c_int, [C.POINTER(Problem1d), C.c_double, C.c_double, C.c_int64,
                                                    C.c_int64, _P, C.c_size_t, _P, _P, _P, _P]),
    "fh_helmholtz1d_spike_reduce_kdev_c128": (C.c_int, [C.POINTER(Problem1d), _P, C.c_int64, C.c_int64, _P, C.c_size_t,
                                                        _P, _P]),
    "fh_helmholtz1d_spike_backsub_kdev_c128": (C.c_int, [C.POINTER(Problem1d), _P, C.c_int64, C.c_int64, _P, C.c_size_t,
                                                         _P, _P, _P, _P]),
    "fh_helmholtz1d_spike_reduce_store_c128": (C.c_int, [C.POINTER(Problem1d), C.c_double, C.c_double, _P, C.c_int64,
                                                         C.c_int64, _P, _P, _P, _P, _P, C.c_size_t, _P, _P]),
    "fh_spike_peer_buffer_bytes": (C.c_size_t, []),
    "fh_helmholtz1d_spike_solve_peer_kdev_c128": (C.c_int, [C.POINTER(Problem1d), _P, C.c_int64, C.c_int64, _P, C.c_size_t,
                                                            _P, C.c_int, C.c_int, _P, C.c_double, _P, _P, _P, _P]),
    "fh_q4_element_matrices_c128": (C.c_int, [_P, _P, C.c_int64, C.c_double, _P, _P, _P, _P]),
    "fh_q4_edge_geometry": (C.c_int, [_P, _P, _P, C.c_int64, C.c_double, _P, _P, _P]),
    "fh_q4_robin_edge_matrices_c128": (C.c_int, [_P, _P, C.c_int64, C.c_double, C.c_double, _P, _P, C.c_int, _P, _P]),
    "fh_q4_neumann_edge_vectors_c128": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P, C.c_int, _P, _P]),
    "fh_q4_scatter_keys": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int, _P, _P]),
    "fh_sorted_segments_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "fh_sorted_segments": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "fh_csr_from_keys": (C.c_int, [_P, _P, C.c_int64, C.c_int64, _P, _P, _P]),
    "fh_segment_sum_c128": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int64, _P, _P]),
    "fh_q4_boundary_flags": (C.c_int, [_P, C.c_int64, C.c_double, C.c_double, C.c_int, _P, _P]),
    "fh_csr_apply_dirichlet_c128": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int, C.c_double, C.c_double,
                                              C.c_double, C.c_double, _P]),
    "fh_q4_boundary_lists": (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_double, C.c_double, _P, _P, _P]),
    "fh_compact_flags_i32": (C.c_int, [_P, C.c_int64, _P, _P, _P]),
    "fh_q4_neumann_load_c128": (C.c_int, [_P, _P, _P, C.c_int64, _P, C.c_double, C.c_int, _P, C.c_int, _P, _P]),
    "fh_q4_gauss_points": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int, _P, _P]),
    "fh_q4_l2_error_workspace_bytes": (C.c_size_t, []),
    "fh_q4_l2_error_c128": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P, _P, C.c_int, _P, _P, C.c_size_t, _P]),
    "fh_csr_to_dense_c128": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P]),
    "fh_dense_lu_solve_c128": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P, _P]),
    "fh_csr_bandwidth": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P]),
    "fh_band_lu_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int64]),
    "fh_csr_band_lu_solve_c128": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int64, _P, _P, C.c_size_t,
                                            _P, _P]),
    "fh_tridiag_solve_pivoted_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "fh_tridiag_solve_pivoted_c128": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64, _P, _P, C.c_size_t, _P]),
    "fh_tridiag_repair_flagged_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int64]),
    "fh_tridiag_repair_flagged_c128": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64, _P, C.c_double, _P, C.c_size_t,
                                                 _P]),
    "fh_tridiag_residual_vector_c128": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64, _P, _P]),
    "fh_tridiag_refine_flagged_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int64]),
    "fh_tridiag_refine_flagged_c128": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64, _P, _P, C.c_size_t, _P]),
    "fh_l2_error_1d_c128": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int64, _P, _P, C.c_int, _P, _P]),
    "fh_tridiag_residual_c128": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64, _P, _P]),
}

_lib = None


def exported_symbols():
    """Names every build of the library must export (checked by the CPU test-suite)."""
    return sorted(_SIGNATURES)


def load():
    """Load the shared library (no GPU needed for this step)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C fem_helmholtz_eqn_b200/csrc`).  There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("fem_helmholtz_eqn_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback")
    return load()


def check(rc, what):
    if rc != 0:
        msg = load().fh_last_error().decode("utf-8", "replace")
        raise FHError(f"{what} failed with code {rc}: {msg}")


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (or NULL for None)."""
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "expected a contiguous CUDA tensor"
    return C.c_void_p(t.data_ptr())


def device():
    return torch.device("cuda", torch.cuda.current_device())


def as_device(a, dtype):
    """numpy / torch / scalar sequence -> contiguous CUDA tensor of `dtype` (H2D copy if needed)."""
    if isinstance(a, torch.Tensor):
        return a.to(device=device(), dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(device())


def device_info():
    lib = require_cuda()
    sm, major, minor, smem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
    check(lib.fh_device_info(C.byref(sm), C.byref(major), C.byref(minor), C.byref(smem)), "fh_device_info")
    return {"sm_count": sm.value, "cc": (major.value, minor.value), "smem_optin": smem.value}


def bind_host_thread_near_device():
    """Restrict this process to the CPU cores NVML reports as local to the current CUDA device, so that pinned host
    buffers allocated afterwards land on the NUMA node the GPU's PCIe root hangs off (one process per GPU: eight
    ranks writing 4 GB each into one socket's memory was measured at a quarter of the per-GPU PCIe rate).  Best
    effort: returns the core list, or None when NVML / the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(torch.cuda.current_device())
        bus = "%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        cores = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return allowed
    except Exception:
        pass
    return None

"""ctypes binding of libidff_b200.so (include/idff_b200.h).  There is no fallback: if the library is
missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IDFF_B200_LIB") or os.path.join(_HERE, "lib", "libidff_b200.so")   # override: A/B experiments
CSRC = os.path.join(_HERE, "csrc")

IDFF_FIELD_MOMENTUM, IDFF_FIELD_SCOREHEAD, IDFF_FIELD_CFM, IDFF_FIELD_MD = 0, 1, 2, 3
IDFF_IMPL_AUTO, IDFF_IMPL_GENERIC, IDFF_IMPL_FUSED, IDFF_IMPL_TENSOR, IDFF_IMPL_LAYERED, IDFF_IMPL_TENSOR16 = 0, 1, 2, 3, 4, 5


class IdffError(RuntimeError):
    pass


class AdamW(C.Structure):
    """idff_adamw_t: torch.optim.AdamW's hyper-parameters + the clip_grad_norm_ threshold."""
    _fields_ = [("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("weight_decay", C.c_double), ("max_grad_norm", C.c_double)]


class FieldParams(C.Structure):
    _fields_ = [("variant", C.c_int32), ("order", C.c_int32), ("dim", C.c_int32), ("t_idx", C.c_int32),
                ("impl", C.c_int32), ("reserved", C.c_int32), ("sigma", C.c_double), ("gamma", C.c_double * 3),
                ("end_div", C.c_double)]


_vp, _i32, _i64, _f32, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double
_PP = C.POINTER(FieldParams)

# name -> (restype, argtypes): every symbol include/idff_b200.h declares
SIGNATURES = {
    "idff_version": (C.c_int, []),
    "idff_last_error": (C.c_char_p, []),
    "idff_source_hash": (C.c_char_p, []),
    "idff_launch_count": (C.c_uint64, []),
    "idff_weights_create": (C.c_int, [C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i32), _i32, _vp, _i32, _i32, _i32,
                                      C.POINTER(_vp)]),
    "idff_weights_destroy": (C.c_int, [_vp]),
    "idff_stage_scalars": (C.c_int, [_PP, _f32, _vp]),
    "idff_debug_tensor_stamps": (C.c_int, [_vp]),
    "idff_debug_tensor16_stamps": (C.c_int, [_vp]),
    "idff_weights_nonfinite": (C.c_int, [_vp, _vp]),
    "idff_mlp_forward": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp]),
    "idff_field_eval": (C.c_int, [_vp, _PP, _f32, _vp, _vp, _vp, _i64, _vp]),
    "idff_sample_rk4": (C.c_int, [_vp, _PP, _vp, _vp, _i32, _vp, _vp, _i64, _vp]),
    "idff_sample_rk4_sweep": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _i32, _vp, _i64, _vp]),
    "idff_sde_num_steps": (C.c_int, [_vp, _i32, _f32]),
    "idff_sample_sde": (C.c_int, [_vp, _PP, _vp, _vp, _i32, _f32, _i32, _vp, _vp, _vp, _i64, _vp]),
    "idff_debug_md_chain_fma": (C.c_int, [_i32]),
    "idff_md_generate": (C.c_int, [_vp, _PP, _i32, _vp, _vp, _i32, _f32, _i32, _vp, _vp, _vp, _i64, _vp]),
    "idff_md_generate_sweep": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, _i32, _f32, _i32, _vp, _vp, _vp, _i64, _vp]),
    "idff_path_sample": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _f32, _i32, _vp, _vp, _i64, _i64, _vp]),
    "idff_philox_plan": (C.c_int, [_i64, _i32, _vp, _vp]),
    "idff_path_sample_philox": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint64, _f32, _i32, _vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    "idff_compute_lambda": (C.c_int, [_vp, _f32, _i32, _i32, _vp, _i64, _vp]),
    "idff_loss_scratch_bytes": (C.c_int64, []),
    "idff_loss": (C.c_int, [_i32, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _vp, _vp, _i64, _i64, _vp, _vp]),
    "idff_mmd_rbf": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _f32, _vp, _i64, _i64, _i64, _i64, _vp]),
    "idff_mmd_rbf_sweep": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _f32, _vp, _vp]),
    "idff_debug_ot_sap": (C.c_int, [_i32]),
    "idff_ot_assign": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "idff_debug_ot_threads": (C.c_int, [_i32]),
    "idff_ot_assign_batch": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "idff_train_flow_state_floats": (C.c_int64, []),
    "idff_train_flow_offsets": (C.c_int, [_vp]),
    "idff_train_flow_init": (C.c_int, [_vp, _vp]),
    "idff_debug_train_stamps": (C.c_int, [_vp, _i32]),
    "idff_debug_train_pair_mode": (C.c_int, [_i32]),
    "idff_train_flow_steps": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _i32, C.POINTER(AdamW), _i64, _vp, _vp,
                                        _vp, _vp]),
    "idff_tune_small_n_route": (C.c_int, [_i64]),
    "idff_debug_wide_dense3": (C.c_int, [_i32]),
    "idff_debug_train_md_stamps": (C.c_int, [_vp]),
    "idff_train_md_state_floats": (C.c_int64, [_i32, _i32]),
    "idff_train_md_offsets": (C.c_int, [_i32, _i32, _vp]),
    "idff_train_md_steps": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _i32, _i32, _f32, _i32, C.POINTER(AdamW), _i64, _vp, _vp,
                                      _vp]),
    "idff_train_md_steps_paired": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _i32, C.POINTER(AdamW),
                                             _i64, _vp, _vp, _vp]),
    "idff_train_state_floats": (C.c_int64, [_i32]),
    "idff_train_offsets": (C.c_int, [_i32, _vp]),
    "idff_train_init": (C.c_int, [_i32, _vp, _vp]),
    "idff_train_steps": (C.c_int, [_i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _i32, _f32, C.POINTER(AdamW), _i64, _vp, _vp,
                                   _vp, _vp]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libidff_b200.so for sm_100a with nvcc (make -C idff_b200/csrc)."""
    res = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if res.returncode != 0:
        raise IdffError("building libidff_b200.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout[-2000:])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IdffError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)        # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().idff_last_error()
        raise IdffError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_of(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise IdffError(f"{name} must be a CUDA tensor: idff_b200 has no CPU path")
    if t.dtype != torch.float32:
        raise IdffError(f"{name} must be float32, got {t.dtype}")


def weights_nonfinite(handle) -> tuple:
    """(particle-evaluations of the fp16-split kernels on this handle whose field value was not finite, rows re-solved in
    fp32 by the safety net of IDFF_IMPL_AUTO)."""
    c = (C.c_uint64 * 2)(0, 0)
    check(lib().idff_weights_nonfinite(handle, c), "idff_weights_nonfinite")
    return int(c[0]), int(c[1])


def source_hash() -> str:
    """sha256 prefix of the sources the loaded library was built from (stamped by the Makefile)."""
    return lib().idff_source_hash().decode()


def expected_source_hash() -> str:
    """the same hash recomputed from the checkout: csrc/*.cu + csrc/*.cuh (sorted), include/idff_b200.h, csrc/Makefile"""
    import glob
    import hashlib
    names = sorted(os.path.basename(p) for p in glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")))
    files = [os.path.join(CSRC, n) for n in names] + [os.path.join(_HERE, "..", "include", "idff_b200.h"),
                                                       os.path.join(CSRC, "Makefile")]
    h = hashlib.sha256()
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def launch_count() -> int:
    return int(lib().idff_launch_count())

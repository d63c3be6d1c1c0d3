"""ctypes binding of libvitb200.so (the C ABI declared in include/vitb200.h).

The library is built in-tree by ``vitb200.build`` (nvcc, sm_100a). There is no fallback: if the shared object is
missing or a call fails, a ``VbError`` is raised.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvitb200.so")
if os.environ.get("VITB200_LIB"):      # experiments: an alternative build of the library (tools only; the product path is the in-tree .so)
    LIB_PATH = os.environ["VITB200_LIB"]


class VbError(RuntimeError):
    pass


class VbGemmDesc(Structure):
    _fields_ = [
        ("M", c_int32), ("N", c_int32), ("K", c_int32), ("batches", c_int32),
        ("a_major", c_int32), ("b_major", c_int32), ("epilogue", c_int32), ("c_dtype", c_int32),
        ("split_k", c_int32), ("c_row_offset", c_int32), ("c_rows", c_int32), ("aux_batch_broadcast", c_int32),
        ("A", c_void_p), ("lda", c_int64), ("batch_stride_a", c_int64),
        ("B", c_void_p), ("ldb", c_int64), ("batch_stride_b", c_int64),
        ("C", c_void_p), ("ldc", c_int64), ("batch_stride_c", c_int64),
        ("C2", c_void_p), ("ldc2", c_int64), ("batch_stride_c2", c_int64),
        ("AUX", c_void_p), ("ldaux", c_int64), ("batch_stride_aux", c_int64),
        ("bias", c_void_p),
        ("max_ctas", c_int32), ("debug_direct_store", c_int32),
        ("drelu_scale", c_float), ("reserved0", c_int32),
        ("a_colsum", c_void_p),
    ]


class VbAttnDesc(Structure):
    _fields_ = [
        ("B", c_int32), ("H", c_int32), ("S", c_int32), ("head_dim", c_int32),
        ("tok_stride", c_int64), ("batch_stride", c_int64),
        ("q", c_void_p), ("k", c_void_p), ("v", c_void_p),
        ("ldq", c_int64), ("ldk", c_int64), ("ldv", c_int64),
        ("o", c_void_p), ("ldo", c_int64),
        ("lse", c_void_p), ("key_padding_mask", c_void_p),
        ("dout", c_void_p), ("lddo", c_int64),
        ("delta", c_void_p),
        ("dq", c_void_p), ("dk", c_void_p), ("dv", c_void_p),
        ("lddq", c_int64), ("lddk", c_int64), ("lddv", c_int64),
        ("dropout_p", c_float), ("dropout_stream", c_uint32), ("dropout_seed", c_void_p),
        ("dqkv_colsum", c_void_p),
        ("S_kv", c_int32), ("reserved0", c_int32),
        ("workspace", c_void_p), ("workspace_bytes", c_int64),
    ]


# symbol -> (restype, argtypes); this table is also what tests/test_abi.py checks against include/vitb200.h
SIGNATURES = {
    "vb_version": (c_int, []),
    "vb_last_error": (c_char_p, []),
    "vb_device_check": (c_int, [c_int]),
    "vb_sm_count": (c_int, [c_int]),
    "vb_gemm_bf16": (c_int, [POINTER(VbGemmDesc), c_void_p]),
    "vb_cast_f32_to_bf16": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "vb_cast_rows_bf16": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int32, c_void_p]),
    "vb_pos_embed_2d_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "vb_pos_embed_2d_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "vb_patchify": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "vb_unpatchify": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "vb_token_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "vb_colsum_bf16": (c_int, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "vb_vecmat_accum": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "vb_embed_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                             c_int32, c_int32, c_void_p]),
    "vb_cross_entropy": (c_int, [c_void_p, c_int64, c_void_p, c_int32, c_int32, c_void_p, c_float, c_void_p, c_int64, c_void_p,
                                 c_int64, c_float, c_void_p, c_void_p]),
    "vb_dwconv3x3_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                 c_void_p]),
    "vb_dwconv3x3_bwd_data": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32,
                                      c_void_p]),
    "vb_dwconv3x3_bwd_weight": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "vb_distill_loss": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int32, c_int32, c_int32, c_float,
                                c_float, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_float,
                                c_void_p, c_void_p]),
    "vb_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float,
                             c_float, c_int32, c_float, c_void_p, c_void_p]),
    "vb_debug_set_gemm_timeline": (c_int, [c_void_p]),
    "vb_debug_set_attn_timeline": (c_int, [c_void_p]),
    "vb_attention_fwd": (c_int, [POINTER(VbAttnDesc), c_void_p]),
    "vb_attention_bwd": (c_int, [POINTER(VbAttnDesc), c_void_p]),
    "vb_attention_workspace_bytes": (c_int64, [POINTER(VbAttnDesc), c_int32]),
    "vb_layernorm_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                 c_void_p, c_int32, c_int32, c_float, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "vb_layernorm_bwd": (c_int, [c_void_p, c_int32, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int32,
                                 c_int32, c_void_p, c_int64, c_void_p]),
    "vb_add_cast_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "vb_add_rows_bcast": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p]),
    "vb_add3": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "vb_dropout_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_float,
                               c_void_p, c_uint32, c_void_p]),
    "vb_dropout_bf16_pair": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_float, c_void_p, c_uint32, c_void_p]),
    "vb_dropout_mask_u8": (c_int, [c_void_p, c_int64, c_float, c_void_p, c_uint32, c_void_p]),
}

_lib = None


def load(build_if_missing=True):
    """Loads (building first if needed and possible) libvitb200.so and declares every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise VbError(f"{LIB_PATH} is missing; run `python -m vitb200.build`")
        from . import build as _build
        _build.build()
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc == 0:
        from . import ops
        ops.LAUNCHES["n"] += ops._KERNELS_PER_CALL.get(what, 1)
        return
    if rc != 0:
        msg = load().vb_last_error()
        raise VbError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")

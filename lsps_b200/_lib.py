"""ctypes binding of liblsps_b200.so (C ABI declared in include/lsps_b200.h).

The library is the product's only compute path: if it is missing and cannot be built the import fails loudly --
there is no torch / CPU fallback.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "liblsps_b200.so")

CONV_S1, CONV_S2, DECONV_S2, DECONV4_S2, CONV1X1 = 0, 1, 2, 3, 4
EP_BIAS, EP_LRELU, EP_MASK, EP_ADD, EP_STATS, EP_INBWD = 1, 2, 4, 8, 16, 32
ACT_NONE, ACT_LRELU, ACT_SOFTPLUS = 0, 1, 2


class ConvShape(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("kind", "n", "h", "w", "cin", "cout")]


class ConvExt(C.Structure):
    """lsps_conv_ext (include/lsps_b200.h): optional extras of lsps_conv_{fwd,dgrad}_ex"""
    _fields_ = [("w2", C.c_void_p), ("bias2", C.c_void_p), ("n_split", C.c_int), ("sums", C.c_void_p),
                ("in_a", C.c_void_p), ("bsums", C.c_void_p), ("w_lo", C.c_void_p),
                ("split", C.c_int), ("groups", C.c_int),
                ("head_w", C.c_void_p), ("head_b", C.c_void_p), ("head_out", C.c_void_p), ("head_target", C.c_void_p),
                ("head_t0", C.c_longlong), ("head_tn", C.c_longlong), ("head_scale", C.c_float),
                ("head_dout", C.c_void_p), ("head_acc", C.c_void_p)]


_vp, _i, _f, _ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong
_SH = C.POINTER(ConvShape)
# name -> argtypes AFTER the leading ctx; every function also ends with the stream (appended below)
_SIGS = {
    "lsps_conv_fwd": [_SH, _vp, _vp, _vp, _vp, _i, _f],
    "lsps_conv_dgrad": [_SH, _vp, _vp, _vp, _vp, _vp, _i, _f],
    "lsps_conv_fwd_grouped": [_SH, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _f],
    "lsps_conv_dgrad_grouped": [_SH, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _f],
    "lsps_conv_fwd_ex": [_SH, _vp, _vp, _vp, _vp, _i, _f, _vp],
    "lsps_conv_dgrad_ex": [_SH, _vp, _vp, _vp, _vp, _vp, _i, _f, _vp],
    "lsps_conv_wgrad": [_SH, _vp, _vp, _vp],
    "lsps_conv_wgrad_split": [_SH, _vp, _vp, _vp],
    "lsps_conv_wgrad_grouped": [_SH, _vp, _vp, _vp, _i],
    "lsps_stem_fwd_split": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f],
    "lsps_stem_wgrad_split": [_vp, _vp, _vp, _vp, _i, _i, _i, _i],
    "lsps_stem_dgrad_split": [_vp, _vp, _vp, _i, _i, _i, _i, _i],
    "lsps_l1_feat_split": [_vp, _vp, _vp, _vp, _f, _vp, _ll, _i],
    "lsps_dhead_fwd_split": [_vp, _vp, _vp, _vp, _ll, _i],
    "lsps_dhead_bwd_split": [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _i],
    "lsps_mask_to_bf16_split": [_vp, _vp, _vp, _f, _ll, _i],
    "lsps_colsum_bf16_split": [_vp, _ll, _i, _vp],
    "lsps_pack_dgrad_multi": [_vp, _vp, _vp, _vp, _i, _i],
    "lsps_adam_ex": [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _f, _i, _f, _vp],
    "lsps_f32_split_bf16": [_vp, _vp, _vp, _ll],
    "lsps_norm_apply_fwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _vp, _vp],
    "lsps_norm_bwd_stats": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp],
    "lsps_norm_bwd_apply": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp],
    "lsps_bn_running_update": [_vp, _vp, _vp, _i, _f, _f],
    "lsps_bn_running_to_sums": [_vp, _vp, _vp, _i, _f],
    "lsps_norm_reduce_images": [_vp, _vp, _i, _i],
    "lsps_colsum_bf16": [_vp, _ll, _i, _vp],
    "lsps_stem_fwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f],
    "lsps_stem_wgrad": [_vp, _vp, _vp, _vp, _i, _i, _i, _i],
    "lsps_stem_dgrad": [_vp, _vp, _vp, _i, _i, _i, _i, _i],
    "lsps_head_fwd": [_vp, _vp, _vp, _vp, _ll],
    "lsps_head_fwd_l1": [_vp, _vp, _vp, _vp, _ll, _vp, _ll, _ll, _f, _vp, _vp],
    "lsps_head_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _f],
    "lsps_instnorm_fwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f],
    "lsps_instnorm_bwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp],
    "lsps_instnorm_bwd_grouped": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _i],
    "lsps_noise_kl_fwd": [_vp, _vp, _vp, _vp, _ll],
    "lsps_noise_kl_philox": [_vp, _vp, _vp, _ll, C.c_ulonglong, C.c_ulonglong],
    "lsps_memset": [_vp, _i, _ll],
    "lsps_memcpy": [_vp, _vp, _ll],
    "lsps_axpy_bf16": [_vp, _vp, _f, _vp, _ll],
    "lsps_l2_bf16": [_vp, _vp, _vp, _f, _vp, _ll],
    "lsps_l1_f32": [_vp, _vp, _vp, _f, _i, _vp, _ll],
    "lsps_l1_feat": [_vp, _vp, _vp, _vp, _f, _vp, _ll],
    "lsps_dhead_fwd": [_vp, _vp, _vp, _vp, _ll, _i],
    "lsps_bce_logits": [_vp, _f, _f, _vp, _vp, _ll],
    "lsps_dhead_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _ll, _i],
    "lsps_dhead_bce": [_vp, _vp, _vp, _ll, _i, _i, _ll, _i, C.POINTER(C.c_float), C.POINTER(C.c_int), _f, _vp, _vp, _vp, _vp,
                       _vp],
    "lsps_mask_to_bf16": [_vp, _vp, _vp, _f, _ll],
    "lsps_linear_fwd": [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _f],
    "lsps_linear_bwd": [_vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i],
    "lsps_act_bwd": [_vp, _vp, _i, _f, _ll],
    "lsps_mse": [_vp, _vp, _vp, _f, _vp, _ll],
    "lsps_vae_reparam": [_vp, _vp, _vp, _vp, _vp, _ll],
    "lsps_vae_reparam_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _f, _ll],
    "lsps_vae_step": [_vp, _vp, C.POINTER(_vp), C.POINTER(_vp), _vp, _vp, _i, _i, _i, _i, _f, _f, _f],
    "lsps_adam": [_vp, _vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _f, _i, _f, _vp],
    "lsps_pack_dgrad": [_vp, _vp, _i, _i, _i],
    "lsps_f32_to_bf16": [_vp, _vp, _ll],
    "lsps_joint_errors": [_vp, _vp, _vp, _i, _i, _f, _f, _f, _vp, _vp, _i],
    "lsps_bf16_to_f32": [_vp, _vp, _ll],
    "lsps_augment_crops": [_vp, _vp, _vp, _vp, _i],
}
EXPORTS = sorted(list(_SIGS) + ["lsps_ctx_create", "lsps_ctx_destroy", "lsps_last_error", "lsps_abi_version",
                                "lsps_launch_count", "lsps_aug_sample_bytes"])


class LspsError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        try:
            from . import build as _build
            _build.build()
        except Exception as e:  # noqa
            raise ImportError("lsps_b200: %s is missing and could not be built (%r); there is no fallback path"
                              % (LIB_PATH, e))
    lib = C.CDLL(LIB_PATH)
    missing = [n for n in _SIGS if not hasattr(lib, n)]
    if missing:
        # a stale library from an older source tree: rebuild once (nvcc cross-compiles without a GPU), load the new
        # file under a fresh handle; still missing -> fail loudly
        from . import build as _build
        _build.build(force=True)
        import shutil
        import tempfile
        tmp = os.path.join(tempfile.mkdtemp(prefix="lsps_lib_"), "liblsps_b200.so")
        shutil.copy(LIB_PATH, tmp)
        lib = C.CDLL(tmp)
        missing = [n for n in _SIGS if not hasattr(lib, n)]
        if missing:
            raise ImportError("lsps_b200: %s lacks %s even after a rebuild" % (LIB_PATH, missing[:4]))
    lib.lsps_last_error.restype = C.c_char_p
    lib.lsps_last_error.argtypes = [_vp]
    lib.lsps_ctx_create.argtypes = [C.POINTER(_vp), _i]
    lib.lsps_ctx_destroy.argtypes = [_vp]
    lib.lsps_launch_count.restype = _ll
    lib.lsps_launch_count.argtypes = [_vp]
    lib.lsps_aug_sample_bytes.restype = _i
    lib.lsps_aug_sample_bytes.argtypes = []
    for name, sig in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = [_vp] + sig + [_vp]
        fn.restype = _i
    return lib


_lib = _load()


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


class Context:
    """One per (process, device).  Methods mirror the C entry points minus the `lsps_` prefix; ctx and the current
    torch stream are supplied automatically; a non-zero return raises LspsError."""

    def __init__(self, device=0):
        if not torch.cuda.is_available():
            raise LspsError("lsps_b200 needs a CUDA device (sm_100a); no CPU fallback exists")
        self.device = device
        self._ctx = _vp()
        rc = _lib.lsps_ctx_create(C.byref(self._ctx), device)
        if rc != 0:
            raise LspsError("lsps_ctx_create failed (%d): device %d is not a usable sm_100 GPU" % (rc, device))

    def __getattr__(self, name):
        fn = getattr(_lib, "lsps_" + name)
        ctx = self._ctx

        def call(*args):
            st = torch.cuda.current_stream().cuda_stream
            rc = fn(ctx, *args, st)
            if rc != 0:
                raise LspsError("lsps_%s failed (%d): %s" % (name, rc, _lib.lsps_last_error(ctx).decode()))
        self.__dict__[name] = call
        return call

    def launch_count(self):
        return int(_lib.lsps_launch_count(self._ctx))

    def __del__(self):
        try:
            if self._ctx:
                _lib.lsps_ctx_destroy(self._ctx)
        except Exception:  # noqa
            pass


_contexts = {}


def context(device=None):
    if device is None:
        device = torch.cuda.current_device()
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]

"""ctypes binding of libdta_b200.so (include/dta_b200.h) and the in-tree nvcc build.

There is deliberately no CPU fallback: if the library cannot be built/loaded or there is
no sm_100 device, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libdta_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

DTA_OK = 0
NET_HANG2020, NET_SPECTRAL, NET_SPATIAL, NET_VANILLA, NET_SPECTRAL_PAIR, NET_SPATIAL_PAIR = 0, 1, 2, 3, 4, 5


class Shape(C.Structure):
    _fields_ = [("net_kind", C.c_int32), ("batch", C.c_int32), ("bands", C.c_int32),
                ("classes", C.c_int32), ("training", C.c_int32)]


class ConvBlock(C.Structure):
    _fields_ = [("conv_w", C.c_void_p), ("conv_b", C.c_void_p), ("bn_w", C.c_void_p), ("bn_b", C.c_void_p),
                ("bn_rm", C.c_void_p), ("bn_rv", C.c_void_p), ("bn_nbt", C.c_void_p)]


class Attention(C.Structure):
    _fields_ = [("pool_w", C.c_void_p), ("pool_b", C.c_void_p), ("w0", C.c_void_p), ("b0", C.c_void_p),
                ("w1", C.c_void_p), ("b1", C.c_void_p)]


class Branch(C.Structure):
    _fields_ = [("conv", ConvBlock * 3), ("attn", Attention * 3), ("fc_w", C.c_void_p * 3), ("fc_b", C.c_void_p * 3)]


class Tensors(C.Structure):
    _fields_ = [("alpha", C.c_void_p), ("branch", Branch * 2)]


class Sizes(C.Structure):
    _fields_ = [("saved_bytes", C.c_size_t), ("workspace_fwd", C.c_size_t), ("workspace_bwd", C.c_size_t),
                ("n_heads", C.c_int32)]


class Plane(C.Structure):
    _fields_ = [("batch", C.c_int32), ("channels", C.c_int32), ("height", C.c_int32), ("width", C.c_int32)]


class AdamHyper(C.Structure):
    _fields_ = [("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("weight_decay", C.c_double), ("step", C.c_int64)]


ATTN_SPECTRAL, ATTN_SPATIAL = 1, 2


class MetadataTensors(C.Structure):
    _fields_ = [("embedding", C.c_void_p), ("bn_w", C.c_void_p), ("bn_b", C.c_void_p), ("bn_rm", C.c_void_p),
                ("bn_rv", C.c_void_p), ("bn_nbt", C.c_void_p), ("mlp_w", C.c_void_p), ("mlp_b", C.c_void_p),
                ("fc_w", C.c_void_p), ("fc_b", C.c_void_p)]


class StageTime(C.Structure):
    _fields_ = [("name", C.c_char * 40), ("total_ms", C.c_double), ("calls", C.c_int64)]


EXPORTS = ["dta_abi_version", "dta_create", "dta_destroy", "dta_last_error", "dta_set_option", "dta_get_option",
           "dta_profile_read", "dta_query_sizes", "dta_saved_region", "dta_forward", "dta_backward", "dta_loss_workspace_bytes",
           "dta_cross_entropy_heads", "dta_preprocess_crops", "dta_grad_allreduce_sizes", "dta_grad_allreduce",
           "dta_plane_mean", "dta_plane_mean_backward", "dta_conv_module_workspace_bytes", "dta_conv_module_forward",
           "dta_conv_module_backward", "dta_attention_sizes", "dta_attention_forward", "dta_attention_backward",
           "dta_classifier_forward", "dta_classifier_backward", "dta_adam_step", "dta_forward_pair", "dta_crops_nonzero",
           "dta_ensemble_mean", "dta_metadata_sizes", "dta_metadata_forward", "dta_metadata_backward", "dta_set_grad_exchange", "dta_set_update_gate", "dta_train_step"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(os.path.dirname(HERE), "include", "dta_b200.h")]
    return any(os.path.getmtime(s) > t for s in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libdta_b200.so next to this file (in-tree, so the
    binary travels with the repo snapshot)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libdta_b200.so (no CPU fallback exists)")
    cu = [s for s in sources() if s.endswith(".cu")]
    tmp = f"{LIB_PATH}.{os.getpid()}.tmp"      # several ranks may find the binary stale at once: private output, atomic replace
    cmd = [nvcc] + NVCC_FLAGS + cu + ["-o", tmp]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB_PATH)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None
_lock = threading.Lock()


def lib():
    """Load (building first when sources are newer and nvcc exists) the C-ABI library."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        override = os.environ.get("DTA_B200_LIB")      # a prebuilt libdta_b200.so to load instead (A/B runs of kernel variants)
        if override:
            if not os.path.exists(override):
                raise RuntimeError(f"DTA_B200_LIB={override} does not exist")
            path = override
        else:
            if needs_build():
                if os.path.exists(LIB_PATH) and not (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
                    pass  # prebuilt binary on a box without nvcc: use it
                else:
                    build()
            path = LIB_PATH
        L = C.CDLL(path)
        L.dta_abi_version.restype = C.c_int
        L.dta_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.dta_create.restype = C.c_int
        L.dta_destroy.argtypes = [C.c_void_p]
        L.dta_destroy.restype = None
        L.dta_last_error.argtypes = [C.c_void_p]
        L.dta_last_error.restype = C.c_char_p
        L.dta_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
        L.dta_set_option.restype = C.c_int
        L.dta_get_option.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64)]
        L.dta_get_option.restype = C.c_int
        L.dta_profile_read.argtypes = [C.c_void_p, C.POINTER(StageTime), C.c_int, C.POINTER(C.c_int), C.c_int]
        L.dta_profile_read.restype = C.c_int
        L.dta_query_sizes.argtypes = [C.POINTER(Shape), C.POINTER(Sizes)]
        L.dta_query_sizes.restype = C.c_int
        L.dta_saved_region.argtypes = [C.POINTER(Shape), C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.dta_saved_region.restype = C.c_int
        L.dta_forward.argtypes = [C.c_void_p, C.POINTER(Shape), C.c_void_p, C.POINTER(Tensors),
                                  C.POINTER(C.c_void_p * 6), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.dta_forward.restype = C.c_int
        L.dta_backward.argtypes = [C.c_void_p, C.POINTER(Shape), C.c_void_p, C.POINTER(Tensors), C.c_void_p,
                                   C.POINTER(C.c_void_p * 6), C.c_void_p, C.POINTER(Tensors), C.c_void_p,
                                   C.c_void_p, C.c_void_p]
        L.dta_backward.restype = C.c_int
        L.dta_train_step.argtypes = [C.c_void_p, C.POINTER(Shape), C.c_void_p, C.POINTER(Tensors), C.c_void_p, C.c_void_p,
                                     C.POINTER(C.c_void_p * 6), C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p * 6),
                                     C.POINTER(Tensors), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.dta_train_step.restype = C.c_int
        L.dta_loss_workspace_bytes.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_size_t)]
        L.dta_loss_workspace_bytes.restype = C.c_int
        L.dta_cross_entropy_heads.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p * 8), C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p * 8), C.c_void_p, C.c_void_p]
        L.dta_cross_entropy_heads.restype = C.c_int
        L.dta_preprocess_crops.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.dta_preprocess_crops.restype = C.c_int
        L.dta_grad_allreduce_sizes.argtypes = [C.c_size_t, C.c_size_t, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                               C.POINTER(C.c_size_t)]
        L.dta_grad_allreduce_sizes.restype = C.c_int
        L.dta_grad_allreduce.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p * 16), C.c_void_p, C.c_size_t, C.c_size_t,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
        L.dta_grad_allreduce.restype = C.c_int
        vp, sz, ci = C.c_void_p, C.c_size_t, C.c_int
        L.dta_plane_mean.argtypes = [vp, vp, sz, ci, vp, vp]
        L.dta_plane_mean_backward.argtypes = [vp, vp, sz, ci, vp, vp]
        L.dta_conv_module_workspace_bytes.argtypes = [C.POINTER(Plane), ci, C.POINTER(sz)]
        L.dta_conv_module_forward.argtypes = [vp, C.POINTER(Plane), ci, ci, ci, ci, vp, C.POINTER(ConvBlock), vp, vp, vp, vp]
        L.dta_conv_module_backward.argtypes = [vp, C.POINTER(Plane), ci, ci, ci, ci, vp, C.POINTER(ConvBlock), vp, vp, vp,
                                               C.POINTER(ConvBlock), vp, vp, vp]
        L.dta_attention_sizes.argtypes = [ci, C.POINTER(Plane), C.POINTER(sz), C.POINTER(sz), C.POINTER(sz)]
        L.dta_attention_forward.argtypes = [vp, ci, C.POINTER(Plane), vp, C.POINTER(Attention), vp, vp, vp, vp]
        L.dta_attention_backward.argtypes = [vp, ci, C.POINTER(Plane), vp, C.POINTER(Attention), vp, vp, vp, vp,
                                             C.POINTER(Attention), vp, vp]
        L.dta_classifier_forward.argtypes = [vp, ci, ci, ci, vp, vp, vp, vp, vp]
        L.dta_classifier_backward.argtypes = [vp, ci, ci, ci, vp, vp, vp, vp, vp, vp, vp]
        L.dta_adam_step.argtypes = [vp, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(AdamHyper), vp, vp, vp]
        L.dta_forward_pair.argtypes = [vp, C.POINTER(Shape), ci, vp, C.POINTER(Tensors), C.POINTER(C.c_void_p * 6), vp, vp, vp]
        L.dta_crops_nonzero.argtypes = [vp, ci, C.POINTER(C.c_void_p * 16), sz, vp, vp, vp]
        L.dta_ensemble_mean.argtypes = [vp, ci, C.POINTER(C.c_void_p * 16), vp, ci, ci, ci, vp, vp]
        L.dta_set_update_gate.argtypes = [vp, vp]
        L.dta_set_grad_exchange.argtypes = [vp, ci, ci, C.POINTER(C.c_void_p * 16), vp, sz, sz, vp]
        L.dta_metadata_sizes.argtypes = [ci, ci, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.dta_metadata_forward.argtypes = [vp, ci, ci, ci, ci, vp, vp, C.POINTER(MetadataTensors), vp, C.c_uint64, vp, vp, vp]
        L.dta_metadata_backward.argtypes = [vp, ci, ci, ci, ci, vp, vp, C.POINTER(MetadataTensors), vp, vp, vp,
                                            C.POINTER(MetadataTensors), vp, vp, vp]
        for name in EXPORTS[16:]:
            getattr(L, name).restype = C.c_int
        _lib = L
        return L


class DtaError(RuntimeError):
    pass


_ctxs = {}


def context(device_index: int):
    """One dta_ctx per CUDA device per process."""
    L = lib()
    h = _ctxs.get(device_index)
    if h is None:
        out = C.c_void_p()
        rc = L.dta_create(C.byref(out), device_index)
        if rc != DTA_OK:
            raise DtaError(f"dta_create failed ({rc}): {L.dta_last_error(None).decode()}")
        h = out
        _ctxs[device_index] = h
    return h


def check(ctx, rc: int, what: str):
    if rc != DTA_OK:
        msg = lib().dta_last_error(ctx).decode()
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise DtaError(f"{what} failed ({rc}): {msg}")


def query_sizes(net_kind: int, batch: int, bands: int, classes: int, training: bool) -> Sizes:
    s = Shape(net_kind, batch, bands, classes, int(training))
    out = Sizes()
    rc = lib().dta_query_sizes(C.byref(s), C.byref(out))
    if rc != DTA_OK:
        raise ValueError(f"dta_query_sizes rejected shape (kind={net_kind}, batch={batch}, bands={bands}, classes={classes})")
    return out


def saved_region(net_kind: int, batch: int, bands: int, classes: int, training: bool, block: int, region: int = 0):
    """(byte offset, float count) of convolution block ``block``'s output (region 0) / BatchNorm scale (1) / shift (2) inside a
    forward's ``saved`` buffer (diagnostic)."""
    s = Shape(net_kind, batch, bands, classes, int(training))
    off, n = C.c_size_t(), C.c_size_t()
    if lib().dta_saved_region(C.byref(s), block, region, C.byref(off), C.byref(n)) != DTA_OK:
        raise ValueError("dta_saved_region rejected the shape / block")
    return off.value, n.value


# Diagnostics (parity tests): when True, every fused forward leaves a reference to its ``saved`` buffer in the module's
# spec (``model.fused_spec().last_saved``) so that the convolution outputs can be read back with ``saved_region``.
KEEP_SAVED = False
# tests: the gradient buffers handed to dta_backward are pre-filled with NaN (default: left uninitialised -- the library writes all
# of them), so a gradient the library failed to write shows up as NaN.  Also switched on by the environment variable DTA_POISON_GRADS=1.
POISON_GRADS = os.environ.get("DTA_POISON_GRADS", "0") == "1"


def set_update_gate(device_index: int, flag_ptr):
    """Registers (or, with None / 0, clears) the device flag that gates BatchNorm running-statistics updates (dta_set_update_gate)."""
    ctx = context(device_index)
    check(ctx, lib().dta_set_update_gate(ctx, flag_ptr or None), "dta_set_update_gate")


def set_option(device_index: int, key: str, value: int):
    ctx = context(device_index)
    check(ctx, lib().dta_set_option(ctx, key.encode(), int(value)), f"dta_set_option({key})")


def get_option(device_index: int, key: str) -> int:
    ctx = context(device_index)
    v = C.c_int64()
    rc = lib().dta_get_option(ctx, key.encode(), C.byref(v))
    if rc != DTA_OK:
        raise ValueError(f"unknown option {key}")
    return v.value


def profile_read(device_index: int, reset: bool = True):
    """{stage: (total_ms, calls)} recorded since the last reset (option "profile" must be 1)."""
    ctx = context(device_index)
    rows = (StageTime * 64)()
    n = C.c_int()
    check(ctx, lib().dta_profile_read(ctx, rows, 64, C.byref(n), int(reset)), "dta_profile_read")
    return {rows[i].name.decode(): (rows[i].total_ms, rows[i].calls) for i in range(min(n.value, 64))}

"""ctypes binding of libdimo_b200.so (include/dimo_b200.h).  No CPU fallback: if the library is
missing or a call fails this raises, it never routes around the CUDA path."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# DIMO_LIB: load another build of the same ABI (instrumented / experimental builds during bring-up)
LIB_PATH = os.environ.get("DIMO_LIB") or os.path.join(_HERE, "lib", "libdimo_b200.so")

_lib = None

c_int, c_i64, c_f32, c_vp, c_sz = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t
c_f64 = ctypes.c_double

# name -> (restype, argtypes); mirrors include/dimo_b200.h one to one
_SIGS = {
    "dimo_abi_version": (c_int, []),
    "dimo_last_error": (ctypes.c_char_p, []),
    "dimo_device_info": (c_int, [c_vp]),
    "dimo_raster_bin_temp_bytes": (c_sz, [c_int] * 4),
    "dimo_raster_preprocess": (c_int, [c_int] * 6 + [c_f32, c_int, c_vp, c_vp] + [c_vp, c_i64] * 6 + [c_vp] * 7 + [c_vp, c_vp]),
    "dimo_raster_bin": (c_int, [c_int] * 4 + [c_i64] + [c_vp] * 4 + [c_vp, c_sz, c_vp, c_vp, c_vp]),
    "dimo_raster_packed_value_bits": (c_int, [c_int] * 4),
    "dimo_raster_blend_fwd": (c_int, [c_int] * 5 + [c_vp] * 11),
    "dimo_raster_blend_bwd": (c_int, [c_int] * 5 + [c_vp] * 12),
    "dimo_raster_preprocess_bwd": (c_int, [c_int] * 6 + [c_f32, c_int, c_vp, c_vp] + [c_vp, c_i64] * 5 + [c_vp] * 9 + [c_int, c_vp]),
    "dimo_knn": (c_int, [c_int] * 3 + [c_vp] * 5),
    "dimo_dist3nn": (c_int, [c_int, c_vp, c_vp, c_vp]),
    "dimo_fps": (c_int, [c_int] * 4 + [c_vp] * 4),
    "dimo_ball_query": (c_int, [c_int] * 4 + [c_f32] + [c_vp] * 5),
    "dimo_chamfer_fwd": (c_int, [c_int] * 2 + [c_vp] * 6 + [c_f32, c_vp]),
    "dimo_chamfer_bwd": (c_int, [c_int] + [c_vp] * 4 + [c_f32] + [c_vp] * 3),
    "dimo_arap_connectivity": (c_int, [c_int] * 3 + [c_f32] + [c_vp] * 4),
    "dimo_arap_energy": (c_int, [c_int] * 3 + [c_vp] * 6),
    "dimo_linear_fwd": (c_int, [c_int] * 3 + [c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_int, c_vp]),
    "dimo_linear_bwd_data": (c_int, [c_int] * 3 + [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_i64, c_int, c_vp]),
    "dimo_linear_bwd_weight": (c_int, [c_int] * 3 + [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "dimo_linear_tc": (c_int, [c_int] * 3 + [c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_vp]),
    "dimo_linear_wgrad_tc": (c_int, [c_int] * 3 + [c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "dimo_linear_wgrad_tc_grouped": (c_int, [c_int, c_int] + [c_vp] * 10 + [c_vp]),
    "dimo_tc_debug_set": (c_int, [c_int, c_int]),
    "dimo_debug_max_sort_clusters": (c_int, []),
    "dimo_set_deterministic": (c_int, [c_int]),
    "dimo_get_deterministic": (c_int, []),
    "dimo_fixed_to_float": (c_int, [c_i64, c_vp, c_vp, c_int, c_vp]),
    "dimo_timenet_workspace_bytes": (c_sz, [c_int] * 3),
    "dimo_timenet_layout": (c_int, [c_int] * 3 + [c_vp]),
    "dimo_timenet_debug_stamps": (c_int, [c_vp]),
    "dimo_timenet_last_launches": (c_int, []),
    "dimo_timenet_fwd": (c_int, [c_int] * 3 + [c_vp] * 6 + [c_sz] + [c_vp] * 3),
    "dimo_timenet_bwd": (c_int, [c_int] * 3 + [c_vp] * 2 + [c_sz] + [c_vp] * 7),
    "dimo_timenet_embed_fwd": (c_int, [c_int] * 3 + [c_vp] * 4 + [c_i64, c_vp]),
    "dimo_timenet_embed_bwd": (c_int, [c_int] * 3 + [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "dimo_lbs_fwd": (c_int, [c_int] * 4 + [c_vp] * 11),
    "dimo_lbs_bwd": (c_int, [c_int] * 4 + [c_vp] * 17),
    "dimo_ssim_fwd": (c_int, [c_int] * 5 + [c_vp] * 6 + [c_f32] * 3 + [c_vp]),
    "dimo_ssim_bwd": (c_int, [c_int] * 5 + [c_vp] * 3 + [c_f32] * 3 + [c_vp] * 4),
    "dimo_smooth_fwd": (c_int, [c_int] * 4 + [c_vp] * 5 + [c_f32] * 4 + [c_vp]),
    "dimo_smooth_bwd": (c_int, [c_int] * 4 + [c_vp] * 3 + [c_f32] * 4 + [c_vp, c_vp, c_int, c_vp, c_vp, c_vp]),
    "dimo_segment_sum": (c_int, [c_int, c_int, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "dimo_sqdiff_sum": (c_int, [c_i64] + [c_vp] * 4 + [c_f32, c_vp]),
    "dimo_adam_step": (c_int, [c_i64] + [c_vp] * 4 + [c_int, c_vp, c_vp, c_f64, c_f64, c_f32, c_int, c_vp, c_vp, c_vp]),
    "dimo_transpose_grouped": (c_int, [c_int] + [c_vp] * 5),
    "dimo_gt_fetch": (c_int, [c_int] * 6 + [c_vp] * 5),
}


def exported_symbols():
    return sorted(_SIGS)


def lib():
    """Loads the library (once).  Raises RuntimeError with build instructions if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m dimo_b200.build` "
                "(dimo_b200 has no CPU or PyTorch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name, None)    # a missing symbol fails loudly at call time (call())
            if fn is not None:
                fn.restype = res
                fn.argtypes = args
        if L.dimo_abi_version() != 2:
            raise RuntimeError("libdimo_b200.so ABI version mismatch")
        # bring-up knobs for A/B runs, e.g. DIMO_KNOBS="3=1,2=148" (see dimo_tc_debug_set in include/dimo_b200.h)
        for kv in filter(None, os.environ.get("DIMO_KNOBS", "").split(",")):
            k, v = kv.split("=")
            if L.dimo_tc_debug_set(int(k), int(v)) != 0:
                raise RuntimeError(f"bad DIMO_KNOBS entry {kv!r}")
        if os.environ.get("DIMO_DETERMINISTIC", "0") not in ("", "0"):
            L.dimo_set_deterministic(1)
        _lib = L
    return _lib


def deterministic():
    """True when gradient accumulation runs in the order-independent fixed-point mode (DIMO_DETERMINISTIC=1 or
    set_deterministic(True)); the host side then hands int64 accumulation buffers to the backward kernels."""
    return bool(lib().dimo_get_deterministic())


def set_deterministic(on):
    lib().dimo_set_deterministic(1 if on else 0)


def grad_sink(p):
    """The tensor a backward kernel may accumulate (+=) the gradient of parameter `p` into, or None: p.grad when it is a
    preallocated contiguous fp32 buffer (the views of dist.FlatGradReducer's flat buffer) and the accumulation is fp32
    (deterministic mode sums in int64 buffers).  Saves the zero-filled temporary and autograd's AccumulateGrad add."""
    if p is None or not p.requires_grad or not p.is_leaf or deterministic():
        return None
    g = p.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.shape != p.shape or not g.is_cuda:
        return None
    return g


def acc_zeros(shape, device):
    """accumulation target for a backward kernel: fp32 zeros, or int64 zeros in deterministic mode"""
    return torch.zeros(shape, dtype=torch.int64 if deterministic() else torch.float32, device=device)


def acc_result(buf, into=None):
    """fp32 view of an accumulation target (deterministic mode: one dimo_fixed_to_float launch); `into`: add to it"""
    if buf.dtype != torch.int64:
        if into is None:
            return buf
        into.add_(buf)
        return into
    out = into if into is not None else torch.empty(buf.shape, dtype=torch.float32, device=buf.device)
    call("dimo_fixed_to_float", buf.numel(), ptr(buf), ptr(out), 1 if into is not None else 0, stream())
    return out


def check(rc):
    if rc != 0:
        raise RuntimeError("libdimo_b200: " + lib().dimo_last_error().decode())


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  Refuses CPU tensors: the product path is CUDA only."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("dimo_b200 ops need CUDA tensors (there is no CPU path)")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


# hand-written kernels launched per C-ABI call
_OWN_LAUNCHES = {
    "dimo_raster_preprocess": 2, "dimo_raster_bin": 4, "dimo_raster_blend_fwd": 1, "dimo_raster_blend_bwd": 1,
    "dimo_raster_preprocess_bwd": 1, "dimo_knn": 1, "dimo_dist3nn": 1,
    "dimo_fps": 1, "dimo_ball_query": 1, "dimo_chamfer_fwd": 1, "dimo_chamfer_bwd": 1,
    "dimo_arap_connectivity": 1, "dimo_arap_energy": 1, "dimo_linear_fwd": 1,
    "dimo_linear_bwd_data": 1, "dimo_linear_tc": 1, "dimo_linear_wgrad_tc": 1, "dimo_linear_wgrad_tc_grouped": 1, "dimo_linear_bwd_weight": 1, "dimo_timenet_embed_fwd": 1,
    "dimo_timenet_embed_bwd": 1, "dimo_lbs_fwd": 1, "dimo_lbs_bwd": 1, "dimo_ssim_fwd": 1, "dimo_ssim_bwd": 1,
    "dimo_timenet_fwd": None, "dimo_timenet_bwd": None,      # asked from the library after the call (chained or per layer)
    "dimo_fixed_to_float": 1,
    "dimo_sqdiff_sum": 1, "dimo_smooth_fwd": 1, "dimo_smooth_bwd": 1, "dimo_segment_sum": 1, "dimo_adam_step": 1, "dimo_transpose_grouped": 1, "dimo_gt_fetch": 1,
}


class _Profile:
    """Optional per-call CUDA-event timing on the launching stream (bench.py's live roofline numbers)."""

    def __init__(self):
        self.reset(False)

    def reset(self, enabled=False):
        self.enabled = enabled
        self.events = []          # (name, start, end)
        self.counts = {}
        self.dyn_launches = 0     # launches of calls whose count depends on the shape (asked from the library)
        self.extra = getattr(self, "extra", {})

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1 in self.events:
            rec = out.setdefault(name, {"ms": 0.0, "calls": 0})
            rec["ms"] += e0.elapsed_time(e1)
            rec["calls"] += 1
        return out

    def kernel_launches_per_step(self, steps):
        n = sum((_OWN_LAUNCHES.get(k, 0) or 0) * v for k, v in self.counts.items()) + self.dyn_launches
        return n // max(steps, 1)


PROFILE = _Profile()
SYNC_EVERY_CALL = False      # debug: device-synchronise after every C-ABI call


def call(name, *args):
    fn = getattr(lib(), name, None)
    if fn is None:
        raise RuntimeError(f"libdimo_b200.so does not export {name}: rebuild with `python -m dimo_b200.build`")
    if PROFILE.enabled and not torch.cuda.is_current_stream_capturing():
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        check(fn(*args))
        e1.record()
        PROFILE.events.append((name, e0, e1))
        PROFILE.counts[name] = PROFILE.counts.get(name, 0) + 1
        if name in _OWN_LAUNCHES and _OWN_LAUNCHES[name] is None:
            PROFILE.dyn_launches += int(lib().dimo_timenet_last_launches())
    else:
        check(fn(*args))
    if SYNC_EVERY_CALL:
        torch.cuda.synchronize()

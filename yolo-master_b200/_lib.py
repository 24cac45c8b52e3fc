"""ctypes binding of libym_b200.so (C ABI: include/ym_b200.h).  Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libym_b200.so")

vp, ci, cf, cll = C.c_void_p, C.c_int, C.c_float, C.c_longlong

# name -> (restype, argtypes).  Kept in one table so tests can check it against the header.
SIGNATURES = {
    "ym_last_error": (C.c_char_p, []),
    "ym_version": (ci, []),
    "ym_device_info": (ci, [C.POINTER(ci), C.POINTER(ci), C.POINTER(ci), C.POINTER(cll)]),
    "ym_conv2d_nhwc": (ci, [vp, ci, ci, ci, ci, ci, vp, ci, vp, ci, ci, ci, ci, ci, vp, ci, ci, vp, ci, ci, vp]),
    "ym_conv2d_tc_supported": (ci, [ci, ci, ci, ci, ci, ci, ci]),
    "ym_set_tc_conv_version": (ci, [ci]),
    "ym_conv2d_tc": (ci, [vp, ci, ci, ci, ci, ci, vp, ci, vp, ci, ci, ci, ci, ci, vp, ci, ci, vp, ci, ci, vp]),
    "ym_stem_conv_nchw": (ci, [vp, ci, ci, ci, ci, ci, vp, vp, ci, vp, ci, vp]),
    "ym_set_stem_impl": (ci, [ci]),
    "ym_kernel_priority": (ci, []),
    "ym_set_kernel_priority": (ci, [ci]),
    "ym_set_small_conv_impl": (ci, [ci]),
    "ym_set_dwconv_tc": (ci, [ci]),
    "ym_set_conv2_epi_groups": (ci, [ci]),
    "ym_set_conv2_debug": (ci, [ci]),
    "ym_dwconv_nhwc": (ci, [vp, ci, ci, ci, ci, vp, vp, ci, ci, ci, ci, ci, ci, vp, ci, vp, ci, vp]),
    "ym_sppf_pool_nhwc": (ci, [vp, ci, ci, ci, ci, ci, ci, vp]),
    "ym_concat2_nhwc": (ci, [vp, ci, ci, ci, vp, ci, ci, vp, ci, ci, ci, ci, vp]),
    "ym_attention_fwd": (ci, [vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, cf, vp, ci, vp]),
    "ym_attention_fwd_tc": (ci, [vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, cf, vp, ci, vp]),
    "ym_attention_fwd_tc2": (ci, [vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, cf, vp, ci, vp]),
    "ym_attention_fwd_tc2_supported": (ci, [ci, ci, ci]),
    "ym_set_attention2_poly": (ci, [ci]),
    "ym_set_attention2_qtiles": (ci, [ci]),
    "ym_set_attention2_variant": (ci, [ci]),
    "ym_attention2_poly": (ci, []),
    "ym_set_attention_impl": (ci, [ci]),
    "ym_moe_ffn_supported": (ci, [ci, ci, ci]),
    "ym_moe_ffn_strips": (ci, [ci, ci]),
    "ym_moe_ffn_stats_floats": (cll, [ci, ci, ci]),
    "ym_moe_ffn": (ci, [ci, vp, ci, ci, ci, ci, ci, ci, vp, vp, ci, vp, vp, vp, vp, vp, ci, vp]),
    "ym_moe_combine_tc_supported": (ci, [ci, ci, ci]),
    "ym_moe_combine_tc": (ci, [vp, ci, ci, ci, ci, vp, vp, vp, vp, vp, ci, vp, ci, ci, vp]),
    "ym_router_blocks": (ci, [ci, ci, ci, vp]),
    "ym_router_partial": (ci, [vp, ci, ci, ci, ci, ci, ci, vp, ci, vp, vp, vp, vp]),
    "ym_moe_ffn_routed": (ci, [vp, ci, ci, ci, ci, ci, ci, vp, ci, vp, ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, ci, vp]),
    "ym_moe_ffn_gn": (ci, [vp, ci, ci, ci, ci, ci, ci, vp, vp, ci, vp, vp, ci, cf, cf, vp, vp, vp, vp, ci, vp]),
    "ym_moe_combine_tc_gn": (ci, [vp, ci, ci, ci, ci, vp, vp, vp, vp, ci, ci, cf, cf, vp, vp, vp, vp, ci, vp, ci, ci, vp]),
    "ym_gn_finalize_tiles": (ci, [vp, ci, ci, ci, ci, cf, cf, vp, vp, vp, vp, vp, vp, vp]),
    "ym_pdl_enabled": (ci, []),
    "ym_set_pdl": (ci, [ci]),
    "ym_set_attention_poly": (ci, [ci]),
    "ym_set_attention_chunked": (ci, [ci]),
    "ym_router_scratch_floats": (cll, [ci, ci, ci, ci, ci, ci]),
    "ym_router_topk": (ci, [vp, ci, ci, ci, ci, ci, ci, vp, ci, vp, vp, vp, vp, vp, ci, ci, vp, vp, vp, vp, vp]),
    "ym_moe_expert_gemm": (ci, [vp, ci, ci, ci, ci, ci, vp, ci, cll, vp, ci, vp, ci, vp, vp, vp, ci, vp]),
    "ym_moe_stats_floats": (cll, [ci, ci, ci]),
    "ym_gn_finalize": (ci, [vp, ci, ci, ci, ci, cf, cf, vp, vp, vp, vp, vp, vp, vp]),
    "ym_moe_combine": (ci, [vp, ci, ci, ci, ci, vp, ci, vp, vp, ci, vp, vp, ci, vp, ci, ci, vp]),
    "ym_esmoe_scratch_floats": (cll, [ci, ci, ci]),
    "ym_esmoe_route": (ci, [vp, ci, ci, ci, ci, vp, vp, ci, vp, vp, ci, ci, cf, vp, vp, vp, vp, vp]),
    "ym_esmoe_dwconv": (ci, [vp, ci, vp, ci, ci, ci, ci, ci, vp, ci, ci, vp, ci, vp]),
    "ym_esmoe_pointwise": (ci, [vp, ci, ci, ci, ci, vp, ci, cll, vp, ci, vp, vp, vp, ci, vp]),
    "ym_esmoe_combine": (ci, [vp, ci, vp, ci, vp, vp, vp, ci, ci, ci, ci, vp]),
    "ym_nms_scratch_bytes": (cll, [ci, ci]),
    "ym_nms_batched": (ci, [vp, ci, ci, ci, cf, cf, ci, ci, cf, ci, cf, cf, cf, vp, vp, vp, vp, vp]),
    "ym_nms_overflowed": (ci, [vp, ci, ci, vp]),
    "ym_nms_large_scratch_bytes": (cll, [ci, ci]),
    "ym_nms_batched_large": (ci, [vp, ci, ci, ci, cf, cf, ci, ci, cf, vp, vp, vp, vp, vp]),
    "ym_tc_gemm_nt": (ci, [vp, ci, vp, ci, vp, vp, ci, vp, ci, ci, ci, ci, ci, vp]),
    "ym_moe_dispatch_tc": (ci, [vp, ci, ci, ci, ci, vp, ci, cll, vp, vp, ci, ci, cf, cf, vp, ci, vp]),
    "ym_moe_dispatch_v2_supported": (ci, [ci, ci, ci, ci, ci, ci, ci]),
    "ym_moe_dispatch_v2": (ci, [vp, ci, ci, ci, ci, vp, ci, ci, vp, vp, ci, ci, cf, cf, vp, ci, vp]),
    "ym_moe_dispatch_v3_supported": (ci, [ci, ci, ci, ci, ci, ci, ci]),
    "ym_moe_dispatch_v3": (ci, [vp, ci, ci, ci, ci, vp, ci, ci, vp, vp, ci, ci, cf, cf, vp, ci, vp]),
    "ym_set_dispatch_debug": (None, [ci]),
    "ym_dispatch_debug_mask": (ci, []),
    "ym_set_dispatch_trace": (None, [vp]),
    "ym_ew_nhwc": (ci, [ci, vp, ci, vp, ci, vp, vp, vp, ci, ci, ci, ci, vp, ci, cll, ci, vp]),
    "ym_groupnorm_stats": (ci, [vp, ci, ci, ci, ci, ci, ci, cf, vp, vp, vp, vp, vp]),
    "ym_layernorm_nhwc": (ci, [vp, ci, vp, vp, cf, vp, ci, cll, ci, vp]),
    "ym_attn_small": (ci, [vp, ci, vp, ci, vp, ci, ci, ci, ci, ci, ci, cf, vp, ci, vp]),
    "ym_attn_window": (ci, [vp, ci, vp, ci, vp, ci, ci, ci, ci, ci, ci, ci, ci, vp, vp, vp, cf, vp, ci, vp]),
    "ym_deform_sample": (ci, [vp, ci, vp, ci, ci, ci, ci, ci, ci, ci, ci, vp, ci, vp]),
    "ym_token_router_scratch_floats": (cll, [ci, ci, ci]),
    "ym_token_router": (ci, [vp, ci, ci, ci, ci, vp, ci, ci, vp, vp, cf, vp, vp, ci, ci, vp, cf, vp, vp, vp, vp]),
    "ym_linear_attn_scratch_floats": (cll, [ci, ci, ci, ci]),
    "ym_linear_attn": (ci, [vp, ci, vp, ci, vp, ci, ci, ci, ci, ci, ci, ci, vp, cf, cf, vp, vp, ci, vp]),
    "ym_adaptive_avgpool_nhwc": (ci, [vp, ci, ci, ci, ci, ci, ci, ci, vp, ci, vp]),
    "ym_gate_router_scratch_floats": (cll, [ci, ci, ci, ci, ci, ci, ci]),
    "ym_gate_router": (ci, [vp, ci, ci, ci, ci, ci, ci, vp, vp, vp, vp, ci, vp, ci, vp, vp, ci, vp, vp, ci, cf, cf, cf, vp, cf, ci,
                            vp, vp, cf, vp, vp, vp, vp, vp, vp]),
    "ym_pixel_router": (ci, [vp, ci, ci, ci, ci, ci, ci, vp, vp, vp, ci, vp, ci, vp, vp, ci, vp, vp, ci, cf, cf, cf, ci, vp, vp, vp, vp, vp]),
    "ym_zero_cost_router_scratch_floats": (cll, [ci, ci]),
    "ym_zero_cost_router": (ci, [vp, ci, ci, ci, ci, ci, vp, ci, cf, vp, cf, ci, vp, vp, vp, vp, vp]),
    "ym_fc_gate": (ci, [vp, ci, ci, ci, vp, ci, vp, vp, ci, cf, cf, vp, vp]),
    "ym_gated_select_scratch_floats": (C.c_longlong, [ci, ci, ci]),
    "ym_gated_select": (ci, [vp, ci, ci, ci, ci, ci, ci, cf, vp, vp, ci, vp, vp, vp, vp, ci, vp]),
    "ym_ctx_mean3": (ci, [vp, ci, vp, ci, vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp, ci, vp]),
    "ym_gap_nhwc": (ci, [vp, ci, ci, ci, ci, vp, ci, vp]),
    "ym_latent_router": (ci, [ci, C.POINTER(vp), C.POINTER(ci), ci, ci, vp, vp, vp, cf, vp, vp, ci, vp, vp, vp, vp, ci, cf, vp, vp, vp]),
    "ym_classify_head": (ci, [vp, ci, ci, ci, vp, vp, ci, vp, vp, vp]),
    "ym_obb_finish": (ci, [ci, C.POINTER(vp), C.POINTER(ci), C.POINTER(ci), C.POINTER(cf), ci, ci, vp, vp, vp]),
    "ym_kpts_decode": (ci, [ci, C.POINTER(vp), C.POINTER(ci), C.POINTER(ci), C.POINTER(cf), ci, ci, ci, vp, vp]),
    "ym_dwconv3_routed_nhwc": (ci, [vp, ci, vp, vp, ci, vp, ci, ci, ci, ci, ci, vp, ci, vp]),
    "ym_route_affine": (ci, [vp, vp, vp, vp, vp, ci, ci, ci, ci, vp, vp]),
    "ym_process_mask_scratch_bytes": (cll, [ci, ci, ci]),
    "ym_process_mask": (ci, [vp, ci, ci, ci, ci, vp, ci, ci, ci, ci, ci, ci, vp, vp, vp]),
    "ym_nms_rotated_scratch_bytes": (cll, [ci, ci]),
    "ym_nms_rotated": (ci, [vp, ci, ci, ci, cf, cf, ci, ci, cf, vp, vp, vp, vp, vp]),
    "ym_letterbox_u8": (ci, [vp, cll, ci, ci, ci, ci, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp, ci, ci, ci, ci, vp]),
    "ym_scale_coords": (ci, [vp, ci, cll, vp, ci, ci, vp]),
    "ym_scale_boxes": (ci, [vp, ci, cll, ci, vp, ci, vp, ci, ci, vp]),
    "ym_detect_topk": (ci, [ci, C.POINTER(vp), C.POINTER(vp), C.POINTER(ci), C.POINTER(ci), C.POINTER(cf), ci, ci, ci, vp, vp, vp, vp]),
    "ym_detect_dense": (ci, [ci, C.POINTER(vp), C.POINTER(vp), C.POINTER(ci), C.POINTER(ci), C.POINTER(cf), ci, ci, ci, ci, vp, vp]),
}

_lib = None


class YMLibraryError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle.  Raises YMLibraryError if the CUDA extension is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise YMLibraryError(
            f"{LIB_PATH} not found: build it with `make -C {os.path.join(_HERE, 'csrc')}` or "
            "`python -c 'import __graft_entry__ as g; g.build()'`.  There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().ym_last_error()
        raise RuntimeError(f"libym_b200 {what} failed (rc={rc}): {msg.decode() if msg else '?'}")

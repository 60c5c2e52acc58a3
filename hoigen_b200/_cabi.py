"""ctypes binding of libhoigen_b200.so (the C ABI declared in include/hoigen_b200.h).

There is NO fallback: if the shared library is missing or a call fails, this raises. Torch is used only to
obtain device pointers and the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

from ._build import LIB_PATH

ACT_NONE, ACT_QUICKGELU, ACT_RELU, ACT_EXP = 0, 1, 2, 3
ABI_VERSION = 2          # HOIGEN_ABI_VERSION of include/hoigen_b200.h that the ctypes mirrors below follow


class HoigenError(RuntimeError):
    pass


class GemmParams(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("w", C.c_void_p),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("lda", C.c_int32), ("ldw", C.c_int32),
        ("bias", C.c_void_p), ("colscale", C.c_void_p),
        ("act", C.c_int32),
        ("residual", C.c_void_p), ("ld_res", C.c_int32),
        ("out_f32", C.c_void_p), ("ld_f32", C.c_int32),
        ("out_bf16", C.c_void_p), ("ld_bf16", C.c_int32),
        ("block_n", C.c_int32), ("split_k", C.c_int32),
        ("act_param", C.c_float), ("ln_stats", C.c_void_p), ("ln_colsum", C.c_void_p),
        ("conv_taps", C.c_int32), ("conv_cin", C.c_int32), ("halo_h", C.c_int32), ("halo_w", C.c_int32),
        ("res_bf16", C.c_void_p), ("ld_resb", C.c_int32),
        ("a2", C.c_void_p), ("lda2", C.c_int32), ("k2", C.c_int32),
        ("conv_stride", C.c_int32),
    ]


CONV_OP_GEMM, CONV_OP_STEM_IM2COL, CONV_OP_MAXPOOL, CONV_OP_GATHER_S2, CONV_OP_AVGPOOL_L2NORM, CONV_OP_STEM_CONV = range(6)


class ConvOp(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("in_", C.c_void_p), ("out", C.c_void_p),
        ("batch", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32), ("taps", C.c_int32),
        ("gemm", GemmParams),
    ]


_lib = None
_inited_devices = set()


def lib_path() -> Path:
    return Path(os.environ.get("HOIGEN_B200_LIB", str(LIB_PATH)))


def load() -> C.CDLL:
    """dlopen the C-ABI library (no CUDA call is made here)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not path.exists():
        raise HoigenError(
            f"{path} not found: build it with `python -m hoigen_b200._build` (or __graft_entry__.build()). "
            "hoigen_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(str(path))
    lib.hoigen_abi_version.restype = C.c_int
    if lib.hoigen_abi_version() != ABI_VERSION:     # a stale build would read the parameter structs with another layout
        raise HoigenError(f"{path} has ABI version {lib.hoigen_abi_version()}, this package expects {ABI_VERSION}: rebuild it "
                          "(python -m hoigen_b200._build --force)")
    lib.hoigen_last_error.restype = C.c_char_p
    lib.hoigen_init.argtypes = [C.c_int]
    for name, sig in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = sig
    lib.hoigen_wire_record_bytes.restype = C.c_int64
    lib.hoigen_wire_record_bytes.argtypes = [C.c_int32, C.c_int64, C.c_int64]
    lib.hoigen_cache_fused_workspace_bytes.restype = C.c_int64
    lib.hoigen_cache_fused_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
    lib.hoigen_launch_count.restype = C.c_longlong
    lib.hoigen_profile_enable.argtypes = [C.c_int]
    lib.hoigen_profile_read.restype = C.c_longlong
    lib.hoigen_profile_read.argtypes = [C.c_char_p, C.c_longlong]
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().hoigen_last_error().decode(errors="replace")
        raise HoigenError(f"{what} failed (status {rc}): {msg}")


def init(device: torch.device | int | None = None) -> C.CDLL:
    lib = load()
    if not torch.cuda.is_available():
        raise HoigenError("hoigen_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    if idx is None:
        idx = torch.cuda.current_device()
    if idx not in _inited_devices:
        check(lib.hoigen_init(idx), "hoigen_init")
        _inited_devices.add(idx)
    return lib


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t: torch.Tensor | None) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


class AdapterWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "wd", "down_b", "wq", "wo", "w1", "w2", "in_proj_b", "out_proj_b", "linear1_b", "linear2_b", "norm2_w", "norm2_b", "norm3_w",
        "norm3_b", "wup")]


ENCODER_WEIGHT_FIELDS = (
    "conv_w", "class_embedding", "positional_embedding", "ln_pre_w", "ln_pre_b", "ln_post_w", "ln_post_b", "proj_t",
    "ln1_w", "ln1_b", "ln2_w", "ln2_b", "qkv_w", "qkv_b", "out_w", "out_b", "fc_w", "fc_b", "proj_w", "proj_b",
    "ad_down_w", "ad_down_b", "ad_up_w", "ad_up_b", "ad_in_proj_w", "ad_in_proj_b", "ad_wq", "ad_wo", "ad_w1", "ad_w2",
    "ad_out_proj_b", "ad_linear1_b", "ad_linear2_b", "ad_norm2_w", "ad_norm2_b", "ad_norm3_w", "ad_norm3_b",
    "qkv_wf", "qkv_colsum", "qkv_bf", "fc_wf", "fc_colsum", "fc_bf")


class EncoderWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ENCODER_WEIGHT_FIELDS]


ENCODER_BUFFER_FIELDS = ("patches", "patch_emb", "x", "xb", "h", "qkv", "attn", "mlp", "delta", "delta2",
                         "adapter_kv", "row_stats", "tokens_out")


class EncoderBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ENCODER_BUFFER_FIELDS]


class ScoreWeights(C.Structure):
    _fields_ = [
        ("num_classes", C.c_int32), ("cache_rows", C.c_int32),
        ("cache_keys", C.c_void_p * 3), ("bias_term", C.c_void_p * 3), ("label_t", C.c_void_p * 3),
        ("colscale", C.c_void_p * 3),
        ("global_keys", C.c_void_p), ("global_bias_term", C.c_void_p), ("colscale_global", C.c_void_p),
        ("dino_keys", C.c_void_p), ("dino_bias_term", C.c_void_p), ("colscale_dino", C.c_void_p),
        ("text_w", C.c_void_p), ("colscale_text", C.c_void_p),
        ("affinity", C.c_int32), ("beta", C.c_float),
        ("cache_bias", C.c_void_p * 3), ("global_bias", C.c_void_p), ("dino_bias", C.c_void_p),
    ]


SCORE_BUFFER_FIELDS = ("pair_feat_bf16", "phi", "phi_img", "g_bf16", "d_bf16", "img_logits", "logits", "ld_logits",
                       "cache_parts")


class ScoreBuffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in SCORE_BUFFER_FIELDS]


class ScoreWeightsFp32(C.Structure):
    _fields_ = [
        ("num_classes", C.c_int32), ("cache_rows", C.c_int32),
        ("cache_keys6", C.c_void_p * 3), ("bias_term", C.c_void_p * 3), ("label3_t", C.c_void_p * 3),
        ("colscale", C.c_void_p * 3),
        ("global_keys6", C.c_void_p), ("global_bias_term", C.c_void_p), ("colscale_global", C.c_void_p),
        ("dino_keys6", C.c_void_p), ("dino_bias_term", C.c_void_p), ("colscale_dino", C.c_void_p),
        ("text_w6", C.c_void_p), ("colscale_text", C.c_void_p),
    ]


class ScoreBuffersFp32(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("feat6", "phi", "phi3", "phi_img", "phi_img3", "g6", "d6", "img_logits", "logits",
                                          "ld_logits")]


class FoldedWeights(C.Structure):
    _fields_ = [("num_classes", C.c_int32), ("pair_w", C.c_void_p), ("global_w", C.c_void_p), ("dino_w", C.c_void_p),
                ("bias_total", C.c_void_p)]


_P, _I, _F, _L = C.c_void_p, C.c_int32, C.c_float, C.c_int64

# name -> argtypes, one entry per symbol in include/hoigen_b200.h
_SIGNATURES: dict[str, list] = {
    "hoigen_gemm_bf16": [C.POINTER(GemmParams), _P],
    "hoigen_debug_gemm_simt": [C.POINTER(GemmParams), _P],
    "hoigen_patchify_bf16": [_P, _P, _I, _P],
    "hoigen_embed_lnpre": [_P, _P, _P, _P, _P, _P, _P, _I, _P],
    "hoigen_layernorm768": [_P, _P, _P, _P, _P, _I, _P],
    "hoigen_add_layernorm768": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _P],
    "hoigen_add_rowstats768": [_P, _P, _P, _P, _P, _P, _I, _P],
    "hoigen_adapter_kv": [_P, _P, _P, _P, _I, _I, _P],
    "hoigen_adapter_block": [_P, _P, _P, _P, C.POINTER(AdapterWeights), _P, _P, _I, _I, _P],
    "hoigen_attention": [_P, _P, _I, _P],
    "hoigen_debug_attention_trace": [_P, _P, _I, _P, _P],
    "hoigen_debug_adapter_trace": [_P],
    "hoigen_encoder_forward": [C.POINTER(EncoderWeights), C.POINTER(EncoderBuffers), _P, _P, _P, _I, _I, _I, _P],
    "hoigen_prior_tokens": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _F, _I, _I, _I, _P, _P, _P],
    "hoigen_roi_pair_features": [_P, _P, _P, _P, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P],
    "hoigen_set_words": [_P, _P, _I, _P],
    "hoigen_copy_words": [_P, _P, _I, _P],
    "hoigen_rows_to_bf16": [_P, _L, _I, _I, _I, _P, _P],
    "hoigen_broadcast_image_logits": [_P, _P, _I, _I, _I, _I, _P, _P],
    "hoigen_score_pairs": [C.POINTER(ScoreWeights), C.POINTER(ScoreBuffers), _P, _P, _P, _I, _I, _P],
    "hoigen_score_cache_fused": [C.POINTER(ScoreWeights), _P, _P, _P, _P, _I, _I, _I, _F, _P, _P, _I, _P],
    "hoigen_rows_split3": [_P, _L, _I, _I, _I, _I, _P, _P],
    "hoigen_score_pairs_fp32": [C.POINTER(ScoreWeightsFp32), C.POINTER(ScoreBuffersFp32), _P, _P, _P, _P, _I, _I, _P],
    "hoigen_score_pairs_folded": [C.POINTER(FoldedWeights), C.POINTER(ScoreBuffers), _P, _P, _P, _I, _I, _P],
    "hoigen_ap_11point": [_P, _P, _P, _P, _I, _P, _P, _P],
    "hoigen_stem_im2col": [_P, _P, _I, _P],
    "hoigen_stem_im2col_hw": [_P, _P, _I, _I, _I, _P],
    "hoigen_stem_conv": [_P, _P, _P, _P, _I, _P],
    "hoigen_stem_conv_hw": [_P, _P, _P, _P, _I, _I, _I, _P],
    "hoigen_maxpool3x3s2_halo": [_P, _P, _I, _I, _I, _I, _P],
    "hoigen_conv_gather_s2": [_P, _P, _I, _I, _I, _I, _I, _P],
    "hoigen_avgpool_l2norm": [_P, _P, _I, _I, _I, _I, _P],
    "hoigen_conv_plan_run": [_P, _I, _P],
    "hoigen_add_layernorm256": [_P, _P, _P, _P, _P, _I, _P, _P, _I, _P],
    "hoigen_attention_heads32": [_P, _I, _P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _I, _F, _P],
    "hoigen_prepare_proposals": [_P, _P, _P, _I, _I, _L, _F, _I, _I, _F, _P, _P, _P, _P, _P],
    "hoigen_associate_pairs": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _I, _I, _F, _P, _P, _P],
    "hoigen_pack_wire": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _L, _P, _P],
    "hoigen_unpack_wire": [_P, _I, _L, _I, _P, _P, _P, _P, _P, _P, _P],
    "hoigen_emit_triplets": [_P, _I, _I, _P, _P, _P, _P, _I, _I, _P, _I, _I, _F, _P, _P, _P, _L, _P, _P, _P, _P, _P, _P],
}

EXPORTED_SYMBOLS = ["hoigen_abi_version", "hoigen_last_error", "hoigen_init", "hoigen_launch_count", "hoigen_wire_record_bytes", "hoigen_cache_fused_workspace_bytes",
                    "hoigen_profile_enable", "hoigen_profile_reset", "hoigen_profile_read", *_SIGNATURES.keys()]


def launch_count() -> int:
    """Kernels launched by this library since the last profile_reset()."""
    return int(load().hoigen_launch_count())


def profile(enable: bool) -> None:
    lib = load()
    lib.hoigen_profile_reset()
    lib.hoigen_profile_enable(1 if enable else 0)


def profile_read() -> list:
    """[(tag, ms, flops, bytes, start_ms)] per recorded launch (synchronises the device)."""
    lib = load()
    buf = C.create_string_buffer(1 << 20)
    n = lib.hoigen_profile_read(buf, len(buf))
    if n < 0:
        raise HoigenError("hoigen_profile_read failed")
    out = []
    for line in buf.raw[:n].decode().splitlines():
        tag, ms, fl, by, t0 = line.split()
        out.append((tag, float(ms), float(fl), float(by), float(t0)))
    return out


def call(name: str, *args) -> None:
    """Invoke a C-ABI entry point on the current stream (the stream pointer is appended) and raise on error."""
    lib = load()
    check(getattr(lib, name)(*args, stream_ptr()), name)


def gemm_params(a: torch.Tensor, w: torch.Tensor, *, bias=None, colscale=None, act=ACT_NONE, residual=None,
                out_f32=None, out_bf16=None, block_n: int = 0, split_k: int = 0, act_param: float = 0.0, ln_stats=None,
                ln_colsum=None, conv_taps: int = 0, halo=None, res_bf16=None, a2=None, conv_stride: int = 1) -> "GemmParams":
    """Fill a hoigen_gemm_params from tensors (no launch).  conv_taps = 9: `a` is the (rows, cin) activation matrix with a
    zero halo, `w` is (N, 9 * cin); halo = (H + 2, W + 2); res_bf16 = bf16 identity added before the activation.
    conv_stride = 2: `a` is the (4 * rows_out, cin) four-phase split of the input (hoigen_conv_gather_s2, taps = 4) and halo the
    OUTPUT's (Ho + 2, Wo + 2)."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    assert a.dim() == 2 and w.dim() == 2 and a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    p = GemmParams()
    if conv_taps == 9:
        assert w.shape[1] == 9 * K and halo is not None
        p.conv_taps, p.conv_cin = 9, K
        K = 9 * K
        if conv_stride == 2:
            assert M % 4 == 0
            M //= 4
            p.conv_stride = 2
    elif a2 is not None:        # second A source: w = [W1 | W2] along K
        assert a2.dtype == torch.bfloat16 and a2.dim() == 2 and a2.stride(1) == 1 and a2.shape[0] == M
        assert w.shape[1] == K + a2.shape[1] and K % 64 == 0
        p.a2, p.lda2, p.k2 = a2.data_ptr(), a2.stride(0), a2.shape[1]
        K = K + a2.shape[1]
    else:
        assert w.shape[1] == K
    p.a, p.w = a.data_ptr(), w.data_ptr()
    p.M, p.N, p.K = M, N, K
    p.lda, p.ldw = a.stride(0), w.stride(0)
    for name, t in (("bias", bias), ("colscale", colscale)):
        if t is not None:
            assert t.dtype == torch.float32 and t.numel() == N and t.is_contiguous()
            setattr(p, name, t.data_ptr())
    p.act = act
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.stride(1) == 1
        p.residual, p.ld_res = residual.data_ptr(), residual.stride(0)
    if out_f32 is not None:
        assert out_f32.dtype == torch.float32 and out_f32.stride(1) == 1 and out_f32.shape[0] >= M
        p.out_f32, p.ld_f32 = out_f32.data_ptr(), out_f32.stride(0)
    if out_bf16 is not None:
        assert out_bf16.dtype == torch.bfloat16 and out_bf16.stride(1) == 1 and out_bf16.shape[0] >= M
        p.out_bf16, p.ld_bf16 = out_bf16.data_ptr(), out_bf16.stride(0)
    p.block_n = block_n
    p.split_k = split_k
    p.act_param = act_param
    if ln_stats is not None:
        assert ln_stats.dtype == torch.float32 and ln_stats.shape == (M, 2) and ln_stats.is_contiguous()
        assert ln_colsum is not None and ln_colsum.dtype == torch.float32 and ln_colsum.numel() == N
        p.ln_stats, p.ln_colsum = ln_stats.data_ptr(), ln_colsum.data_ptr()
    if halo is not None:
        p.halo_h, p.halo_w = int(halo[0]), int(halo[1])
    if res_bf16 is not None:
        assert res_bf16.dtype == torch.bfloat16 and res_bf16.stride(1) == 1 and res_bf16.shape[0] >= M
        p.res_bf16, p.ld_resb = res_bf16.data_ptr(), res_bf16.stride(0)
    return p


def gemm_bf16(a: torch.Tensor, w: torch.Tensor, *, simt: bool = False, **kw) -> None:
    """out = epi(a[M,K] @ w[N,K]^T); see hoigen_gemm_bf16 in include/hoigen_b200.h."""
    lib = init(a.device)
    p = gemm_params(a, w, **kw)
    fn = lib.hoigen_debug_gemm_simt if simt else lib.hoigen_gemm_bf16
    check(fn(C.byref(p), stream_ptr()), "hoigen_gemm_bf16")

"""ctypes binding of include/v2w.h (libv2w.so).

Loading never falls back: if the shared library is missing or a symbol is absent the import
of a native entry point raises ``NativeLibraryError``.  The library itself loads on a CPU-only
box (CUDA runtime is linked statically); any device entry point then fails with
``V2W_ERR_CUDA`` -- there is no CPU execution path.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional

MAX_STAGES = 8
MAX_KERNELS = 4
MAX_DILATIONS = 4
MAX_TAPS = 16
MAX_TILES = 16

PREC_FP32 = 0
PREC_BF16 = 1
PREC_FP16 = 2
PREC_FP32_TC = 3
HALF_BF16, HALF_FP16, HALF_FP16_SPLIT, HALF_BF16_SPLIT = 0, 1, 2, 3
PACKED_HEADER_BYTES = 1024
TC_FLAG_FP16 = 0x100
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "fp16": PREC_FP16, "fp32tc": PREC_FP32_TC}
# 16-bit operand format of the packed weight buffer a plan of each precision binds (the FFMA plan reads the
# fp32 arrays, which every buffer holds).  A BF16 plan of a ResBlock1 generator binds HALF_BF16_SPLIT: ask the
# plan (v2w_plan_packed_format) rather than this table.
HALF_OF_PRECISION = {"fp32": HALF_BF16, "bf16": HALF_BF16, "fp16": HALF_FP16, "fp32tc": HALF_FP16_SPLIT}
ENGINE_SIMT = 0
ENGINE_TC = 1
OUT_F32, OUT_PCM16 = 0, 1
OP_TRANSPOSE_IN, OP_COND_AFFINE, OP_TAPGEMM, OP_CONV_POST, OP_MRF_FUSED, OP_PAIR_FUSED = 0, 1, 2, 3, 4, 5
OP_MRF_WIDE1, OP_MRF_WIDE2 = 6, 7
#: op kinds that run on the tcgen05 engines and carry algorithmic FLOPs
GEMM_OPS = (OP_TAPGEMM, OP_MRF_FUSED, OP_PAIR_FUSED, OP_MRF_WIDE1, OP_MRF_WIDE2)
FUSED_MAX_SEG, FUSED_MAX_TAPS = 16, 96
SEG_CONV1, SEG_SUM, SEG_CONV2 = 0, 1, 2
ROLE_CONV_PRE, ROLE_UPSAMPLE, ROLE_RESBLOCK, ROLE_CONV_POST = 0, 1, 2, 3
TC_FLAG_PER_TAP = 1
TC_FLAG_BASE_OFFSET = 2

import os as _os
LIB_PATH = Path(_os.environ.get("V2W_LIB") or (Path(__file__).resolve().parent / "libv2w.so"))


class NativeLibraryError(RuntimeError):
    pass


class V2WError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libv2w error {code}: {message}")
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("num_wv_feat", C.c_int32),
        ("upsample_initial_channel", C.c_int32),
        ("num_upsamples", C.c_int32),
        ("upsample_rates", C.c_int32 * MAX_STAGES),
        ("upsample_kernel_sizes", C.c_int32 * MAX_STAGES),
        ("resblock_type", C.c_int32),
        ("num_kernels", C.c_int32),
        ("resblock_kernel_sizes", C.c_int32 * MAX_KERNELS),
        ("num_dilations", C.c_int32),
        ("resblock_dilation_sizes", (C.c_int32 * MAX_DILATIONS) * MAX_KERNELS),
        ("spk_dim", C.c_int32),
        ("noise_dim", C.c_int32),
    ]


class ConvParams(C.Structure):
    _fields_ = [("v", C.c_void_p), ("g", C.c_void_p), ("bias", C.c_void_p)]


class CondParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "fc_weight", "fc_bias", "sn_weight_orig", "sn_u", "sn_v", "sn_bias", "bn_mean", "bn_var")]


class Params(C.Structure):
    _fields_ = [
        ("num_convs", C.c_int32),
        ("convs", C.POINTER(ConvParams)),
        ("num_stages", C.c_int32),
        ("cond", C.POINTER(CondParams)),
    ]


class LayerDesc(C.Structure):
    _fields_ = [
        ("role", C.c_int32), ("stage", C.c_int32), ("conv_index", C.c_int32),
        ("cin", C.c_int32), ("cout", C.c_int32), ("up", C.c_int32), ("n_total", C.c_int32),
        ("kernel_size", C.c_int32), ("dilation", C.c_int32), ("padding", C.c_int32),
        ("num_taps", C.c_int32),
        ("tap_off", C.c_int32 * MAX_TAPS),
        ("tap_mask", C.c_uint32 * MAX_TAPS),
        ("w16_off", C.c_int64), ("w32_off", C.c_int64), ("bias_off", C.c_int64),
        ("macs_per_row", C.c_int64),
    ]


class CondDesc(C.Structure):
    _fields_ = [("channels", C.c_int32), ("in_dim", C.c_int32)] + [
        (n, C.c_int64) for n in ("fcw_off", "fcb_off", "snw_off", "snb_off", "rstd_off", "mean_off")]


class LaunchDesc(C.Structure):   # v2w_launch_desc
    _fields_ = [(n, C.c_int32) for n in (
        "grid", "smem_bytes", "cta_pair", "k_chunk", "msub", "stages_a", "stages_w", "tap_group", "y_cols", "h_cols",
        "epi_mode", "epi_wide", "epi_phased", "split")]


class OpDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("engine", C.c_int32), ("layer", C.c_int32),
        ("batch", C.c_int32), ("rows", C.c_int32),
        ("n_tile", C.c_int32), ("n_tiles", C.c_int32),
        ("num_taps", C.c_int32 * MAX_TILES),
        ("tap_index", (C.c_int32 * MAX_TAPS) * MAX_TILES),
        ("in_act", C.c_int64), ("in_is_bf16", C.c_int32), ("in_slope", C.c_float),
        ("cond_stage", C.c_int32),
        ("residual", C.c_int64), ("accum_in", C.c_int64), ("out_scale", C.c_float),
        ("out_f32", C.c_int64), ("out_bf16", C.c_int64), ("out_slope", C.c_float),
        ("flops", C.c_int64),
        ("stage", C.c_int32),
        ("pair", C.c_int32),
        ("post", C.c_int32),
        ("out_bf16_pitch", C.c_int32),
        ("in_pitch", C.c_int32),
        ("f32_blocked", C.c_int32),
        ("out_blocked", C.c_int32),
        ("acc_blocked", C.c_int32),
        ("split", C.c_int32),
    ]


class WideDesc(C.Structure):
    _fields_ = [
        ("ok", C.c_int32), ("channels", C.c_int32), ("num_branches", C.c_int32), ("first_layer", C.c_int32),
        ("sum_taps", C.c_int32), ("sum_off", C.c_int32 * MAX_TAPS),
        ("wsum_off", C.c_int64), ("b1cat_off", C.c_int64), ("bsum_off", C.c_int64),
    ]


class FusedDesc(C.Structure):
    _fields_ = [
        ("ok", C.c_int32), ("channels", C.c_int32), ("num_branches", C.c_int32),
        ("first_layer", C.c_int32), ("hc", C.c_int32), ("h1", C.c_int32), ("rc", C.c_int32),
        ("a0_boxes", C.c_int32), ("num_segments", C.c_int32), ("num_taps", C.c_int32),
        ("stacked", C.c_int32), ("pair", C.c_int32), ("preload", C.c_int32), ("post", C.c_int32), ("order", C.c_int32 * 4), ("t_bufs", C.c_int32), ("a0_bufs", C.c_int32),
        ("seg_type", C.c_int32 * FUSED_MAX_SEG), ("seg_branch", C.c_int32 * FUSED_MAX_SEG),
        ("seg_ntaps", C.c_int32 * FUSED_MAX_SEG), ("seg_first", C.c_int32 * FUSED_MAX_SEG),
        ("seg_flags", C.c_int32 * FUSED_MAX_SEG),
        ("seg_dcol", C.c_int32 * FUSED_MAX_SEG), ("seg_n", C.c_int32 * FUSED_MAX_SEG),
        ("seg_acc", C.c_int32 * FUSED_MAX_SEG),
        ("seg_wait", C.c_int32 * FUSED_MAX_SEG), ("seg_commit", C.c_int32 * FUSED_MAX_SEG),
        ("shift", C.c_int32 * FUSED_MAX_TAPS), ("offset", C.c_int32 * FUSED_MAX_TAPS),
        ("w_row", C.c_int32 * FUSED_MAX_TAPS), ("w_rows", C.c_int32),
        ("w_off", C.c_int64), ("bsum_off", C.c_int64),
    ]


class TrainStage(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "fc_weight", "fc_bias", "sn_weight_orig", "sn_bias", "sn_u", "sn_v",
        "bn_running_mean", "bn_running_var")]


class TrainState(C.Structure):
    _fields_ = [("num_stages", C.c_int32), ("momentum", C.c_float), ("stage", TrainStage * MAX_STAGES)]


class MelConfig(C.Structure):
    _fields_ = [("n_fft", C.c_int32), ("hop_size", C.c_int32), ("win_size", C.c_int32),
                ("num_mels", C.c_int32), ("sampling_rate", C.c_int32),
                ("fmin", C.c_float), ("fmax", C.c_float)]


class TapGemmArgs(C.Structure):
    _fields_ = [
        ("engine", C.c_int32),
        ("batch", C.c_int32), ("rows", C.c_int32), ("cin", C.c_int32), ("n_total", C.c_int32),
        ("cmod", C.c_int32), ("n_tile", C.c_int32), ("n_tiles", C.c_int32),
        ("num_taps", C.c_int32 * MAX_TILES),
        ("tap_off", (C.c_int32 * MAX_TAPS) * MAX_TILES),
        ("tap_row", (C.c_int32 * MAX_TAPS) * MAX_TILES),
        ("in_act", C.c_void_p), ("in_is_bf16", C.c_int32), ("in_slope", C.c_float),
        ("w16", C.c_void_p), ("w32", C.c_void_p),
        ("bias", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
        ("residual", C.c_void_p), ("accum_in", C.c_void_p), ("out_scale", C.c_float),
        ("out_f32", C.c_void_p), ("out_bf16", C.c_void_p), ("out_slope", C.c_float),
        ("tc_flags", C.c_int32),
    ]


#: every symbol include/v2w.h declares: (name, restype, argtypes)
_SIGNATURES = [
    ("v2w_abi_version", C.c_int32, []),
    ("v2w_last_error", C.c_char_p, []),
    ("v2w_device_count", C.c_int32, []),
    ("v2w_model_create", C.c_int32, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    ("v2w_model_destroy", None, [C.c_void_p]),
    ("v2w_model_num_layers", C.c_int32, [C.c_void_p]),
    ("v2w_model_get_layer", C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(LayerDesc)]),
    ("v2w_model_get_cond", C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(CondDesc)]),
    ("v2w_model_get_fused", C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(FusedDesc)]),
    ("v2w_model_get_wide", C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(WideDesc)]),
    ("v2w_model_num_pairs", C.c_int32, [C.c_void_p]),
    ("v2w_model_get_pair", C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(FusedDesc)]),
    ("v2w_model_num_convs", C.c_int32, [C.c_void_p]),
    ("v2w_model_packed_bytes", C.c_int64, [C.c_void_p]),
    ("v2w_model_pack", C.c_int32, [C.c_void_p, C.POINTER(Params), C.c_void_p]),
    ("v2w_model_pack_half", C.c_int32, [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_int32]),
    ("v2w_model_cond_region", C.c_int32, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("v2w_model_pack_cond", C.c_int32, [C.c_void_p, C.POINTER(CondParams), C.c_int32, C.c_void_p]),
    ("v2w_model_flops_per_frame", C.c_int64, [C.c_void_p]),
    ("v2w_plan_create", C.c_int32,
     [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    ("v2w_plan_destroy", None, [C.c_void_p]),
    ("v2w_plan_workspace_bytes", C.c_int64, [C.c_void_p]),
    ("v2w_plan_packed_format", C.c_int32, [C.c_void_p]),
    ("v2w_plan_num_ops", C.c_int32, [C.c_void_p]),
    ("v2w_plan_get_op", C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(OpDesc)]),
    ("v2w_plan_get_launch", C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(LaunchDesc)]),
    ("v2w_plan_set_engine", C.c_int32, [C.c_void_p, C.c_int32, C.c_int32]),
    ("v2w_plan_bind", C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("v2w_plan_refresh_weights", C.c_int32, [C.c_void_p, C.c_void_p]),
    ("v2w_packer_create", C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p)]),
    ("v2w_packer_run", C.c_int32, [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_int32, C.c_void_p,
                                   C.POINTER(C.c_int32)]),
    ("v2w_packer_run_cond", C.c_int32, [C.c_void_p, C.POINTER(CondParams), C.c_int32, C.c_void_p, C.c_void_p]),
    ("v2w_packer_destroy", None, [C.c_void_p]),
    ("v2w_forward", C.c_int32,
     [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
      C.POINTER(C.c_int32)]),
    ("v2w_forward_graph", C.c_int32,
     [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32),
      C.POINTER(C.c_int32)]),
    ("v2w_forward_host", C.c_int32,
     [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
      C.POINTER(C.c_int32)]),
    ("v2w_forward_train", C.c_int32,
     [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(TrainState), C.c_void_p,
      C.POINTER(C.c_int32)]),
    ("v2w_forward_ragged", C.c_int32, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int32, C.c_void_p, C.POINTER(C.c_int32)]),
    ("v2w_model_halo_frames", C.c_int32, [C.c_void_p]),
    ("v2w_mel_create", C.c_int32, [C.POINTER(MelConfig), C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    ("v2w_mel_destroy", None, [C.c_void_p]),
    ("v2w_mel_num_frames", C.c_int32, [C.c_void_p]),
    ("v2w_mel_workspace_bytes", C.c_int64, [C.c_void_p]),
    ("v2w_mel_get_filterbank", C.c_int32, [C.c_void_p, C.c_void_p]),
    ("v2w_mel_bind", C.c_int32, [C.c_void_p, C.c_void_p]),
    ("v2w_mel_forward", C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    ("v2w_npy_shape", C.c_int32, [C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("v2w_npy_read_f32", C.c_int32, [C.c_char_p, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(C.c_int64)]),
    ("v2w_npy_shapes", C.c_int32, [C.POINTER(C.c_char_p), C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int32]),
    ("v2w_npy_read_batch", C.c_int32, [C.POINTER(C.c_char_p), C.c_int32, C.c_void_p, C.POINTER(C.c_int64),
                                       C.POINTER(C.c_int64), C.c_int64, C.c_int32]),
    ("v2w_wav_write_batch", C.c_int32, [C.POINTER(C.c_char_p), C.c_int32, C.c_void_p, C.POINTER(C.c_int64),
                                        C.POINTER(C.c_int64), C.c_int32, C.c_int32]),
    ("v2w_pipeline_create", C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p)]),
    ("v2w_pipeline_submit", C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("v2w_pipeline_wait", C.c_int32, [C.c_void_p, C.POINTER(C.c_int64)]),
    ("v2w_pipeline_destroy", None, [C.c_void_p]),
    ("v2w_plan_set_profiling", C.c_int32, [C.c_void_p, C.c_int32]),
    ("v2w_plan_op_time_ms", C.c_float, [C.c_void_p, C.c_int32]),
    ("v2w_tapgemm", C.c_int32, [C.POINTER(TapGemmArgs), C.c_void_p]),
]

EXPORTED_SYMBOLS = [s[0] for s in _SIGNATURES]

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load libv2w.so (once) and type every entry point.  Raises, never degrades."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m wavthruvec_pytorch_b200.build` "
            "(or __graft_entry__.build()). There is no fallback path.")
    try:
        lib = C.CDLL(str(LIB_PATH))
    except OSError as e:  # pragma: no cover - depends on the box
        raise NativeLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, restype, argtypes in _SIGNATURES:
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise NativeLibraryError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.v2w_abi_version() != 1:
        raise NativeLibraryError(f"ABI version mismatch: library {lib.v2w_abi_version()}, binding 1")
    _lib = lib
    return lib


def last_error() -> str:
    return load().v2w_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != 0:
        raise V2WError(rc, last_error())


def make_config(h) -> Config:
    """Attribute bag `h` (vec2wav/hparams.py names) -> v2w_config."""
    cfg = Config()
    cfg.num_wv_feat = int(h.num_wv_feat)
    cfg.upsample_initial_channel = int(h.upsample_initial_channel)
    rates, ksz = list(h.upsample_rates), list(h.upsample_kernel_sizes)
    if len(rates) != len(ksz):
        raise ValueError("upsample_rates and upsample_kernel_sizes differ in length")
    if len(rates) > MAX_STAGES:
        raise ValueError(f"at most {MAX_STAGES} upsample stages")
    cfg.num_upsamples = len(rates)
    for i, (u, k) in enumerate(zip(rates, ksz)):
        cfg.upsample_rates[i] = int(u)
        cfg.upsample_kernel_sizes[i] = int(k)
    # vec2wav/models.py:84 -- only the *string* '1' selects ResBlock1
    cfg.resblock_type = 1 if h.resblock == '1' else 2
    rks = list(h.resblock_kernel_sizes)
    rds = [list(d) for d in h.resblock_dilation_sizes]
    # the reference zips kernel sizes with dilation rows (models.py:97)
    n = min(len(rks), len(rds))
    if n > MAX_KERNELS:
        raise ValueError(f"at most {MAX_KERNELS} resblock kernel sizes")
    cfg.num_kernels = n
    nd = min(len(d) for d in rds[:n]) if n else 0
    if nd > MAX_DILATIONS:
        nd = MAX_DILATIONS
    cfg.num_dilations = nd
    for j in range(n):
        cfg.resblock_kernel_sizes[j] = int(rks[j])
        for d in range(nd):
            cfg.resblock_dilation_sizes[j][d] = int(rds[j][d])
    cfg.spk_dim = int(h.spk_dim)
    cfg.noise_dim = int(h.noise_dim)
    return cfg

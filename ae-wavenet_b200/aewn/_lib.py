"""ctypes binding of libaewn.so -- mirrors include/aewn.h one to one.

The library is REQUIRED: there is no CPU or eager-PyTorch fallback for the hot path.  Importing this module never
touches the GPU; the first kernel call fails loudly if the shared object is missing or a launch fails.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libaewn.so")

MAX_ACTS, MAX_SEGS, MAX_NTILES = 6, 6, 4
WGRAD_MAX_ACTS, WGRAD_MAX_ITEMS = 6, 32

EPI_LINEAR, EPI_GATE_FWD, EPI_GATE_BWD = 0, 1, 2
F_ACCUM, F_RELU, F_MASKPOS, F_RELU_FIRST, F_MERGE_NEXT, F_AB16, F_NO_OUT32 = 1, 2, 4, 8, 16, 32, 64
ERR_TIMEOUT = -1003
ERR_RANGE = -1004     # an activation left the fp16 operand range of the fused layer kernel
CLUSTER_PAIR_MMA = 102   # aewn.h AEWN_CLUSTER_PAIR_MMA: 2-CTA clusters issuing cta_group::2 MMAs


class Act(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("t_extent", C.c_int), ("channels", C.c_int), ("batch", C.c_int),
                ("row_pitch", C.c_longlong), ("batch_stride", C.c_longlong)]


class Seg(C.Structure):
    _fields_ = [("act", C.c_int), ("shift", C.c_int), ("channels", C.c_int), ("w_koff", C.c_int)]


class NTile(C.Structure):
    _fields_ = [("w_row", C.c_int), ("n", C.c_int), ("n_valid", C.c_int), ("mode", C.c_int), ("flags", C.c_int),
                ("seg_mask", C.c_int), ("t_lo", C.c_int), ("t_hi", C.c_int), ("t_zero_lo", C.c_int),
                ("out", C.c_void_p), ("out2", C.c_void_p), ("out3", C.c_void_p),
                ("out_bs", C.c_longlong), ("out_cs", C.c_longlong), ("out_toff", C.c_int),
                ("dup_toff", C.c_int), ("dup_t_hi", C.c_int), ("zero_count", C.c_void_p),
                ("add", C.c_void_p), ("add2", C.c_void_p), ("add_bs", C.c_longlong), ("add_cs", C.c_longlong),
                ("add_toff", C.c_int), ("add_t_lo", C.c_int), ("bias", C.c_void_p),
                ("out16", C.c_void_p), ("out16_bs", C.c_longlong), ("out16_cp", C.c_int), ("out16_scale", C.c_void_p)]


class TGemmDesc(C.Structure):
    _fields_ = [("acts", Act * MAX_ACTS), ("n_acts", C.c_int), ("segs", Seg * MAX_SEGS), ("n_segs", C.c_int),
                ("w", C.c_void_p), ("w_rows", C.c_int), ("w_kpad", C.c_int),
                ("ntiles", NTile * MAX_NTILES), ("n_ntiles", C.c_int), ("batch", C.c_int),
                ("t_begin", C.c_int), ("t_end", C.c_int), ("err", C.c_void_p), ("max_ctas", C.c_int),
                ("dbg_lbo", C.c_int), ("dbg_sbo", C.c_int), ("cluster", C.c_int), ("no_tma_store", C.c_int)]


class WGradItem(C.Structure):
    _fields_ = [("g_act", C.c_int), ("x_act", C.c_int), ("g_row", C.c_int), ("x_row", C.c_int),
                ("m_valid", C.c_int), ("n", C.c_int), ("n_valid", C.c_int), ("shift", C.c_int),
                ("t_lo", C.c_int), ("t_hi", C.c_int), ("n_split", C.c_int),
                ("out", C.c_void_p), ("out_rs", C.c_longlong), ("out_cs", C.c_longlong)]


class WGradDesc(C.Structure):
    _fields_ = [("acts", Act * WGRAD_MAX_ACTS), ("n_acts", C.c_int), ("items", WGradItem * WGRAD_MAX_ITEMS),
                ("n_items", C.c_int), ("batch", C.c_int), ("err", C.c_void_p), ("max_ctas", C.c_int),
                ("pair_x", C.c_int)]


WGW_MAX_CHUNKS, WGW_MAX_UNITS = 3, 8


class WGWChunk(C.Structure):
    _fields_ = [("x_act", C.c_int), ("x_row", C.c_int), ("n", C.c_int), ("n_valid", C.c_int), ("shift", C.c_int),
                ("reserved", C.c_int), ("out", C.c_void_p), ("out_rs", C.c_longlong), ("out_cs", C.c_longlong)]


class WGWUnit(C.Structure):
    _fields_ = [("g_act", C.c_int), ("g_row", C.c_int), ("m_valid", C.c_int), ("t_lo", C.c_int), ("t_hi", C.c_int),
                ("n_chunks", C.c_int), ("n_split", C.c_int), ("reserved", C.c_int), ("chunk", WGWChunk * WGW_MAX_CHUNKS)]


class WGradWDesc(C.Structure):
    _fields_ = [("acts", Act * WGRAD_MAX_ACTS), ("n_acts", C.c_int), ("units", WGWUnit * WGW_MAX_UNITS),
                ("n_units", C.c_int), ("batch", C.c_int), ("err", C.c_void_p), ("max_ctas", C.c_int)]


class Act16(C.Structure):
    """aewn_act16: fp16 channels-last tensor (batch, t_rows, row_pitch)."""
    _fields_ = [("ptr", C.c_void_p), ("t_rows", C.c_int), ("channels", C.c_int), ("batch", C.c_int),
                ("row_pitch", C.c_longlong), ("batch_stride", C.c_longlong)]


class WGradHDesc(C.Structure):
    _fields_ = [("acts", Act16 * WGRAD_MAX_ACTS), ("n_acts", C.c_int), ("units", WGWUnit * WGW_MAX_UNITS),
                ("n_units", C.c_int), ("batch", C.c_int), ("err", C.c_void_p), ("max_ctas", C.c_int),
                ("inv_scale", C.c_void_p)]


class CopyBlock(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("ni", C.c_int), ("nj", C.c_int), ("si", C.c_longlong),
                ("sj", C.c_longlong), ("di", C.c_longlong)]


class GrccFwdDesc(C.Structure):
    """aewn_grcc_fwd_desc: one fused dilation layer (include/aewn.h)."""
    _fields_ = [("x16", C.c_void_p), ("x16_bs", C.c_longlong), ("x16_cp", C.c_int),
                ("c16", C.c_void_p), ("c16_bs", C.c_longlong), ("c16_cp", C.c_int), ("t_rows", C.c_int),
                ("w1h", C.c_void_p), ("w1_k", C.c_int), ("w2h", C.c_void_p),
                ("x32", C.c_void_p), ("xo32", C.c_void_p), ("x_bs", C.c_longlong), ("x_cs", C.c_longlong),
                ("xo16", C.c_void_p), ("dup", C.c_void_p), ("dup_toff", C.c_int), ("dup_t_hi", C.c_int),
                ("th", C.c_void_p), ("sg", C.c_void_p), ("z", C.c_void_p), ("a_bs", C.c_longlong),
                ("a_cs", C.c_longlong), ("save", C.c_int),
                ("skp", C.c_void_p), ("s_bs", C.c_longlong), ("s_cs", C.c_longlong), ("skp_mode", C.c_int),
                ("batch", C.c_int), ("R", C.c_int), ("D", C.c_int), ("S", C.c_int), ("n_cond1", C.c_int),
                ("dil", C.c_int), ("final_layer", C.c_int), ("t_lo", C.c_int), ("t_zero_lo", C.c_int),
                ("t_hi", C.c_int), ("skp_t_lo", C.c_int), ("skp_zero_lo", C.c_int), ("err", C.c_void_p),
                ("max_ctas", C.c_int), ("dbg_clock", C.c_void_p),
                ("z16", C.c_void_p), ("z16_bs", C.c_longlong), ("z16_cp", C.c_int)]


class GrccDgradDesc(C.Structure):
    """aewn_grcc_dgrad_desc: data gradient of one dilation layer on the fused-layer engine (include/aewn.h)."""
    _fields_ = [("g16", C.c_void_p), ("g16_bs", C.c_longlong), ("g16_cp", C.c_int), ("t_rows", C.c_int),
                ("w1t16", C.c_void_p), ("w_k", C.c_int), ("g_sig", C.c_void_p), ("gx", C.c_void_p),
                ("x_bs", C.c_longlong), ("x_cs", C.c_longlong), ("add_t_lo", C.c_int),
                ("g_cond", C.c_void_p), ("c_bs", C.c_longlong), ("c_cs", C.c_longlong), ("n_cond", C.c_int),
                ("batch", C.c_int), ("R", C.c_int), ("dil", C.c_int),
                ("t_lo", C.c_int), ("t_zero_lo", C.c_int), ("t_hi", C.c_int),
                ("cond_t_lo", C.c_int), ("cond_zero_lo", C.c_int), ("err", C.c_void_p), ("max_ctas", C.c_int),
                ("gx16", C.c_void_p), ("gx16_bs", C.c_longlong), ("gx16_cp", C.c_int), ("g_inv_scale", C.c_void_p)]


class MfccDesc(C.Structure):
    """aewn_mfcc_desc: geometry and tables of the MFCC + delta kernels (include/aewn.h, csrc/loader.cu)."""
    _fields_ = [("n_fft", C.c_int), ("hop", C.c_int), ("n_mels", C.c_int), ("n_mfcc", C.c_int),
                ("left_pad", C.c_int), ("trim_left", C.c_int), ("n_frames_all", C.c_int), ("n_frames", C.c_int),
                ("top_db", C.c_float), ("twiddle", C.c_void_p), ("window", C.c_void_p), ("melw", C.c_void_p),
                ("dctm", C.c_void_p), ("sg", C.c_void_p)]


class GrccGzDesc(C.Structure):
    """aewn_grcc_gz_desc: gate derivative of one dilation layer on the fused-layer engine (include/aewn.h)."""
    _fields_ = [("gx16", C.c_void_p), ("gx16_bs", C.c_longlong), ("gx16_cp", C.c_int),
                ("gs16", C.c_void_p), ("gs16_bs", C.c_longlong), ("gs16_cp", C.c_int), ("t_rows", C.c_int),
                ("w2t16", C.c_void_p), ("w_k", C.c_int), ("w_koff_skp", C.c_int),
                ("ab", C.c_void_p), ("a_bs", C.c_longlong), ("a_cs", C.c_longlong),
                ("g16", C.c_void_p), ("g16_bs", C.c_longlong), ("g16_cp", C.c_int), ("gg_off", C.c_int),
                ("batch", C.c_int), ("D", C.c_int), ("t_lo", C.c_int), ("t_zero_lo", C.c_int), ("t_hi", C.c_int),
                ("err", C.c_void_p), ("max_ctas", C.c_int)]


GEN_MAX_LAYERS = 64
GEN_MAX_BLOCKS = 2 * GEN_MAX_LAYERS + 2
GEN_MAX_REP = 4


class GenBlock(C.Structure):
    _fields_ = [("kind", C.c_int), ("rows", C.c_int), ("rowf", C.c_int), ("off", C.c_int)]


class GenDesc(C.Structure):
    _fields_ = [("n_layers", C.c_int), ("R", C.c_int), ("D", C.c_int), ("S", C.c_int), ("P", C.c_int), ("Q", C.c_int),
                ("cluster", C.c_int), ("n_rep", C.c_int), ("n_groups", C.c_int),
                ("t_begin", C.c_int), ("t_end", C.c_int), ("t_prime", C.c_int),
                ("dil", C.c_int * GEN_MAX_LAYERS), ("hist_off", C.c_int * (GEN_MAX_LAYERS + 1)),
                ("n_blocks", C.c_int), ("blocks", GenBlock * GEN_MAX_BLOCKS),
                ("wstream", C.c_void_p), ("stream_stride", C.c_longlong),
                ("cond", C.c_void_p), ("cond_pitch", C.c_int), ("cond_len", C.c_int),
                ("base_t", C.c_void_p), ("base_pitch", C.c_int),
                ("hist", C.c_void_p), ("wav", C.c_void_p), ("wav_pitch", C.c_int),
                ("uniforms", C.c_void_p), ("logits_out", C.c_void_p),
                ("stage_bytes", C.c_int), ("n_stages", C.c_int), ("err", C.c_void_p),
                ("dbg_clock", C.c_void_p)]


_lib = None

# every symbol include/aewn.h declares (tests/test_capi.py checks the shared object exports all of them)
SYMBOLS = ["aewn_version", "aewn_last_error_string", "aewn_launch_count", "aewn_tgemm", "aewn_wgrad", "aewn_wgradw",
           "aewn_base_embed_fwd", "aewn_base_embed_bwd", "aewn_fill", "aewn_relu_mask_bwd",
           "aewn_vq_fwd", "aewn_vq_commit_bwd", "aewn_ema_update", "aewn_pack_blocks", "aewn_add_blocks", "aewn_nll_fwd", "aewn_nll_bwd",
           "aewn_gen_smem_bytes", "aewn_gen_max_clusters", "aewn_gen_run",
           "aewn_grcc_fwd", "aewn_cvt_f16_cl", "aewn_pack_blocks_f16", "aewn_conv1x1_f32", "aewn_conv1x1_wgrad_f32",
           "aewn_grcc_dgrad", "aewn_pack_blocks_bf16",
           "aewn_mu_encode", "aewn_mu_decode", "aewn_jitter_indices", "aewn_mfcc", "aewn_amax_pow2_scale", "aewn_wgradh", "aewn_cvt_f16_cl_scaled", "aewn_grcc_gz"]


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make` or `python -c 'import __graft_entry__ as g; g.build()'`."
                " The aewn hot path has no CPU / eager fallback.")
        _lib = C.CDLL(LIB_PATH)
        _lib.aewn_last_error_string.restype = C.c_char_p
        _lib.aewn_launch_count.restype = C.c_longlong
        _lib.aewn_version.restype = C.c_int
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().aewn_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"aewn: {what} failed (code {rc}): {msg}")


def launch_count():
    return int(lib().aewn_launch_count())

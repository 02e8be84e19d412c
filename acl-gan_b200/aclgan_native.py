"""ctypes binding of libaclgan_b200.so (the C ABI declared in include/aclgan_b200.h).

The shared library is built IN-TREE by `build()` (nvcc, sm_100a) so it travels to the GPU box with
the repo snapshot.  There is no CPU or eager fallback: if the library is missing or a launch
fails, the caller gets an exception.
"""
import ctypes as C
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.environ.get("ACLGAN_LIB") or os.path.join(HERE, "libaclgan_b200.so")      # (ACLGAN_LIB: perf triage with a variant build)

MAX_TAPS = 64
MAX_AVARIANTS = 4

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3
OUT_BF16, OUT_F32, OUT_SPLIT, OUT_F32_ATOMIC = 0, 1, 2, 3


class TmapSpec(C.Structure):
    _fields_ = [("base", C.c_uint64), ("rank", C.c_uint32), ("elem_bytes", C.c_uint32),
                ("dims", C.c_uint64 * 5), ("strides", C.c_uint64 * 5), ("box", C.c_uint32 * 5)]


class OutSpec(C.Structure):
    _fields_ = [("ptr", C.c_uint64 * 2), ("kind", C.c_int32), ("act", C.c_int32), ("slope", C.c_float),
                ("mirror", C.c_int32), ("off", C.c_int64), ("sn", C.c_int64), ("sy", C.c_int64),
                ("sx", C.c_int64), ("sc", C.c_int64), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("C", C.c_int32), ("bias", C.c_uint64), ("bias_n", C.c_int32), ("stats", C.c_uint64),
                ("d2s_c", C.c_int32), ("ring", C.c_int32), ("d2s_sy", C.c_int64), ("d2s_sx", C.c_int64),
                ("z_mod", C.c_int32), ("stats_c", C.c_int32), ("z_off", C.c_int64)]


class IgemmPlan(C.Structure):
    _fields_ = [("a", (TmapSpec * MAX_AVARIANTS) * 2), ("b", TmapSpec * 2),
                ("planes", C.c_int32), ("nseg", C.c_int32), ("n_avariants", C.c_int32),
                ("block_n", C.c_int32), ("n_tiles", C.c_int32),
                ("box_x", C.c_int32), ("box_y", C.c_int32), ("box_z", C.c_int32),
                ("tiles_x", C.c_int32), ("tiles_y", C.c_int32), ("tiles_z", C.c_int32),
                ("cchunks", C.c_int32), ("num_taps", C.c_int32),
                ("tap_dx", C.c_int32 * MAX_TAPS), ("tap_dy", C.c_int32 * MAX_TAPS),
                ("tap_var", C.c_int32 * MAX_TAPS), ("tap_bk", C.c_int32 * MAX_TAPS),
                ("flat", C.c_int32), ("flat_w", C.c_int32), ("flat_img", C.c_int32),
                ("n_groups", C.c_int32), ("group_taps", C.c_int32), ("group_off", C.c_int64 * 4),
                ("seg_mode", C.c_int32), ("seg_rows", C.c_int32), ("num_segs", C.c_int32), ("seg_taps", C.c_int32),
                ("seg_dx", C.c_int32 * 16), ("seg_dy", C.c_int32 * 16), ("tap_row", C.c_int32 * MAX_TAPS),
                ("a_seg", TmapSpec * 2), ("fold", C.c_int32), ("tile_step", C.c_int32),
                ("out", OutSpec)]


class WgradPlan(C.Structure):
    _fields_ = [("mop", (TmapSpec * MAX_AVARIANTS) * 2), ("nop", (TmapSpec * MAX_AVARIANTS) * 2),
                ("planes", C.c_int32), ("nseg", C.c_int32),
                ("n_mvariants", C.c_int32), ("n_nvariants", C.c_int32),
                ("m_chunks", C.c_int32), ("n_chunks", C.c_int32), ("m_tiles", C.c_int32), ("n_tiles", C.c_int32),
                ("box_x", C.c_int32), ("box_y", C.c_int32), ("box_z", C.c_int32),
                ("blocks_x", C.c_int32), ("blocks_y", C.c_int32), ("blocks_z", C.c_int32),
                ("ksplit", C.c_int32), ("num_taps", C.c_int32),
                ("m_dx", C.c_int32 * MAX_TAPS), ("m_dy", C.c_int32 * MAX_TAPS), ("m_var", C.c_int32 * MAX_TAPS),
                ("n_dx", C.c_int32 * MAX_TAPS), ("n_dy", C.c_int32 * MAX_TAPS), ("n_var", C.c_int32 * MAX_TAPS),
                ("tap_out", C.c_int32 * MAX_TAPS),
                ("dw", C.c_uint64), ("dw_sm", C.c_int64), ("dw_st", C.c_int64), ("M", C.c_int32), ("Nn", C.c_int32),
                ("seg_mode", C.c_int32), ("seg_rows", C.c_int32), ("seg_taps", C.c_int32), ("seg_on_m", C.c_int32),
                ("seg_kw0", C.c_int32 * MAX_TAPS), ("seg_cnt", C.c_int32 * MAX_TAPS),
                ("seg_map", TmapSpec * 2), ("seg_step", C.c_int32), ("pad_", C.c_int32)]


class Act(C.Structure):
    _fields_ = [("data", C.c_uint64 * 2), ("planes", C.c_int32), ("n", C.c_int32), ("h", C.c_int32),
                ("w", C.c_int32), ("c", C.c_int32), ("pad", C.c_int32)]


class ConvDesc(C.Structure):
    _fields_ = [("cin", C.c_int32), ("cout", C.c_int32), ("k", C.c_int32), ("stride", C.c_int32),
                ("pad", C.c_int32), ("window", C.c_int32)]


NORM_NONE, NORM_IN, NORM_ADAIN, NORM_LN = 0, 1, 2, 3
MASK_NONE, MASK_FROM_Z, MASK_FROM_OUT = 0, 1, 2
WINDOW_NONE, WINDOW_IN, WINDOW_OUT = 0, 1, 2


class Tensor4(C.Structure):
    _fields_ = [("ptr", C.c_uint64), ("kind", C.c_int32), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("c", C.c_int32)]


class PackImgArgs(C.Structure):
    _fields_ = [("src0", C.c_uint64), ("src1", C.c_uint64), ("c0", C.c_int32), ("c1", C.c_int32),
                ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("dst", Act)]


class NormFinalizeArgs(C.Structure):
    _fields_ = [("mode", C.c_int32), ("n", C.c_int32), ("c", C.c_int32), ("hw", C.c_int32),
                ("c_valid", C.c_int32), ("eps", C.c_float), ("sums", C.c_uint64), ("w", C.c_uint64),
                ("b", C.c_uint64), ("scale", C.c_uint64), ("shift", C.c_uint64), ("mean", C.c_uint64),
                ("inv", C.c_uint64), ("sigma", C.c_uint64), ("wb_stride", C.c_int64), ("stat_groups", C.c_int32),
                ("pad_", C.c_int32)]


class ApplyArgs(C.Structure):
    _fields_ = [("y", Tensor4), ("scale", C.c_uint64), ("shift", C.c_uint64), ("act", C.c_int32),
                ("slope", C.c_float), ("has_res", C.c_int32), ("res", Act), ("upsample", C.c_int32), ("dst", Act)]


class BlockBwdArgs(C.Structure):
    _fields_ = [("gp", C.c_uint64), ("g_kind", C.c_int32), ("gp_pad", C.c_int32), ("upsample", C.c_int32),
                ("gr", C.c_uint64), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("mask_mode", C.c_int32), ("slope", C.c_float), ("y", Tensor4), ("scale", C.c_uint64),
                ("shift", C.c_uint64), ("out", Act), ("norm", C.c_int32), ("mean", C.c_uint64),
                ("inv", C.c_uint64), ("sums", C.c_uint64), ("ca", C.c_uint64), ("cb", C.c_uint64),
                ("cc", C.c_uint64), ("dy", Act), ("dbias", C.c_uint64), ("dbias_n", C.c_int32)]


class NormBwdFinalizeArgs(C.Structure):
    _fields_ = [("mode", C.c_int32), ("n", C.c_int32), ("c", C.c_int32), ("hw", C.c_int32),
                ("c_valid", C.c_int32), ("sums", C.c_uint64), ("inv", C.c_uint64), ("sigma", C.c_uint64),
                ("w", C.c_uint64), ("ca", C.c_uint64), ("cb", C.c_uint64), ("cc", C.c_uint64),
                ("dw", C.c_uint64), ("db", C.c_uint64), ("wb_stride", C.c_int64),
                ("fsums", C.c_uint64), ("mean", C.c_uint64), ("dbias", C.c_uint64), ("fstat_groups", C.c_int32),
                ("pad2_", C.c_int32)]


class ImgGradPackArgs(C.Structure):
    _fields_ = [("dimg", C.c_uint64), ("out_img", C.c_uint64), ("n", C.c_int32), ("c", C.c_int32),
                ("h", C.c_int32), ("w", C.c_int32), ("dy", Act), ("dbias", C.c_uint64)]


class ImgGradUnpackArgs(C.Structure):
    _fields_ = [("src", C.c_uint64), ("n", C.c_int32), ("c", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("cs", C.c_int32), ("pad", C.c_int32), ("c_off", C.c_int32), ("dst", C.c_uint64),
                ("accumulate", C.c_int32)]


class PackWeightArgs(C.Structure):
    _fields_ = [("w", C.c_uint64), ("co", C.c_int32), ("ci", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
                ("base", C.c_int64), ("s_co", C.c_int64), ("s_ci", C.c_int64), ("s_kh", C.c_int64),
                ("s_kw", C.c_int64), ("dst", C.c_uint64 * 2), ("planes", C.c_int32)]


class AdamTensor(C.Structure):
    _fields_ = [("p", C.c_uint64), ("m", C.c_uint64), ("v", C.c_uint64), ("g", C.c_uint64),
                ("d", C.c_int32 * 4), ("gs", C.c_int64 * 4), ("goff", C.c_int64),
                ("pk", (C.c_uint64 * 2) * 2), ("aff", (C.c_int64 * 5) * 2), ("planes", C.c_int32), ("pad_", C.c_int32)]


class AvgPoolArgs(C.Structure):
    _fields_ = [("src", C.c_uint64), ("dst", C.c_uint64), ("planes", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("accumulate", C.c_int32)]


class StyleHeadArgs(C.Structure):
    _fields_ = [("x", Act), ("c_valid", C.c_int32), ("style_dim", C.c_int32), ("weight", C.c_uint64), ("bias", C.c_uint64),
                ("pooled", C.c_uint64), ("style", C.c_uint64), ("dstyle", C.c_uint64), ("dweight", C.c_uint64),
                ("dbias", C.c_uint64), ("gr", C.c_uint64), ("g_kind", C.c_int32), ("pad_", C.c_int32)]


class MlpArgs(C.Structure):
    _fields_ = [("n", C.c_int32), ("n_layers", C.c_int32), ("dims", C.c_int32 * 5), ("pad_", C.c_int32),
                ("w", C.c_uint64 * 4), ("b", C.c_uint64 * 4), ("h", C.c_uint64 * 5), ("h0_stride", C.c_int64),
                ("dh", C.c_uint64 * 5), ("dw", C.c_uint64 * 4), ("db", C.c_uint64 * 4)]


class DisHeadArgs(C.Structure):
    _fields_ = [("x", Act), ("c_valid", C.c_int32), ("groups", C.c_int32), ("weight", C.c_uint64), ("bias", C.c_uint64),
                ("logits", C.c_uint64), ("dlogits", C.c_uint64), ("loss", C.c_uint64), ("target", C.c_float * 4),
                ("gweight", C.c_float * 4), ("loss_slot", C.c_int32 * 4), ("gan_kind", C.c_int32), ("pad_", C.c_int32)]


class DisHeadBwdArgs(C.Structure):
    _fields_ = [("x", Act), ("c_valid", C.c_int32), ("g_kind", C.c_int32), ("weight", C.c_uint64), ("dlogits", C.c_uint64),
                ("dweight", C.c_uint64), ("dbias", C.c_uint64), ("gr", C.c_uint64)]


class BlendArgs(C.Structure):
    _fields_ = [("out4", C.c_uint64), ("bg", C.c_uint64), ("dst", C.c_uint64), ("n", C.c_int32), ("h", C.c_int32),
                ("w", C.c_int32), ("acc_out4", C.c_int32), ("acc_bg", C.c_int32), ("pad_", C.c_int32),
                ("ddst", C.c_uint64), ("dout4", C.c_uint64), ("dbg", C.c_uint64)]


LOSS_L1, LOSS_FOCUS = 0, 1
GAN_LSGAN, GAN_NSGAN = 0, 1         # aclgan_dis_head_args.gan_kind


class LossReduceArgs(C.Structure):
    _fields_ = [("mode", C.c_int32), ("n", C.c_int32), ("ca", C.c_int32), ("c", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("a", C.c_uint64), ("b", C.c_uint64), ("acc", C.c_uint64), ("slot", C.c_int32), ("acc_da", C.c_int32),
                ("da", C.c_uint64), ("gscale", C.c_float), ("upper", C.c_float), ("lower", C.c_float), ("eps", C.c_float)]


class FocusGradArgs(C.Structure):
    _fields_ = [("out4", C.c_uint64), ("dout4", C.c_uint64), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("slot", C.c_int32), ("size_slot", C.c_int32), ("acc", C.c_int32), ("sums", C.c_uint64),
                ("delta", C.c_float), ("eps", C.c_float), ("gscale", C.c_float), ("pad_", C.c_int32)]


class UpDeriveArgs(C.Structure):
    _fields_ = [("w5", C.c_uint64), ("bias", C.c_uint64), ("co", C.c_int32), ("ci", C.c_int32), ("planes", C.c_int32),
                ("pad_", C.c_int32), ("pk", (C.c_uint64 * 2) * 2), ("aff", (C.c_int64 * 5) * 2), ("bias4", C.c_uint64)]


class UpStripsArgs(C.Structure):
    _fields_ = [("src", Act), ("rows", Act), ("cols", Act)]


class UpDyPackArgs(C.Structure):
    _fields_ = [("dy", Act), ("cout", C.c_int32), ("pad_", C.c_int32), ("s2d", Act), ("rows", Act), ("cols", Act)]


class UpScatterArgs(C.Structure):
    _fields_ = [("g", C.c_uint64), ("grows", C.c_uint64), ("gcols", C.c_uint64), ("kind", C.c_int32), ("n", C.c_int32),
                ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32), ("pad_", C.c_int32)]


class UpFoldWgradArgs(C.Structure):
    _fields_ = [("dwp", C.c_uint64), ("affp", C.c_int64 * 5), ("dw5", C.c_uint64), ("aff5", C.c_int64 * 5),
                ("co", C.c_int32), ("ci", C.c_int32)]


class NativeError(RuntimeError):
    pass


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into libaclgan_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = sources()
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "aclgan_b200.h")]
    manifest = os.path.join(CSRC, ".linked_sources")     # a source added to / removed from csrc/ forces a relink
    names = "\n".join(os.path.basename(s) for s in srcs)
    linked = open(manifest).read() if os.path.exists(manifest) else None
    if not force and os.path.exists(LIB_PATH) and linked == names:
        if os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(d) for d in deps):
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    for s in srcs:
        o = os.path.join(CSRC, os.path.basename(s)[:-3] + ".o")
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(d) for d in deps if not d.endswith(".cu") or d == s):
            cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                   "-Xcompiler", "-fPIC", "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas")
                cmd.insert(2, "-v")
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise NativeError("nvcc failed for %s:\n%s\n%s" % (s, r.stdout, r.stderr))
            if verbose:
                print(r.stderr)
        objs.append(o)
    cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise NativeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(manifest, "w") as f:
        f.write(names)
    return LIB_PATH


PROBE_SRC = os.path.join(HERE, "..", "tools", "probe", "probe.cu")
PROBE_LIB = os.path.join(HERE, "..", "tools", "probe", "libaclgan_probe.so")


def build_probe(force=False):
    """The tcgen05 issue-rate / TMEM read micro-benchmarks (tools/probe/probe.cu, driven by tools/probe_umma.py) are a
    measurement tool, not part of the product: they build into their own library next to their source."""
    if not os.path.exists(PROBE_SRC):
        return None
    deps = [PROBE_SRC] + glob.glob(os.path.join(CSRC, "*.cuh"))
    if force or not os.path.exists(PROBE_LIB) or os.path.getmtime(PROBE_LIB) < max(os.path.getmtime(d) for d in deps):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler",
                            "-fPIC", "-I", CSRC, "-shared", PROBE_SRC, "-o", PROBE_LIB], capture_output=True, text=True)
        if r.returncode != 0:
            raise NativeError("nvcc failed for %s:\n%s\n%s" % (PROBE_SRC, r.stdout, r.stderr))
    return PROBE_LIB


_lib = None


def lib():
    """Loads the extension; raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError("libaclgan_b200.so not built: run `python __graft_entry__.py` (build()) first")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(L):
    L.aclgan_version.restype = C.c_int
    L.aclgan_build_info.restype = C.c_char_p
    L.aclgan_plan_conv_fwd.argtypes = [C.POINTER(ConvDesc), C.POINTER(Act), C.c_uint64 * 2, C.POINTER(OutSpec),
                                       C.POINTER(IgemmPlan)]
    L.aclgan_plan_conv_dgrad.argtypes = [C.POINTER(ConvDesc), C.POINTER(Act), C.c_uint64 * 2, C.c_int,
                                         C.POINTER(OutSpec), C.POINTER(IgemmPlan)]
    L.aclgan_packed_weight_shape.argtypes = [C.POINTER(ConvDesc), C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.aclgan_packed_weight_index.argtypes = [C.POINTER(ConvDesc), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.aclgan_packed_weight_index.restype = C.c_int64
    L.aclgan_igemm_launch.argtypes = [C.POINTER(IgemmPlan), C.c_void_p]
    L.aclgan_igemm_stats_supported.argtypes = [C.POINTER(IgemmPlan)]
    L.aclgan_fold_launch.argtypes = [C.POINTER(IgemmPlan), C.c_int, C.c_void_p]
    L.aclgan_plan_conv_wgrad.argtypes = [C.POINTER(ConvDesc), C.POINTER(Act), C.POINTER(Act), C.c_uint64,
                                         C.POINTER(WgradPlan)]
    L.aclgan_wgrad_layout.argtypes = [C.POINTER(ConvDesc)]
    L.aclgan_wgrad_launch.argtypes = [C.POINTER(WgradPlan), C.c_void_p]
    L.aclgan_igemm_launch_repeat.argtypes = [C.POINTER(IgemmPlan), C.c_int, C.c_void_p]
    L.aclgan_wgrad_launch_repeat.argtypes = [C.POINTER(WgradPlan), C.c_int, C.c_void_p]
    L.aclgan_pack_img.argtypes = [C.POINTER(PackImgArgs), C.c_void_p]
    L.aclgan_norm_stats.argtypes = [C.POINTER(Tensor4), C.c_uint64, C.c_void_p]
    L.aclgan_norm_finalize.argtypes = [C.POINTER(NormFinalizeArgs), C.c_void_p]
    L.aclgan_norm_apply.argtypes = [C.POINTER(ApplyArgs), C.c_void_p]
    L.aclgan_norm_finalize_apply.argtypes = [C.POINTER(NormFinalizeArgs), C.POINTER(ApplyArgs), C.c_void_p]
    L.aclgan_block_bwd_reduce.argtypes = [C.POINTER(BlockBwdArgs), C.c_void_p]
    L.aclgan_block_bwd_apply.argtypes = [C.POINTER(BlockBwdArgs), C.c_void_p]
    L.aclgan_norm_bwd_finalize.argtypes = [C.POINTER(NormBwdFinalizeArgs), C.c_void_p]
    L.aclgan_norm_bwd_finalize_apply.argtypes = [C.POINTER(NormBwdFinalizeArgs), C.POINTER(BlockBwdArgs), C.c_void_p]
    L.aclgan_img_grad_pack.argtypes = [C.POINTER(ImgGradPackArgs), C.c_void_p]
    L.aclgan_img_grad_unpack.argtypes = [C.POINTER(ImgGradUnpackArgs), C.c_void_p]
    L.aclgan_pack_weight.argtypes = [C.POINTER(PackWeightArgs), C.c_void_p]
    L.aclgan_adam_step.argtypes = [C.c_uint64, C.c_uint64, C.c_int32, C.c_uint64, C.c_void_p]
    L.aclgan_adam_advance.argtypes = [C.c_uint64, C.c_void_p]
    L.aclgan_adam_units.argtypes = [C.POINTER(AdamTensor)]
    L.aclgan_stream_wait_external_event.argtypes = [C.c_void_p, C.c_void_p]
    L.aclgan_avgpool3x3s2_fwd.argtypes = [C.POINTER(AvgPoolArgs), C.c_void_p]
    L.aclgan_avgpool3x3s2_bwd.argtypes = [C.POINTER(AvgPoolArgs), C.c_void_p]
    L.aclgan_style_head_fwd.argtypes = [C.POINTER(StyleHeadArgs), C.c_void_p]
    L.aclgan_style_head_bwd.argtypes = [C.POINTER(StyleHeadArgs), C.c_void_p]
    L.aclgan_mlp_fwd.argtypes = [C.POINTER(MlpArgs), C.c_void_p]
    L.aclgan_mlp_bwd.argtypes = [C.POINTER(MlpArgs), C.c_void_p]
    L.aclgan_dis_head_fwd.argtypes = [C.POINTER(DisHeadArgs), C.c_void_p]
    L.aclgan_dis_head_bwd.argtypes = [C.POINTER(DisHeadBwdArgs), C.c_void_p]
    L.aclgan_focus_blend_fwd.argtypes = [C.POINTER(BlendArgs), C.c_void_p]
    L.aclgan_focus_blend_bwd.argtypes = [C.POINTER(BlendArgs), C.c_void_p]
    L.aclgan_loss_reduce.argtypes = [C.POINTER(LossReduceArgs), C.c_void_p]
    L.aclgan_focus_grad.argtypes = [C.POINTER(FocusGradArgs), C.c_void_p]
    L.aclgan_loss_combine.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, C.c_void_p]
    L.aclgan_axpby.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_float, C.c_float, C.c_int64, C.c_int32, C.c_void_p]
    L.aclgan_up_derive_weights.argtypes = [C.POINTER(UpDeriveArgs), C.c_void_p]
    L.aclgan_up_gather_strips.argtypes = [C.POINTER(UpStripsArgs), C.c_void_p]
    L.aclgan_up_dy_pack.argtypes = [C.POINTER(UpDyPackArgs), C.c_void_p]
    L.aclgan_up_scatter_strips.argtypes = [C.POINTER(UpScatterArgs), C.c_void_p]
    L.aclgan_up_fold_wgrad.argtypes = [C.POINTER(UpFoldWgradArgs), C.c_void_p]
    L.aclgan_pack_nchw.argtypes = [C.c_uint64, C.c_int32, C.POINTER(Act), C.c_void_p]
    L.aclgan_unpack_plane.argtypes = [C.POINTER(Act), C.c_int32, C.c_uint64, C.c_void_p]
    L.aclgan_augment_u8.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.aclgan_stats_to_bias.argtypes = [C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.aclgan_zero.argtypes = [C.c_uint64, C.c_int64, C.c_void_p]
    L.aclgan_copy.argtypes = [C.c_uint64, C.c_uint64, C.c_int64, C.c_void_p]


launch_count = 0


def check(rc, what):
    global launch_count
    if not what.startswith("plan"):
        launch_count += 1
    if rc != 0:
        raise NativeError("%s failed with status %d (%s)" % (what, rc, status_name(rc)))


_STATUS = {-1: "ACLGAN_ERR_SHAPE: a geometry the kernels do not take, e.g. an input smaller than its reflect padding",
           -2: "ACLGAN_ERR_ALIGN: a pointer / stride that is not aligned for TMA", -3: "ACLGAN_ERR_UNSUPPORTED",
           -4: "ACLGAN_ERR_DRIVER: a CUDA driver entry point (tensor-map encode) failed"}


def status_name(rc):
    """include/aclgan_b200.h: 0 ok, negative ACLGAN_ERR_*, positive a cudaError_t"""
    return _STATUS.get(rc, "cudaError_t %d" % rc if rc > 0 else "unknown status")


def exported_symbols():
    """Symbols include/aclgan_b200.h declares (parsed from the header), for the ABI test."""
    import re
    hdr = open(os.path.join(HERE, "..", "include", "aclgan_b200.h")).read()
    return sorted(set(re.findall(r"\b(aclgan_[a-z0-9_]+)\s*\(", hdr)))

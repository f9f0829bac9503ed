"""ctypes binding of libdlsg.so (include/dlsg.h).  No fallback: a missing library is an error."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('DLSG_LIB') or os.path.join(HERE, 'libdlsg.so')     # DLSG_LIB: an experimental build (tools only)

F32, BF16 = 0, 1
GEMM_SIMT, GEMM_TC = 0, 1
EPI_BIAS_N, EPI_BIAS_M, EPI_TANH, EPI_ACCUM, EPI_STORE_T, EPI_ATOMIC = 1, 2, 4, 8, 16, 32
GEMM_A_STATIC, GEMM_B_STATIC = 64, 128
NORM_PRE_TANH, NORM_POST_TANH, NORM_IN_IS_TANH = 1, 2, 4

i32, i64, u32, u64, f32, vp = C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_float, C.c_void_p


class GemmT(C.Structure):
    _fields_ = [('A', vp), ('B', vp), ('D', vp), ('bias', vp),
                ('M', i32), ('N', i32), ('K', i32), ('batch', i32),
                ('sam', i64), ('sak', i64), ('sbn', i64), ('sbk', i64), ('ldd', i64),
                ('stride_a', i64), ('stride_b', i64), ('stride_d', i64),
                ('a_dtype', i32), ('b_dtype', i32), ('d_dtype', i32), ('impl', i32), ('flags', i32), ('_pad0', i32),
                ('splitk', i32), ('_pad', i32), ('stride_split', i64), ('alpha', f32), ('_pad1', i32),
                ('workspace', vp), ('workspace_bytes', i64)]


class NormFwdT(C.Structure):
    _fields_ = [('x', vp), ('res', vp), ('y', vp), ('y2', vp), ('gamma', vp), ('beta', vp), ('stats', vp),
                ('rows', i64), ('D', i32), ('flags', i32),
                ('ldx', i64), ('ldres', i64), ('ldy', i64), ('ldy2', i64),
                ('x_dtype', i32), ('res_dtype', i32), ('y_dtype', i32), ('y2_dtype', i32),
                ('drop_p', f32), ('_pad', u32), ('seed', u64), ('offset', u64)]


class NormBwdT(C.Structure):
    _fields_ = [('dy', vp), ('x', vp), ('res', vp), ('gamma', vp), ('beta', vp), ('stats', vp),
                ('dx', vp), ('dgamma', vp), ('dbeta', vp),
                ('rows', i64), ('D', i32), ('flags', i32),
                ('lddy', i64), ('ldx', i64), ('ldres', i64), ('lddx', i64),
                ('dy_dtype', i32), ('x_dtype', i32), ('res_dtype', i32), ('dx_dtype', i32),
                ('drop_p', f32), ('dx_accum', i32), ('seed', u64), ('offset', u64), ('dxsum', vp)]


class CellFwdT(C.Structure):
    _fields_ = [('gates', vp), ('nsplit', i32), ('_pad0', i32), ('stride_split', i64),
                ('row_bias', vp), ('ld_row_bias', i64), ('bias', vp),
                ('c_prev', vp), ('c_out', vp), ('h_out', vp),
                ('h2', vp), ('ldh2', i64), ('h2_dtype', i32), ('_pad1', i32),
                ('h3', vp), ('ldh3', i64), ('h3_dtype', i32), ('_pad2', i32),
                ('B', i32), ('H', i32),
                ('drop_p', f32), ('_pad3', i32), ('seed', u64), ('offset', u64)]


class CellBwdT(C.Structure):
    _fields_ = [('acts', vp), ('c_prev', vp), ('c_new', vp), ('dh', vp), ('lddh', i64), ('dh2', vp), ('lddh2', i64), ('dc_next', vp),
                ('dgates', vp), ('dgates2', vp), ('ld_dgates2', i64), ('dgates2_dtype', i32), ('_pad0', i32),
                ('dgatesT', vp), ('ld_dgatesT', i64), ('dgatesT_dtype', i32), ('_pad1', i32),
                ('dc_prev', vp), ('B', i32), ('H', i32),
                ('drop_p', f32), ('_pad2', i32), ('seed', u64), ('offset', u64),
                ('dh2_nsplit', i32), ('_pad3', i32), ('dh2_stride_split', i64),
                ('dc_next2', vp), ('dgates_add', vp), ('dh_total', vp), ('ld_dgates', i64)]


class AdamSegT(C.Structure):
    _fields_ = [('p', vp), ('g', vp), ('m', vp), ('v', vp), ('dst16', vp), ('rows', i64), ('cols', i64), ('ld', i64), ('ld_dst', i64),
                ('ld_g', i64), ('g_dtype', i32), ('_pad', i32), ('step', vp)]


class NormBwd2T(C.Structure):
    _fields_ = [('x', vp), ('dy', vp), ('u', vp), ('gamma', vp), ('stats', vp), ('g_dy', vp), ('g_x', vp), ('g_gamma', vp),
                ('rows', i64), ('D', i32), ('_pad', i32)]


class SegT(C.Structure):
    _fields_ = [('src', vp), ('src2', vp), ('dst', vp), ('rows', i64), ('cols', i64), ('ld_src', i64), ('ld_dst', i64),
                ('src_dtype', i32), ('dst_dtype', i32)]


class CellBwd2T(C.Structure):
    _fields_ = [('acts', vp), ('c_prev', vp), ('c_new', vp), ('dh', vp), ('dc_next', vp), ('u', vp), ('w', vp),
                ('g_dh', vp), ('g_dc', vp), ('g_pre', vp), ('g_cprev', vp), ('B', i32), ('H', i32),
                ('u2', vp), ('u2_stride_split', i64), ('u2_nsplit', i32), ('g_dh2_dtype', i32), ('g_dh2', vp), ('ld_g_dh2', i64),
                ('ld_u', i64), ('ld_g_dh', i64)]


class CellNormFwdT(C.Structure):
    _fields_ = [('cell', CellFwdT), ('gamma', vp), ('beta', vp), ('stats', vp),
                ('y', vp), ('ldy', i64), ('y2', vp), ('ldy2', i64),
                ('y_dtype', i32), ('y2_dtype', i32), ('post_tanh', i32), ('_pad', i32),
                ('ydrop_p', f32), ('_pad2', i32), ('yseed', u64), ('yoffset', u64)]


class NormCellBwdT(C.Structure):
    _fields_ = [('cell', CellBwdT), ('dy', vp), ('lddy', i64), ('x', vp), ('ldx', i64),
                ('gamma', vp), ('beta', vp), ('stats', vp), ('dgamma', vp), ('dbeta', vp), ('ld_dparam', i64), ('dgates_sum', vp),
                ('post_tanh', i32), ('_pad', i32), ('ydrop_p', f32), ('_pad2', i32), ('yseed', u64), ('yoffset', u64)]


class EwT(C.Structure):
    _fields_ = [('inp', vp * 5), ('out', vp * 3), ('n', i64), ('cols', i64), ('op', i32), ('_pad', i32)]


EW_TANH_BWD, EW_TANH_BWD2, EW_MUL_BWD, EW_MUL_BWD2, EW_LERP_ROWS, EW_LERP_ROWS_BWD = range(6)


class SoftmaxT(C.Structure):
    _fields_ = [('x', vp), ('y', vp), ('mask', vp), ('outer', i64), ('n', i64), ('inner', i64),
                ('so', i64), ('sn', i64), ('si', i64), ('scale', f32), ('mask_mode', i32)]


class AttnFwdT(C.Structure):
    _fields_ = [('Kp', vp), ('Vp', vp), ('qp', vp), ('alpha', vp), ('ctx', vp),
                ('rows', i32), ('nh', i32), ('P', i32), ('H', i32), ('rows_per_node', i32), ('ctx_dtype', i32),
                ('ldctx', i64), ('ldalpha', i64), ('nodes', i32), ('_pad', i32)]


class AttnBwdT(C.Structure):
    _fields_ = [('Kp', vp), ('Vp', vp), ('qp', vp), ('alpha', vp), ('dctx', vp), ('dalpha_ext', vp),
                ('dqp', vp), ('dKp', vp), ('dVp', vp),
                ('rows', i32), ('nh', i32), ('P', i32), ('H', i32), ('lddctx', i64), ('ldalpha', i64),
                ('dqp_dtype', i32), ('_pad', i32)]


class Attn2FwdT(C.Structure):
    _fields_ = [('KW', vp), ('VW', vp), ('q', vp), ('alpha', vp), ('co', vp),
                ('ldq', i64), ('ldalpha', i64), ('ldco', i64),
                ('rows', i32), ('nh', i32), ('P', i32), ('Hk', i32), ('Hv', i32), ('rows_per_node', i32), ('nodes', i32),
                ('scale', f32),
                ('y', vp), ('ldy', i64), ('y_dtype', i32), ('_pad0', i32),
                ('gamma', vp * 2), ('beta', vp * 2), ('stats', vp), ('stats_head_stride', i64),
                ('drop_p', f32), ('_pad1', i32), ('seed', u64), ('offset', u64), ('offset_head_stride', u64)]


class CellNormAttn2FwdT(C.Structure):
    _fields_ = [('cn', CellNormFwdT), ('at', Attn2FwdT)]


class Attn2BwdT(C.Structure):
    _fields_ = [('KW', vp), ('VW', vp), ('q', vp), ('alpha', vp), ('dco', vp), ('dalpha_ext', vp),
                ('dq', vp), ('dKW', vp), ('dVW', vp),
                ('ldq', i64), ('ldalpha', i64), ('lddco', i64), ('lddq', i64),
                ('rows', i32), ('nh', i32), ('P', i32), ('Hk', i32), ('Hv', i32), ('scale', f32),
                ('dy', vp), ('lddy', i64), ('co', vp), ('ldco', i64),
                ('gamma', vp * 2), ('stats', vp), ('stats_head_stride', i64),
                ('dgamma_rows', vp), ('dbeta_rows', vp), ('ld_dparam', i64),
                ('drop_p', f32), ('_pad1', i32), ('seed', u64), ('offset', u64), ('offset_head_stride', u64),
                ('dl_save', vp), ('ld_dl_save', i64), ('dco_save', vp), ('ld_dco_save', i64)]


class LstmStepT(C.Structure):
    _fields_ = [('W', vp * 4), ('h_in', vp * 4), ('ldh_in', i64), ('gin', vp * 4), ('ldgin', i64), ('c_in', vp * 4), ('c_out', vp * 4),
                ('acts', vp * 4), ('h_out', vp * 4), ('ldh_out', i64), ('h_op', vp * 4), ('ldh_op', i64),
                ('B', i32), ('H', i32), ('ndir', i32), ('_pad', i32)]


class RegionAggFwdT(C.Structure):
    _fields_ = [('Y', vp * 2), ('ldy', i64), ('F', vp * 2), ('ldf', i64), ('gamma', vp * 2), ('beta', vp * 2),
                ('agg', vp * 2), ('ldagg', i64), ('U', vp * 2), ('ldu', i64),
                ('stats', vp * 2), ('St', vp * 2), ('tconst', vp * 2),
                ('B', i32), ('E', i32), ('T', i32), ('TR', i32), ('H', i32), ('scores_only', i32), ('scale', f32), ('_pad', i32)]


class RegionAggBwdT(C.Structure):
    _fields_ = [('Y', vp * 2), ('ldy', i64), ('stats', vp * 2), ('St', vp * 2), ('dSm', vp * 2),
                ('F', vp * 2), ('ldf', i64), ('dA', vp * 2), ('ldda', i64), ('U', vp * 2), ('ldu', i64),
                ('tcF', vp * 2), ('tcA', vp * 2), ('gamma', vp * 2), ('beta', vp * 2),
                ('dpre', vp * 2), ('ldd', i64), ('dF', vp * 2), ('lddf', i64),
                ('dgamma', vp * 2), ('dbeta', vp * 2), ('dbias', vp * 2), ('work', vp * 2),
                ('B', i32), ('E', i32), ('T', i32), ('TR', i32), ('H', i32), ('_pad', i32), ('scale', f32), ('_pad2', i32)]


SIGNATURES = {
    'dlsg_lstm_step_supported': (i32, [i32, i32]),
    'dlsg_lstm_step_fwd': (i32, [C.POINTER(LstmStepT), vp]),
    'dlsg_region_aggregate_supported': (i32, [i32, i32, i32]),
    'dlsg_region_aggregate_fwd': (i32, [C.POINTER(RegionAggFwdT), vp]),
    'dlsg_region_aggregate_bwd_workspace': (i64, [i32, i32, i32]),
    'dlsg_region_aggregate_bwd': (i32, [C.POINTER(RegionAggBwdT), vp]),
    'dlsg_attn2_supported': (i32, [i32, i32, i32, i32]),
    'dlsg_attn2_fwd': (i32, [C.POINTER(Attn2FwdT), vp]),
    'dlsg_cell_norm_attn2_supported': (i32, [C.POINTER(CellNormAttn2FwdT)]),
    'dlsg_cell_norm_attn2_fwd': (i32, [C.POINTER(CellNormAttn2FwdT), vp]),
    'dlsg_attn2_bwd': (i32, [C.POINTER(Attn2BwdT), vp]),
    'dlsg_attn2_bwd_nodes': (i32, [vp, i64, i64, vp, i64, i64, vp, i64, i64, vp, i64, i64, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]),
    'dlsg_version': (i32, []),
    'dlsg_sm_arch': (i32, []),
    'dlsg_last_error': (C.c_char_p, []),
    'dlsg_debug_gemm_trace': (None, [vp]),
    'dlsg_gemm': (i32, [C.POINTER(GemmT), vp]),
    'dlsg_convert2d': (i32, [vp, i32, i64, vp, i32, i64, vp, i64, i64, i64, vp]),
    'dlsg_convert2d_batched': (i32, [vp, i32, i64, vp, i32, i64, vp, i64, i64, i64, i64, i64, i64, i64, vp]),
    'dlsg_multi_convert': (i32, [vp, vp, i32, vp]),
    'dlsg_multi_convert_host': (i32, [vp, i32, i32, vp]),
    'dlsg_adam_multi': (i32, [vp, i32, i32, vp, vp, f32, f32, f32, f32, vp]),
    'dlsg_colsum': (i32, [vp, i32, i64, i64, i64, vp, vp]),
    'dlsg_norm_fwd': (i32, [C.POINTER(NormFwdT), vp]),
    'dlsg_norm_bwd': (i32, [C.POINTER(NormBwdT), vp]),
    'dlsg_norm_bwd_streaming': (i32, [C.POINTER(NormBwdT)]),
    'dlsg_norm_bwd2': (i32, [C.POINTER(NormBwd2T), vp]),
    'dlsg_lstm_cell_fwd': (i32, [C.POINTER(CellFwdT), vp]),
    'dlsg_lstm_cell_bwd': (i32, [C.POINTER(CellBwdT), vp]),
    'dlsg_lstm_cell_bwd2': (i32, [C.POINTER(CellBwd2T), vp]),
    'dlsg_lstm_cell_norm_fwd': (i32, [C.POINTER(CellNormFwdT), vp]),
    'dlsg_norm_lstm_cell_bwd': (i32, [C.POINTER(NormCellBwdT), vp]),
    'dlsg_fused_step_supported': (i32, [i32]),
    'dlsg_softmax_fwd': (i32, [C.POINTER(SoftmaxT), vp]),
    'dlsg_softmax_bwd': (i32, [C.POINTER(SoftmaxT), vp, vp, vp]),
    'dlsg_softmax_bwd2': (i32, [C.POINTER(SoftmaxT), vp, vp, vp, vp, vp]),
    'dlsg_ew': (i32, [C.POINTER(EwT), vp]),
    'dlsg_node_attn_fwd': (i32, [C.POINTER(AttnFwdT), vp]),
    'dlsg_node_attn_bwd': (i32, [C.POINTER(AttnBwdT), vp]),
    'dlsg_latent_psl_fwd': (i32, [vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    'dlsg_latent_psl_bwd': (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    'dlsg_latent_psl_fwd_multi': (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    'dlsg_latent_psl_bwd_multi': (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    'dlsg_embedding_gather': (i32, [vp, vp, i64, i32, i32, vp, i32, i64, vp, i32, i64, f32, u64, u64, vp]),
    'dlsg_embedding_scatter_add': (i32, [vp, vp, i64, i32, i32, vp, i64, f32, u64, u64, vp]),
    'dlsg_mean_nodes_fwd': (i32, [vp, i32, i32, i32, vp, i64, vp]),
    'dlsg_mean_nodes_bwd': (i32, [vp, i64, i32, i32, i32, vp, vp]),
    'dlsg_axpby': (i32, [vp, f32, vp, f32, i64, vp]),
    'dlsg_add_rowbcast': (i32, [vp, vp, vp, i64, i64, f32, u64, u64, vp]),
    'dlsg_dropout': (i32, [vp, vp, i64, f32, u64, u64, vp]),
    'dlsg_relu': (i32, [vp, i64, vp]),
    'dlsg_relu_bwd': (i32, [vp, vp, vp, i64, vp]),
    'dlsg_mul': (i32, [vp, vp, vp, i64, vp]),
    'dlsg_row_argmax': (i32, [vp, i64, i32, i32, vp, i64, vp]),
    'dlsg_log_softmax': (i32, [vp, i64, i32, i32, vp, i64, vp]),
    'dlsg_ce_masked': (i32, [vp, vp, vp, i32, i32, i32, vp, vp, f32, vp, vp, vp]),
    'dlsg_beam_topk': (i32, [vp, i64, i32, i32, vp, i32, i32, vp, vp, i32, vp]),
    'dlsg_beam_merge': (i32, [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, i32, vp]),
    'dlsg_beam_gather': (i32, [vp, vp, vp, i32, i32, i32, vp]),
    'dlsg_beam_gather_multi': (i32, [vp, vp, vp, i32, vp, i32, i32, vp]),
    'dlsg_beam_backtrack': (i32, [vp, vp, i32, i32, i32, vp, vp]),
}

_lib = None


class DlsgError(RuntimeError):
    pass


def load():
    """Load libdlsg.so (built by __graft_entry__.build() / dlsg/build.py). Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DlsgError('libdlsg.so not found at %s - run `python __graft_entry__.py` (build) first; '
                        'there is no CPU / PyTorch fallback for the D-LSG hot path' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().dlsg_last_error()
        raise DlsgError('%s failed (rc=%d): %s' % (what, rc, msg.decode() if msg else ''))

"""ctypes binding of include/pcvae_b200.h (libpcvae_b200.so).

The library is the product: there is no Python/torch fallback for any entry
point.  Loading fails loudly if the shared object is missing, and every call
raises `PcvaeError` on a non-zero return code (for example on a non-sm_100
device).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpcvae_b200.so")

FAMILY_MLP, FAMILY_PNP, FAMILY_MLP_MASK = 0, 1, 2     # MLP_MASK: first encoder layer reads [x*mask, mask]
MASK_U8, MASK_F32 = 0, 1
DEC_FWD, DEC_BWD, DEC_TRAIN, DEC_EVAL = 0, 1, 2, 3
NSUMS = 8
S_RE_Q, S_RE_P, S_KL_Q, S_KL_P, S_KL_REG, S_RE_D, S_RE_IMP, S_SSE_UNOBS = range(8)

#: every symbol include/pcvae_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "pcvae_abi_version", "pcvae_last_error", "pcvae_param_count", "pcvae_param_offsets",
    "pcvae_decoder_offset", "pcvae_enc_act_ws_floats", "pcvae_enc_fwd", "pcvae_enc_bwd", "pcvae_dec",
    "pcvae_loss_terms", "pcvae_grid_ctas", "pcvae_reduce_sums", "pcvae_reduce_grads", "pcvae_adam_step",
    "pcvae_reward_workspace_bytes", "pcvae_reward_chain", "pcvae_ffma_probe", "pcvae_gather_rows",
    "pcvae_draw_submask", "pcvae_draw_normal", "pcvae_dense_fwd", "pcvae_dense_bwd", "pcvae_mnar_sample_z",
    "pcvae_mnar_sample_z_bwd", "pcvae_mnar_loss_workspace_bytes", "pcvae_mnar_loss", "pcvae_set_reward_tensor_cores",
    "pcvae_dec_tc_workspace_floats", "pcvae_set_train_tensor_cores", "pcvae_enc_tc_workspace_floats",
    "pcvae_weight_images_floats", "pcvae_build_weight_images",
    "pcvae_prep_batch", "pcvae_reduce_adam", "pcvae_profile_events", "pcvae_prep_packed",
    "pcvae_dp_exchange_bytes", "pcvae_dp_exchange_alloc", "pcvae_dp_exchange_open", "pcvae_dp_exchange_close",
    "pcvae_dp_exchange_free", "pcvae_dp_reduce_adam", "pcvae_prep_batch_dev", "pcvae_reduce_adam_dev",
    "pcvae_dp_reduce_adam_emulated", "pcvae_miwae_heads", "pcvae_miwae_heads_bwd", "pcvae_miwae_sample_z",
    "pcvae_miwae_sample_z_bwd", "pcvae_miwae_loss_workspace_bytes", "pcvae_miwae_loss",
    "pcvae_mnar_impute_workspace_bytes", "pcvae_mnar_impute",
]


class PcvaeError(RuntimeError):
    pass


DP_MAX_WORLD = 16


class DpParams(C.Structure):
    """pcvae_dp_params (include/pcvae_b200.h)."""
    _fields_ = [("grad_partials", C.c_void_p), ("grid", C.c_int), ("param_count", C.c_long),
                ("grad", C.c_void_p), ("theta", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("step", C.c_int), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("sums_partials", C.c_void_p), ("rows", C.c_int), ("obs_dim", C.c_int), ("sums", C.c_void_p),
                ("world", C.c_int), ("rank", C.c_int), ("seq", C.c_uint),
                ("peer_buffers", C.c_void_p * DP_MAX_WORLD), ("status", C.c_void_p), ("step_state", C.c_void_p)]


class Model(C.Structure):
    _fields_ = [("family", C.c_int), ("obs_dim", C.c_int), ("emb_dim", C.c_int), ("latent_dim", C.c_int)]


_P2 = C.c_void_p * 2


class EncFwdParams(C.Structure):
    _fields_ = [("model", Model), ("rows", C.c_int), ("n_branch", C.c_int), ("mask_kind", C.c_int),
                ("theta", C.c_void_p), ("x", C.c_void_p), ("mask", _P2), ("eps", _P2), ("mean", _P2),
                ("logvar", _P2), ("z", _P2), ("act_ws", C.c_void_p), ("pnp_ac", C.c_void_p),
                ("tc_workspace", C.c_void_p), ("tc_workspace_floats", C.c_long), ("weight_images", C.c_void_p)]


class EncBwdParams(C.Structure):
    _fields_ = [("model", Model), ("rows", C.c_int), ("n_branch", C.c_int), ("mask_kind", C.c_int),
                ("theta", C.c_void_p), ("x", C.c_void_p), ("mask", _P2), ("act_ws", C.c_void_p),
                ("d_mean", _P2), ("d_logvar", _P2), ("pnp_ac", C.c_void_p), ("grad_partials", C.c_void_p),
                ("d_z", _P2), ("eps", _P2), ("logvar", _P2),
                ("tc_workspace", C.c_void_p), ("tc_workspace_floats", C.c_long), ("weight_images", C.c_void_p)]


class DecParams(C.Structure):
    _fields_ = [("model", Model), ("mode", C.c_int), ("rows", C.c_int), ("n_branch", C.c_int),
                ("mask_kind", C.c_int), ("theta", C.c_void_p), ("z", _P2), ("xhat", _P2), ("x", C.c_void_p),
                ("mask", _P2), ("mean", _P2), ("logvar", _P2), ("eps", _P2),
                ("alpha", C.c_float), ("beta_w", C.c_float), ("x_logvar", C.c_float), ("loss_scale", C.c_float),
                ("sums_partials", C.c_void_p), ("d_mean", _P2), ("d_logvar", _P2), ("d_xhat", _P2),
                ("d_z", _P2), ("grad_partials", C.c_void_p), ("tc_workspace", C.c_void_p), ("tc_workspace_floats", C.c_long),
                ("weight_images", C.c_void_p)]


class LossParams(C.Structure):
    _fields_ = [("rows", C.c_int), ("obs_dim", C.c_int), ("latent_dim", C.c_int), ("n_branch", C.c_int),
                ("mask_kind", C.c_int), ("x", C.c_void_p), ("mask", _P2), ("xhat", _P2), ("mean", _P2),
                ("logvar", _P2), ("alpha", C.c_float), ("beta_w", C.c_float), ("x_logvar", C.c_float),
                ("loss_scale", C.c_float), ("sums_partials", C.c_void_p), ("d_xhat", _P2), ("d_mean", _P2),
                ("d_logvar", _P2)]


class RewardParams(C.Structure):
    _fields_ = [("model", Model), ("rows", C.c_int), ("samples", C.c_int), ("mask_kind", C.c_int),
                ("theta", C.c_void_p), ("x", C.c_void_p), ("mask", C.c_void_p), ("im", C.c_void_p),
                ("im_sample_stride", C.c_long), ("R", C.c_void_p), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_size_t), ("pnp_ac", C.c_void_p)]


ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_ELU, ACT_HARDTANH_M10_0 = range(5)


class DenseFwdParams(C.Structure):
    _fields_ = [("rows", C.c_int), ("in_dim", C.c_int), ("out_dim", C.c_int), ("act", C.c_int), ("x", C.c_void_p),
                ("mask", C.c_void_p), ("W", C.c_void_p), ("b", C.c_void_p), ("y", C.c_void_p)]


class DenseBwdParams(C.Structure):
    _fields_ = [("rows", C.c_int), ("in_dim", C.c_int), ("out_dim", C.c_int), ("act", C.c_int), ("x", C.c_void_p),
                ("mask", C.c_void_p), ("y", C.c_void_p), ("dy", C.c_void_p), ("W", C.c_void_p), ("dx", C.c_void_p),
                ("dW_partials", C.c_void_p), ("db_partials", C.c_void_p), ("dW", C.c_void_p), ("db", C.c_void_p)]


class MnarLossParams(C.Structure):
    _fields_ = [("rows", C.c_int), ("samples", C.c_int), ("obs_dim", C.c_int), ("latent_dim", C.c_int),
                ("regularised", C.c_int), ("x", C.c_void_p), ("mask", C.c_void_p), ("mask_p", C.c_void_p),
                ("xm", _P2), ("xlv", _P2), ("mean", _P2), ("logvar", _P2), ("eps_kl", C.c_void_p), ("W", C.c_void_p),
                ("b", C.c_void_p), ("alpha", C.c_float), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("out", C.c_void_p), ("xm_imputed", C.c_void_p), ("d_xm", _P2), ("d_xlv", _P2), ("d_mean", _P2),
                ("d_logvar", _P2), ("d_W", C.c_void_p), ("d_b", C.c_void_p)]


class MiwaeLossParams(C.Structure):
    """pcvae_miwae_loss_params (include/pcvae_b200.h)."""
    _fields_ = [("rows", C.c_int), ("samples", C.c_int), ("obs_dim", C.c_int), ("latent_dim", C.c_int),
                ("regularised", C.c_int), ("mask_kind", C.c_int), ("rowwise", C.c_int), ("x", C.c_void_p),
                ("mask", C.c_void_p), ("mask_p", C.c_void_p), ("xm", _P2), ("xs", _P2), ("df", _P2), ("mean", _P2),
                ("scale", _P2), ("eps2", _P2), ("alpha", C.c_float), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_size_t), ("out", C.c_void_p), ("xm_imputed", C.c_void_p), ("d_xm", _P2),
                ("d_xs", _P2), ("d_df", _P2), ("d_mean", _P2), ("d_scale", _P2)]


class MnarImputeParams(C.Structure):
    """pcvae_mnar_impute_params (include/pcvae_b200.h)."""
    _fields_ = [("rows", C.c_int), ("samples", C.c_int), ("obs_dim", C.c_int), ("latent_dim", C.c_int),
                ("regularised", C.c_int), ("dec0_W", C.c_void_p), ("dec0_b", C.c_void_p), ("dec2_W", C.c_void_p),
                ("dec2_b", C.c_void_p), ("xmean_W", C.c_void_p), ("xmean_b", C.c_void_p), ("xlogvar_W", C.c_void_p),
                ("xlogvar_b", C.c_void_p), ("W", C.c_void_p), ("b", C.c_void_p), ("x", C.c_void_p), ("mask", C.c_void_p),
                ("mean", C.c_void_p), ("logvar", C.c_void_p), ("eps", C.c_void_p), ("eps_kl", C.c_void_p),
                ("seed", C.c_ulonglong), ("offset", C.c_ulonglong), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_size_t), ("xm_imputed", C.c_void_p)]


MIWAE_HEADS_ENC, MIWAE_HEADS_DEC = 0, 1

_lib = None


def load():
    """dlopen libpcvae_b200.so (built in-tree by build.py / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PcvaeError(f"{LIB_PATH} is missing: run `python -m vae_posterior_consistency_b200.build` "
                         "(there is no CPU or PyTorch fallback for this path)")
    try:
        import torch  # noqa: F401  -- makes libcudart.so.12 resident so the soname resolves
    except Exception:  # pragma: no cover
        pass
    lib = C.CDLL(LIB_PATH)
    lib.pcvae_last_error.restype = C.c_char_p
    lib.pcvae_param_count.restype = C.c_long
    lib.pcvae_param_count.argtypes = [C.POINTER(Model)]
    lib.pcvae_param_offsets.argtypes = [C.POINTER(Model), C.POINTER(C.c_long)]
    lib.pcvae_decoder_offset.restype = C.c_long
    lib.pcvae_decoder_offset.argtypes = [C.POINTER(Model)]
    lib.pcvae_enc_act_ws_floats.restype = C.c_size_t
    lib.pcvae_enc_act_ws_floats.argtypes = [C.POINTER(Model), C.c_int, C.c_int]
    lib.pcvae_enc_tc_workspace_floats.restype = C.c_long
    lib.pcvae_enc_tc_workspace_floats.argtypes = [C.POINTER(Model), C.c_int, C.c_int]
    lib.pcvae_enc_fwd.argtypes = [C.POINTER(EncFwdParams), C.c_void_p]
    lib.pcvae_enc_bwd.argtypes = [C.POINTER(EncBwdParams), C.c_void_p]
    lib.pcvae_dec.argtypes = [C.POINTER(DecParams), C.c_void_p]
    lib.pcvae_dec_tc_workspace_floats.restype = C.c_long
    lib.pcvae_dec_tc_workspace_floats.argtypes = [C.POINTER(Model), C.c_int, C.c_int]
    lib.pcvae_set_train_tensor_cores.argtypes = [C.c_int]
    lib.pcvae_weight_images_floats.restype = C.c_long
    lib.pcvae_weight_images_floats.argtypes = [C.POINTER(Model)]
    lib.pcvae_build_weight_images.argtypes = [C.POINTER(Model), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pcvae_set_reward_tensor_cores.argtypes = [C.c_int]
    lib.pcvae_profile_events.argtypes = [C.c_void_p, C.c_int]
    lib.pcvae_loss_terms.argtypes = [C.POINTER(LossParams), C.c_void_p]
    lib.pcvae_reduce_sums.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.pcvae_reduce_grads.argtypes = [C.c_void_p, C.c_int, C.c_long, C.c_long, C.c_long, C.c_void_p, C.c_int,
                                       C.c_void_p]
    lib.pcvae_adam_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_float,
                                    C.c_float, C.c_float, C.c_float, C.c_void_p]
    lib.pcvae_reward_workspace_bytes.restype = C.c_size_t
    lib.pcvae_reward_workspace_bytes.argtypes = [C.POINTER(Model), C.c_int, C.c_int]
    lib.pcvae_reward_chain.argtypes = [C.POINTER(RewardParams), C.c_void_p]
    lib.pcvae_gather_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.c_int, C.c_void_p]
    lib.pcvae_draw_submask.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_float, C.c_ulonglong, C.c_ulonglong,
                                       C.c_void_p]
    lib.pcvae_draw_normal.argtypes = [C.c_void_p, C.c_long, C.c_ulonglong, C.c_ulonglong, C.c_void_p]
    lib.pcvae_prep_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_int, C.c_int, C.c_int, C.c_float, C.c_ulonglong, C.c_ulonglong, C.c_void_p]
    lib.pcvae_prep_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_int, C.c_int, C.c_float, C.c_ulonglong, C.c_ulonglong, C.c_void_p]
    lib.pcvae_prep_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int, C.c_int, C.c_int, C.c_float, C.c_ulonglong, C.c_ulonglong, C.c_void_p, C.c_void_p]
    lib.pcvae_reduce_adam_dev.argtypes = [C.c_void_p, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p]
    lib.pcvae_dp_exchange_bytes.restype = C.c_size_t
    lib.pcvae_dp_exchange_bytes.argtypes = [C.c_long, C.c_int]
    lib.pcvae_dp_exchange_alloc.argtypes = [C.c_long, C.c_int, C.POINTER(C.c_void_p), C.c_char_p]
    lib.pcvae_dp_exchange_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.pcvae_dp_exchange_close.argtypes = [C.c_void_p]
    lib.pcvae_dp_exchange_free.argtypes = [C.c_void_p]
    lib.pcvae_dp_reduce_adam.argtypes = [C.POINTER(DpParams), C.c_void_p]
    lib.pcvae_dp_reduce_adam_emulated.argtypes = [C.POINTER(C.POINTER(DpParams)), C.c_int, C.c_void_p]
    lib.pcvae_reduce_adam.argtypes = [C.c_void_p, C.c_int, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p]
    lib.pcvae_dense_fwd.argtypes = [C.POINTER(DenseFwdParams), C.c_void_p]
    lib.pcvae_dense_bwd.argtypes = [C.POINTER(DenseBwdParams), C.c_void_p]
    lib.pcvae_mnar_sample_z.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p]
    lib.pcvae_mnar_sample_z_bwd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                            C.c_int, C.c_int, C.c_void_p]
    lib.pcvae_mnar_loss_workspace_bytes.restype = C.c_size_t
    lib.pcvae_mnar_loss_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.pcvae_mnar_loss.argtypes = [C.POINTER(MnarLossParams), C.c_void_p]
    lib.pcvae_miwae_heads.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pcvae_miwae_heads_bwd.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]
    lib.pcvae_miwae_sample_z.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p]
    lib.pcvae_miwae_sample_z_bwd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p]
    lib.pcvae_miwae_loss_workspace_bytes.restype = C.c_size_t
    lib.pcvae_miwae_loss_workspace_bytes.argtypes = [C.c_int, C.c_int]
    lib.pcvae_miwae_loss.argtypes = [C.POINTER(MiwaeLossParams), C.c_void_p]
    lib.pcvae_mnar_impute_workspace_bytes.restype = C.c_size_t
    lib.pcvae_mnar_impute_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.pcvae_mnar_impute.argtypes = [C.POINTER(MnarImputeParams), C.c_void_p]
    lib.pcvae_ffma_probe.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_void_p]
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().pcvae_last_error().decode(errors="replace")
        raise PcvaeError(f"{what}: error {rc}: {msg}")


def model(family: int, obs_dim: int, emb_dim: int = 0, latent_dim: int = 10) -> Model:
    return Model(family, obs_dim, emb_dim if family == FAMILY_PNP else 0, latent_dim)


def param_count(m: Model) -> int:
    n = load().pcvae_param_count(C.byref(m))
    if n < 0:
        check(1, "pcvae_param_count")
    return n


def param_offsets(m: Model):
    buf = (C.c_long * 17)()
    n = load().pcvae_param_offsets(C.byref(m), buf)
    if n < 0:
        check(1, "pcvae_param_offsets")
    return list(buf[: n + 1])


def decoder_offset(m: Model) -> int:
    return load().pcvae_decoder_offset(C.byref(m))

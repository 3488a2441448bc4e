"""PyTorch custom ops (`torch.library`, namespace `pcvae::`) over the C ABI, with autograd.

These are what the nn.Module mirror in VAE.py composes, so that `model.forward`,
`model.loss` and `train_loss.backward()` keep working exactly as the reference's callers
expect (src/experiment_main/train.py:87-116) while every arithmetic step runs in
libpcvae_b200.so.  All ops take the FLAT parameter vector `theta` (a differentiable
`torch.cat` of the module's nn.Parameters), so gradients flow back to the individual
parameters through autograd's own cat/view bookkeeping.

There is no CPU implementation: calling an op with CPU tensors raises PcvaeError.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from . import kernels as KR
from . import lib as L

_ENGINES = {}


def engine(family: int, D: int, K: int, device) -> KR.Engine:
    device = torch.device(device)
    if device.type != "cuda":
        raise L.PcvaeError("pcvae ops need CUDA tensors: there is no CPU fallback for this path")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = (family, D, K if family == L.FAMILY_PNP else 0, device.index)
    if key not in _ENGINES:
        _ENGINES[key] = KR.Engine(family, D, K, device)
    return _ENGINES[key]


# ------------------------------------------------------------------------------------------
# encoder
# ------------------------------------------------------------------------------------------

@torch.library.custom_op("pcvae::encoder", mutates_args=())
def encoder_op(theta: Tensor, x: Tensor, mask: Tensor, eps: Optional[Tensor], family: int, D: int,
               K: int) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """(z, mean, logvar, saved activations).  VAE.py:387-395 / 719-741."""
    eng = engine(family, D, K, x.device)
    mean, logvar, z, ws = eng.enc_fwd(theta, x, [mask], [eps], save=True)
    return z[0], mean[0], logvar[0], ws


@encoder_op.register_fake
def _(theta, x, mask, eps, family, D, K):
    B = x.shape[0]
    mk = lambda: x.new_empty(B, KR.LATENT, dtype=torch.float32)
    return mk(), mk(), mk(), x.new_empty(1, dtype=torch.float32)


@torch.library.custom_op("pcvae::encoder_bwd", mutates_args=())
def encoder_bwd_op(theta: Tensor, x: Tensor, mask: Tensor, eps: Optional[Tensor], logvar: Tensor, ws: Tensor,
                   d_z: Tensor, d_mean: Tensor, d_logvar: Tensor, family: int, D: int, K: int) -> Tensor:
    eng = engine(family, D, K, x.device)
    eng.enc_bwd(theta, x, [mask], ws, [d_mean], [d_logvar], d_z=[d_z], eps=[eps], logvar=[logvar])
    grad = torch.zeros_like(theta)
    eng.reduce_grads(grad, 0, eng.dec_off)
    return grad


@encoder_bwd_op.register_fake
def _(theta, x, mask, eps, logvar, ws, d_z, d_mean, d_logvar, family, D, K):
    return torch.empty_like(theta)


def _enc_setup(ctx, inputs, output):
    theta, x, mask, eps, family, D, K = inputs
    z, mean, logvar, ws = output
    ctx.save_for_backward(theta, x, mask, eps, logvar, ws)
    ctx.cfg = (family, D, K)


def _enc_backward(ctx, d_z, d_mean, d_logvar, d_ws):
    theta, x, mask, eps, logvar, ws = ctx.saved_tensors
    family, D, K = ctx.cfg
    zeros = lambda: torch.zeros(x.shape[0], KR.LATENT, device=x.device)
    d_z = zeros() if d_z is None else d_z
    d_mean = zeros() if d_mean is None else d_mean
    d_logvar = zeros() if d_logvar is None else d_logvar
    g = encoder_bwd_op(theta, x, mask, eps, logvar, ws, d_z.contiguous(), d_mean.contiguous(),
                       d_logvar.contiguous(), family, D, K)
    return g, None, None, None, None, None, None


torch.library.register_autograd("pcvae::encoder", _enc_backward, setup_context=_enc_setup)


# ------------------------------------------------------------------------------------------
# decoder
# ------------------------------------------------------------------------------------------

@torch.library.custom_op("pcvae::decoder", mutates_args=())
def decoder_op(theta: Tensor, z: Tensor, family: int, D: int, K: int) -> Tensor:
    """xhat = Sigmoid(seq_decoder(z)), VAE.py:397-401."""
    eng = engine(family, D, K, z.device)
    return eng.dec(L.DEC_FWD, theta, [z])["xhat"][0]


@decoder_op.register_fake
def _(theta, z, family, D, K):
    return z.new_empty(z.shape[0], D, dtype=torch.float32)


@torch.library.custom_op("pcvae::decoder_bwd", mutates_args=())
def decoder_bwd_op(theta: Tensor, z: Tensor, d_xhat: Tensor, family: int, D: int, K: int) -> Tuple[Tensor, Tensor]:
    eng = engine(family, D, K, z.device)
    out = eng.dec(L.DEC_BWD, theta, [z], d_xhat=[d_xhat])
    grad = torch.zeros_like(theta)
    eng.reduce_grads(grad, eng.dec_off, eng.P)
    return grad, out["d_z"][0]


@decoder_bwd_op.register_fake
def _(theta, z, d_xhat, family, D, K):
    return torch.empty_like(theta), torch.empty_like(z)


def _dec_setup(ctx, inputs, output):
    theta, z, family, D, K = inputs
    ctx.save_for_backward(theta, z)
    ctx.cfg = (family, D, K)


def _dec_backward(ctx, d_xhat):
    theta, z = ctx.saved_tensors
    family, D, K = ctx.cfg
    g, dz = decoder_bwd_op(theta, z, d_xhat.contiguous(), family, D, K)
    return g, dz, None, None, None


torch.library.register_autograd("pcvae::decoder", _dec_backward, setup_context=_dec_setup)


# ------------------------------------------------------------------------------------------
# loss
# ------------------------------------------------------------------------------------------

@torch.library.custom_op("pcvae::vae_loss", mutates_args=())
def vae_loss_op(x: Tensor, mask: Tensor, mask_p: Optional[Tensor], xhat_q: Tensor, xhat_p: Optional[Tensor],
                mean_q: Tensor, logvar_q: Tensor, mean_p: Optional[Tensor], logvar_p: Optional[Tensor],
                alpha: float, beta_w: float, loss_scale: float, x_logvar: float, want_grads: bool
                ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """(loss * loss_scale as a 0-dim fp32 tensor, sums[8] float64, and the gradients of that loss
    w.r.t. xhat_q, mean_q, logvar_q, xhat_p, mean_p, logvar_p -- empty tensors when not computed).
    VAE.py:403-467 / 749-817 / 933-964 / 1171-1208."""
    reg = mask_p is not None
    D = x.shape[1]
    eng = engine(L.FAMILY_MLP, D, 0, x.device)          # the loss kernel does not depend on the family
    masks = [mask, mask_p] if reg else [mask]
    xh = [xhat_q, xhat_p] if reg else [xhat_q]
    mu = [mean_q, mean_p] if reg else [mean_q]
    lv = [logvar_q, logvar_p] if reg else [logvar_q]
    sums, d_xh, d_mu, d_lv = eng.loss_terms(x, masks, xh, mu, lv, alpha if reg else 0.0, beta_w, loss_scale,
                                            want_grads, x_logvar=x_logvar)
    loss = (KR.loss_from_sums(sums, 1, alpha, beta_w, reg) * loss_scale).to(torch.float32)
    e = lambda: x.new_empty(0, dtype=torch.float32)
    if not want_grads:
        return loss, sums, e(), e(), e(), e(), e(), e()
    if not reg:
        return loss, sums, d_xh[0], d_mu[0], d_lv[0], e(), e(), e()
    return loss, sums, d_xh[0], d_mu[0], d_lv[0], d_xh[1], d_mu[1], d_lv[1]


@vae_loss_op.register_fake
def _(x, mask, mask_p, xhat_q, xhat_p, mean_q, logvar_q, mean_p, logvar_p, alpha, beta_w, loss_scale, x_logvar,
      want_grads):
    e = lambda: x.new_empty(0, dtype=torch.float32)
    g = [e() for _ in range(6)]
    if want_grads:
        g[:3] = [torch.empty_like(xhat_q), torch.empty_like(mean_q), torch.empty_like(logvar_q)]
        if mask_p is not None:
            g[3:] = [torch.empty_like(xhat_p), torch.empty_like(mean_p), torch.empty_like(logvar_p)]
    return (x.new_empty((), dtype=torch.float32), x.new_empty(L.NSUMS, dtype=torch.float64), *g)


def _loss_setup(ctx, inputs, output):
    ctx.reg = inputs[2] is not None
    ctx.has_grads = inputs[13]
    ctx.save_for_backward(*output[2:])


def _loss_backward(ctx, d_loss, *unused):
    if not ctx.has_grads:
        raise RuntimeError("pcvae::vae_loss was run with want_grads=False")
    g = ctx.saved_tensors
    sc = lambda t: t * d_loss
    out = [None] * 14
    out[3], out[5], out[6] = sc(g[0]), sc(g[1]), sc(g[2])
    if ctx.reg:
        out[4], out[7], out[8] = sc(g[3]), sc(g[4]), sc(g[5])
    return tuple(out)


torch.library.register_autograd("pcvae::vae_loss", _loss_backward, setup_context=_loss_setup)


# ------------------------------------------------------------------------------------------
# reward
# ------------------------------------------------------------------------------------------

@torch.library.custom_op("pcvae::reward_chain", mutates_args=())
def reward_chain_op(theta: Tensor, x: Tensor, mask: Tensor, im: Tensor, family: int, D: int, K: int) -> Tensor:
    """R[N, D-1] of one acquisition step, evaluate.py:416-425 + 514-634."""
    eng = engine(family, D, K, x.device)
    R, _ = eng.reward(theta, x, mask, im)
    return R


@reward_chain_op.register_fake
def _(theta, x, mask, im, family, D, K):
    return x.new_empty(x.shape[0], D - 1, dtype=torch.float32)


# ------------------------------------------------------------------------------------------
# generic dense layer + not-MIWAE MNAR pieces (REG_notMIWAE_v2 / notMIWAE_myversion)
# ------------------------------------------------------------------------------------------

@torch.library.custom_op("pcvae::dense", mutates_args=())
def dense_op(x: Tensor, W: Tensor, b: Tensor, mask: Optional[Tensor], act: int) -> Tensor:
    """y = act((x*mask) W^T + b) -- nn.Linear + activation of VAE.py:2342-2363."""
    return KR.dense_fwd(x, W, b, act, mask)


@dense_op.register_fake
def _(x, W, b, mask, act):
    return x.new_empty(x.shape[0], W.shape[0], dtype=torch.float32)


@torch.library.custom_op("pcvae::dense_bwd", mutates_args=())
def dense_bwd_op(x: Tensor, W: Tensor, y: Tensor, dy: Tensor, mask: Optional[Tensor], act: int,
                 need_dx: bool) -> Tuple[Tensor, Tensor, Tensor]:
    dx, dW, db = KR.dense_bwd(x, W, y, dy, act, mask, need_dx)
    if dx is None:
        dx = x.new_empty(0, dtype=torch.float32)
    return dx, dW, db


@dense_bwd_op.register_fake
def _(x, W, y, dy, mask, act, need_dx):
    return (torch.empty_like(x) if need_dx else x.new_empty(0)), torch.empty_like(W), W.new_empty(W.shape[0])


def _dense_setup(ctx, inputs, output):
    x, W, b, mask, act = inputs
    ctx.save_for_backward(x, W, output, mask)
    ctx.act = act
    ctx.need_dx = x.requires_grad


def _dense_backward(ctx, dy):
    x, W, y, mask = ctx.saved_tensors
    dx, dW, db = dense_bwd_op(x, W, y, dy.contiguous(), mask, ctx.act, ctx.need_dx)
    return (dx if ctx.need_dx else None), dW, db, None, None


torch.library.register_autograd("pcvae::dense", _dense_backward, setup_context=_dense_setup)


@torch.library.custom_op("pcvae::mnar_sample_z", mutates_args=())
def mnar_sample_z_op(mean: Tensor, logvar: Tensor, eps: Optional[Tensor], samples: int) -> Tensor:
    """z [B,S,L] from per-row (mean, logvar) and eps [B,S,L]; VAE.py:2382-2387, 2753-2760."""
    return KR.mnar_sample_z(mean, logvar, eps, samples)


@mnar_sample_z_op.register_fake
def _(mean, logvar, eps, samples):
    return mean.new_empty(mean.shape[0], samples, mean.shape[1])


@torch.library.custom_op("pcvae::mnar_sample_z_bwd", mutates_args=())
def mnar_sample_z_bwd_op(d_z: Tensor, logvar: Tensor, eps: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    return KR.mnar_sample_z_bwd(d_z, logvar, eps)


@mnar_sample_z_bwd_op.register_fake
def _(d_z, logvar, eps):
    return torch.empty_like(logvar), torch.empty_like(logvar)


def _sz_setup(ctx, inputs, output):
    mean, logvar, eps, samples = inputs
    ctx.save_for_backward(logvar, eps)


def _sz_backward(ctx, d_z):
    logvar, eps = ctx.saved_tensors
    dm, dv = mnar_sample_z_bwd_op(d_z.contiguous(), logvar, eps)
    return dm, dv, None, None


torch.library.register_autograd("pcvae::mnar_sample_z", _sz_backward, setup_context=_sz_setup)


@torch.library.custom_op("pcvae::mnar_loss", mutates_args=())
def mnar_loss_op(x: Tensor, mask: Tensor, mask_p: Optional[Tensor], xm_q: Tensor, xlv_q: Tensor,
                 xm_p: Optional[Tensor], xlv_p: Optional[Tensor], mean_q: Tensor, logvar_q: Tensor,
                 mean_p: Optional[Tensor], logvar_p: Optional[Tensor], W: Tensor, b: Tensor,
                 eps_kl: Optional[Tensor], alpha: float, want_grads: bool, want_imputed: bool
                 ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor,
                            Tensor, Tensor]:
    """(loss fp32, out[4] f64 = loss/RE_q.mean()/loss_q/loss_p, xm_imputed, then gradients of the loss w.r.t.
    xm_q, xlv_q, mean_q, logvar_q, xm_p, xlv_p, mean_p, logvar_p, W, b).  VAE.py:2398-2471 / 2772-2823."""
    reg = mask_p is not None
    r = KR.mnar_loss(x, mask, mask_p, [xm_q, xm_p] if reg else [xm_q], [xlv_q, xlv_p] if reg else [xlv_q],
                     [mean_q, mean_p] if reg else [mean_q], [logvar_q, logvar_p] if reg else [logvar_q], W, b, alpha,
                     reg, eps_kl=eps_kl, want_grads=want_grads, want_imputed=want_imputed)
    e = lambda: x.new_empty(0, dtype=torch.float32)
    o = lambda t: e() if t is None else t
    loss = r["out"][0].to(torch.float32)
    g = [e() for _ in range(10)]
    if want_grads:
        g = [r["d_xm"][0], r["d_xlv"][0], r["d_mean"][0], r["d_logvar"][0]]
        g += [r["d_xm"][1], r["d_xlv"][1], r["d_mean"][1], r["d_logvar"][1]] if reg else [e(), e(), e(), e()]
        g += [r["d_W"].reshape(W.shape), r["d_b"].reshape(b.shape)]
    return (loss, r["out"], o(r["xm_imputed"]), *g)


@mnar_loss_op.register_fake
def _(x, mask, mask_p, xm_q, xlv_q, xm_p, xlv_p, mean_q, logvar_q, mean_p, logvar_p, W, b, eps_kl, alpha, want_grads,
      want_imputed):
    e = lambda: x.new_empty(0, dtype=torch.float32)
    return (x.new_empty((), dtype=torch.float32), x.new_empty(4, dtype=torch.float64),
            torch.empty_like(x) if want_imputed else e(), *[e() for _ in range(10)])


def _mnar_setup(ctx, inputs, output):
    ctx.reg = inputs[2] is not None
    ctx.has_grads = inputs[15]
    ctx.save_for_backward(*output[3:])
    # 12 of the 13 outputs never take part in a loss: without this autograd fills a zero tensor for each of them
    # (among them four [B, S, D] ones) on every backward
    ctx.set_materialize_grads(False)


def _mnar_backward(ctx, d_loss, *unused):
    if not ctx.has_grads:
        raise RuntimeError("pcvae::mnar_loss was run with want_grads=False")
    out = [None] * 17
    if d_loss is None:
        return tuple(out)
    g = ctx.saved_tensors
    n = 8 if ctx.reg else 4
    live = list(g[:n]) + [g[8], g[9]]
    scaled = torch._foreach_mul(live, d_loss)           # one multi-tensor launch instead of ten element-wise ones
    out[3], out[4], out[7], out[8] = scaled[0], scaled[1], scaled[2], scaled[3]
    if ctx.reg:
        out[5], out[6], out[9], out[10] = scaled[4], scaled[5], scaled[6], scaled[7]
    out[11], out[12] = scaled[n], scaled[n + 1]
    return tuple(out)


torch.library.register_autograd("pcvae::mnar_loss", _mnar_backward, setup_context=_mnar_setup)


# ------------------------------------------------------------------------------------------
# MIWAE / Reg_MIWAE pieces (reference VAE.py:3011-3301)
# ------------------------------------------------------------------------------------------

@torch.library.custom_op("pcvae::miwae_enc_heads", mutates_args=())
def miwae_enc_heads_op(raw: Tensor) -> Tuple[Tensor, Tensor]:
    """(mean, scale = softplus) from the encoder's raw [B, 2L] output, VAE.py:3047-3049."""
    o = KR.miwae_heads(raw, L.MIWAE_HEADS_ENC)
    return o[0], o[1]


@miwae_enc_heads_op.register_fake
def _(raw):
    W = raw.shape[1] // 2
    return raw.new_empty(raw.shape[0], W), raw.new_empty(raw.shape[0], W)


@torch.library.custom_op("pcvae::miwae_dec_heads", mutates_args=())
def miwae_dec_heads_op(raw: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """(sigmoid, softplus + 0.001, softplus + 3) from the decoder's raw [R, 3D] output, VAE.py:3061-3066."""
    o = KR.miwae_heads(raw, L.MIWAE_HEADS_DEC)
    return o[0], o[1], o[2]


@miwae_dec_heads_op.register_fake
def _(raw):
    W = raw.shape[1] // 3
    return tuple(raw.new_empty(raw.shape[0], W) for _ in range(3))


@torch.library.custom_op("pcvae::miwae_heads_bwd", mutates_args=())
def miwae_heads_bwd_op(raw: Tensor, mode: int, d0: Optional[Tensor], d1: Optional[Tensor], d2: Optional[Tensor]) -> Tensor:
    return KR.miwae_heads_bwd(raw, mode, [d0, d1, d2])


@miwae_heads_bwd_op.register_fake
def _(raw, mode, d0, d1, d2):
    return torch.empty_like(raw)


def _heads_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0])


def _c(t):
    return None if t is None else t.contiguous()


def _enc_heads_backward(ctx, d0, d1):
    (raw,) = ctx.saved_tensors
    return miwae_heads_bwd_op(raw, L.MIWAE_HEADS_ENC, _c(d0), _c(d1), None)


def _dec_heads_backward(ctx, d0, d1, d2):
    (raw,) = ctx.saved_tensors
    return miwae_heads_bwd_op(raw, L.MIWAE_HEADS_DEC, _c(d0), _c(d1), _c(d2))


torch.library.register_autograd("pcvae::miwae_enc_heads", _enc_heads_backward, setup_context=_heads_setup)
torch.library.register_autograd("pcvae::miwae_dec_heads", _dec_heads_backward, setup_context=_heads_setup)


@torch.library.custom_op("pcvae::miwae_sample_z", mutates_args=())
def miwae_sample_z_op(mean: Tensor, scale: Tensor, eps: Optional[Tensor], samples: int) -> Tensor:
    """z [B,S,L] = mean + scale * eps (Normal(mean, scale).rsample()), VAE.py:3050-3058."""
    return KR.miwae_sample_z(mean, scale, eps, samples)


@miwae_sample_z_op.register_fake
def _(mean, scale, eps, samples):
    return mean.new_empty(mean.shape[0], samples, mean.shape[1])


@torch.library.custom_op("pcvae::miwae_sample_z_bwd", mutates_args=())
def miwae_sample_z_bwd_op(d_z: Tensor, eps: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    return KR.miwae_sample_z_bwd(d_z, eps)


@miwae_sample_z_bwd_op.register_fake
def _(d_z, eps):
    return d_z.new_empty(d_z.shape[0], d_z.shape[2]), d_z.new_empty(d_z.shape[0], d_z.shape[2])


def _msz_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[2])
    ctx.has_eps = inputs[2] is not None


def _msz_backward(ctx, d_z):
    (eps,) = ctx.saved_tensors if ctx.has_eps else (None,)
    dm, ds = miwae_sample_z_bwd_op(d_z.contiguous(), eps)
    return dm, (ds if ctx.has_eps else None), None, None


torch.library.register_autograd("pcvae::miwae_sample_z", _msz_backward, setup_context=_msz_setup)


@torch.library.custom_op("pcvae::miwae_loss", mutates_args=())
def miwae_loss_op(x: Tensor, mask: Tensor, mask_p: Optional[Tensor], xm_q: Tensor, xs_q: Tensor, df_q: Tensor,
                  mean_q: Tensor, scale_q: Tensor, eps2_q: Tensor, xm_p: Optional[Tensor], xs_p: Optional[Tensor],
                  df_p: Optional[Tensor], mean_p: Optional[Tensor], scale_p: Optional[Tensor], eps2_p: Optional[Tensor],
                  alpha: float, rowwise: bool, want_grads: bool, want_imputed: bool
                  ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor,
                             Tensor]:
    """(loss fp32, out[6] f64, xm_imputed, then the gradients of the loss w.r.t. xm_q, xs_q, df_q, mean_q, scale_q, xm_p,
    xs_p, df_p, mean_p, scale_p).  MIWAE.loss / Reg_MIWAE.loss, VAE.py:3068-3110 / 3197-3263."""
    reg = mask_p is not None
    br = lambda q, p_: [q, p_] if reg else [q]
    r = KR.miwae_loss(x, mask, mask_p, br(xm_q, xm_p), br(xs_q, xs_p), br(df_q, df_p), br(mean_q, mean_p),
                      br(scale_q, scale_p), br(eps2_q, eps2_p), alpha, reg, rowwise=rowwise, want_grads=want_grads,
                      want_imputed=want_imputed)
    e = lambda: x.new_empty(0, dtype=torch.float32)
    loss = r["out"][0].to(torch.float32)
    g = [e() for _ in range(10)]
    if want_grads:
        g[:5] = [r["d_xm"][0], r["d_xs"][0], r["d_df"][0], r["d_mean"][0], r["d_scale"][0]]
        if reg:
            g[5:] = [r["d_xm"][1], r["d_xs"][1], r["d_df"][1], r["d_mean"][1], r["d_scale"][1]]
    return (loss, r["out"], e() if r["xm_imputed"] is None else r["xm_imputed"], *g)


@miwae_loss_op.register_fake
def _(x, mask, mask_p, xm_q, xs_q, df_q, mean_q, scale_q, eps2_q, xm_p, xs_p, df_p, mean_p, scale_p, eps2_p, alpha, rowwise,
      want_grads, want_imputed):
    e = lambda: x.new_empty(0, dtype=torch.float32)
    return (x.new_empty((), dtype=torch.float32), x.new_empty(6, dtype=torch.float64),
            torch.empty_like(x) if want_imputed else e(), *[e() for _ in range(10)])


def _miwae_setup(ctx, inputs, output):
    ctx.reg = inputs[2] is not None
    ctx.has_grads = inputs[17]
    ctx.save_for_backward(*output[3:])


def _miwae_backward(ctx, d_loss, *unused):
    if not ctx.has_grads:
        raise RuntimeError("pcvae::miwae_loss was run with want_grads=False")
    g = ctx.saved_tensors
    sc = lambda t: t * d_loss
    out = [None] * 19
    out[3], out[4], out[5], out[6], out[7] = sc(g[0]), sc(g[1]), sc(g[2]), sc(g[3]), sc(g[4])
    if ctx.reg:
        out[9], out[10], out[11], out[12], out[13] = sc(g[5]), sc(g[6]), sc(g[7]), sc(g[8]), sc(g[9])
    return tuple(out)


torch.library.register_autograd("pcvae::miwae_loss", _miwae_backward, setup_context=_miwae_setup)

"""CPU oracle for the partial-VAE posterior-consistency hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`vae_posterior_consistency_b200/`) imports this file; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may use it, and there only as the checker / the timed CPU baseline.

It is a closed-form restatement (torch CPU tensors, dtype-generic so the same
code runs in fp32 and fp64) of the reference's arithmetic.  All citations are
relative to the reference repo root (`src/...`).

Parity status: the reference ships no tests, golden vectors or seeds
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, generated in the build container by `tests/golden/make_golden.py`
(imports the unmodified reference, records its inputs, noise draws and outputs)
and committed as `tests/golden/*.pt`.  `tests/test_oracle_golden.py` checks the
oracle against every one of those fixtures.

Parameter containers are plain dicts keyed by the reference's state_dict names
(SURVEY.md A.1), e.g. `seq_encoder.0.weight`.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]

HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)
#: fixed decoder log-variance log((0.1*sqrt(2))^2)=log 0.02, src/models/VAE.py:379
X_LOGVAR = math.log((0.1 * math.sqrt(2.0)) ** 2)
LATENT = 10
BETA_MAX_EPOCH = 2800  # src/models/VAE.py:384


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------

def _lin(h: Tensor, w: Tensor, b: Tensor) -> Tensor:
    return h @ w.t() + b


def _as(m: Tensor, like: Tensor) -> Tensor:
    return m.to(like.dtype)


def is_pnp(p: Params) -> bool:
    return "type_pars1" in p


def init_params(family: str, obs_dim: int, K: int = 20, latent: int = LATENT,
                seed: int = 0, dtype=torch.float32) -> Params:
    """Random parameters with the reference's shapes and init distributions
    (nn.Linear default init; Xavier-uniform for the PNP embeddings,
    src/models/VAE.py:366-376, 687-709).  Used for synthetic benchmarks and
    tests; the values are not meant to match any particular torch seed."""
    g = torch.Generator().manual_seed(seed)

    def linear(out_f, in_f):
        bound = 1.0 / math.sqrt(in_f)
        w = (torch.rand(out_f, in_f, generator=g, dtype=torch.float64) * 2 - 1) * bound
        b = (torch.rand(out_f, generator=g, dtype=torch.float64) * 2 - 1) * bound
        return w.to(dtype), b.to(dtype)

    def xavier(r, c):
        bound = math.sqrt(6.0 / (r + c))
        return ((torch.rand(r, c, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)

    p: Params = {}
    if family in ("pnp", "eddi"):
        p["type_pars1"] = xavier(obs_dim, K)
        p["type_bias1"] = xavier(obs_dim, 1)
        p["pnp_encoder1.0.weight"], p["pnp_encoder1.0.bias"] = linear(K, K + 2)
        enc, first_in = "pnp_encoder2", K
    elif family in ("mlp", "vae"):
        enc, first_in = "seq_encoder", obs_dim
    elif family in ("mlp_mask", "vae_mask"):
        # Reg_VAE_mask / vanilla_VAE_mask: first layer reads [x*mask, mask], src/models/VAE.py:526-527, 1011-1012
        enc, first_in = "seq_encoder", 2 * obs_dim
    else:
        raise ValueError(family)
    p["prior_mean"] = torch.zeros(latent, dtype=dtype)
    p["prior_std"] = torch.ones(latent, dtype=dtype)
    for idx, (o, i) in zip((0, 2, 4), ((100, first_in), (50, 100), (2 * latent, 50))):
        p[f"{enc}.{idx}.weight"], p[f"{enc}.{idx}.bias"] = linear(o, i)
    for idx, (o, i) in zip((0, 2, 4), ((50, latent), (100, 50), (obs_dim, 100))):
        p[f"seq_decoder.{idx}.weight"], p[f"seq_decoder.{idx}.bias"] = linear(o, i)
    return p


# --------------------------------------------------------------------------
# encoders / decoder
# --------------------------------------------------------------------------

def mlp_first_layer_pre(p: Params, x: Tensor, mask: Tensor) -> Tensor:
    """W1 (x*mask) + b1, the pre-activation the reward's incremental form builds on.  The mask-augmented
    family (first-layer fan-in 2D) feeds `stack([x*mask, mask], 1).reshape(-1, 2D)` = [x*mask | mask],
    src/models/VAE.py:547, 1032."""
    m = _as(mask, x)
    h = x * m
    if p["seq_encoder.0.weight"].shape[1] == 2 * x.shape[1]:
        h = torch.cat([h, m.expand_as(h)], 1)
    return _lin(h, p["seq_encoder.0.weight"], p["seq_encoder.0.bias"])


def mlp_tail(p: Params, h_pre: Tensor, prefix: str = "seq_encoder") -> Tuple[Tensor, Tensor]:
    h1 = torch.relu(h_pre)
    h2 = torch.relu(_lin(h1, p[f"{prefix}.2.weight"], p[f"{prefix}.2.bias"]))
    o = _lin(h2, p[f"{prefix}.4.weight"], p[f"{prefix}.4.bias"])
    L = o.shape[1] // 2
    return o[:, :L], o[:, L:]


def mlp_encoder_stats(p: Params, x: Tensor, mask: Tensor) -> Tuple[Tensor, Tensor]:
    """Zero-imputation encoder, src/models/VAE.py:387-388 (nets :366-372)."""
    return mlp_tail(p, mlp_first_layer_pre(p, x, mask))


def pnp_embed_as_written(p: Params, x: Tensor) -> Tensor:
    """Per-feature embedding ReLU(W_e [x_d, x_d E_d, b_d] + b_e), as written at
    src/models/VAE.py:726-733.  Returns [B, D, K]."""
    B, D = x.shape
    E, bE = p["type_pars1"], p["type_bias1"]
    We, be = p["pnp_encoder1.0.weight"], p["pnp_encoder1.0.bias"]
    xf = x.reshape(B, D, 1)
    feat = torch.cat([xf, xf * E.unsqueeze(0), bE.unsqueeze(0).expand(B, D, 1)], dim=2)
    return torch.relu(feat @ We.t() + be)


def pnp_collapse(p: Params) -> Tuple[Tensor, Tensor]:
    """A, C of SURVEY.md A.3: W_e [x, x E_d, b_d] + b_e == x*A_d + C_d."""
    E, bE = p["type_pars1"], p["type_bias1"]
    We, be = p["pnp_encoder1.0.weight"], p["pnp_encoder1.0.bias"]
    K = E.shape[1]
    A = We[:, 0].unsqueeze(0) + E @ We[:, 1:K + 1].t()
    C = bE * We[:, K + 1].unsqueeze(0) + be.unsqueeze(0)
    return A, C


def pnp_aggregate(p: Params, x: Tensor, mask: Tensor, collapsed: bool = False) -> Tensor:
    """agg_b = sum_d mask_bd * e_bd, src/models/VAE.py:731-733."""
    if collapsed:
        A, C = pnp_collapse(p)
        emb = torch.relu(x.unsqueeze(2) * A.unsqueeze(0) + C.unsqueeze(0))
    else:
        emb = pnp_embed_as_written(p, x)
    return (_as(mask, x).unsqueeze(2) * emb).sum(1)


def pnp_tail(p: Params, agg: Tensor) -> Tuple[Tensor, Tensor]:
    h1_pre = _lin(agg, p["pnp_encoder2.0.weight"], p["pnp_encoder2.0.bias"])
    return mlp_tail(p, h1_pre, prefix="pnp_encoder2")


def pnp_encoder_stats(p: Params, x: Tensor, mask: Tensor, collapsed: bool = False):
    """PNP/EDDI set encoder, src/models/VAE.py:719-734."""
    return pnp_tail(p, pnp_aggregate(p, x, mask, collapsed))


def encoder_stats(p: Params, x: Tensor, mask: Tensor, collapsed: bool = False):
    if is_pnp(p):
        return pnp_encoder_stats(p, x, mask, collapsed)
    return mlp_encoder_stats(p, x, mask)


def reparam(mu: Tensor, logvar: Tensor, eps: Tensor) -> Tensor:
    """Normal(mu, exp(logvar/2)).rsample() == mu + eps*std, src/models/VAE.py:390-392."""
    return mu + eps * torch.exp(logvar / 2)


def decoder(p: Params, z: Tensor) -> Tensor:
    """src/models/VAE.py:397-401 (nets :374-376)."""
    g1 = torch.relu(_lin(z, p["seq_decoder.0.weight"], p["seq_decoder.0.bias"]))
    g2 = torch.relu(_lin(g1, p["seq_decoder.2.weight"], p["seq_decoder.2.bias"]))
    return torch.sigmoid(_lin(g2, p["seq_decoder.4.weight"], p["seq_decoder.4.bias"]))


# --------------------------------------------------------------------------
# loss terms
# --------------------------------------------------------------------------

def masked_nll(x: Tensor, xhat: Tensor, m: Tensor, x_logvar: float = X_LOGVAR) -> Tensor:
    """sum(-Normal(xhat*m, exp(lv*m/2)).log_prob(x*m)), src/models/VAE.py:422-423,488-490.
    Masked-out entries still contribute 0.5*log(2*pi) each."""
    mf = _as(m, x)
    lv = x_logvar * mf
    scale = torch.exp(lv / 2)
    diff = x * mf - xhat * mf
    return torch.sum(diff * diff / (2 * scale * scale) + torch.log(scale) + HALF_LOG_2PI)


def kl_std_normal(mu: Tensor, logvar: Tensor) -> Tensor:
    """sum KL(N(mu, e^{lv/2}) || N(0,1)), src/models/VAE.py:476-478."""
    return 0.5 * torch.sum(torch.exp(logvar) + mu * mu - 1.0 - logvar)


def kl_normal_normal(mu_q, lv_q, mu_p, lv_p) -> Tensor:
    """sum KL(N_q || N_p), src/models/VAE.py:469-474."""
    var_ratio = torch.exp(lv_q - lv_p)
    t1 = (mu_q - mu_p) ** 2 / torch.exp(lv_p)
    return 0.5 * torch.sum(var_ratio + t1 - 1.0 - (lv_q - lv_p))


def reg_loss_terms(x, xh_p, mu_p, lv_p, xh_q, mu_q, lv_q, mask, mask_p):
    mb, mpb = mask.bool(), mask_p.bool()
    return dict(
        RE_q=masked_nll(x, xh_q, mb), RE_p=masked_nll(x, xh_p, mpb),
        KL_q=kl_std_normal(mu_q, lv_q), KL_p=kl_std_normal(mu_p, lv_p),
        KL_reg=kl_normal_normal(mu_q, lv_q, mu_p, lv_p),
        RE_d=masked_nll(x, xh_q, mb & ~mpb), RE_imp=masked_nll(x, xh_q, ~mb))


def reg_loss(x, xh_p, mu_p, lv_p, xh_q, mu_q, lv_q, mask, mask_p, epoch=1,
             beta=1.0, alpha=1.0, beta_annealing=False, stage="train"):
    """Reg_VAE.loss / Reg_EDDI.loss with reg_type='kl_reg', src/models/VAE.py:403-467, 749-817.
    Returns (train_loss, RE_q/B, RE_q_imputed/B)."""
    B = x.shape[0]
    t = reg_loss_terms(x, xh_p, mu_p, lv_p, xh_q, mu_q, lv_q, mask, mask_p)
    bw = beta * (epoch / BETA_MAX_EPOCH) if beta_annealing else beta
    loss_q = t["RE_q"] + bw * t["KL_q"]
    if stage == "evaluate":
        return loss_q / B, t["RE_q"] / B, t["RE_imp"] / B
    loss_p = t["RE_p"] + bw * t["KL_p"]
    loss = loss_q + alpha * (t["KL_reg"] - loss_q + loss_p + t["RE_d"])
    return loss / B, t["RE_q"] / B, torch.zeros((), dtype=x.dtype)


def vanilla_loss(x, xh_q, mu_q, lv_q, mask, epoch=1, beta=1.0, beta_annealing=False):
    """vanilla_VAE.loss / vanilla_EDDI.loss, src/models/VAE.py:1171-1208, 933-964.
    Returns (train_loss, RE_q/B, RE_q_imputed/B)."""
    B = x.shape[0]
    mb = mask.bool()
    RE_q = masked_nll(x, xh_q, mb)
    RE_imp = masked_nll(x, xh_q, ~mb)
    bw = beta * (epoch / BETA_MAX_EPOCH) if beta_annealing else beta
    return (RE_q + bw * kl_std_normal(mu_q, lv_q)) / B, RE_q / B, RE_imp / B


# --------------------------------------------------------------------------
# whole training step (forward + loss + backward) and Adam
# --------------------------------------------------------------------------

def trainable_names(p: Params) -> Sequence[str]:
    return [k for k in p if not k.startswith("prior_")]


def train_step(p: Params, x, mask, mask_p, eps_q, eps_p, alpha=1.0, beta=1.0,
               regularised=True, collapsed=False):
    """One reference step: model.forward -> model.loss -> backward
    (src/experiment_main/train.py:87-116).  Returns (train_loss, grads dict, aux)."""
    names = trainable_names(p)
    q = {k: (v.detach().clone().requires_grad_(True) if k in names else v) for k, v in p.items()}
    mu_q, lv_q = encoder_stats(q, x, mask, collapsed)
    xh_q = decoder(q, reparam(mu_q, lv_q, eps_q))
    if regularised:
        mu_p, lv_p = encoder_stats(q, x, mask_p, collapsed)
        xh_p = decoder(q, reparam(mu_p, lv_p, eps_p))
        loss, _, _ = reg_loss(x, xh_p, mu_p, lv_p, xh_q, mu_q, lv_q, mask, mask_p,
                              beta=beta, alpha=alpha)
        aux = dict(mu_q=mu_q, lv_q=lv_q, xh_q=xh_q, mu_p=mu_p, lv_p=lv_p, xh_p=xh_p)
    else:
        loss, _, _ = vanilla_loss(x, xh_q, mu_q, lv_q, mask, beta=beta)
        aux = dict(mu_q=mu_q, lv_q=lv_q, xh_q=xh_q)
    grads = torch.autograd.grad(loss, [q[k] for k in names], allow_unused=True)
    gd = {k: (g if g is not None else torch.zeros_like(q[k])) for k, g in zip(names, grads)}
    return loss.detach(), gd, {k: v.detach() for k, v in aux.items()}


def adam_step(param, grad, m, v, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults (src/experiment_main/train.py:21), single-tensor form."""
    m = b1 * m + (1 - b1) * grad
    v = b2 * v + (1 - b2) * grad * grad
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return param - (lr / bc1) * (m / denom), m, v


# --------------------------------------------------------------------------
# evaluation metrics
# --------------------------------------------------------------------------

def rmse_unobserved(xh_q, x, mask):
    """src/experiment_main/evaluate.py:232-234."""
    nm = ~mask.bool()
    return torch.sqrt(torch.sum((xh_q * nm - x * nm) ** 2) / torch.sum(nm))


# --------------------------------------------------------------------------
# active-selection reward
# --------------------------------------------------------------------------

def _reward_kl(mu_i, lv_i, mu, lv):
    """0.5*sum((mu_i-mu)^2/std + var_i/var - 1 - lv_i + lv), src/experiment_main/evaluate.py:582-583,631-632.
    NOTE divides by std, not variance."""
    return 0.5 * torch.sum((mu_i - mu) ** 2 / torch.exp(lv / 2)
                           + torch.exp(lv_i) / torch.exp(lv) - 1.0 - lv_i + lv, 1)


def reward_chain_as_written(p: Params, i: int, x, mask, im, loc) -> Tensor:
    """R_lindley_chain with chaini_I / chaini_II exactly as the reference loops
    (src/experiment_main/evaluate.py:514-634): 4 full encoder calls per sample."""
    M = im.shape[0]
    temp_x = x.clone()
    acc = torch.zeros(len(loc), dtype=x.dtype)
    mloc = mask[loc].to(x.dtype)
    for m in range(M):
        temp_x[loc, i] = im[m, loc, i]
        xs = temp_x[loc]
        m0 = mloc.clone()
        mu, lv = encoder_stats(p, xs, m0)
        m0[:, i] = 1
        mu_i, lv_i = encoder_stats(p, xs, m0)
        acc = acc + _reward_kl(mu_i, lv_i, mu, lv)
        temp_x[loc, -1] = im[m, loc, -1]
        xs = temp_x[loc]
        m1 = mloc.clone()
        m1[:, -1] = 1
        mu, lv = encoder_stats(p, xs, m1)
        m1[:, i] = 1
        mu_i, lv_i = encoder_stats(p, xs, m1)
        acc = acc - _reward_kl(mu_i, lv_i, mu, lv)
    return acc / M


def reward_all(p: Params, x, mask, im, incremental: bool = True) -> Tensor:
    """Reward matrix R[N, D-1] for one active-learning step: candidates with
    mask[n,u]==0 get the chain reward, the rest keep -1e4
    (src/experiment_main/evaluate.py:391,416-425).  `incremental=True` uses the
    rank-1 restatement of SURVEY.md A.5 (what the CUDA kernel implements)."""
    M, N, D = im.shape
    R = torch.full((N, D - 1), -1e4, dtype=x.dtype)
    maskf = mask.to(x.dtype)
    if not incremental:
        for u in range(D - 1):
            loc = torch.nonzero(maskf[:, u] == 0).flatten()
            if len(loc):
                R[loc, u] = reward_chain_as_written(p, u, x, maskf, im, loc)
        return R
    pnp = is_pnp(p)
    if pnp:
        A, C = pnp_collapse(p)
        base_in = pnp_aggregate(p, x, maskf, collapsed=True)       # agg0 [N,K]
        tail = lambda h: pnp_tail(p, h)
        delta = lambda v, col: torch.relu(v.unsqueeze(1) * A[col].unsqueeze(0) + C[col].unsqueeze(0))
    else:
        W1 = p["seq_encoder.0.weight"]
        base_in = mlp_first_layer_pre(p, x, maskf)                  # h0 [N,100]
        tail = lambda h: mlp_tail(p, h)
        delta = lambda v, col: v.unsqueeze(1) * W1[:, col].unsqueeze(0)
    mu0, lv0 = tail(base_in)
    acc = torch.zeros(N, D - 1, dtype=x.dtype)
    for m in range(M):
        hT = base_in + delta(im[m, :, D - 1], D - 1)
        muT, lvT = tail(hT)
        for u in range(D - 1):
            du = delta(im[m, :, u], u)
            mu_a, lv_a = tail(base_in + du)
            mu_b, lv_b = tail(hT + du)
            acc[:, u] = acc[:, u] + _reward_kl(mu_a, lv_a, mu0, lv0)
            acc[:, u] = acc[:, u] - _reward_kl(mu_b, lv_b, muT, lvT)
    sel = maskf[:, :D - 1] == 0
    R[sel] = (acc / M)[sel]
    return R


# --------------------------------------------------------------------------
# not-MIWAE (MNAR self-masking) family: REG_notMIWAE_v2 / notMIWAE_myversion
# --------------------------------------------------------------------------

def elu(h: Tensor) -> Tensor:
    return torch.where(h > 0, h, torch.exp(h) - 1.0)


def mnar_encoder_stats(p: Params, x: Tensor, mask: Tensor) -> Tuple[Tensor, Tensor]:
    """seq_encoder (D->128->128, ELU) + q_mu / q_logstd heads, src/models/VAE.py:2377-2380, 2748-2751.
    Returns per-row (mean, log_var) [B, L] (the reference then repeats them over the S samples)."""
    h = elu(_lin(x * _as(mask, x), p["seq_encoder.0.weight"], p["seq_encoder.0.bias"]))
    h = elu(_lin(h, p["seq_encoder.2.weight"], p["seq_encoder.2.bias"]))
    return _lin(h, p["q_mu.0.weight"], p["q_mu.0.bias"]), _lin(h, p["q_logstd.0.weight"], p["q_logstd.0.bias"])


def mnar_decoder(p: Params, z: Tensor) -> Tuple[Tensor, Tensor]:
    """seq_decoder (L->128->128, ELU) + x_mean (Sigmoid) / x_logvar (Hardtanh[-10,0]) heads, VAE.py:2392-2396."""
    g = elu(_lin(z, p["seq_decoder.0.weight"], p["seq_decoder.0.bias"]))
    g = elu(_lin(g, p["seq_decoder.2.weight"], p["seq_decoder.2.bias"]))
    xm = torch.sigmoid(_lin(g, p["x_mean.0.weight"], p["x_mean.0.bias"]))
    xlv = torch.clamp(_lin(g, p["x_logvar.0.weight"], p["x_logvar.0.bias"]), -10.0, 0.0)
    return xm, xlv


def _nll_rows(x, xm, xlv, m):
    """per-(b,s) sum over features of -Normal(xm*m, exp(xlv*m/2)).log_prob(x*m), VAE.py:2491-2493."""
    scale = torch.exp(xlv * m / 2)
    diff = x * m - xm * m
    return torch.sum(diff * diff / (2 * scale * scale) + torch.log(scale) + HALF_LOG_2PI, 2)


def _selfmask_logp(p: Params, x, xm, m):
    """sum_d Bernoulli(logits=-softplus(W)(x~ - b)).log_prob(m), x~ = xm(1-m) + x m; VAE.py:2407,2417,2434-2435."""
    mixed = xm * (1 - m) + x * m
    logits = -torch.nn.functional.softplus(p["W"]) * (mixed - p["b"])
    return torch.sum(-torch.nn.functional.binary_cross_entropy_with_logits(logits, m.expand_as(logits),
                                                                           reduction="none"), 2)


def mnar_reg_loss(p: Params, x, mask, mask_p, mu_q, lv_q, mu_p, lv_p, xm_q, xlv_q, xm_p, xlv_p, alpha=1.0):
    """REG_notMIWAE_v2.loss, VAE.py:2398-2471.  mu/lv are [B, L]; xm/xlv are [B, S, D].
    Returns (loss, xm_imputed [B, D], RE_q.mean())."""
    S = xm_q.shape[1]
    X, M_, MP = x.unsqueeze(1), mask.to(x.dtype).unsqueeze(1), mask_p.to(x.dtype).unsqueeze(1)
    RE_q = _nll_rows(X, xm_q, xlv_q, M_)
    RE_p = _nll_rows(X, xm_p, xlv_p, MP)
    kl = lambda mu, lv: 0.5 * torch.sum(torch.exp(lv) + mu * mu - 1.0 - lv, 1, keepdim=True)
    l_w_q = RE_q + kl(mu_q, lv_q) - _selfmask_logp(p, X, xm_q, M_)
    l_w_p = RE_p + kl(mu_p, lv_p)
    loss_q = torch.mean(torch.logsumexp(l_w_q, 1) - math.log(float(S)))
    loss_p = torch.mean(torch.logsumexp(l_w_p, 1) - math.log(float(S)))
    kl_el = 0.5 * (torch.exp(lv_q - lv_p) + (mu_q - mu_p) ** 2 / torch.exp(lv_p) - 1.0 - (lv_q - lv_p))
    KL_reg = kl_el.mean()                       # mean over [B,S,L] == mean over [B,L] (copies over S)
    RE_d = _nll_rows(X, xm_q, xlv_q, M_ * (1 - MP)).mean()
    loss = loss_q + alpha * (KL_reg - loss_q + loss_p + RE_d)
    wl = torch.softmax(-l_w_q, 1)
    return loss, torch.sum(xm_q * wl.unsqueeze(2), 1), RE_q.mean()


def mnar_vanilla_loss(p: Params, x, mask, mu, lv, xm, xlv, eps_kl):
    """notMIWAE_myversion.loss, VAE.py:2772-2823: Monte-Carlo KL from a SECOND draw z' = mu + std*eps_kl
    (eps_kl [B, S, L]).  Returns (loss, xm_imputed, RE.mean())."""
    S = xm.shape[1]
    X, M_ = x.unsqueeze(1), mask.to(x.dtype).unsqueeze(1)
    RE = _nll_rows(X, xm, xlv, M_)
    std = torch.exp(lv / 2).unsqueeze(1)
    z2 = mu.unsqueeze(1) + eps_kl * std
    log_q = torch.sum(-(z2 - mu.unsqueeze(1)) ** 2 / (2 * std * std) - torch.log(std) - HALF_LOG_2PI, 2)
    log_p = torch.sum(-z2 * z2 / 2 - HALF_LOG_2PI, 2)
    l_w = RE + (log_q - log_p) - _selfmask_logp(p, X, xm, M_)
    loss = torch.mean(torch.logsumexp(l_w, 1) - math.log(float(S)))
    wl = torch.softmax(-l_w, 1)
    return loss, torch.sum(xm * wl.unsqueeze(2), 1), RE.mean()


def mnar_trainable_names(p: Params) -> Sequence[str]:
    return [k for k in p if not k.startswith("logits.")]


def mnar_train_step(p: Params, x, mask, mask_p, eps_q, eps_p, alpha=1.0, regularised=True, eps_kl=None):
    """forward + loss + backward of the MNAR families (train.py:87-101); eps_* are [B, S, L]."""
    names = mnar_trainable_names(p)
    q = {k: (v.detach().clone().requires_grad_(True) if k in names else v) for k, v in p.items()}
    mu_q, lv_q = mnar_encoder_stats(q, x, mask)
    z_q = mu_q.unsqueeze(1) + eps_q * torch.exp(lv_q / 2).unsqueeze(1)
    xm_q, xlv_q = mnar_decoder(q, z_q)
    if regularised:
        mu_p, lv_p = mnar_encoder_stats(q, x, mask_p)
        z_p = mu_p.unsqueeze(1) + eps_p * torch.exp(lv_p / 2).unsqueeze(1)
        xm_p, xlv_p = mnar_decoder(q, z_p)
        loss, xm_imp, re = mnar_reg_loss(q, x, mask, mask_p, mu_q, lv_q, mu_p, lv_p, xm_q, xlv_q, xm_p, xlv_p, alpha)
    else:
        loss, xm_imp, re = mnar_vanilla_loss(q, x, mask, mu_q, lv_q, xm_q, xlv_q, eps_kl)
    grads = torch.autograd.grad(loss, [q[k] for k in names], allow_unused=True)
    gd = {k: (g if g is not None else torch.zeros_like(q[k])) for k, g in zip(names, grads)}
    return loss.detach(), gd, dict(xm_imp=xm_imp.detach(), re=re.detach(), xm_q=xm_q.detach(), xlv_q=xlv_q.detach(),
                                   mu_q=mu_q.detach(), lv_q=lv_q.detach())


def init_mnar_params(obs_dim: int, latent: int = LATENT, seed: int = 0, with_logits: bool = True) -> Params:
    """Random parameters with the shapes of REG_notMIWAE_v2."""
    g = torch.Generator().manual_seed(seed)

    def linear(out_f, in_f):
        bound = 1.0 / math.sqrt(in_f)
        return ((torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound, (torch.rand(out_f, generator=g) * 2 - 1) * bound)

    p: Params = {}
    bound = math.sqrt(6.0 / (obs_dim + obs_dim))
    p["W"] = (torch.rand(1, 1, obs_dim, generator=g) * 2 - 1) * bound
    p["b"] = (torch.rand(1, 1, obs_dim, generator=g) * 2 - 1) * bound
    for name, (o, i) in (("seq_encoder.0", (128, obs_dim)), ("seq_encoder.2", (128, 128)), ("q_mu.0", (latent, 128)),
                         ("q_logstd.0", (latent, 128)), ("seq_decoder.0", (128, latent)), ("seq_decoder.2", (128, 128)),
                         ("x_mean.0", (obs_dim, 128)), ("x_logvar.0", (obs_dim, 128))):
        p[name + ".weight"], p[name + ".bias"] = linear(o, i)
    if with_logits:
        w, b = linear(obs_dim, obs_dim)
        p["logits.0.weight"], p["logits.0.bias"] = w.double(), b.double()
    return p


# --------------------------------------------------------------------------
# MIWAE (Student-t decoder + importance-weighted bound): MIWAE / Reg_MIWAE, src/models/VAE.py:3011-3301
# (SURVEY.md section 8f item 4).  Restated here so that the family's oracle is pinned to the reference before its
# kernels exist; the product package does not build this family yet.
# --------------------------------------------------------------------------

HALF_LOG_PI = 0.5 * math.log(math.pi)


def miwae_encoder_stats(p: Params, x: Tensor, mask: Tensor) -> Tuple[Tensor, Tensor]:
    """mean [B, L], scale = softplus(raw) [B, L]  (VAE.py:3047-3049, 3178-3180); nets :3026-3032."""
    h = torch.relu(_lin(x * _as(mask, x), p["seq_encoder.0.weight"], p["seq_encoder.0.bias"]))
    h = torch.relu(_lin(h, p["seq_encoder.2.weight"], p["seq_encoder.2.bias"]))
    o = _lin(h, p["seq_encoder.4.weight"], p["seq_encoder.4.bias"])
    L = o.shape[1] // 2
    return o[:, :L], torch.nn.functional.softplus(o[:, L:])


def miwae_decoder(p: Params, z: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """z [B, S, L] -> (mean = sigmoid, scale = softplus + 0.001, df = softplus + 3), each [B, S, D]  (VAE.py:3061-3066)."""
    h = torch.relu(_lin(z, p["seq_decoder.0.weight"], p["seq_decoder.0.bias"]))
    h = torch.relu(_lin(h, p["seq_decoder.2.weight"], p["seq_decoder.2.bias"]))
    o = _lin(h, p["seq_decoder.4.weight"], p["seq_decoder.4.bias"])
    D = o.shape[-1] // 3
    sp = torch.nn.functional.softplus
    return torch.sigmoid(o[..., :D]), sp(o[..., D:2 * D]) + 0.001, sp(o[..., 2 * D:]) + 3.0


def student_t_log_prob(x: Tensor, loc: Tensor, scale: Tensor, df: Tensor) -> Tensor:
    """torch.distributions.StudentT(df, loc, scale).log_prob(x)."""
    y = (x - loc) / scale
    Z = torch.log(scale) + 0.5 * torch.log(df) + HALF_LOG_PI + torch.lgamma(0.5 * df) - torch.lgamma(0.5 * (df + 1.0))
    return -0.5 * (df + 1.0) * torch.log1p(y * y / df) - Z


def _miwae_bound_terms(x, m, xm, xs, df, mean, scale, eps2):
    """Per-branch pieces of the bound as the reference writes them.  NOTE the reference's indexing quirk
    (VAE.py:3080-3081): the [B*S] vector of per-(row, sample) log-likelihoods, laid out row-major in (b, s), is
    reshaped to [S, B] WITHOUT a transpose, so entry (i, j) of the "[samples, rows]" matrix is the likelihood of
    (row, sample) = divmod(i * B + j, S); log p(z) - log q(z|x) IS transposed properly.  Kept for parity."""
    B, S, D = xm.shape
    logp = student_t_log_prob(x.unsqueeze(1), xm, xs, df)                   # [B, S, D]
    mf = _as(m, x).unsqueeze(1)
    lpx = (logp * mf).sum(2).reshape(S, B)                                  # the mis-indexed view
    z = mean.unsqueeze(1) + scale.unsqueeze(1) * eps2                       # fresh draw inside loss(), VAE.py:3086-3088
    logpz = (-0.5 * z * z - HALF_LOG_2PI).sum(2).t()                        # [S, B]
    sc = scale.unsqueeze(1)
    logq = (-((z - mean.unsqueeze(1)) ** 2) / (2 * sc * sc) - torch.log(sc) - HALF_LOG_2PI).sum(2).t()
    lw = lpx + logpz - logq
    return logp, lw


def miwae_loss(x, mask, xm, xs, df, mean, scale, eps2):
    """MIWAE.loss (VAE.py:3068-3110): returns (neg_bound, xm_imputed [B, D], imputed-likelihood scalar of llh_eval)."""
    B, S, D = xm.shape
    logp, lw = _miwae_bound_terms(x, mask, xm, xs, df, mean, scale, eps2)
    neg_bound = -torch.mean(torch.logsumexp(lw, 0))
    w = torch.softmax(lw, 0)                                                # [S, B]
    xm_imp = torch.einsum('ki,kij->ij', w, xm.permute(1, 0, 2))
    imp = (logp * (1.0 - _as(mask, x)).unsqueeze(1)).sum() / (B * 5000)
    return neg_bound, xm_imp, imp


def reg_miwae_loss(x, mask, mask_p, q, pbr, eps2_q, eps2_p, alpha=1.0):
    """Reg_MIWAE.loss (VAE.py:3197-3263).  q / pbr = (xm, xs, df, mean, scale) of the two branches."""
    xm_q, xs_q, df_q, mean_q, scale_q = q
    xm_p, xs_p, df_p, mean_p, scale_p = pbr
    B, S, D = xm_q.shape
    logp_q, lw_q = _miwae_bound_terms(x, mask, xm_q, xs_q, df_q, mean_q, scale_q, eps2_q)
    nb_q = -torch.mean(torch.logsumexp(lw_q, 0))
    _, lw_p = _miwae_bound_terms(x, mask_p, xm_p, xs_p, df_p, mean_p, scale_p, eps2_p)
    nb_p = -torch.mean(torch.logsumexp(lw_p, 0))
    only_q = (_as(mask, x) * (1.0 - _as(mask_p, x))).unsqueeze(1)
    reg_like = (logp_q * only_q).sum(2).mean()                              # mean over all (row, sample) pairs
    # KL(N(mean_q, scale_q) || N(mean_p, scale_p)) per latent, mean over [B, S, L] == mean over [B, L]
    var_ratio = (scale_q / scale_p) ** 2
    t1 = ((mean_q - mean_p) / scale_p) ** 2
    kl_reg = (0.5 * (var_ratio + t1 - 1.0 - torch.log(var_ratio))).mean()
    loss = nb_q + alpha * (kl_reg - nb_q + nb_p - reg_like)
    w = torch.softmax(lw_q, 0)
    xm_imp = torch.einsum('ki,kij->ij', w, xm_q.permute(1, 0, 2))
    return loss, xm_imp


def miwae_trainable_names(p: Params) -> Sequence[str]:
    return [k for k in p if k.startswith("seq_encoder") or k.startswith("seq_decoder")]


def miwae_train_step(p: Params, x, mask, mask_p, draws, alpha=1.0, regularised=True):
    """forward + loss + backward.  `draws` in the reference's order (SURVEY.md A.6 style): regularised
    [eps_q, eps_p, eps2_q, eps2_p], vanilla [eps, eps2]; each [B, S, L]."""
    names = miwae_trainable_names(p)
    w = {k: (v.detach().clone().requires_grad_(True) if k in names else v) for k, v in p.items()}
    mean_q, scale_q = miwae_encoder_stats(w, x, mask)
    z_q = mean_q.unsqueeze(1) + scale_q.unsqueeze(1) * draws[0]
    xm_q, xs_q, df_q = miwae_decoder(w, z_q)
    if regularised:
        mean_p, scale_p = miwae_encoder_stats(w, x, mask_p)
        z_p = mean_p.unsqueeze(1) + scale_p.unsqueeze(1) * draws[1]
        xm_p, xs_p, df_p = miwae_decoder(w, z_p)
        loss, xm_imp = reg_miwae_loss(x, mask, mask_p, (xm_q, xs_q, df_q, mean_q, scale_q),
                                      (xm_p, xs_p, df_p, mean_p, scale_p), draws[2], draws[3], alpha)
        imp = None
    else:
        loss, xm_imp, imp = miwae_loss(x, mask, xm_q, xs_q, df_q, mean_q, scale_q, draws[1])
    grads = torch.autograd.grad(loss, [w[k] for k in names], allow_unused=True)
    gd = {k: (g if g is not None else torch.zeros_like(w[k])) for k, g in zip(names, grads)}
    return loss.detach(), gd, dict(xm_q=xm_q.detach(), xs_q=xs_q.detach(), df_q=df_q.detach(), mean_q=mean_q.detach(),
                                   scale_q=scale_q.detach(), xm_imp=xm_imp.detach(),
                                   imp=None if imp is None else imp.detach())


def miwae_loss_closed_form_grads(x, mask, raw, mean, scale, eps2):
    """What a fused MIWAE loss kernel has to produce (the kernel of SURVEY.md 8f item 4 is to follow this line by line):
    the loss and its gradients with respect to the RAW decoder output `raw` [B, S, 3D] (before the sigmoid / softplus
    heads, VAE.py:3061-3066) and the direct gradients with respect to the encoder statistics mean / scale [B, L] through
    log p(z) - log q(z|x) of the loss-internal draw z = mean + scale * eps2 (VAE.py:3086-3092).  Closed forms only (no
    autograd); tests/test_oracle_golden.py checks them against autograd of miwae_loss."""
    B, S, D3 = raw.shape
    D = D3 // 3
    sig = torch.sigmoid
    xm = sig(raw[..., :D])
    xs = torch.nn.functional.softplus(raw[..., D:2 * D]) + 0.001
    df = torch.nn.functional.softplus(raw[..., 2 * D:]) + 3.0
    xb = x.unsqueeze(1)
    y = (xb - xm) / xs
    A = 1.0 + y * y / df
    logp = -0.5 * (df + 1.0) * torch.log(A) - (torch.log(xs) + 0.5 * torch.log(df) + HALF_LOG_PI + torch.lgamma(0.5 * df)
                                               - torch.lgamma(0.5 * (df + 1.0)))
    mf = _as(mask, x).unsqueeze(1)
    lpx_flat = (logp * mf).sum(2).reshape(-1)                               # index k = b * S + s
    z = mean.unsqueeze(1) + scale.unsqueeze(1) * eps2
    logpz = (-0.5 * z * z - HALF_LOG_2PI).sum(2)                            # [B, S]
    logq = (-0.5 * eps2 * eps2 - torch.log(scale).unsqueeze(1) - HALF_LOG_2PI).sum(2)
    lw = lpx_flat.reshape(S, B) + (logpz - logq).t()                        # [S, B], the reference's un-transposed reshape
    loss = -torch.mean(torch.logsumexp(lw, 0))
    w = torch.softmax(lw, 0)                                                # [S, B]
    g_lw = -w / B                                                           # dloss / dlw[i, j]
    # likelihood side: lw[i, j] reads lpx_flat[i * B + j], i.e. (row, sample) = divmod(i * B + j, S)
    g_lpx = g_lw.reshape(-1).reshape(B, S)                                  # back in (b, s) order
    dl_loc = (df + 1.0) * y / (xs * df * A)
    dl_scale = (df + 1.0) * y * y / (xs * df * A) - 1.0 / xs
    dl_df = (-0.5 * torch.log(A) + 0.5 * (df + 1.0) * y * y / (df * df * A)
             - (0.5 / df + 0.5 * torch.digamma(0.5 * df) - 0.5 * torch.digamma(0.5 * (df + 1.0))))
    gm = (g_lpx.unsqueeze(2) * mf)
    d_raw = torch.cat([gm * dl_loc * xm * (1.0 - xm),
                       gm * dl_scale * sig(raw[..., D:2 * D]),
                       gm * dl_df * sig(raw[..., 2 * D:])], 2)
    # prior / posterior side: lw[i, j] reads (row j, sample i)
    gt = g_lw.t().unsqueeze(2)                                              # [B, S, 1]
    d_mean = (gt * (-z)).sum(1)
    d_scale = (gt * (-z * eps2 + 1.0 / scale.unsqueeze(1))).sum(1)
    return loss, d_raw, d_mean, d_scale


def reg_miwae_loss_closed_form_grads(x, mask, mask_p, raw_q, mean_q, scale_q, eps2_q, raw_p, mean_p, scale_p, eps2_p, alpha=1.0):
    """Reg_MIWAE.loss (VAE.py:3197-3263) and its gradients in closed form:
        loss = nb_q + alpha (KL_reg - nb_q + nb_p - reg_like)
             = (1 - alpha) nb_q + alpha nb_p + alpha KL_reg - alpha reg_like
    nb_* as in miwae_loss_closed_form_grads (masks: mask for q, mask_p for p), reg_like the mean over (row, sample) of the
    q-branch log-likelihood on mask & ~mask_p, KL_reg the mean over [B, L] of KL(N(mean_q, scale_q) || N(mean_p, scale_p)).
    Returns (loss, d_raw_q, d_mean_q, d_scale_q, d_raw_p, d_mean_p, d_scale_p)."""
    B, S, D3 = raw_q.shape
    D = D3 // 3
    nb_q, dq_raw, dq_mean, dq_scale = miwae_loss_closed_form_grads(x, mask, raw_q, mean_q, scale_q, eps2_q)
    nb_p, dp_raw, dp_mean, dp_scale = miwae_loss_closed_form_grads(x, mask_p, raw_p, mean_p, scale_p, eps2_p)
    # reg_like and its gradient with respect to the q branch's raw decoder output
    sig = torch.sigmoid
    xm = sig(raw_q[..., :D])
    xs = torch.nn.functional.softplus(raw_q[..., D:2 * D]) + 0.001
    df = torch.nn.functional.softplus(raw_q[..., 2 * D:]) + 3.0
    y = (x.unsqueeze(1) - xm) / xs
    A = 1.0 + y * y / df
    logp = -0.5 * (df + 1.0) * torch.log(A) - (torch.log(xs) + 0.5 * torch.log(df) + HALF_LOG_PI + torch.lgamma(0.5 * df)
                                               - torch.lgamma(0.5 * (df + 1.0)))
    only_q = (_as(mask, x) * (1.0 - _as(mask_p, x))).unsqueeze(1)
    reg_like = (logp * only_q).sum(2).mean()
    gw = only_q / (B * S)
    dl_loc = (df + 1.0) * y / (xs * df * A)
    dl_scale = (df + 1.0) * y * y / (xs * df * A) - 1.0 / xs
    dl_df = (-0.5 * torch.log(A) + 0.5 * (df + 1.0) * y * y / (df * df * A)
             - (0.5 / df + 0.5 * torch.digamma(0.5 * df) - 0.5 * torch.digamma(0.5 * (df + 1.0))))
    d_reg_raw = torch.cat([gw * dl_loc * xm * (1.0 - xm), gw * dl_scale * sig(raw_q[..., D:2 * D]),
                           gw * dl_df * sig(raw_q[..., 2 * D:])], 2)
    # KL_reg = mean over [B, L] of  log(sp/sq) + (sq^2 + (mq - mp)^2) / (2 sp^2) - 1/2
    L = mean_q.shape[1]
    dm = mean_q - mean_p
    kl = (torch.log(scale_p / scale_q) + (scale_q ** 2 + dm ** 2) / (2 * scale_p ** 2) - 0.5).mean()
    n = B * L
    dk_mq = dm / scale_p ** 2 / n
    dk_mp = -dk_mq
    dk_sq = (-1.0 / scale_q + scale_q / scale_p ** 2) / n
    dk_sp = (1.0 / scale_p - (scale_q ** 2 + dm ** 2) / scale_p ** 3) / n
    loss = (1.0 - alpha) * nb_q + alpha * nb_p + alpha * kl - alpha * reg_like
    return (loss,
            (1.0 - alpha) * dq_raw - alpha * d_reg_raw, (1.0 - alpha) * dq_mean + alpha * dk_mq, (1.0 - alpha) * dq_scale + alpha * dk_sq,
            alpha * dp_raw, alpha * dp_mean + alpha * dk_mp, alpha * dp_scale + alpha * dk_sp)

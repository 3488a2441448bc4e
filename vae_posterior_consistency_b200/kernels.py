"""Tensor-level wrappers around the C ABI (no autograd here; see ops.py).

PyTorch supplies device memory and the current CUDA stream only; every
arithmetic step of the hot path runs inside libpcvae_b200.so.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import List, Optional, Sequence

import torch

from . import lib as L

#: fixed decoder log-variance log 0.02 (reference src/models/VAE.py:379)
X_LOGVAR = math.log((0.1 * math.sqrt(2.0)) ** 2)
LATENT = 10

MLP_KEYS = ["seq_encoder.0.weight", "seq_encoder.0.bias", "seq_encoder.2.weight", "seq_encoder.2.bias",
            "seq_encoder.4.weight", "seq_encoder.4.bias"]
PNP_KEYS = ["type_pars1", "type_bias1", "pnp_encoder1.0.weight", "pnp_encoder1.0.bias",
            "pnp_encoder2.0.weight", "pnp_encoder2.0.bias", "pnp_encoder2.2.weight", "pnp_encoder2.2.bias",
            "pnp_encoder2.4.weight", "pnp_encoder2.4.bias"]
DEC_KEYS = ["seq_decoder.0.weight", "seq_decoder.0.bias", "seq_decoder.2.weight", "seq_decoder.2.bias",
            "seq_decoder.4.weight", "seq_decoder.4.bias"]


def param_keys(family: int) -> List[str]:
    return (PNP_KEYS if family == L.FAMILY_PNP else MLP_KEYS) + DEC_KEYS


def flatten_params(sd, family: int, device=None) -> torch.Tensor:
    """state_dict (reference key names) -> flat fp32 theta in the C-ABI layout."""
    parts = [sd[k].detach().reshape(-1).to(torch.float32) for k in param_keys(family)]
    flat = torch.cat(parts)
    return flat.to(device) if device is not None else flat


def unflatten_params(theta: torch.Tensor, family: int, obs_dim: int, emb_dim: int = 0):
    m = L.model(family, obs_dim, emb_dim)
    offs = L.param_offsets(m)
    D, K = obs_dim, emb_dim
    in1 = 2 * D if family == L.FAMILY_MLP_MASK else D        # Reg_VAE_mask / vanilla_VAE_mask, reference VAE.py:526
    shapes = ([(D, K), (D, 1), (K, K + 2), (K,), (100, K), (100,)] if family == L.FAMILY_PNP else [(100, in1), (100,)])
    shapes += [(50, 100), (50,), (20, 50), (20,), (50, 10), (50,), (100, 50), (100,), (D, 100), (D,)]
    keys = param_keys(family)
    return {k: theta[offs[i]:offs[i + 1]].view(*shapes[i]) for i, k in enumerate(keys)}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _pair(ts: Sequence[Optional[torch.Tensor]]):
    a = (C.c_void_p * 2)()
    for i, t in enumerate(ts[:2]):
        a[i] = _p(t)
    return a


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def prep_masks(masks: Sequence[torch.Tensor]):
    """Masks stay bit-exact: torch.bool/uint8 are passed as bytes, anything else as float32."""
    if all(m.dtype in (torch.bool, torch.uint8) for m in masks):
        return L.MASK_U8, [m.contiguous() for m in masks]
    return L.MASK_F32, [_f32(m) for m in masks]


def pack_mask_bits(mask: torch.Tensor) -> torch.Tensor:
    """Host side of pcvae_prep_packed: a [rows, D] boolean mask -> [rows, ceil(D / 32)] int32 words, bit j of word w =
    mask[row][32 w + j].  Done once per table by the loader (like its min-max pass), not per step."""
    import numpy as np
    m = mask.detach().cpu().numpy().astype(bool)
    rows, D = m.shape
    W = (D + 31) // 32
    by = np.packbits(m, axis=1, bitorder="little")                       # [rows, ceil(D / 8)] bytes
    out = np.zeros((rows, W * 4), dtype=np.uint8)
    out[:, :by.shape[1]] = by
    return torch.from_numpy(out.view("<i4").reshape(rows, W).copy())       # the bit pattern of the uint32 words


def compact_rows(x: torch.Tensor, mask: torch.Tensor):
    """Observed-entry form of a host table: (vals [nnz] fp32 in row-major order, row_off [rows] int32 (as uint32 on the
    device), mask_bits).  Entries under mask == 0 are dropped: they never reach the training loss (reference VAE.py:388,
    411-445), so a training step fed from this form equals one fed from the dense table."""
    m = mask.detach().cpu().bool()
    vals = x.detach().cpu().float()[m].contiguous()
    cnt = m.sum(1, dtype=torch.int64)
    row_off = (torch.cumsum(cnt, 0) - cnt).to(torch.int32)
    return vals, row_off, pack_mask_bits(m)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.PcvaeError("pcvae ops need CUDA tensors: there is no CPU fallback for this path")


class Engine:
    """Per-(family, D, K, device) launcher with cached workspaces."""

    def __init__(self, family: int, obs_dim: int, emb_dim: int = 0, device=None):
        self.lib = L.load()
        self.family, self.D, self.K = family, obs_dim, (emb_dim if family == L.FAMILY_PNP else 0)
        self.model = L.model(family, obs_dim, emb_dim)
        self.P = L.param_count(self.model)
        self.dec_off = L.decoder_offset(self.model)
        self.device = torch.device(device if device is not None else "cuda")
        with torch.cuda.device(self.device):
            self.grid = self.lib.pcvae_grid_ctas()
        if self.grid <= 0:
            L.check(2, "pcvae_grid_ctas")
        self._gp = None
        self._sp = None
        self._ac = None

    # ---- workspaces -------------------------------------------------------------
    def grad_partials(self):
        if self._gp is None:
            self._gp = torch.zeros(self.grid, self.P, device=self.device, dtype=torch.float32)
        return self._gp

    def sums_partials(self):
        if self._sp is None:
            self._sp = torch.zeros(self.grid, L.NSUMS, device=self.device, dtype=torch.float32)
        return self._sp

    def build_weight_images(self, theta, out=None):
        """Operand images of the weights for the tensor-core training kernels, built once from `theta`
        (pcvae_build_weight_images); pass the result as `wimg` to enc_fwd / dec / enc_bwd calls that use the same theta.
        None when this model has no tensor-core training kernels."""
        n = self.lib.pcvae_weight_images_floats(C.byref(self.model))
        if n <= 0:
            return None
        if out is None:
            out = torch.empty(n, device=theta.device, dtype=torch.float32)
        with torch.cuda.device(theta.device):
            L.check(self.lib.pcvae_build_weight_images(C.byref(self.model), _p(theta), _p(out), _stream()),
                    "pcvae_build_weight_images")
        return out

    def pnp_ac(self):
        if self.family != L.FAMILY_PNP:
            return None
        if self._ac is None:
            k4 = (self.K + 3) // 4 * 4
            self._ac = torch.empty(2 * self.D * k4, device=self.device, dtype=torch.float32)
        return self._ac

    def act_ws(self, rows, n_branch):
        n = self.lib.pcvae_enc_act_ws_floats(C.byref(self.model), rows, n_branch)
        return torch.empty(max(n, 1), device=self.device, dtype=torch.float32)

    # ---- host-streamed batches ----------------------------------------------------
    def prep_packed(self, bits, mask, mask_p=None, eps=None, vals=None, row_off=None, x=None, keep=0.7, seed=0xC0FFEE,
                    offset=0):
        """pcvae_prep_packed: bit-packed mask (+ optionally the observed-entry stream of x) -> dense uint8 mask, x,
        sub-mask and noise in one launch.  All tensors on the device; `mask`, `mask_p` uint8/bool [rows, D]."""
        _need_cuda(bits, mask, mask_p, eps, vals, row_off, x)
        rows = mask.shape[0]
        n_eps = 0 if eps is None else eps.shape[0]
        with torch.cuda.device(self.device):
            L.check(self.lib.pcvae_prep_packed(_p(bits), _p(vals), _p(row_off), _p(x), _p(mask), _p(mask_p), _p(eps), rows,
                                               self.D, n_eps, float(keep), seed, offset, _stream()), "pcvae_prep_packed")

    # ---- encoder ----------------------------------------------------------------
    def enc_fwd(self, theta, x, masks, eps=None, save=False, want_z=True, wimg=None):
        _need_cuda(theta, x, *masks)
        x = _f32(x)
        nb = len(masks)
        kind, masks = prep_masks(masks)
        B = x.shape[0]
        mk = lambda: [torch.empty(B, LATENT, device=x.device, dtype=torch.float32) for _ in range(nb)]
        mean, logvar = mk(), mk()
        z = mk() if want_z else [None] * nb
        eps = [None] * nb if eps is None else [None if e is None else _f32(e) for e in eps]
        # what the backward needs: the tcgen05 encoder's scratch when this shape has one, else the FFMA kernel's
        # saved activations; `ws` is that one tensor (enc_bwd tells the two apart by their sizes, which never coincide)
        act, tcw = None, None
        if save and B > 0:
            n = self.lib.pcvae_enc_tc_workspace_floats(C.byref(self.model), B, nb)   # 0: no tensor-core encoder here
            if n > 0:
                tcw = torch.empty(n, device=x.device, dtype=torch.float32)
            else:
                act = self.act_ws(B, nb)
        elif save:
            act = self.act_ws(B, nb)
        p = L.EncFwdParams(model=self.model, rows=B, n_branch=nb, mask_kind=kind, theta=_p(theta), x=_p(x),
                           mask=_pair(masks), eps=_pair(eps), mean=_pair(mean), logvar=_pair(logvar), z=_pair(z),
                           act_ws=_p(act), pnp_ac=_p(self.pnp_ac()),
                           tc_workspace=_p(tcw), tc_workspace_floats=0 if tcw is None else tcw.numel(),
                           weight_images=_p(wimg))
        with torch.cuda.device(x.device):
            L.check(self.lib.pcvae_enc_fwd(C.byref(p), _stream()), "pcvae_enc_fwd")
        return mean, logvar, z, (tcw if tcw is not None else act)

    def enc_bwd(self, theta, x, masks, ws, d_mean, d_logvar, d_z=None, eps=None, logvar=None, wimg=None):
        x = _f32(x)
        kind, masks = prep_masks(masks)
        gp = self.grad_partials()
        nb = len(masks)
        d_mean = [_f32(t) for t in d_mean]
        d_logvar = [_f32(t) for t in d_logvar]
        opt = lambda ts: [None] * nb if ts is None else [None if t is None else _f32(t) for t in ts]
        d_z, eps, logvar = opt(d_z), opt(eps), opt(logvar)
        n_tc = self.lib.pcvae_enc_tc_workspace_floats(C.byref(self.model), x.shape[0], nb) if x.shape[0] > 0 else 0
        act, tcw = (None, ws) if (n_tc > 0 and ws.numel() == n_tc) else (ws, None)
        p = L.EncBwdParams(model=self.model, rows=x.shape[0], n_branch=nb, mask_kind=kind, theta=_p(theta),
                           x=_p(x), mask=_pair(masks), act_ws=_p(act), d_mean=_pair(d_mean),
                           d_logvar=_pair(d_logvar), pnp_ac=_p(self.pnp_ac()), grad_partials=_p(gp),
                           d_z=_pair(d_z), eps=_pair(eps), logvar=_pair(logvar),
                           tc_workspace=_p(tcw), tc_workspace_floats=0 if tcw is None else tcw.numel(),
                           weight_images=_p(wimg))
        with torch.cuda.device(x.device):
            L.check(self.lib.pcvae_enc_bwd(C.byref(p), _stream()), "pcvae_enc_bwd")
        return gp

    # ---- decoder ----------------------------------------------------------------
    def dec(self, mode, theta, z, *, x=None, masks=(), mean=(), logvar=(), eps=(), alpha=0.0, beta_w=1.0,
            loss_scale=1.0, d_xhat=(), want_xhat=False, x_logvar=X_LOGVAR, wimg=None):
        _need_cuda(theta, *z)
        z = [_f32(t) for t in z]
        nb = len(z)
        B = z[0].shape[0]
        dev = z[0].device
        kind, masks = prep_masks(masks) if len(masks) else (L.MASK_U8, [])
        mk = lambda w: [torch.empty(B, w, device=dev, dtype=torch.float32) for _ in range(nb)]
        xhat = mk(self.D) if (want_xhat or mode == L.DEC_FWD) else [None] * nb
        d_mean = mk(LATENT) if mode == L.DEC_TRAIN else [None] * nb
        d_logvar = mk(LATENT) if mode == L.DEC_TRAIN else [None] * nb
        d_z = mk(LATENT) if mode == L.DEC_BWD else [None] * nb
        pad = lambda ts: list(ts) + [None] * (2 - len(ts))
        conv = lambda ts: pad([None if t is None else _f32(t) for t in ts])
        tcw = None
        if mode == L.DEC_TRAIN:      # scratch for the tcgen05 decoder (0 floats when this shape has none)
            n = self.lib.pcvae_dec_tc_workspace_floats(C.byref(self.model), B, nb)
            if n > 0:
                tcw = torch.empty(n, device=dev, dtype=torch.float32)
        # converted inputs are held in locals until the launch is enqueued
        xc = None if x is None else _f32(x)
        masks_c, mean_c, logvar_c, eps_c, dxh_c = pad(masks), conv(mean), conv(logvar), conv(eps), conv(d_xhat)
        p = L.DecParams(model=self.model, mode=mode, rows=B, n_branch=nb, mask_kind=kind, theta=_p(theta),
                        z=_pair(z), xhat=_pair(xhat), x=_p(xc), mask=_pair(masks_c), mean=_pair(mean_c),
                        logvar=_pair(logvar_c), eps=_pair(eps_c),
                        alpha=alpha, beta_w=beta_w, x_logvar=x_logvar, loss_scale=loss_scale,
                        sums_partials=_p(self.sums_partials()), d_mean=_pair(d_mean), d_logvar=_pair(d_logvar),
                        d_xhat=_pair(dxh_c), d_z=_pair(d_z),
                        grad_partials=_p(self.grad_partials() if mode in (L.DEC_TRAIN, L.DEC_BWD) else None),
                        tc_workspace=_p(tcw), tc_workspace_floats=0 if tcw is None else tcw.numel(),
                        weight_images=_p(wimg))
        with torch.cuda.device(dev):
            L.check(self.lib.pcvae_dec(C.byref(p), _stream()), "pcvae_dec")
        self._last_tcw = tcw      # kept alive until the next call (also lets tests inspect the scratch)
        del xc, masks_c, mean_c, logvar_c, eps_c, dxh_c
        return dict(xhat=xhat, d_mean=d_mean, d_logvar=d_logvar, d_z=d_z)

    # ---- stand-alone loss ---------------------------------------------------------
    def loss_terms(self, x, masks, xhat, mean, logvar, alpha, beta_w, loss_scale, want_grads, x_logvar=X_LOGVAR):
        _need_cuda(x, *xhat)
        x = _f32(x)
        nb = len(xhat)
        kind, masks = prep_masks(masks)
        B, D = x.shape
        xhat = [_f32(t) for t in xhat]
        mean = [_f32(t) for t in mean]
        logvar = [_f32(t) for t in logvar]
        mk = lambda like: [torch.empty_like(t) for t in like] if want_grads else [None] * nb
        d_xhat, d_mean, d_logvar = mk(xhat), mk(mean), mk(logvar)
        p = L.LossParams(rows=B, obs_dim=D, latent_dim=mean[0].shape[1], n_branch=nb, mask_kind=kind, x=_p(x),
                         mask=_pair(masks), xhat=_pair(xhat), mean=_pair(mean), logvar=_pair(logvar), alpha=alpha,
                         beta_w=beta_w, x_logvar=x_logvar, loss_scale=loss_scale,
                         sums_partials=_p(self.sums_partials()), d_xhat=_pair(d_xhat), d_mean=_pair(d_mean),
                         d_logvar=_pair(d_logvar))
        with torch.cuda.device(x.device):
            L.check(self.lib.pcvae_loss_terms(C.byref(p), _stream()), "pcvae_loss_terms")
        return self.reduce_sums(B, D), d_xhat, d_mean, d_logvar

    # ---- reductions / optimiser -----------------------------------------------------
    def reduce_sums(self, rows, obs_dim=None):
        sums = torch.empty(L.NSUMS, device=self.device, dtype=torch.float64)
        with torch.cuda.device(self.device):
            L.check(self.lib.pcvae_reduce_sums(_p(self.sums_partials()), self.grid, rows,
                                               self.D if obs_dim is None else obs_dim, _p(sums), _stream()),
                    "pcvae_reduce_sums")
        return sums

    def reduce_grads(self, grad, begin=0, end=None, accumulate=False):
        end = self.P if end is None else end
        with torch.cuda.device(self.device):
            L.check(self.lib.pcvae_reduce_grads(_p(self.grad_partials()), self.grid, self.P, begin, end, _p(grad),
                                                int(accumulate), _stream()), "pcvae_reduce_grads")
        return grad

    def adam_step(self, theta, grad, exp_avg, exp_avg_sq, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
        with torch.cuda.device(self.device):
            L.check(self.lib.pcvae_adam_step(_p(theta), _p(grad), _p(exp_avg), _p(exp_avg_sq), theta.numel(), step,
                                             lr, b1, b2, eps, _stream()), "pcvae_adam_step")

    def reduce_adam(self, grad, theta, exp_avg, exp_avg_sq, step, rows, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
        """reduce_grads over all parameters + adam_step + reduce_sums in one launch (single-GPU training)."""
        sums = torch.empty(L.NSUMS, device=self.device, dtype=torch.float64)
        with torch.cuda.device(self.device):
            L.check(self.lib.pcvae_reduce_adam(_p(self.grad_partials()), self.grid, self.P, _p(grad), _p(theta),
                                               _p(exp_avg), _p(exp_avg_sq), step, lr, b1, b2, eps,
                                               _p(self.sums_partials()), rows, self.D, _p(sums), _stream()),
                    "pcvae_reduce_adam")
        return sums

    def dp_reduce_adam(self, xch, grad, theta, exp_avg, exp_avg_sq, step, rows, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8,
                       step_state=None, sums=None):
        """Data-parallel tail of a step in one launch (pcvae_dp_reduce_adam): reduce, exchange the reduced gradient with
        the other ranks over NVLink peer memory (`xch`: dist.PeerExchange), rank-ordered sum, Adam, this rank's sums."""
        if sums is None:
            sums = torch.empty(L.NSUMS, device=self.device, dtype=torch.float64)
        p = L.DpParams()
        p.grad_partials, p.grid, p.param_count = _p(self.grad_partials()), self.grid, self.P
        p.grad, p.theta, p.exp_avg, p.exp_avg_sq = _p(grad), _p(theta), _p(exp_avg), _p(exp_avg_sq)
        p.step, p.lr, p.beta1, p.beta2, p.eps = step, lr, b1, b2, eps
        p.sums_partials, p.rows, p.obs_dim, p.sums = _p(self.sums_partials()), rows, self.D, _p(sums)
        p.world, p.rank, p.seq = xch.world, xch.rank, xch.next_seq()
        p.step_state = _p(step_state)                     # device counter (graph replay): seq / step come from it
        for r in range(xch.world):
            p.peer_buffers[r] = xch.ptrs[r]
        p.status = _p(xch.status)
        with torch.cuda.device(self.device):
            L.check(self.lib.pcvae_dp_reduce_adam(C.byref(p), _stream()), "pcvae_dp_reduce_adam")
        return sums

    # ---- reward -----------------------------------------------------------------------
    def reward(self, theta, x, mask, im, workspace=None):
        """R[N, D-1] for one acquisition step (evaluate.py:416-425)."""
        _need_cuda(theta, x, mask, im)
        x = _f32(x)
        im = _f32(im)
        kind, (mask,) = prep_masks([mask])
        M, N, D = im.shape
        assert D == self.D and x.shape == (N, D)
        nbytes = self.lib.pcvae_reward_workspace_bytes(C.byref(self.model), N, M)
        if workspace is None or workspace.numel() < nbytes:
            workspace = torch.empty(max(nbytes, 1), device=x.device, dtype=torch.uint8)
        R = torch.empty(N, D - 1, device=x.device, dtype=torch.float32)
        p = L.RewardParams(model=self.model, rows=N, samples=M, mask_kind=kind, theta=_p(theta), x=_p(x),
                           mask=_p(mask), im=_p(im), im_sample_stride=N * D, R=_p(R), workspace=_p(workspace),
                           workspace_bytes=workspace.numel(), pnp_ac=_p(self.pnp_ac()))
        with torch.cuda.device(x.device):
            L.check(self.lib.pcvae_reward_chain(C.byref(p), _stream()), "pcvae_reward_chain")
        return R, workspace


    def reward_streamed(self, theta, hx, hmask, him, hR, chunks=8, copy_stream=None):
        """The reward of one acquisition step with x / mask / im in PINNED HOST memory and R returned to pinned host
        memory: rows are independent, so they are taken in `chunks` row blocks whose host-to-device
        copies (copy stream, double-buffered; im [M, N, D] is copied sample by sample, each a contiguous block) run
        under the reward kernel of the previous block.  Returns the number of bytes copied in each direction."""
        M, N, D = him.shape
        dev = theta.device
        main = torch.cuda.current_stream(dev)
        cs = copy_stream if copy_stream is not None else torch.cuda.Stream(device=dev)
        nmax = (N + chunks - 1) // chunks
        bufs = getattr(self, "_rs_bufs", None)
        if bufs is None or bufs[0][0].shape[0] < nmax or bufs[0][2].shape[0] != M:
            bufs = [(torch.empty(nmax, D, device=dev), torch.empty(nmax, D, device=dev, dtype=hmask.dtype),
                     torch.empty(M, nmax, D, device=dev), torch.empty(nmax, D - 1, device=dev)) for _ in range(2)]
            self._rs_bufs = bufs
            self._rs_ws = None
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        cs.wait_stream(main)
        h2d = d2h = 0
        for c in range(chunks):
            lo, hi = (c * N) // chunks, ((c + 1) * N) // chunks
            n, b = hi - lo, c % 2
            if n == 0:
                continue
            dx, dm, dim, dR = bufs[b]
            with torch.cuda.stream(cs):
                if c >= 2:
                    cs.wait_event(freed[b])
                dx[:n].copy_(hx[lo:hi], non_blocking=True)
                dm[:n].copy_(hmask[lo:hi], non_blocking=True)
                for m in range(M):
                    dim[m, :n].copy_(him[m, lo:hi], non_blocking=True)
                ready[b].record(cs)
            h2d += 2 * n * D * 4 + M * n * D * 4
            main.wait_event(ready[b])
            # the kernel reads im[m][row] at im + m * stride: the chunk buffer keeps the full-chunk sample stride
            kind, (mk,) = prep_masks([dm[:n]])
            nbytes = self.lib.pcvae_reward_workspace_bytes(C.byref(self.model), n, M)
            if self._rs_ws is None or self._rs_ws.numel() < nbytes:
                self._rs_ws = torch.empty(max(nbytes, 1), device=dev, dtype=torch.uint8)
            p = L.RewardParams(model=self.model, rows=n, samples=M, mask_kind=kind, theta=_p(theta), x=_p(dx), mask=_p(mk),
                               im=_p(dim), im_sample_stride=nmax * D, R=_p(dR), workspace=_p(self._rs_ws),
                               workspace_bytes=self._rs_ws.numel(), pnp_ac=_p(self.pnp_ac()))
            with torch.cuda.device(dev):
                L.check(self.lib.pcvae_reward_chain(C.byref(p), _stream()), "pcvae_reward_chain")
            hR[lo:hi].copy_(dR[:n], non_blocking=True)
            freed[b].record(main)
            d2h += n * (D - 1) * 4
        return h2d, d2h


# ------------------------------------------------------------------------------------------------
# generic dense layers and the not-MIWAE MNAR pieces
# ------------------------------------------------------------------------------------------------

_GRID = {}


def grid_ctas(device) -> int:
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _GRID:
        with torch.cuda.device(idx):
            g = L.load().pcvae_grid_ctas()
        if g <= 0:
            L.check(2, "pcvae_grid_ctas")
        _GRID[idx] = g
    return _GRID[idx]


def dense_fwd(x, W, b, act, mask=None):
    """y = act((x*mask) W^T + b); x [R, K] -> y [R, N]."""
    _need_cuda(x, W, b)
    x, W, b = _f32(x), _f32(W), _f32(b)
    mask = None if mask is None else _f32(mask)
    R, K = x.shape
    N = W.shape[0]
    y = torch.empty(R, N, device=x.device, dtype=torch.float32)
    p = L.DenseFwdParams(rows=R, in_dim=K, out_dim=N, act=act, x=_p(x), mask=_p(mask), W=_p(W), b=_p(b), y=_p(y))
    with torch.cuda.device(x.device):
        L.check(L.load().pcvae_dense_fwd(C.byref(p), _stream()), "pcvae_dense_fwd")
    return y


def dense_bwd(x, W, y, dy, act, mask=None, need_dx=True):
    """(dx, dW, db) of dense_fwd."""
    x, W, y, dy = _f32(x), _f32(W), _f32(y), _f32(dy)
    mask = None if mask is None else _f32(mask)
    R, K = x.shape
    N = W.shape[0]
    lib = L.load()
    g = grid_ctas(x.device)
    dWp = torch.empty(g, N * K, device=x.device, dtype=torch.float32)
    dbp = torch.empty(g, N, device=x.device, dtype=torch.float32)
    dx = torch.empty(R, K, device=x.device, dtype=torch.float32) if need_dx else None
    p = L.DenseBwdParams(rows=R, in_dim=K, out_dim=N, act=act, x=_p(x), mask=_p(mask), y=_p(y), dy=_p(dy), W=_p(W),
                         dx=_p(dx), dW_partials=_p(dWp), db_partials=_p(dbp))
    dW = torch.empty(N, K, device=x.device, dtype=torch.float32)
    db = torch.empty(N, device=x.device, dtype=torch.float32)
    p.dW, p.db = _p(dW), _p(db)          # the call reduces the per-CTA partials itself
    with torch.cuda.device(x.device):
        L.check(lib.pcvae_dense_bwd(C.byref(p), _stream()), "pcvae_dense_bwd")
    return dx, dW, db


def mnar_sample_z(mean, logvar, eps, samples):
    """z[b,s,:] = mean[b] + exp(logvar[b]/2) * eps[b,s]; eps None -> z = mean (sample=False)."""
    mean, logvar = _f32(mean), _f32(logvar)
    eps = None if eps is None else _f32(eps)
    B, Lt = mean.shape
    z = torch.empty(B, samples, Lt, device=mean.device, dtype=torch.float32)
    with torch.cuda.device(mean.device):
        L.check(L.load().pcvae_mnar_sample_z(_p(mean), _p(logvar), _p(eps), _p(z), B, samples, Lt, _stream()),
                "pcvae_mnar_sample_z")
    return z


def mnar_sample_z_bwd(d_z, logvar, eps):
    d_z, logvar = _f32(d_z), _f32(logvar)
    eps = None if eps is None else _f32(eps)
    B, S, Lt = d_z.shape
    d_mean, d_logvar = torch.empty_like(logvar), torch.empty_like(logvar)
    with torch.cuda.device(d_z.device):
        L.check(L.load().pcvae_mnar_sample_z_bwd(_p(d_z), _p(logvar), _p(eps), _p(d_mean), _p(d_logvar), B, S, Lt,
                                                 _stream()), "pcvae_mnar_sample_z_bwd")
    return d_mean, d_logvar


def mnar_loss(x, mask, mask_p, xm, xlv, mean, logvar, W, b, alpha, regularised, eps_kl=None, want_grads=False,
              want_imputed=False):
    """REG_notMIWAE_v2.loss / notMIWAE_myversion.loss.  xm/xlv/mean/logvar are lists over branches (q[, p]).
    Returns dict(out=[loss, RE_q.mean(), loss_q, loss_p] float64, xm_imputed, grads...)."""
    x, mask = _f32(x), _f32(mask)
    mask_p = None if mask_p is None else _f32(mask_p)
    xm, xlv = [_f32(t) for t in xm], [_f32(t) for t in xlv]
    mean, logvar = [_f32(t) for t in mean], [_f32(t) for t in logvar]
    W, b = _f32(W).reshape(-1), _f32(b).reshape(-1)
    eps_kl = None if eps_kl is None else _f32(eps_kl)
    B, S, D = xm[0].shape
    Lt = mean[0].shape[1]
    dev = x.device
    lib = L.load()
    nbytes = lib.pcvae_mnar_loss_workspace_bytes(B, S, D)
    ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    out = torch.empty(4, device=dev, dtype=torch.float64)
    xm_imp = torch.empty(B, D, device=dev) if want_imputed else None
    nb = 2 if regularised else 1
    g = lambda like: [torch.empty_like(t) for t in like[:nb]] if want_grads else [None] * nb
    d_xm, d_xlv, d_mean, d_logvar = g(xm), g(xlv), g(mean), g(logvar)
    d_W = torch.empty(D, device=dev) if want_grads else None
    d_b = torch.empty(D, device=dev) if want_grads else None
    pad = lambda ts: list(ts) + [None] * (2 - len(ts))
    p = L.MnarLossParams(rows=B, samples=S, obs_dim=D, latent_dim=Lt, regularised=int(regularised), x=_p(x),
                         mask=_p(mask), mask_p=_p(mask_p), xm=_pair(pad(xm)), xlv=_pair(pad(xlv)),
                         mean=_pair(pad(mean)), logvar=_pair(pad(logvar)), eps_kl=_p(eps_kl), W=_p(W), b=_p(b),
                         alpha=float(alpha), workspace=_p(ws), workspace_bytes=nbytes, out=_p(out),
                         xm_imputed=_p(xm_imp), d_xm=_pair(pad(d_xm)), d_xlv=_pair(pad(d_xlv)),
                         d_mean=_pair(pad(d_mean)), d_logvar=_pair(pad(d_logvar)), d_W=_p(d_W), d_b=_p(d_b))
    with torch.cuda.device(dev):
        L.check(lib.pcvae_mnar_loss(C.byref(p), _stream()), "pcvae_mnar_loss")
    return dict(out=out, xm_imputed=xm_imp, d_xm=d_xm, d_xlv=d_xlv, d_mean=d_mean, d_logvar=d_logvar, d_W=d_W, d_b=d_b)


IMPUTE_MAX_D = 64


def mnar_impute(model, x, mask, mean, logvar, samples, regularised, eps=None, eps_kl=None, seed=0x1A9E, offset=0):
    """x_imputed [N, D] of a not-MIWAE model for N rows with `samples` importance samples each, in one pass
    (pcvae_mnar_impute).  `model` supplies the decoder / self-masking parameters (nn.Linear layouts); mean / logvar are
    its encoder statistics for (x, mask).  eps / eps_kl [N, S, L] (parity mode) or None (Philox in the kernel)."""
    _need_cuda(x, mask, mean, logvar)
    x, mask, mean, logvar = _f32(x), _f32(mask), _f32(mean), _f32(logvar)
    N, D = x.shape
    Lt = mean.shape[1]
    lib = L.load()
    prm = lambda t: _f32(t.detach())
    w = [prm(model.seq_decoder[0].weight), prm(model.seq_decoder[0].bias), prm(model.seq_decoder[2].weight),
         prm(model.seq_decoder[2].bias), prm(model.x_mean[0].weight), prm(model.x_mean[0].bias),
         prm(model.x_logvar[0].weight), prm(model.x_logvar[0].bias), prm(model.W).reshape(-1), prm(model.b).reshape(-1)]
    eps = None if eps is None else _f32(eps)
    eps_kl = None if eps_kl is None else _f32(eps_kl)
    with torch.cuda.device(x.device):
        nbytes = lib.pcvae_mnar_impute_workspace_bytes(N, samples, D)
    ws = torch.empty(max(nbytes, 4), device=x.device, dtype=torch.uint8)
    out = torch.empty(N, D, device=x.device, dtype=torch.float32)
    p = L.MnarImputeParams(rows=N, samples=samples, obs_dim=D, latent_dim=Lt, regularised=int(regularised),
                           dec0_W=_p(w[0]), dec0_b=_p(w[1]), dec2_W=_p(w[2]), dec2_b=_p(w[3]), xmean_W=_p(w[4]),
                           xmean_b=_p(w[5]), xlogvar_W=_p(w[6]), xlogvar_b=_p(w[7]), W=_p(w[8]), b=_p(w[9]), x=_p(x),
                           mask=_p(mask), mean=_p(mean), logvar=_p(logvar), eps=_p(eps), eps_kl=_p(eps_kl), seed=seed,
                           offset=offset, workspace=_p(ws), workspace_bytes=ws.numel(), xm_imputed=_p(out))
    with torch.cuda.device(x.device):
        L.check(lib.pcvae_mnar_impute(C.byref(p), _stream()), "pcvae_mnar_impute")
    return out


# ------------------------------------------------------------------------------------------------
# MIWAE / Reg_MIWAE pieces (Student-t decoder + importance-weighted bound), reference VAE.py:3011-3301
# ------------------------------------------------------------------------------------------------

def miwae_heads(raw, mode):
    """Encoder heads (mode ENC): raw [R, 2W] -> (mean, softplus); decoder heads (mode DEC): raw [R, 3W] ->
    (sigmoid, softplus + 0.001, softplus + 3).  VAE.py:3047-3049, 3061-3066."""
    _need_cuda(raw)
    raw = _f32(raw)
    C_ = 2 if mode == L.MIWAE_HEADS_ENC else 3
    R, W = raw.shape[0], raw.shape[1] // C_
    outs = [torch.empty(R, W, device=raw.device, dtype=torch.float32) for _ in range(C_)]
    with torch.cuda.device(raw.device):
        L.check(L.load().pcvae_miwae_heads(_p(raw), R, W, mode, _p(outs[0]), _p(outs[1]), _p(outs[2]) if C_ == 3 else None,
                                           _stream()), "pcvae_miwae_heads")
    return outs


def miwae_heads_bwd(raw, mode, grads):
    raw = _f32(raw)
    C_ = 2 if mode == L.MIWAE_HEADS_ENC else 3
    R, W = raw.shape[0], raw.shape[1] // C_
    g = [None if t is None else _f32(t) for t in grads] + [None] * (3 - len(grads))
    d_raw = torch.empty_like(raw)
    with torch.cuda.device(raw.device):
        L.check(L.load().pcvae_miwae_heads_bwd(_p(raw), R, W, mode, _p(g[0]), _p(g[1]), _p(g[2]), _p(d_raw), _stream()),
                "pcvae_miwae_heads_bwd")
    return d_raw


def miwae_sample_z(mean, scale, eps, samples):
    """z[b,s,:] = mean[b] + scale[b] * eps[b,s]; eps None -> z = mean (sample=False).  VAE.py:3050-3058."""
    _need_cuda(mean, scale)
    mean, scale = _f32(mean), _f32(scale)
    eps = None if eps is None else _f32(eps)
    B, Lt = mean.shape
    z = torch.empty(B, samples, Lt, device=mean.device, dtype=torch.float32)
    with torch.cuda.device(mean.device):
        L.check(L.load().pcvae_miwae_sample_z(_p(mean), _p(scale), _p(eps), _p(z), B, samples, Lt, _stream()),
                "pcvae_miwae_sample_z")
    return z


def miwae_sample_z_bwd(d_z, eps):
    d_z = _f32(d_z)
    eps = None if eps is None else _f32(eps)
    B, S, Lt = d_z.shape
    d_mean = torch.empty(B, Lt, device=d_z.device, dtype=torch.float32)
    d_scale = torch.empty_like(d_mean)
    with torch.cuda.device(d_z.device):
        L.check(L.load().pcvae_miwae_sample_z_bwd(_p(d_z), _p(eps), _p(d_mean), _p(d_scale), B, S, Lt, _stream()),
                "pcvae_miwae_sample_z_bwd")
    return d_mean, d_scale


def miwae_loss(x, mask, mask_p, xm, xs, df, mean, scale, eps2, alpha, regularised, rowwise=False, want_grads=False,
               want_imputed=False):
    """MIWAE.loss / Reg_MIWAE.loss (VAE.py:3068-3110, 3197-3263).  xm / xs / df / mean / scale / eps2 are lists over the
    branches (q[, p]).  Returns dict(out=[loss, nb_q, nb_p, kl_reg, reg_like, imp] float64, xm_imputed, gradients)."""
    _need_cuda(x, *xm)
    x = _f32(x)
    nb = 2 if regularised else 1
    kind, masks = prep_masks([mask, mask_p] if regularised else [mask])
    xm, xs, df = [_f32(t) for t in xm[:nb]], [_f32(t) for t in xs[:nb]], [_f32(t) for t in df[:nb]]
    mean, scale, eps2 = [_f32(t) for t in mean[:nb]], [_f32(t) for t in scale[:nb]], [_f32(t) for t in eps2[:nb]]
    B, S, D = xm[0].shape
    Lt = mean[0].shape[1]
    dev = x.device
    lib = L.load()
    nbytes = lib.pcvae_miwae_loss_workspace_bytes(B, S)
    ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    out = torch.empty(6, device=dev, dtype=torch.float64)
    xm_imp = torch.empty(B, D, device=dev) if want_imputed else None
    g = lambda like: [torch.empty_like(t) for t in like] if want_grads else [None] * nb
    d_xm, d_xs, d_df, d_mean, d_scale = g(xm), g(xs), g(df), g(mean), g(scale)
    pad = lambda ts: list(ts) + [None] * (2 - len(ts))
    p = L.MiwaeLossParams(rows=B, samples=S, obs_dim=D, latent_dim=Lt, regularised=int(regularised), mask_kind=kind,
                          rowwise=int(rowwise), x=_p(x), mask=_p(masks[0]), mask_p=_p(masks[1]) if regularised else None,
                          xm=_pair(pad(xm)), xs=_pair(pad(xs)), df=_pair(pad(df)), mean=_pair(pad(mean)),
                          scale=_pair(pad(scale)), eps2=_pair(pad(eps2)), alpha=float(alpha), workspace=_p(ws),
                          workspace_bytes=nbytes, out=_p(out), xm_imputed=_p(xm_imp), d_xm=_pair(pad(d_xm)),
                          d_xs=_pair(pad(d_xs)), d_df=_pair(pad(d_df)), d_mean=_pair(pad(d_mean)),
                          d_scale=_pair(pad(d_scale)))
    with torch.cuda.device(dev):
        L.check(lib.pcvae_miwae_loss(C.byref(p), _stream()), "pcvae_miwae_loss")
    return dict(out=out, xm_imputed=xm_imp, d_xm=d_xm, d_xs=d_xs, d_df=d_df, d_mean=d_mean, d_scale=d_scale)


def loss_from_sums(sums: torch.Tensor, rows: int, alpha: float, beta_w: float, regularised: bool):
    """train_loss = L / B with L as in VAE.py:441-452 (device tensor, float64)."""
    loss_q = sums[L.S_RE_Q] + beta_w * sums[L.S_KL_Q]
    if not regularised:
        return loss_q / rows
    loss_p = sums[L.S_RE_P] + beta_w * sums[L.S_KL_P]
    return (loss_q + alpha * (sums[L.S_KL_REG] - loss_q + loss_p + sums[L.S_RE_D])) / rows


class FusedTrainer:
    """The fused training step of train.py:87-116 on flat device vectors:
    encoder fwd (both branches) -> decoder + loss + decoder bwd -> encoder bwd ->
    deterministic partial reduce -> Adam.  Data parallel (dist_group, world_size > 1): the tail is one
    launch, reduce + gradient exchange over NVLink peer memory + Adam (pcvae_dp_reduce_adam, dist.PeerExchange);
    with PCVAE_DP=nccl, or when peer access cannot be set up, reduce -> NCCL all-reduce -> Adam."""

    def __init__(self, family, obs_dim, emb_dim, theta, regularised=True, alpha=1.0, beta_w=1.0, lr=1e-3,
                 dist_group=None, world_size=1):
        self.eng = Engine(family, obs_dim, emb_dim, theta.device)
        self.theta = theta
        self.grad = torch.zeros_like(theta)
        self.exp_avg = torch.zeros_like(theta)
        self.exp_avg_sq = torch.zeros_like(theta)
        self.step_count = 0
        self.regularised, self.alpha, self.beta_w, self.lr = regularised, alpha, beta_w, lr
        self.dist_group, self.world_size = dist_group, world_size
        # data parallel: the fused reduce + NVLink exchange + Adam kernel unless PCVAE_DP=nccl asks for the plain
        # reduce -> NCCL all-reduce -> Adam sequence (kept as the cross-check of the fused kernel)
        # operand images of the weights for the tensor-core kernels, rebuilt from theta once per step (two tiny launches)
        # instead of by every CTA of the four row-tile kernels (PCVAE_WEIGHT_IMAGES=0: per launch, as before)
        self.wimg = None
        if os.environ.get("PCVAE_WEIGHT_IMAGES", "1") != "0":
            n = self.eng.lib.pcvae_weight_images_floats(C.byref(self.eng.model))
            if n > 0:
                self.wimg = torch.empty(n, device=theta.device, dtype=torch.float32)
        self.xch = None
        if dist_group is not None and world_size > 1 and os.environ.get("PCVAE_DP", "peer") != "nccl":
            from .dist import PeerExchange
            self.xch = PeerExchange.create_or_none(self.eng.P, dist_group, theta.device)

    def forward_backward(self, x, mask, mask_p, eps_q, eps_p, global_rows=None, reduce=True, images_ready=False):
        """`images_ready`: the caller has already rebuilt self.wimg from the current theta (on a forked stream, joined)."""
        e = self.eng
        if self.wimg is not None and not images_ready:
            e.build_weight_images(self.theta, self.wimg)
        B = x.shape[0]
        rows = B if global_rows is None else global_rows
        masks = [mask, mask_p] if self.regularised else [mask]
        eps = [eps_q, eps_p] if self.regularised else [eps_q]
        alpha = self.alpha if self.regularised else 0.0
        mean, logvar, z, ws = e.enc_fwd(self.theta, x, masks, eps, save=True, wimg=self.wimg)
        out = e.dec(L.DEC_TRAIN, self.theta, z, x=x, masks=masks, mean=mean, logvar=logvar, eps=eps, alpha=alpha,
                    beta_w=self.beta_w, loss_scale=1.0 / rows, wimg=self.wimg)
        e.enc_bwd(self.theta, x, masks, ws, out["d_mean"], out["d_logvar"], wimg=self.wimg)
        if not reduce:
            return None
        e.reduce_grads(self.grad)
        sums = e.reduce_sums(B)
        return sums

    def step(self, x, mask, mask_p, eps_q, eps_p, global_rows=None):
        self.step_count += 1
        if self.xch is not None:
            self.forward_backward(x, mask, mask_p, eps_q, eps_p, global_rows, reduce=False)
            sums = self.eng.dp_reduce_adam(self.xch, self.grad, self.theta, self.exp_avg, self.exp_avg_sq, self.step_count,
                                           x.shape[0], lr=self.lr)
        elif self.dist_group is not None and self.world_size > 1:
            sums = self.forward_backward(x, mask, mask_p, eps_q, eps_p, global_rows)
            torch.distributed.all_reduce(self.grad, group=self.dist_group)
            self.eng.adam_step(self.theta, self.grad, self.exp_avg, self.exp_avg_sq, self.step_count, lr=self.lr)
        else:       # single GPU: partial reduce, Adam and the loss sums in one launch
            self.forward_backward(x, mask, mask_p, eps_q, eps_p, global_rows, reduce=False)
            sums = self.eng.reduce_adam(self.grad, self.theta, self.exp_avg, self.exp_avg_sq, self.step_count,
                                        x.shape[0], lr=self.lr)
        rows = x.shape[0] if global_rows is None else global_rows
        return loss_from_sums(sums, rows, self.alpha, self.beta_w, self.regularised)


class GraphedFusedTrainer(FusedTrainer):
    """FusedTrainer with the whole step -- batch gather + sub-mask + noise (pcvae_prep_batch_dev), the six training
    kernels, reduce + Adam (pcvae_reduce_adam_dev) and the running loss total -- captured once in a CUDA graph and
    replayed (throughput mode, single GPU, uint8 masks, obs_dim % 4 == 0).  The per-step scalars (batch number, Philox
    offset, Adam step) come from a device counter, so a replay is an exact repetition of the launch sequence an eager
    `prep_batch -> step` would issue for that step number (tests/test_gpu_parity.py checks bit-identity).  At the
    reference's default batch of 64 the step is launch-bound: 210 -> 96 us; at 65 536 rows the launch gaps go: 386 -> 358 us.

    `idx_batches` holds n_batches index lists of `batch_rows` rows; step number s uses list s % n_batches, so the caller
    writes the lists of an epoch rotated by the step number the epoch starts at (`set_batches`)."""

    def __init__(self, family, obs_dim, emb_dim, theta, table, mask_table, batch_rows, n_batches, keep=0.7, seed=0xC0FFEE,
                 regularised=True, alpha=1.0, beta_w=1.0, lr=1e-3, dist_group=None, world_size=1, global_rows=None):
        super().__init__(family, obs_dim, emb_dim, theta, regularised=regularised, alpha=alpha, beta_w=beta_w, lr=lr,
                         dist_group=dist_group, world_size=world_size)
        if world_size > 1 and self.xch is None:
            raise L.PcvaeError("GraphedFusedTrainer: data parallelism needs the NVLink peer exchange (an NCCL all-reduce "
                               "inside the captured step is not supported)")
        self.global_rows = global_rows
        if mask_table.dtype not in (torch.bool, torch.uint8) or obs_dim % 4 or obs_dim > 128:
            raise L.PcvaeError("GraphedFusedTrainer: needs uint8 / bool masks and obs_dim % 4 == 0, obs_dim <= 128")
        dev = theta.device
        self.table, self.mtable = _f32(table), mask_table.contiguous()
        self.B, self.n_batches, self.keep, self.seed = int(batch_rows), int(n_batches), float(keep), int(seed)
        self.idx = torch.zeros(self.n_batches, self.B, dtype=torch.int64, device=dev)
        self.state = torch.zeros(2, dtype=torch.int64, device=dev)          # [completed steps, ticket]
        # Prepare-ahead (default; PCVAE_PREP_AHEAD=0 for the plain prep -> step order): the batch of step n + 1 is gathered on a
        # forked stream beside the last launch of step n -- reduce [+ NVLink gradient exchange] + Adam, which reads none of
        # x / masks / noise and, under data parallelism, spends most of its time waiting for the other ranks' flags.  The
        # preparation counts its own steps (`prep_count`, advanced on the forked stream) because the step counter proper is
        # advanced by the kernel it runs beside.  Same lists, same Philox offsets, same buffers: the step sequence is
        # unchanged (tests/test_gpu_parity.py: bit-identical to the prep -> step order).
        # 1 GPU: 0.3432 -> 0.3372 ms at cfg4; 2 GPUs: 0.3680 -> 0.3401 ms (the exchange is hidden under the gather)
        self.ahead = os.environ.get("PCVAE_PREP_AHEAD", "1") == "1"
        self.prep_count = torch.zeros(1, dtype=torch.int64, device=dev)
        self._prepped = False
        # prepare-ahead mode also rebuilds the weight images for the NEXT step right after Adam, under the tail of the forked
        # gather, instead of at the start of the step where nothing hides the launch (4 us); whatever changes theta between
        # steps (host-counted steps, new lists at an epoch boundary, a caller writing into theta) goes through
        # invalidate_images() / sync_counter() / set_batches(), which make the next step rebuild them first
        self._imgs_ready = False
        self._imgs_in_tail = self.ahead and os.environ.get("PCVAE_IMAGES_IN_TAIL", "1") == "1"
        self._fork = torch.cuda.Stream(device=dev) if self.ahead else None
        self._img_stream = torch.cuda.Stream(device=dev) if self.wimg is not None else None
        self.x = torch.empty(self.B, obs_dim, device=dev)
        self.mask = torch.empty(self.B, obs_dim, device=dev, dtype=mask_table.dtype)
        self.mask_p = torch.empty_like(self.mask)
        self.n_eps = 2 if regularised else 1
        self.eps = torch.empty(self.n_eps, self.B, LATENT, device=dev)
        self.sums2 = torch.zeros(2 * L.NSUMS, dtype=torch.float64, device=dev)  # [step sums | totals since reset_total()]
        self.sums = self.sums2[:L.NSUMS]
        self.graph = None

    def set_batches(self, idx_batches: torch.Tensor):
        """Index lists of the coming n_batches steps, in order; rotated here so that list j is used at step number
        step_count + j."""
        assert idx_batches.shape == (self.n_batches, self.B)
        rot = self.step_count % self.n_batches
        self.idx.copy_(torch.roll(idx_batches.to(self.idx.device), rot, 0))
        self._imgs_ready = False
        if self.ahead:                                    # the batch of the coming step, from the new lists
            self._reprep()

    def invalidate_images(self):
        """Call after writing into `theta` between steps: the next step rebuilds the weight operand images first."""
        self._imgs_ready = False

    def reset_total(self):
        self.sums2[L.NSUMS:].zero_()

    @property
    def total(self):
        """Sum of the step losses since reset_total() (every step of the sum had B rows: the loss is linear in the sums)."""
        return loss_from_sums(self.sums2[L.NSUMS:], self.global_rows or self.B, self.alpha, self.beta_w, self.regularised)

    def _prep(self, counter):
        """Batch of the step whose number `counter` (a device word) holds into x / mask / mask_p / eps."""
        with torch.cuda.device(self.theta.device):
            L.check(self.eng.lib.pcvae_prep_batch_dev(_p(self.table), _p(self.mtable), _p(self.idx), self.n_batches,
                                                      _p(self.x), _p(self.mask), _p(self.mask_p), _p(self.eps), self.B,
                                                      self.eng.D, self.n_eps, self.keep, self.seed, 0, _p(counter), _stream()),
                    "pcvae_prep_batch_dev")

    def _reprep(self):
        """Prepare-ahead mode: (re)prepare the batch of the step about to run and point the preparation at the one after."""
        self.prep_count.copy_(self.state[:1])
        self._prep(self.prep_count)
        self.prep_count += 1
        self._prepped = True

    def _launch_step(self):
        args = (self.x, self.mask, self.mask_p if self.regularised else None, self.eps[0],
                self.eps[1] if self.regularised else None)
        main = torch.cuda.current_stream()
        build_now = self.wimg is not None and not (self._imgs_in_tail and self._imgs_ready)
        if build_now:                                     # the weight images of this step, beside the batch preparation
            self._img_stream.wait_stream(main)
            with torch.cuda.stream(self._img_stream):
                self.eng.build_weight_images(self.theta, self.wimg)
        if self.ahead:
            if not self._prepped:
                self._reprep()
        else:
            self._prep(self.state)
        if build_now:
            main.wait_stream(self._img_stream)
        self.forward_backward(*args, global_rows=self.global_rows, reduce=False, images_ready=True)
        if self.ahead:                                    # the next batch, beside reduce [+ exchange] + Adam
            self._fork.wait_stream(main)
            with torch.cuda.stream(self._fork):
                self._prep(self.prep_count)
                self.prep_count += 1
        sums = self._launch_tail()
        if self.ahead:
            if self.wimg is not None and self._imgs_in_tail:  # images of the next step from the theta Adam just wrote
                self.eng.build_weight_images(self.theta, self.wimg)
                self._imgs_ready = True
            main.wait_stream(self._fork)
        return sums

    def _launch_tail(self):
        e = self.eng
        sums = self.sums2
        if self.xch is not None:                          # data parallel: reduce + NVLink exchange + Adam, device-counted
            e.dp_reduce_adam(self.xch, self.grad, self.theta, self.exp_avg, self.exp_avg_sq, 1, self.B, lr=self.lr,
                             step_state=self.state, sums=sums)
            return self.sums
        with torch.cuda.device(e.device):
            L.check(e.lib.pcvae_reduce_adam_dev(_p(e.grad_partials()), e.grid, e.P, _p(self.grad), _p(self.theta), _p(self.exp_avg),
                                                _p(self.exp_avg_sq), _p(self.state), self.lr, 0.9, 0.999, 1e-8,
                                                _p(e.sums_partials()), self.B, e.D, _p(sums), _stream()), "pcvae_reduce_adam_dev")
        return self.sums

    def capture(self, warmup=3):
        """Warm up on a side stream (`warmup` real steps, they count), then capture one step."""
        # high priority: the kernels of the step are placed before the blocks of the forked batch preparation, which then
        # fill what the weight-gradient kernel leaves free on every SM (the stream priority is recorded in the graph's nodes)
        side = torch.cuda.Stream(device=self.theta.device, priority=-1)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._launch_step()
                self.step_count += 1
                if self.xch is not None:
                    self.xch.seq = self.step_count
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.theta.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):
            self._launch_step()                           # recorded, not executed: neither counter moves
        if self.xch is not None:
            self.xch.seq = self.step_count
        return self

    def step_graph(self):
        """One training step (replay).  Returns the step's loss sums (a static buffer: read or clone before the next step)."""
        if self.graph is None:
            raise L.PcvaeError("GraphedFusedTrainer: capture() first")
        if self.ahead and not self._prepped:              # after sync_counter() without new lists
            self._reprep()
        if self._imgs_in_tail and self.wimg is not None and not self._imgs_ready:
            self.eng.build_weight_images(self.theta, self.wimg)     # the recorded step builds the NEXT step's images only
            self._imgs_ready = True
        self.graph.replay()
        self.step_count += 1
        if self.xch is not None:
            self.xch.seq = self.step_count                # host mirror of the device-counted call number
        return self.sums

    def step_eager_dev(self):
        """The same step launched kernel by kernel (cross-check of the replay)."""
        sums = self._launch_step()
        self.step_count += 1
        if self.xch is not None:
            self.xch.seq = self.step_count
        return sums

    def sync_counter(self):
        """After steps taken through the inherited host-counted `step` (a ragged last batch): publish the host step count."""
        self.state[0] = self.step_count
        self._imgs_ready = False                          # the host-counted steps rebuilt the images BEFORE their Adam update
        self._prepped = False                             # the buffers hold the batch of a step number that has passed

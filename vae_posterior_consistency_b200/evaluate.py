"""Mirror of the hot-path callers in the reference's `src/experiment_main/evaluate.py`:
`eval_vae` (ELBO / RMSE / NLL, evaluate.py:136-297), `active_learning_func` (sequential
feature acquisition, evaluate.py:300-511) and `R_lindley_chain` (evaluate.py:514-542), with
identical signatures and identical saved artefacts.

The active-selection loop is embarrassingly parallel by test row: with torch.distributed
initialised every rank handles a contiguous block of rows, the reward kernel needs no
collective, the information curve takes one scalar all-reduce per step and the histories are
gathered to rank 0 at the end.  In parity mode every rank draws the full
host noise with the same seed and slices it, so N-GPU results are bit-identical to 1 GPU.
"""
import os

import numpy as np
import torch

from . import lib as L
from . import ops
from .dist import all_reduce_sum, merge_row_blocks, row_block, world
from .loaders import model_loader
from .utils import create_missing_uci
from . import kernels as KR
from .VAE import draw_noise, draw_noise_bsl


#: device time and (row, candidate, sample) count of the reward calls of the last active_learning_func run on this rank
LAST_TIMING = {}


def _family_dir(vae_type):
    return ''.join(c for c in '_'.join(vae_type.split('_')[:2]) if not c.isdigit())


def _result_path(kind, experiment_type, data_type, vae_type, stem, missing_rate, alpha, p_missingness, reg_type):
    base = os.path.join('experiments', experiment_type, data_type, kind, _family_dir(vae_type))
    os.makedirs(base, exist_ok=True)
    if 'vanilla' in vae_type:
        return os.path.join(base, f'{stem}_{missing_rate}_missing_rate_test.pt')
    return os.path.join(base, f'{stem}_{alpha}_{p_missingness}_{reg_type}_{missing_rate}_missing_rate_full_reg_test.pt')


def eval_vae(list_loaders, missing_rate, obs_dim, hid_dim, K, M, latent_dim, data_type, training_parameters,
             experiment_type, vae_type, max_epochs, valid_k, num_estimates, device=torch.device('cpu'), alpha=0.5,
             stage='evaluate', p_missingness=30, reg_type='ml_reg', beta=1.0, beta_annealing=False,
             alpha_annealing=True):
    """M repeats over every loader; per batch: forward (both branches draw noise, only q is scored),
    ELBO = (RE_q + beta KL_q)/B, NLL on observed / unobserved entries, RMSE on unobserved entries of the
    stochastic decoder mean (evaluate.py:209-245); saves four scalars per loader (evaluate.py:247-297)."""
    device = torch.device(device)
    with torch.no_grad():
        model = model_loader('test', obs_dim, hid_dim, K, latent_dim, missing_rate, data_type, training_parameters,
                             max_epochs, valid_k, num_estimates, experiment_type, reg_type, vae_type, alpha=alpha,
                             p_missingness=p_missingness)
        model.to(device)
        if not hasattr(model, 'flat_theta'):
            # the reference's else-branch (evaluate.py:218-231) would also take importance-sampling models here; its
            # drivers never do (imputation.py:40-59 sends MIWAE types to eval_miwae, imputation_mnar.py to eval_vae_mnar)
            raise NotImplementedError(f"eval_vae: vae_type {vae_type!r} is an importance-sampling model; use eval_miwae "
                                      "(MIWAE / Reg_MIWAE) or eval_vae_mnar (not-MIWAE)")
        regularised = 'reg' in vae_type
        theta = model.flat_theta()
        eng = ops.engine(model.FAMILY, obs_dim, model._emb(), device)
        beta_w = model._beta_w(beta, beta_annealing, max_epochs)
        results = {}
        for loader, loader_stage in list_loaders:
            recon, res, res_negll, res_negll_imp = [], [], [], []
            for _ in range(M):
                rmses, elbos, negls, negls_imp = [], [], [], []
                for data_sample, mask in loader:
                    data_sample, mask = data_sample.to(device), mask.to(device)
                    B = data_sample.shape[0]
                    create_missing_uci(data_sample.shape, p_missingness)       # drawn for every family, evaluate.py:173
                    eps_q = draw_noise(B, latent_dim, device, model.noise)
                    if regularised:
                        draw_noise(B, latent_dim, device, model.noise)         # the p-branch draw of forward()
                    mean, logvar, z, _ = eng.enc_fwd(theta, data_sample, [mask], [eps_q])
                    eng.dec(L.DEC_EVAL, theta, z, x=data_sample, masks=[mask], mean=mean, logvar=logvar,
                            beta_w=beta_w)
                    s = eng.reduce_sums(B)
                    elbos.append((s[L.S_RE_Q] + beta_w * s[L.S_KL_Q]) / B)
                    negls.append(s[L.S_RE_Q] / B)
                    negls_imp.append(s[L.S_RE_IMP] / B)
                    n_unobs = torch.sum(~mask.bool())
                    rmses.append(torch.sqrt(s[L.S_SSE_UNOBS] / n_unobs))
                recon.append(torch.stack(rmses).float().mean())
                res.append(torch.stack(elbos).float().mean())
                res_negll.append(torch.stack(negls).float().mean())
                res_negll_imp.append(torch.stack(negls_imp).float().mean())
            out = {'rmse': torch.stack(recon).mean().cpu(), 'vae_elbo': torch.stack(res).mean().cpu(),
                   'negative_llh': torch.stack(res_negll).mean().cpu(),
                   'negative_llh_imputed': torch.stack(res_negll_imp).mean().cpu()}
            results[loader_stage] = out
            q = '' if 'vanilla' in vae_type else '_q'
            names = {'rmse': ('rest', '_rmse'), 'vae_elbo': ('elbos', '_vae_elbo'),
                     'negative_llh': ('rest', f'_negative_llh{q}'),
                     'negative_llh_imputed': ('rest', f'_negative_llh{q}_imputed')}
            for key, (kind, suffix) in names.items():
                torch.save(out[key], _result_path(kind, experiment_type, data_type, vae_type,
                                                  f'{loader_stage}_{vae_type}{suffix}', missing_rate, alpha,
                                                  p_missingness, reg_type))
        return results


class _NoiseQueue:
    """Stand-in for a model's `noise` mode that hands out pre-drawn [rows, S, L] tensors in call order."""

    def __init__(self, tensors):
        self.tensors, self.i = list(tensors), 0

    def __call__(self, rows, samples, latent, device):
        t = self.tensors[self.i]
        self.i += 1
        assert t.shape == (rows, samples, latent), (t.shape, rows, samples, latent)
        return t.to(device, non_blocking=True)


def eval_miwae(list_loaders, missing_rate, obs_dim, hid_dim, K, M, latent_dim, data_type, training_parameters,
               experiment_type, vae_type, max_epochs, valid_k, num_estimates, device=torch.device('cpu'), alpha=0.5,
               stage='evaluate', p_missingness=30, reg_type='ml_reg', beta=1.0, beta_annealing=False,
               alpha_annealing=True):
    """Importance-weighted imputation RMSE of MIWAE / Reg_MIWAE (reference evaluate.py:72-133).  The reference calls
    forward + loss(llh_eval=True) once per ROW with S = valid_k samples; here the rows of a batch go through the
    kernels in blocks with `rowwise` semantics (every row its own batch, which is what the per-row calls amount to:
    the [B*S] -> [S, B] reshape of VAE.py:3078-3081 is the identity for B = 1).  In parity mode the host noise is
    drawn row by row in the reference's order (Reg: eps_q, eps_p of forward, then the two loss-internal draws;
    vanilla: eps, then the loss-internal draw), so a seeded run consumes the RNG exactly like the reference."""
    device = torch.device(device)
    with torch.no_grad():
        model = model_loader('test', obs_dim, hid_dim, K, latent_dim, missing_rate, data_type, training_parameters,
                             max_epochs, valid_k, num_estimates, experiment_type, reg_type, vae_type, alpha=alpha,
                             p_missingness=p_missingness)
        model.to(device)
        reg = 'reg_MIWAE' in vae_type
        S, Lt = model.num_samples, latent_dim
        n_draws = 4 if reg else 2
        results = {}
        throughput = os.environ.get('PCVAE_MODE', 'parity') == 'throughput'
        for loader, loader_stage in list_loaders:
            recon = []
            for _ in range(M):
                XM = []
                for data_sample, mask in loader:
                    data_sample, mask = data_sample.to(device), mask.to(device)
                    B = data_sample.shape[0]
                    temp_mask = create_missing_uci(data_sample.shape, p_missingness)      # evaluate.py:94, per batch
                    mask_p = mask * temp_mask.to(device)
                    temp_XM = torch.zeros(B, obs_dim, device=device)
                    block = max(1, (1 << 18) // max(S, 1))                                # ~256k decoder rows per block
                    for lo in range(0, B, block):
                        hi = min(B, lo + block)
                        if throughput:
                            model.noise = 'device'
                        else:
                            draws = [[] for _ in range(n_draws)]
                            for _i in range(lo, hi):
                                for d in draws:
                                    d.append(torch.empty(1, S, Lt).normal_())
                            model.noise = _NoiseQueue([torch.cat(d) for d in draws])
                        try:
                            xb, mb = data_sample[lo:hi], mask[lo:hi]
                            if reg:
                                out = model.forward(xb, mb, mask_p[lo:hi])
                                (mean_p, scale_p, x_mean_p, x_scale_p, deg_free_p, mean_q, scale_q, x_mean_q, x_scale_q,
                                 deg_free_q) = out
                                xm, _, _ = model.loss(xb, x_mean_p, x_scale_p, deg_free_p, mean_p, scale_p, x_mean_q,
                                                      x_scale_q, deg_free_q, mean_q, scale_q, mb, mask_p[lo:hi], 1,
                                                      llh_eval=True, beta_annealing=beta_annealing, beta=beta, alpha=alpha,
                                                      rowwise=True)
                            else:
                                mean, scale, x_mean, x_scale, deg_free = model.forward(xb, mb)
                                xm, _, _ = model.loss(xb, x_mean, x_scale, deg_free, mean, scale, mb, max_epochs,
                                                      llh_eval=True, beta_annealing=beta_annealing, beta=beta, rowwise=True)
                        finally:
                            model.noise = 'host'
                        temp_XM[lo:hi] = xm
                    miss = ~mask.bool()
                    XM.append(torch.sqrt(torch.sum(torch.square(temp_XM * miss - data_sample.view(-1, obs_dim) * miss))
                                         / torch.sum(miss)))
                recon.append(torch.stack(XM).mean())
            recon = torch.stack(recon).mean().cpu()
            results[loader_stage] = recon
            base = os.path.join('experiments', experiment_type, data_type, 'rest', _family_dir(vae_type))
            os.makedirs(base, exist_ok=True)
            if 'vanilla' in vae_type:
                fname = f'{loader_stage}_{vae_type}_rmse_50_missing_rate_test.pt'
            else:
                fname = f'{loader_stage}_{vae_type}_rmse_{alpha}_{p_missingness}_{reg_type}_full_reg_50_missing_rate_test.pt'
            torch.save(recon, os.path.join(base, fname))
        return results


def eval_vae_mnar(data_test, mask_test, missing_rate, obs_dim, hid_dim, K, M, latent_dim, data_type,
                  training_parameters, experiment_type, vae_type, max_epochs, valid_k, num_estimates,
                  device=torch.device('cpu'), alpha=0.5, stage='evaluate', p_missingness=30, reg_type='ml_reg',
                  beta=1.0, beta_annealing=False, alpha_annealing=True, not_miwae_type='changed'):
    """Importance-weighted MNAR imputation RMSE (evaluate.py:13-69).  The reference loops row by row with
    S = valid_k samples; here all rows go through ONE kernel over the [rows, samples] grid with an online softmax
    (pcvae_mnar_impute).  Parity mode: the host noise is still drawn row by row in the
    reference's order (sub-mask, eps_q, eps_p / eps, eps_kl) so that the inputs match bit for bit; PCVAE_MODE=throughput:
    the noise is drawn inside the kernel (Philox), no per-row host work at all."""
    device = torch.device(device)
    with torch.no_grad():
        model = model_loader('test', obs_dim, hid_dim, K, latent_dim, missing_rate, data_type, training_parameters,
                             max_epochs, valid_k, num_estimates, experiment_type, reg_type, vae_type, alpha=alpha,
                             p_missingness=p_missingness, not_miwae_type=not_miwae_type)
        model.to(device)
        reg = 'reg_notMIWAE' in vae_type
        S, Lt = model.num_samples, latent_dim
        N = data_test.shape[0]
        x_all, m_all = data_test.float().to(device), mask_test.float().to(device)
        throughput = os.environ.get('PCVAE_MODE', 'parity') == 'throughput'
        # one pass over the [rows, samples] grid with an online softmax (pcvae_mnar_impute): nothing of size [rows, S, D]
        # is materialised.  Shapes the kernel does not take (obs_dim > 64) and PCVAE_IMPUTE=blocked (the cross-check) go
        # through the generic dense / loss kernels in row blocks.
        fused = (obs_dim <= KR.IMPUTE_MAX_D and Lt <= 16 and os.environ.get('PCVAE_IMPUTE', 'fused') != 'blocked')
        if fused:
            block = N if throughput else max(1, min(N, (1 << 24) // max(S * Lt, 1)))   # parity: <= 64 MB of host noise per draw
        else:
            block = max(1, min(N, (1 << 22) // max(S * obs_dim, 1)))                    # ~4M decoder outputs per block
        recons = []
        for rep in range(M):
            XM = torch.zeros(N, obs_dim, device=device)
            for lo in range(0, N, block):
                hi = min(N, lo + block)
                eps_q = eps_kl = None
                if not throughput:
                    # the reference's per-row order: sub-mask (evaluate.py:31, a full-table draw per row), then the two
                    # [1, S, L] normal draws of the row (reg: q and p encoder; vanilla: encoder, then the loss' z')
                    eq, ek = [], []
                    for i in range(lo, hi):
                        create_missing_uci(data_test.shape, p_missingness)
                        eq.append(torch.empty(1, S, Lt).normal_())
                        ek.append(torch.empty(1, S, Lt).normal_())
                    eps_q, eps_kl = torch.cat(eq).to(device), torch.cat(ek).to(device)
                xb, mb = x_all[lo:hi], m_all[lo:hi]
                mean, log_var = model._stats(xb, mb)
                if fused:
                    XM[lo:hi] = KR.mnar_impute(model, xb, mb, mean, log_var, S, reg, eps=eps_q,
                                               eps_kl=None if (reg or throughput) else eps_kl, seed=0x1A9E + rep,
                                               offset=0)
                    continue
                if throughput:
                    eps_q = draw_noise_bsl(hi - lo, S, Lt, device, 'device')
                    eps_kl = draw_noise_bsl(hi - lo, S, Lt, device, 'device')
                z = ops.mnar_sample_z_op(mean, log_var, eps_q, S)
                xm, xlv = model.decoder(z)
                if reg:
                    # llh_eval only needs the q branch (VAE.py:2458-2461); zero p inputs keep the kernel's contract
                    out = ops.mnar_loss_op(xb, mb, mb, xm, xlv, xm, xlv, mean, log_var, mean, log_var, model.W,
                                           model.b, None, float(alpha), False, True)
                else:
                    out = ops.mnar_loss_op(xb, mb, None, xm, xlv, None, None, mean, log_var, None, None, model.W,
                                           model.b, eps_kl, 1.0, False, True)
                XM[lo:hi] = out[2]
            miss = 1 - m_all
            recons.append(torch.sqrt(torch.sum((XM * miss - x_all * miss) ** 2) / torch.sum(miss)))
        recon = torch.stack(recons).mean().cpu()
        base = os.path.join('experiments', experiment_type, data_type, 'rest',
                            ''.join(c for c in vae_type if not c.isdigit()))
        os.makedirs(base, exist_ok=True)
        if 'vanilla' in vae_type:
            fname = f'{vae_type}_rmse_{not_miwae_type}_large_batch_test.pt'
        else:
            fname = f'{vae_type}_rmse_{alpha}_{p_missingness}_{reg_type}_full_reg_large_batch_v2_test.pt'
        torch.save(recon, os.path.join(base, fname))
        return recon


def R_lindley_chain(i, x, mask, M, vae, im, loc):
    """Reward of candidate `i` for the rows `loc` (evaluate.py:514-542).  Kept for API compatibility; it
    evaluates the all-candidates kernel on the selected rows and returns column `i`."""
    loc_t = torch.as_tensor(np.asarray(loc), dtype=torch.long, device=x.device)
    if loc_t.numel() == 0:
        return torch.empty(0, device=x.device)
    xs, base = x[loc_t].float(), mask[loc_t].float()
    R = ops.reward_chain_op(vae.flat_theta().detach(), xs, base, im[:, loc_t].contiguous().float(), vae.FAMILY,
                            vae.obs_dim, vae._emb())
    return R[:, i]


def active_learning_func(data_loader_train, test_data, test_mask, missing_rate, obs_dim, hid_dim, K, M, latent_dim,
                         data_type, training_parameters, experiment_type, vae_type, max_epochs, valid_k,
                         num_estimates, device=torch.device('cpu'), alpha=1.0, stage='evaluate', p_missingness=30,
                         reg_type='ml_reg', beta=1.0, beta_annealing=False, alpha_annealing=True, Repeat=5):
    device = torch.device(device)
    world_size, rank, group = world()
    n_test = test_data.shape[0]
    lo, hi = row_block(n_test, world_size, rank)                           # this rank's row block
    n_loc = hi - lo
    C = obs_dim - 1
    reward_events = []                                                   # (start, stop, triples) of every reward call
    info = torch.zeros(Repeat, n_test, obs_dim)
    action = torch.zeros(Repeat, n_test, C)
    R_hist = torch.zeros(Repeat, C, n_test, C)
    im_hist = torch.zeros(Repeat, C, M, n_test, obs_dim)
    with torch.no_grad():
        for r in range(Repeat):
            model = model_loader('test', obs_dim, hid_dim, K, latent_dim, missing_rate, data_type,
                                 training_parameters, max_epochs, valid_k, num_estimates, experiment_type, reg_type,
                                 vae_type, alpha=alpha, p_missingness=p_missingness, alpha_annealing=alpha_annealing)
            model.to(device)
            # PCVAE_MODE=throughput: Philox noise on the device, no host draws (statistically, not bitwise, the same loop)
            model.noise = 'device' if os.environ.get('PCVAE_MODE', 'parity') == 'throughput' else 'host'
            regularised = 'reg' in vae_type
            theta = model.flat_theta()
            eng = ops.engine(model.FAMILY, obs_dim, model._emb(), device)
            create_missing_uci(test_data.shape, p_missingness)       # evaluate.py:351 (feeds the unused p branch)
            x = test_data[lo:hi].float().to(device)
            target = x[:, -1].clone()
            mask = torch.zeros(n_loc, obs_dim, device=device)        # float mask; target column never observed
            ws = None

            def sample_means():
                """M x model.forward(x, mask, mask_p)[x_mean_q] (evaluate.py:365-386, 394-415).  The noise is drawn as the
                reference draws it -- per sample a full-size host draw, sliced to this rank's rows, the p-branch draw made
                and discarded -- but the M encoder / decoder passes go to the GPU as ONE launch each over the M stacked
                copies of the rows (rows are independent, so the results are the same bit for bit)."""
                eps = []
                for _ in range(M):
                    eps.append(draw_noise(n_test, latent_dim, device, model.noise)[lo:hi])
                    if regularised:
                        draw_noise(n_test, latent_dim, device, model.noise)
                outs = []
                group = max(1, min(M, (1 << 20) // max(n_loc, 1)))                 # <= ~1M stacked rows per launch
                for m0 in range(0, M, group):
                    k = min(group, M - m0)
                    e = torch.cat(eps[m0:m0 + k]).contiguous()
                    _, _, z, _ = eng.enc_fwd(theta, x.repeat(k, 1), [mask.repeat(k, 1)], [e])
                    outs.append(eng.dec(L.DEC_FWD, theta, z)["xhat"][0].view(k, n_loc, obs_dim))
                return torch.cat(outs, 0)

            def target_mse(im):
                # [M] sums of squared errors over this rank's rows, accumulated in float64 so that the total does not
                # depend on how the rows are split over the ranks (fp32 squares add exactly enough in fp64)
                se = ((im[:, :, -1] - target.unsqueeze(0)) ** 2).double().sum(1)
                if world_size > 1:
                    all_reduce_sum(se, group)
                return (se / n_test).float().mean()                               # mean over rows, then over M

            # The histories of this repeat stay on the device while the loop runs and go to the host once at its end (one
            # transfer per repeat instead of three synchronising ones per acquisition step; im alone is M * rows * obs_dim
            # floats per step): unless they would not fit comfortably, then they are copied step by step as before.
            dev_hist = C * M * n_loc * obs_dim * 4 <= (1 << 30)
            if dev_hist:
                im_dev = torch.empty(C, M, n_loc, obs_dim, device=device)
                R_dev = torch.empty(C, n_loc, C, device=device)
                act_dev = torch.empty(n_loc, C, device=device)
                info_dev = torch.empty(C + 1, device=device)
            first = target_mse(sample_means())
            if dev_hist:
                info_dev[0] = first
            else:
                info[r, :, 0] = first.cpu()
            for t in range(C):
                print("Repeat = {:.1f}".format(r))
                print("Strategy = {:.1f}".format(2))
                print("Step = {:.1f}".format(t))
                im = sample_means()
                if dev_hist:
                    im_dev[t].copy_(im)
                else:
                    im_hist[r, t, :, lo:hi] = im.cpu()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record()
                R, ws = eng.reward(theta, x, mask, im, ws)
                ev1.record()
                reward_events.append((ev0, ev1, (mask[:, :C] == 0).sum() * M))      # counted on the device: no sync per step
                if model.noise == 'host':
                    # the reference's chaini_I / chaini_II call encoder(sample=True): 4*M discarded [|loc|, L]
                    # draws per candidate (evaluate.py:562-626); burn them so the next `im` sees the same RNG state
                    unsel = (mask[:, :C] == 0).sum(0)
                    if world_size > 1:
                        all_reduce_sum(unsel, group)
                    for cnt in unsel.cpu().tolist():
                        for _ in range(4 * M):
                            torch.empty(int(cnt), latent_dim).normal_()
                i_opt = R.argmax(dim=1)
                if dev_hist:
                    R_dev[t].copy_(R)
                    act_dev[:, t] = i_opt.float()
                else:
                    R_hist[r, t, lo:hi] = R.cpu()
                    action[r, lo:hi, t] = i_opt.float().cpu()
                mask = mask + torch.eye(obs_dim, device=device)[i_opt]
                nxt = target_mse(sample_means())
                if dev_hist:
                    info_dev[t + 1] = nxt
                else:
                    info[r, :, t + 1] = nxt.cpu()
            if dev_hist:
                im_hist[r, :, :, lo:hi].copy_(im_dev)           # straight into the history (no intermediate host tensor)
                R_hist[r, :, lo:hi].copy_(R_dev)
                action[r, lo:hi].copy_(act_dev)
                info[r] = info_dev.cpu().unsqueeze(0).expand(n_test, C + 1)
                del im_dev, R_dev, act_dev, info_dev
    torch.cuda.synchronize(device)
    LAST_TIMING.update(reward_ms=sum(a.elapsed_time(b) for a, b, _ in reward_events),
                       triples=sum(int(n) for _, _, n in reward_events), reward_calls=len(reward_events))
    if world_size > 1:
        for tns in (action, R_hist, im_hist):
            merge_row_blocks(tns, group, device)                       # row blocks are disjoint: sum == gather
    if rank == 0:
        stems = {'information_curve_CHAI': info, 'action_CHAI': action, 'R_hist_CHAI': R_hist, 'im_CHAI': im_hist}
        for name, tns in stems.items():
            if 'vanilla' in vae_type:
                sep = '_' if name == 'information_curve_CHAI' else '__'
                fname = f'{vae_type}_{missing_rate}_missing_rate{sep}UCI_{name}_default_test.pt'
            else:
                fname = (f'{vae_type}_UCI_{name}_{alpha}_{p_missingness}_{reg_type}_{missing_rate}'
                         '_missing_rate_default_full_reg_test.pt')
            base = os.path.join('experiments', experiment_type, data_type, 'rest', _family_dir(vae_type))
            os.makedirs(base, exist_ok=True)
            torch.save(tns, os.path.join(base, fname))
    return info, action, R_hist

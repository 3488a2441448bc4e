"""Mirror of the reference's `src/experiment_main/train.py` (same `train(...)` signature,
positional order and checkpoint side effect), re-supplied because the reference version
mixes CPU and device tensors (train.py:54-58) and has no data-parallel hook.

Two execution paths, both through libpcvae_b200.so:
 * fused  -- kl_reg / vanilla training without beta annealing (everything the drivers
             select): FusedTrainer = encoder fwd -> decoder+loss+decoder bwd -> encoder bwd ->
             deterministic gradient reduce -> [NCCL all-reduce] -> fused Adam, flat vectors.
 * module -- anything else (ml_reg, beta annealing): the nn.Module API with autograd and
             torch.optim.Adam, exactly the reference's loop.

RNG parity: per epoch the DataLoader iterator draws its base seed and the
RandomSampler its permutation seed from torch's global generator; per regularised step
`np.random.rand(B, D)` (sub-mask) then two `[B, L]` normal draws (q, then p).  In parity mode
(`noise='host'`, the default) all of these are drawn on the host in that order.  Setting the
environment variable PCVAE_MODE=throughput switches to device-side gather + Philox draws.
"""
import os

import torch
from torch import optim
from tqdm import tqdm

from . import kernels as KR
from . import lib as L
from .dist import row_block, world
from .loaders import checkpoint_path, model_loader
from .utils import create_missing_uci, create_missing_uci_drop_eddi
from .VAE import draw_noise, fill_normal_
from .graphed import GraphedTrainer


def _family_dir(vae_type):
    return ''.join(c for c in '_'.join(vae_type.split('_')[:2]) if not c.isdigit())


def _save(model, experiment_type, data_type, vae_type, missing_rate, alpha, p_missingness, reg_type):
    path = checkpoint_path(experiment_type, data_type, vae_type, missing_rate, alpha, p_missingness, reg_type,
                           _family_dir(vae_type))
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, path)
    return path


def _epoch_batches(loader, device, throughput, prep=None, world_size=1, rank=0):
    """Yield (x, mask, global_rows) batches on `device` in the order the reference's DataLoader would; under data
    parallelism every rank gets ITS row block of each global batch (dist.row_block) and `global_rows` is the size of the
    whole batch.  In throughput mode the rank slices the sampler's index list first and gathers only its own rows; with
    `prep` = dict(keep=..., n_eps=..., step=...) and a shape pcvae_prep_batch takes, the sub-mask and the noise of the
    step are drawn by the same launch and left in prep['mask_p'] / prep['eps']."""
    table = getattr(loader, 'pcvae_table', None)
    if not throughput or table is None or not table[0].is_cuda:
        for data_sample, mask in loader:
            n = data_sample.shape[0]
            lo, hi = row_block(n, world_size, rank)
            yield data_sample[lo:hi].to(device), mask[lo:hi].to(device), n
        return
    # same RNG consumption as iter(DataLoader): base seed first (dataloader.py), then the sampler's own seed
    torch.empty((), dtype=torch.int64).random_()
    lib = L.load()
    data, mask = table
    kind = L.MASK_U8 if mask.dtype in (torch.bool, torch.uint8) else L.MASK_F32
    mask_src = mask if kind == L.MASK_U8 else mask.float()
    D = data.shape[1]
    for idx in loader.batch_sampler:
        n = len(idx)
        lo, hi = row_block(n, world_size, rank)
        idx = torch.as_tensor(idx[lo:hi], dtype=torch.int64).to(device, non_blocking=True)
        B = idx.numel()
        x = torch.empty(B, D, device=device)
        m = torch.empty(B, D, device=device, dtype=mask_src.dtype)
        if prep is not None and kind == L.MASK_U8 and D % 4 == 0 and D <= 128:
            prep['step'] += 1
            mp = torch.empty_like(m)
            eps = torch.empty(prep['n_eps'], B, 10, device=device)
            with torch.cuda.device(device):
                L.check(lib.pcvae_prep_batch(data.data_ptr(), mask_src.data_ptr(), idx.data_ptr(), x.data_ptr(),
                                             m.data_ptr(), mp.data_ptr(), eps.data_ptr(), B, D, prep['n_eps'],
                                             prep['keep'], 0xC0FFEE + rank, prep['step'] * 8,
                                             torch.cuda.current_stream().cuda_stream), "pcvae_prep_batch")
            prep['mask_p'], prep['eps'] = mp, eps
            yield x, m, n
            continue
        with torch.cuda.device(device):
            L.check(lib.pcvae_gather_rows(data.data_ptr(), mask_src.data_ptr(), idx.data_ptr(), x.data_ptr(),
                                          m.data_ptr(), B, D, kind, torch.cuda.current_stream().cuda_stream),
                    "pcvae_gather_rows")
        yield x, m, n


def _graph_epoch(trainer, loader, device, regularised, latent_dim, world_size=1, rank=0):
    """One epoch through GraphedFusedTrainer: the sampler's full batches are uploaded as one index table and replayed
    from the CUDA graph, a ragged last batch is gathered and stepped eagerly.  Same RNG consumption on the host as
    iter(DataLoader) (base seed, then the sampler's permutation).  Data parallel: every rank holds the same
    permutation (train() checks the ranks' RNG states) and takes its row block of every batch."""
    torch.empty((), dtype=torch.int64).random_()
    batches = list(loader.batch_sampler)
    Bg = trainer.B * world_size                              # rows of a full GLOBAL batch
    lo, hi = row_block(Bg, world_size, rank)
    full = [b[lo:hi] for b in batches if len(b) == Bg]
    ragged = [b for b in batches if len(b) != Bg]
    trainer.reset_total()
    trainer.set_batches(torch.as_tensor(full, dtype=torch.int64))
    done = 0
    if trainer.graph is None:
        done = max(1, min(3, len(full)))
        trainer.capture(warmup=done)
    for _ in range(len(full) - done):
        trainer.step_graph()
    total = trainer.total.clone()
    lib = L.load()
    data, mask = trainer.table, trainer.mtable
    for b in ragged:                                     # at most one
        rlo, rhi = row_block(len(b), world_size, rank)
        if rhi == rlo:
            raise L.PcvaeError(f"train(): the last batch of the epoch has {len(b)} rows for {world_size} ranks")
        idx = torch.as_tensor(b[rlo:rhi], dtype=torch.int64).to(device)
        n = idx.numel()
        x = torch.empty(n, data.shape[1], device=device)
        m = torch.empty(n, data.shape[1], device=device, dtype=mask.dtype)
        mp = torch.empty_like(m)
        eps = torch.empty(2 if regularised else 1, n, latent_dim, device=device)
        with torch.cuda.device(device):
            L.check(lib.pcvae_prep_batch(data.data_ptr(), mask.data_ptr(), idx.data_ptr(), x.data_ptr(), m.data_ptr(),
                                         mp.data_ptr(), eps.data_ptr(), n, data.shape[1], eps.shape[0], trainer.keep,
                                         trainer.seed, 8 * trainer.step_count, torch.cuda.current_stream().cuda_stream),
                    "pcvae_prep_batch")
        total += trainer.step(x, m, mp if regularised else None, eps[0], eps[1] if regularised else None,
                              global_rows=len(b))
        trainer.sync_counter()
    return total


def _device_submask(mask, keep, step):
    out = torch.empty_like(mask)
    lib = L.load()
    with torch.cuda.device(mask.device):
        L.check(lib.pcvae_draw_submask(mask.data_ptr(), out.data_ptr(), mask.numel(), keep, 0xC0FFEE, step * 8,
                                       torch.cuda.current_stream().cuda_stream), "pcvae_draw_submask")
    return out


def train(data_loader_train, missing_rate, obs_dim, hid_dim, K, M, latent_dim, data_type,
          training_parameters, experiment_type, vae_type, train_k, num_estimates, max_epochs=1000,
          device=torch.device('cpu'), alpha=1.0, stage='train', p_missingness=30, reg_type='ml_reg', beta=1.0,
          beta_annealing=False, alpha_annealing=True, not_miwae_type='changed'):
    device = torch.device(device)
    if device.type != 'cuda':
        raise L.PcvaeError("train(): this implementation runs on a CUDA device only (no CPU fallback); "
                           "the reference's own CPU path is the oracle")
    model = model_loader('train', obs_dim, hid_dim, K, latent_dim, missing_rate, data_type,
                         training_parameters, max_epochs, train_k, num_estimates, experiment_type, reg_type, vae_type,
                         alpha=alpha, p_missingness=p_missingness)
    model.to(device)
    if 'notMIWAE' not in vae_type:
        data_loader_train, _ = data_loader_train
    throughput = os.environ.get('PCVAE_MODE', 'parity') == 'throughput'
    model.noise = 'device' if throughput else 'host'
    regularised = 'reg' in vae_type
    # the Student-t family: everything model_loader's last two branches build (loaders.py:135-147, 234-245)
    miwae = 'MIWAE' in vae_type and 'notMIWAE' not in vae_type
    fused = ('MIWAE' not in vae_type) and (not beta_annealing) and (not regularised or reg_type == 'kl_reg')
    world_size, rank, group = world()
    if world_size > 1:
        if not fused:
            # the module / autograd paths (not-MIWAE, MIWAE, ml_reg, beta annealing) have no gradient exchange: running
            # them on several ranks would be N-fold redundant work on replicas that only agree if seeded identically
            raise L.PcvaeError(f"train(): data parallelism is built for the fused partial-VAE step (kl_reg / vanilla); "
                               f"run vae_type {vae_type!r} on one GPU")
        _sync_replicas(model, group, device)

    keep = 1 - p_missingness / 100
    use_graph = False
    trainer = None
    if fused:
        theta = model.flat_theta().detach().clone()
        table = getattr(data_loader_train, 'pcvae_table', None)
        bs = getattr(data_loader_train, 'batch_size', None)
        # throughput mode: the whole step is replayed from a CUDA graph (GraphedFusedTrainer); the ragged last batch of an
        # epoch takes the eager launches.  At the reference's batch of 64 the step is launch-bound.  Data parallel: each
        # rank replays its own graph over its row block of every global batch, the tail of the step is the fused
        # reduce + NVLink exchange + Adam kernel (needs peer access; otherwise eager launches + NCCL all-reduce).
        use_graph = (throughput and table is not None and table[0].is_cuda and bs is not None and bs % world_size == 0
                     and table[1].dtype in (torch.bool, torch.uint8) and obs_dim % 4 == 0 and obs_dim <= 128
                     and model.FAMILY == L.FAMILY_MLP and 'with_drop' not in vae_type and table[0].shape[0] >= bs
                     and os.environ.get('PCVAE_GRAPH', '1') != '0'
                     and (world_size == 1 or os.environ.get('PCVAE_DP', 'peer') != 'nccl'))
        if use_graph:
            try:
                trainer = KR.GraphedFusedTrainer(model.FAMILY, obs_dim, model._emb(), theta, table[0], table[1],
                                                 bs // world_size, table[0].shape[0] // bs, keep=keep, seed=0xC0FFEE + rank,
                                                 regularised=regularised, alpha=float(alpha), beta_w=float(beta), lr=0.001,
                                                 dist_group=group if world_size > 1 else None, world_size=world_size,
                                                 global_rows=bs if world_size > 1 else None)
            except L.PcvaeError:
                if world_size == 1:
                    raise
                use_graph = False                            # no peer access (every rank agrees, PeerExchange.create_or_none)
        if not use_graph:
            trainer = KR.FusedTrainer(model.FAMILY, obs_dim, model._emb(), theta, regularised=regularised,
                                      alpha=float(alpha), beta_w=float(beta), lr=0.001, dist_group=group, world_size=world_size)
    else:
        # throughput mode: the launch-bound MNAR step is replayed from a CUDA graph (graphed.py)
        graphed = throughput and 'notMIWAE' in vae_type and not beta_annealing
        # the graph replays ONE fused multi-tensor Adam launch instead of the ~14 of the for-each implementation
        optimizer = (optim.Adam(model.parameters(), lr=0.001, capturable=True, fused=True) if graphed
                     else optim.Adam(model.parameters(), lr=0.001))
        if graphed:
            def make_fn():
                if regularised:
                    def fn(x, m, mp):
                        mean_p, logvar_p, xm_p, xlv_p, mean_q, logvar_q, xm_q, xlv_q = model.forward(x, m, mp, stage=stage)
                        return model.loss(x, xm_p, xlv_p, mean_p, logvar_p, xm_q, xlv_q, mean_q, logvar_q, m, mp, 1,
                                          beta_annealing=False, beta=beta, alpha=alpha, alpha_annealing=alpha_annealing,
                                          stage=stage)[1]
                else:
                    def fn(x, m):
                        mean_q, logvar_q, xm_q, xlv_q = model.forward(x, m)
                        return model.loss(x, xm_q, xlv_q, mean_q, logvar_q, 1, m, beta_annealing=False, beta=beta,
                                          stage=stage)[1]
                return fn
            graph_trainer = GraphedTrainer(model, make_fn, optimizer, fill_normal_)
    try:
        _train_epochs(locals())
    finally:
        if trainer is not None and getattr(trainer, 'xch', None) is not None:
            trainer.xch.close()                              # CUDA IPC handles and the exchange buffer of this rank

    if fused:
        with torch.no_grad():
            flat = KR.unflatten_params(trainer.theta, model.FAMILY, obs_dim, model._emb())
            params = dict(model.named_parameters())
            for k, v in flat.items():
                params[k].copy_(v.view_as(params[k]))
    if rank == 0:
        _save(model, experiment_type, data_type, vae_type, missing_rate, alpha, p_missingness, reg_type)
    print('Training is over!')
    return model


def _sync_replicas(model, group, device):
    """Data parallel start-up: every replica takes rank 0's freshly initialised parameters, and the ranks' host RNG
    states (torch CPU generator: DataLoader permutation and parity-mode noise; NumPy: sub-masks) must be identical --
    every rank slices the SAME global batch, which only holds when the launcher seeded all ranks alike.  The states are
    compared (not overwritten: rank 0's stream stays what a single-GPU run would see) and a mismatch raises."""
    import hashlib
    import numpy as np
    import torch.distributed as dist
    with torch.no_grad():
        for p_ in model.parameters():
            dist.broadcast(p_.data, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    h = hashlib.sha256(torch.get_rng_state().numpy().tobytes())
    st = np.random.get_state()
    h.update(np.asarray(st[1]).tobytes() + str(st[2:]).encode())
    mine = torch.tensor(list(h.digest()[:8]), dtype=torch.int64, device=device)
    lo, hi = mine.clone(), mine.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    if not torch.equal(lo, hi):
        raise L.PcvaeError("train(): the ranks' host RNG states differ.  Data-parallel training slices ONE global batch "
                           "over the ranks, so every rank must be seeded identically (torch.manual_seed, np.random.seed) "
                           "before train() is called")


def _train_epochs(v):
    """The epoch loop of train() (reference train.py:23-117); `v` is train()'s local namespace."""
    model, trainer, device, vae_type = v['model'], v['trainer'], v['device'], v['vae_type']
    loader, throughput, fused, regularised, miwae = v['data_loader_train'], v['throughput'], v['fused'], v['regularised'], v['miwae']
    world_size, rank, group, use_graph = v['world_size'], v['rank'], v['group'], v['use_graph']
    keep, latent_dim, max_epochs, p_missingness = v['keep'], v['latent_dim'], v['max_epochs'], v['p_missingness']
    beta, beta_annealing, alpha, alpha_annealing, stage = v['beta'], v['beta_annealing'], v['alpha'], v['alpha_annealing'], v['stage']
    graphed = v.get('graphed', False)
    optimizer, graph_trainer = v.get('optimizer'), v.get('graph_trainer')
    step = 0
    # throughput mode, fused regularised step: gather + sub-mask + noise in one launch (pcvae_prep_batch)
    prep = dict(keep=keep, n_eps=2, step=0) if (throughput and fused and regularised and latent_dim == 10
                                                and 'with_drop' not in vae_type) else None
    for i in tqdm(range(max_epochs)):
        if use_graph:
            total = _graph_epoch(trainer, loader, device, regularised, latent_dim, world_size, rank)
        else:
            total = torch.zeros((), device=device, dtype=torch.float64)
            for data_sample, mask, n in _epoch_batches(loader, device, throughput, prep, world_size, rank):
                step += 1
                lo, hi = row_block(n, world_size, rank)              # this rank's rows of the global batch of n rows
                sl = slice(lo, hi)
                shape = (n, data_sample.shape[1])
                mask_p = None
                prepped = prep.pop('mask_p', None) if prep is not None else None
                if prepped is not None:
                    mask_p, eps_pair = prepped, prep.pop('eps')
                    total += trainer.step(data_sample, mask, mask_p, eps_pair[0], eps_pair[1], global_rows=n)
                    continue
                # host draws are made for the WHOLE global batch on every rank (the same RNG consumption as one GPU), then
                # sliced; device draws are made for the local rows only
                if 'with_drop' in vae_type:
                    mask_drop = create_missing_uci_drop_eddi(shape)[sl].to(device)
                else:
                    if regularised:
                        if throughput and mask.dtype in (torch.bool, torch.uint8):
                            mask_p = _device_submask(mask, keep, step)
                        elif throughput:
                            mask_p = (torch.rand(data_sample.shape, device=device) < keep).to(mask.dtype) * mask
                        else:
                            temp_mask = create_missing_uci(shape, p_missingness)              # host, NumPy RNG
                            mask_p = temp_mask[sl].to(device) * mask
                    mask_drop = torch.ones(data_sample.shape, device=device)
                if fused:
                    def noise():
                        if model.noise == 'host':
                            return draw_noise(n, latent_dim, device, 'host')[sl]
                        return draw_noise(hi - lo, latent_dim, device, model.noise)
                    if regularised:
                        eps_q = noise()
                        eps_p = noise()
                        loss = trainer.step(data_sample, mask, mask_p, eps_q, eps_p, global_rows=n)
                    else:
                        eps_q = noise()
                        mk = (mask * mask_drop)                                                   # float32, train.py:97
                        loss = trainer.step(data_sample, mk, None, eps_q, None, global_rows=n)
                    total += loss
                elif graphed:
                    if regularised:
                        total += graph_trainer.step(data_sample, mask, mask_p)
                    else:
                        total += graph_trainer.step(data_sample, mask * mask_drop)
                elif miwae:
                    if regularised:                                           # train.py:102-108
                        (mean_p, scale_p, x_mean_p, x_scale_p, deg_free_p, mean_q, scale_q, x_mean_q, x_scale_q,
                         deg_free_q) = model.forward(data_sample, mask, mask_p)
                        print_loss, train_loss = model.loss(data_sample, x_mean_p, x_scale_p, deg_free_p, mean_p, scale_p,
                                                            x_mean_q, x_scale_q, deg_free_q, mean_q, scale_q, mask, mask_p,
                                                            i + 1, beta_annealing=beta_annealing, beta=beta, alpha=alpha)
                    else:                                                     # train.py:109-113
                        mean, scale, x_mean, x_scale, deg_free = model.forward(data_sample, mask)
                        print_loss, train_loss = model.loss(data_sample, x_mean, x_scale, deg_free, mean, scale, mask, i + 1)
                    optimizer.zero_grad()
                    train_loss.backward()
                    optimizer.step()
                    total += train_loss.detach()
                else:
                    if regularised:
                        out = model.forward(data_sample, mask, mask_p, stage=stage)
                        mean_p, logvar_p, x_mean_p, x_logvar_p, mean_q, logvar_q, x_mean_q, x_logvar_q = out
                        print_loss, train_loss = model.loss(
                            data_sample, x_mean_p, x_logvar_p, mean_p, logvar_p, x_mean_q, x_logvar_q, mean_q, logvar_q,
                            mask, mask_p, i + 1, beta_annealing=beta_annealing, beta=beta, alpha=alpha,
                            alpha_annealing=alpha_annealing, stage=stage)
                    else:
                        mean_q, logvar_q, x_mean_q, x_logvar_q = model.forward(data_sample, mask * mask_drop)
                        print_loss, train_loss = model.loss(data_sample, x_mean_q, x_logvar_q, mean_q, logvar_q, i + 1,
                                                            mask * mask_drop, beta_annealing=beta_annealing, beta=beta,
                                                            stage=stage)
                    optimizer.zero_grad()
                    train_loss.backward()
                    optimizer.step()
                    total += train_loss.detach()
        if world_size > 1 and fused:
            torch.distributed.all_reduce(total, group=group)
        # float(total) synchronises the stream; the data-parallel exchange's bounded wait reports here (a late rank: no
        # Adam update was applied on the waiting ranks, pcvae_dp.cu) instead of letting the replicas drift apart
        epoch_total = float(total)
        if trainer is not None and getattr(trainer, 'xch', None) is not None:
            trainer.xch.check()
        tqdm.write('Epoch: [{}/{}], Total Loss: {}'.format(i, max_epochs, epoch_total))

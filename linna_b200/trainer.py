"""Host side of emulator training: ``Predictor.train``'s epoch loop and heuristics
(``linna/predictor_gpu.py:201-449``), ``train_nn`` (``linna/util.py:1272-1306``) and ``train_NN``
(``linna/util.py:1315-1472``: file loading, clipping, normalisation statistics, pickled
transforms).  The per-step work is ``train.FusedTrainer.step``; the whole training and validation
sets live on the GPU, batches are gathered by index on the device, losses are read back once per
epoch.
"""
import os

import numpy as np
import torch

from . import nnutils
from .train import FusedTrainer


def _find_lr(pred, X, Y, cmd, aux, batch, start_lr=1e-4, end_lr=5e-3, num_iter=100):
    """LR range test (the reference delegates to torch_lr_finder, linna/predictor_gpu.py:223-238):
    exponential sweep start_lr -> end_lr with AdamW(wd=1e-4) on a copy of the model, smoothed loss,
    learning rate at the steepest descent."""
    import copy
    scratch = copy.deepcopy(pred.model)
    tr = FusedTrainer(scratch, pred.X_transform, pred.y_transform, aux, batch, lr=start_lr, weight_decay=1e-4)
    n = X.shape[0]
    lrs, losses = [], []
    gamma = (end_lr / start_lr) ** (1.0 / max(num_iter - 1, 1))
    gen = torch.Generator(device="cpu").manual_seed(0)
    smooth, best = None, None
    for it in range(num_iter):
        idx = torch.randint(0, n, (min(batch, n),), generator=gen).to(X.device)
        tr.lr = start_lr * gamma ** it
        loss = float(tr.step(X[idx], Y[idx], cmd[idx]).item())
        smooth = loss if smooth is None else 0.05 * loss + 0.95 * smooth
        lrs.append(tr.lr), losses.append(smooth)
        best = smooth if best is None else min(best, smooth)
        if not np.isfinite(smooth) or smooth > 5 * best:
            break
    tr.engine.close()
    if len(losses) < 3:
        return start_lr
    lr = lrs[int(np.gradient(np.array(losses)).argmin())]
    return lr / 1e2 if lr > 1e0 else lr


def _as_device_set(loader, device):
    """(X, Y) of a DataLoader's underlying ArrayDataset on the GPU, or None when it is something else."""
    ds = getattr(loader, "dataset", None)
    if ds is not None and hasattr(ds, "X") and hasattr(ds, "y"):
        return (torch.from_numpy(np.ascontiguousarray(ds.X, np.float32)).to(device),
                torch.from_numpy(np.ascontiguousarray(ds.y, np.float32)).to(device))
    xs, ys = [], []
    for X, y in loader:
        xs.append(torch.as_tensor(X, dtype=torch.float32)), ys.append(torch.as_tensor(y, dtype=torch.float32))
    return torch.cat(xs).to(device), torch.cat(ys).to(device)


def run_training(pred, dataset, num_epochs, loss_fn, val_dataset=None, val_metric_fn=None, initfrombest=False,
                 pool=None, nocpu=False, rank=0, size=1):
    """``Predictor.train``.  Returns (train_losses[, val_metrics]) as numpy arrays like the reference."""
    from .predictor_gpu import EarlyStopping
    from tqdm.auto import tqdm
    if not torch.cuda.is_available():
        raise RuntimeError("Predictor.train: no CUDA device -- linna_b200 has no CPU fallback")
    torch.manual_seed(1234)                                              # predictor_gpu.py:221
    dev = torch.device("cuda", torch.cuda.current_device())
    aux = getattr(loss_fn, "auxileryfunction", None)
    if aux is None:
        raise TypeError("Predictor.train needs a linna.util.Loss_fn (its constants feed the fused loss kernel)")
    X, Y = _as_device_set(dataset, dev)
    batch = int(getattr(dataset, "batch_size", None) or X.shape[0])
    drop_last = bool(getattr(dataset, "drop_last", False))
    dist_on = size > 1 and torch.distributed.is_available() and torch.distributed.is_initialized()
    world = torch.distributed.get_world_size() if dist_on else 1
    wrank = torch.distributed.get_rank() if dist_on else 0
    # Data-parallel training (one process per GPU under torchrun): every rank runs this function with the same data set
    # and the same seeds.  The rank that writes files / prints is the torch.distributed rank -- the reference's `rank`
    # argument is never forwarded by train_NN / train_nn (linna/util.py:1287), so it cannot be trusted here.
    rank = wrank if dist_on else rank

    outdir = pred.outdir
    lr_path = os.path.join(outdir, "lr.npy") if outdir is not None else None
    probe = FusedTrainer(pred.model, pred.X_transform, pred.y_transform, aux, max(batch, 1), lr=1e-3)
    cmd = probe.chisq_md(X, Y)
    if isinstance(pred.optim, str) and pred.optim == "automatic":
        if lr_path is not None and os.path.isfile(lr_path):
            lr = float(np.load(lr_path))
        else:
            lr = _find_lr(pred, X, Y, cmd, aux, batch) if rank == 0 else None
            if dist_on:
                box = [lr]
                torch.distributed.broadcast_object_list(box, src=0)
                lr = box[0]
            if lr_path is not None and rank == 0:
                np.save(lr_path, lr)
            if dist_on:
                torch.distributed.barrier()
    elif isinstance(pred.optim, torch.optim.Optimizer):
        lr = float(pred.optim.param_groups[0]["lr"])
    else:
        lr = 1e-3
    # predictor_gpu.py:246 scales lr by `size` because every DDP rank of the reference would take a FULL batch (an
    # effective batch of size x batch).  Here the rows of each batch are sharded over the ranks and the gradients
    # averaged: the effective batch -- and therefore the learning rate -- is that of the single-GPU run.
    if not dist_on:
        lr = lr * size
    tr = probe
    tr.lr, tr.weight_decay = lr, 1e-4                                    # AdamW(lr, weight_decay=1E-4), :267
    tr.pg, tr.world = None, world
    if initfrombest:
        # weights only: the reference builds a fresh AdamW(lr, weight_decay=1e-4) after loading them (predictor_gpu.py:247-267)
        if not _load_best(pred, tr, with_optimizer=False):
            print("best.pth.tar does not exsit")

    have_val = val_dataset is not None
    if have_val:
        Xv, Yv = _as_device_set(val_dataset, dev)
        cmd_v = tr.chisq_md(Xv, Yv)
    train_losses, val_metrics = [], []
    es = EarlyStopping(patience=500)
    pbar = tqdm(range(num_epochs)) if rank == 0 else range(num_epochs)
    n = X.shape[0]
    nb = n // batch if drop_last else (n + batch - 1) // batch
    nb = max(nb, 1)
    losses_dev = torch.zeros(nb, dtype=torch.float32, device=dev)
    old = told = 0.0
    is_best = False
    ckpt_every = int(os.environ.get("LINNA_CHECKPOINT_EVERY", "10"))
    gen = torch.Generator(device="cpu")
    gen.manual_seed(int(torch.empty((), dtype=torch.int64).random_().item()))
    for i in pbar:
        perm = torch.randperm(n, generator=gen).to(dev)                   # DataLoader(shuffle=True)
        for b in range(nb):
            idx = perm[b * batch:(b + 1) * batch]
            if world > 1:                                                # data-parallel: this rank's rows of the batch
                idx = idx[wrank::world]
            tr.step(X[idx], Y[idx], cmd[idx], loss_out=losses_dev[b:b + 1])
        if world > 1:     # the heuristics below must see the SAME numbers on every rank: the global batch means
            torch.distributed.all_reduce(losses_dev, op=torch.distributed.ReduceOp.AVG)
        ep_losses = losses_dev.cpu().numpy().astype(np.float64)          # ONE device->host read per epoch
        train_losses.extend(ep_losses.tolist())
        loss = float(ep_losses[-1])
        if have_val:
            vm = tr.val_metric(Xv, Yv, cmd_v).cpu().numpy().astype(np.float64)
            if val_metric_fn is not None and getattr(val_metric_fn, "auxileryfunction", None) is None:
                raise TypeError("val_metric_fn must be a linna.util.Val_metric_fn")
            val_metrics.append(vm)
            if rank == 0 and hasattr(pbar, "set_description"):
                pbar.set_description("Train/val Loss: {0:.5e}, {1:.5e}    Epoch".format(loss, vm[0]))
            is_best = False
            if outdir is not None:
                is_best = bool(vm[0] < pred.best_val_loss)
                if is_best:
                    pred.best_val_loss = vm[0]
            recent = np.array(val_metrics)[-10:, 0]
            stalled = (np.std(recent) < 0.01 * np.mean(recent)) and 10 <= i < 120 and i % 10 == 0
            if stalled:                                                  # :319-337 "bad trainning": start over
                print("bad trainning: {0}".format(i), flush=True)
                pred.model.init_weight()
                tr.load_from_module(), tr.reset_optimizer()
                if i > 10 and tr.lr > 2e-4 and tr.lr > 2e-6:
                    print("learning rate too large: {0}".format(tr.lr), flush=True)
                    tr.lr /= 2.0
            blown = (np.isnan(vm[0]) or vm[0] > 1e10 or (vm[0] - old > 5 * old and i != 0) or
                     (loss - told > 5 * told and i != 0))
            if blown:                                                    # :339-373 divergence: back to the best model
                if not _load_best(pred, tr):
                    pred.model.init_weight()
                    tr.load_from_module()
                tr.reset_optimizer()
                if (np.isnan(vm[0]) or vm[0] > 1e10 or vm[0] - old > 10 * old) and i > 10 and tr.lr > 2e-6:
                    print("learning rate too large: {0}".format(tr.lr), flush=True)
                    tr.lr /= 2.0
                if not np.isnan(vm[0]) and vm[0] - old > 5 * old:
                    val_metrics[-1][0] = old
            else:
                crit = es.step(val_metrics[-1][0], loss)
                if crit == 1:
                    if tr.lr > 2e-6:
                        print("\n learning rate too large: {0}\n".format(tr.lr), flush=True)
                        tr.lr /= 2.0
                        tr.weight_decay /= 2
                    else:
                        es.cooling = 0
                elif crit == 2:
                    # every rank sees the same losses and validation metrics (all-reduced gradients, identical val set),
                    # so every rank takes this branch in the same epoch: all of them leave the loop together and none is
                    # left waiting in the next all-reduce
                    if rank == 0:
                        print("early stop", flush=True)
                        print("learning rate", tr.lr, flush=True)
                        _checkpoint(pred, tr, i, is_best, force=True)
                    break
                elif crit == 3:
                    print("\n weight decay too small: {0}\n".format(tr.weight_decay), flush=True)
                    if tr.weight_decay < 1e0:
                        tr.weight_decay *= 2
            old, told = val_metrics[-1][0], loss
        force = (i % ckpt_every == 0 or i == num_epochs - 1)
        if outdir is not None and rank == 0:
            _checkpoint(pred, tr, i, is_best, force=force)
        if world > 1 and outdir is not None and (is_best or force):
            torch.distributed.barrier()      # best.pth.tar is complete before any rank may roll back to it
    tr.commit()
    pred._engine = None
    tr.engine.close()
    if have_val:
        return np.array(train_losses), np.array(val_metrics)
    return np.array(train_losses)


def _checkpoint(pred, tr, epoch, is_best, force):
    """``last.pth.tar`` (+ ``best.pth.tar``) in the reference layout (predictor_gpu.py:405-419).  The
    reference writes every epoch; here only when the model improved or every LINNA_CHECKPOINT_EVERY
    epochs, because one epoch is a few milliseconds of GPU work."""
    if not (is_best or force):
        return
    tr.sync_to_module()
    nnutils.save_checkpoint({"epoch": epoch + 1, "state_dict": pred.model.state_dict(),
                             "optim_dict": tr.optim_state_dict()}, is_best=is_best, checkpoint=pred.outdir)


def _load_best(pred, tr, with_optimizer=False):
    """Weights of best.pth.tar into the module and the device trainer.  The reference's re-initialisations all build a
    NEW AdamW(lr, weight_decay=1e-4) afterwards (predictor_gpu.py:267, :329, :358): the moments are not restored unless
    asked for."""
    path = os.path.join(pred.outdir, "best.pth.tar") if pred.outdir is not None else None
    if path is None or not os.path.isfile(path):
        return False
    ckpt = nnutils.load_checkpoint(path, pred.model, None, device="cpu")
    tr.load_from_module()
    if with_optimizer:
        try:
            tr.load_optim_state_dict(ckpt.get("optim_dict", {}))
        except Exception:
            tr.reset_optimizer()
    return True


def train_nn(outdir, model, train_x, train_y, val_x, val_y, X_transform, y_transform, loss_fn, val_metric_fn,
             dev="cpu", verbose=False, retrain=True, pool=None, nocpu=False, size=0, rank=0, params=None):
    """linna/util.py:1272-1306."""
    from . import predictor_gpu
    from .util import ArrayDataset
    from torch.utils.data import DataLoader
    if not retrain and os.path.isfile(os.path.join(outdir, "best.pth.tar")):
        return
    pred = predictor_gpu.Predictor(train_x.shape[-1], train_y.shape[-1], X_transform=X_transform,
                                   y_transform=y_transform, device=dev, optim="automatic", model=model,
                                   scheduler=None, outdir=outdir)
    train_loader = DataLoader(ArrayDataset(train_x, train_y), batch_size=params["batch_size"], shuffle=True,
                              drop_last=True, num_workers=0)
    val_loader = DataLoader(ArrayDataset(val_x, val_y), batch_size=len(val_y))
    pred.train(train_loader, params["num_epochs"], loss_fn, val_loader, val_metric_fn, initfrombest=True, pool=None,
               nocpu=nocpu, rank=rank, size=max(size, 1))
    return pred


def _load_sets(outdir_list):
    """Concatenate the training / validation sets of all iterations so far (util.py:1346-1373)."""
    def cat(name, loader):
        parts = []
        for d in outdir_list:
            a = loader(os.path.join(d, name))
            if len(a) > 1:
                parts.append(a)
        return np.concatenate(parts) if parts else np.array(parts)
    tx, ty = cat("train_samples_x.txt", np.loadtxt), cat("train_samples_y.npy", np.load)
    vx, vy = cat("val_samples_x.txt", np.loadtxt), cat("val_samples_y.npy", np.load)
    last = np.load(os.path.join(outdir_list[0], "train_samples_y.npy"))
    if len(last) == 0:
        last = ty
    return tx, ty, vx, vy, np.array(last)


def train_NN(nnsampler, cov, inv_cov, sigma, outdir_in, outdir_list, data, dolog10index=None, ypositive=False,
             retrain=True, norder=2, temperature=None, docuda=False, pool=None, tsize=1, nnmodel_in=None,
             params=None, usebest=False):
    """Build the normalisation statistics, pickle the transforms (on-disk layout of SURVEY 8b) and train
    the emulator.  linna/util.py:1315-1472; ``docuda`` is accepted and ignored -- training always runs
    on the GPU here."""
    from . import util as U
    device = "cpu"                      # where the pickled transform tensors live; compute is on the GPU
    # under torchrun (tsize > 1, one process per GPU) every rank computes the same statistics; one of them writes the files
    dist_on = torch.distributed.is_available() and torch.distributed.is_initialized()
    writer = (not dist_on) or torch.distributed.get_rank() == 0

    class _NoWrite:
        def __init__(self, obj):
            self._o = obj

        def __getattr__(self, name):
            return getattr(self._o, name)

        def __call__(self, *a, **k):
            return self._o(*a, **k)

        def pickle(self, path):
            if writer:
                self._o.pickle(path)
    inv_cov_tensor = torch.tensor(inv_cov, dtype=torch.float64)
    cov_tensor = torch.tensor(cov, dtype=torch.float64)
    y_transform_data = U.Y_transform_data(sigma, device=device)
    _NoWrite(y_transform_data).pickle(os.path.join(outdir_in, "y_transform_data.pkl"))
    y_invtransform_data = U.Y_invtransform_data(sigma, device=device)
    _NoWrite(y_invtransform_data).pickle(os.path.join(outdir_in, "y_invtransform_data.pkl"))
    data_tensor = torch.from_numpy(data.astype(np.float32)).clone().requires_grad_()

    train_x, train_y, val_x, val_y, train_y_last = _load_sets(outdir_list)
    if usebest:
        bx = [np.loadtxt(d + "best_samples_x.txt") for d in outdir_list]
        by = [np.load(d + "best_samples_y.npy") for d in outdir_list]
        bx, by = [a for a in bx if len(a) > 1], [a for a in by if len(a) > 1]
        if bx:
            bx, by = np.concatenate(bx), np.concatenate(by)
            if train_x.ndim > 1:
                train_x, train_y = np.concatenate([bx, train_x]), np.concatenate([by, train_y])
            else:
                train_x, train_y, train_y_last = bx, by, by
        vbx = np.concatenate([np.loadtxt(d + "best_samples_x_val.txt") for d in outdir_list])
        vby = np.concatenate([np.load(d + "best_samples_y_val.npy") for d in outdir_list])
        if vbx.ndim > 1:
            val_x = np.concatenate([vbx, val_x]) if val_x.ndim > 1 else vbx
            val_y = np.concatenate([vby, val_y]) if val_y.ndim > 1 else vby
    print(train_x.shape, train_y.shape, val_x.shape, val_y.shape)

    # outlier clipping (util.py:1404-1438)
    if ypositive:
        np.clip(train_y, 1e-30, 1e10, out=train_y)
        np.clip(val_y, 1e-30, 1e10, out=val_y)
        train_y_last[train_y_last < 1e-30] = 1e-30
        keep = np.mean(train_y, axis=1) != 1e-30
        train_x, train_y = train_x[keep], train_y[keep]
        train_y_last = train_y_last[np.mean(train_y_last, axis=1) != 1e-30]
        keep = np.mean(val_y, axis=1) != 1e-30
        val_x, val_y = val_x[keep], val_y[keep]
    else:
        np.clip(train_y, -1e5, 1e10, out=train_y)
        np.clip(val_y, -1e5, 1e8, out=val_y)
        np.clip(train_y_last, -1e5, 1e10, out=train_y_last)

    def log10_cols(X):
        X1 = torch.tensor(X, dtype=torch.float32)
        if dolog10index is not None:
            for ind in dolog10index:
                X1[:, ind] = torch.log10(X1[:, ind])
        return X1
    X_mean = log10_cols(train_x).mean(axis=0)
    X_std = log10_cols(train_x).std(axis=0)
    X_transform = U.X_transform_class(X_mean, X_std, device, dolog10index)
    _NoWrite(X_transform).pickle(os.path.join(outdir_in, "X_transform.pkl"))
    # y_mean = median(y / sigma), y_std = median |y / sigma - y_mean| (linna/util.py:1440-1450): radix selection on the
    # GPU (engine.column_median_mad) instead of two CPU sorts of the whole training set; same lower-median convention
    from . import engine as _eng
    ysrc = train_y if ypositive else train_y_last
    y_dev = torch.as_tensor(np.ascontiguousarray(ysrc, np.float32)).cuda()
    y_mean, y_std = _eng.column_median_mad(y_dev, np.asarray(sigma, np.float32), take_log=bool(ypositive))
    y_mean, y_std = y_mean.cpu(), y_std.cpu()
    del y_dev
    if not ypositive:
        y_std[y_std < 1e-10] = 1e0
    y_transform = U.Y_transform_class(y_mean, y_std, device, ypositive=ypositive)
    _NoWrite(y_transform).pickle(os.path.join(outdir_in, "y_transform.pkl"))
    y_inv_transform = U.Y_invtransform_class(y_mean, y_std, data_tensor, device, ypositive=ypositive)
    _NoWrite(y_inv_transform).pickle(os.path.join(outdir_in, "y_invtransform.pkl"))
    if dist_on:
        torch.distributed.barrier()

    loss_fn = U.Loss_fn(data_tensor, cov_tensor, inv_cov_tensor, y_transform_data, y_inv_transform, device)
    val_metric_fn = U.Val_metric_fn(data_tensor, cov_tensor, inv_cov_tensor, y_transform_data, y_inv_transform, device)
    nnmodel = nnmodel_in(len(train_x[0]), len(train_y[0]), None, docpu=False)
    nnsampler.model = nnmodel
    return train_nn(outdir_in, nnsampler.model, train_x, train_y, val_x, val_y, X_transform, y_transform, loss_fn,
                    val_metric_fn, dev=device, verbose=True, retrain=retrain, pool=pool, nocpu=True, size=tsize,
                    params=params)


train_nn.__module__ = "linna.util"
train_NN.__module__ = "linna.util"

"""Multi-rank worker for tests/test_gpu_multi.py, run under torchrun (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P \
        tests/dist/worker.py <task> <outfile.json>

  train     a 2-rank data-parallel FusedTrainer.step (rows of the batch sharded, ONE NCCL all-reduce of the flat gradient,
            AdamW on every rank) against the single-GPU step on the same global batch
  ensemble  the sharded ensemble sampler (sampler.HMCSampler.sample: sub-ensembles per rank, chains gathered at every
            convergence check, decision broadcast) on an analytic Gaussian
  hmc       HMCSampler.sample_chains with the chains sharded over the ranks, on the reference's fixture emulator
"""
import json
import os
import pickle
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def task_train(out):
    from linna_b200 import synthetic
    from linna_b200.train import FusedTrainer
    import linna.nn as N
    import linna.util as U
    rank, world = dist.get_rank(), dist.get_world_size()
    p = synthetic.make_problem(6, 8, seed=4)
    rng = np.random.default_rng(0)
    theta = synthetic.training_set(p, 256, seed=3, spread=0.5)
    A = rng.standard_normal((6, 8))
    target = np.tanh(theta @ A) * p.sigma + 0.3 * p.sigma
    p.data = target[0].copy()
    sig = np.asarray(p.sigma, np.float32)
    ytd = U.Y_transform_data(sig, "cpu")
    ymean = torch.tensor(np.median(target / sig, axis=0).astype(np.float32))
    ystd = torch.tensor((np.median(np.abs(target / sig - ymean.numpy()), axis=0)).astype(np.float32))
    yinv = U.Y_invtransform_class(ymean, ystd, torch.tensor(p.data.astype(np.float32)), "cpu")
    loss_fn = U.Loss_fn(torch.tensor(p.data.astype(np.float32)), torch.tensor(p.cov), torch.tensor(p.inv_cov), ytd, yinv, "cpu")
    xt = U.X_transform_class(torch.tensor(theta.mean(0).astype(np.float32)), torch.tensor(theta.std(0).astype(np.float32)), "cpu")
    yt = U.Y_transform_class(ymean, ystd, "cpu")
    X = torch.from_numpy(theta.astype(np.float32)).cuda()
    Y = torch.from_numpy(target.astype(np.float32)).cuda()
    res = {}
    for path in ("tc", "ffma"):
        torch.manual_seed(3)
        model = N.ChtoModelv2(6, 8, None)
        tr = FusedTrainer(model, xt, yt, loss_fn.auxileryfunction, 128, lr=2e-3, world_size=world)
        tr.engine.set_train_path(path)
        cmd = tr.chisq_md(X, Y)
        losses = []
        for s in range(4):
            idx = torch.arange(s * 64, s * 64 + 128, device="cuda") % 256       # the global batch of this step
            mine = idx[rank::world]                                            # this rank's rows (trainer.run_training)
            l = tr.step(X[mine], Y[mine], cmd[mine]).clone()
            dist.all_reduce(l, op=dist.ReduceOp.AVG)
            losses.append(float(l.item()))
        w_dp = tr.p.clone()
        gathered = [torch.zeros_like(w_dp) for _ in range(world)]
        dist.all_gather(gathered, w_dp)
        res[path + "_ranks_identical"] = bool(all(torch.equal(gathered[0], g) for g in gathered))
        res[path + "_peer_reduce"] = tr._peer is not None      # gradient average read from peer memory inside the AdamW kernel
        if rank == 0:
            torch.manual_seed(3)
            model1 = N.ChtoModelv2(6, 8, None)
            tr1 = FusedTrainer(model1, xt, yt, loss_fn.auxileryfunction, 128, lr=2e-3, world_size=1)
            tr1.engine.set_train_path(path)
            l1 = []
            for s in range(4):
                idx = torch.arange(s * 64, s * 64 + 128, device="cuda") % 256
                l1.append(float(tr1.step(X[idx], Y[idx], cmd[idx]).item()))
            d = (w_dp - tr1.p).abs()
            res[path] = {"max_abs_diff": float(d.max()), "frac_gt_1e-5": float((d > 1e-5).float().mean()),
                         "w_scale": float(tr1.p.abs().max()), "loss_dp": losses, "loss_1gpu": l1, "kernel": tr.kernel_path()}
        tr.engine.close()
    if rank == 0:
        with open(out, "w") as f:
            json.dump(res, f)


def task_ensemble(out):
    from linna_b200 import sampler
    rank, world = dist.get_rank(), dist.get_world_size()
    d, W = 4, 64
    rng = np.random.default_rng(0)
    A = rng.standard_normal((d, d))
    cov = A @ A.T / d + np.eye(d)
    icov = torch.from_numpy(np.linalg.inv(cov).astype(np.float32)).cuda()
    mu = torch.from_numpy(rng.standard_normal(d).astype(np.float32)).cuda()

    def lnp(x):
        r = x - mu
        return -0.5 * torch.einsum("ij,jk,ik->i", r, icov, r)
    np.random.seed(5)                      # the same x0 on every rank: each takes its own rows
    x0 = mu.cpu().numpy() + 0.1 * np.random.randn(W, d)
    outdir = os.path.dirname(out)
    s = sampler.HMCSampler(lnp, None, None, d, W, x0=x0)
    store = s.sample(None, 4000, outdir=outdir, overwrite=True, ntimes=30, tautol=0.05, meanshift=0.1, stdshift=0.1, nk=2)
    chain = np.asarray(store.chain)
    res = {"steps": int(chain.shape[0]), "walkers": int(chain.shape[1]), "mean": chain[200:].reshape(-1, d).mean(0).tolist(),
           "cov": np.cov(chain[200:].reshape(-1, d), rowvar=False).tolist(), "mu": mu.cpu().numpy().tolist(), "true_cov": cov.tolist(),
           "rank": rank, "world": world}
    # the two sub-ensembles are different chains (own RNG streams), and every rank reads the same stored chain
    res["halves_differ"] = bool(not np.allclose(chain[-1, :W // world], chain[-1, W // world:2 * (W // world)]))
    allres = [None] * world
    dist.all_gather_object(allres, (res["steps"], float(chain.sum())))
    res["ranks_agree"] = bool(all(a == allres[0] for a in allres))
    if rank == 0:
        with open(out, "w") as f:
            json.dump(res, f)


def task_hmc(out):
    import shutil
    import linna.util as U
    from linna.HMCSampler import HMCSampler
    rank, world = dist.get_rank(), dist.get_world_size()
    fix = os.path.join(ROOT, "tests", "golden", "ref_fixture_iter_0")
    d = os.path.join(os.path.dirname(out), "fix_rank%d" % rank)
    shutil.copytree(fix, d)
    pred, yinv = U.retrieve_model(d, 2, 2)
    with open(os.path.join(d, "model_args.pkl"), "rb") as f:
        args = pickle.load(f)
    priors = [dict(param="x%d" % i, dist="flat", arg1=-2.0, arg2=2.0) for i in range(2)]
    lp = U.Log_prob(np.asarray(args[6]), np.asarray(args[2]), pred, yinv, U.Transform(priors), 1.0, U.gaussianlogliklihood, nograd=False)
    C = 2048
    g = torch.Generator().manual_seed(5)
    x0 = 0.1 * torch.randn(C, 2, generator=g)
    samp = HMCSampler(lp, x0, torch.ones(2), device="cuda")
    xs, ls, acc = samp.sample_chains(60, 5, 0.15, seed=11)
    flat = xs[20:].reshape(-1, 2).cpu().numpy().astype(np.float64)
    res = {"shape": list(xs.shape), "mean": flat.mean(0).tolist(), "std": flat.std(0).tolist(), "acc": acc, "world": world,
           "first_half_mean": xs[20:, :C // 2].reshape(-1, 2).mean(0).cpu().numpy().tolist(),
           "second_half_mean": xs[20:, C // 2:].reshape(-1, 2).mean(0).cpu().numpy().tolist()}
    if rank == 0:
        with open(out, "w") as f:
            json.dump(res, f)


if __name__ == "__main__":
    task, out = sys.argv[1], sys.argv[2]
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        {"train": task_train, "ensemble": task_ensemble, "hmc": task_hmc}[task](out)
    finally:
        dist.barrier()
        dist.destroy_process_group()

"""Multi-GPU tests (need >= 2 GPUs on the box; skipped otherwise): the product's own multi-rank paths, one process per
GPU under torchrun with NCCL -- data-parallel training (the only collective of the design: ONE all-reduce of the flat
gradient per step) against the single-GPU step, and the sharded samplers (independent walkers / chains per rank, chains
gathered by NCCL) against the analytic posterior and the 1-rank run."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(task, tmp_path, nproc=2):
    if torch.cuda.device_count() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / (task + ".json"))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist", "worker.py"), task, out]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    with open(out) as f:
        return json.load(f)


def test_data_parallel_step_equals_single_gpu_step(tmp_path):
    res = _run("train", tmp_path)
    for path in ("tc", "ffma"):
        assert res[path + "_ranks_identical"]              # every rank holds the same weights after the reduction + AdamW
        assert res[path + "_peer_reduce"], "the NVLink peer-memory reduction (linna_train_adamw_peer) was not used"
        r = res[path]
        assert r["kernel"] == path
        # mean of the shard gradients == gradient of the global batch up to float32 summation order; AdamW can turn a
        # rounding-level difference of a near-zero gradient into ~lr in a single weight (tests/test_gpu_train.py)
        np.testing.assert_allclose(r["loss_dp"], r["loss_1gpu"], rtol=2e-4)
        assert r["frac_gt_1e-5"] < 0.02 and r["max_abs_diff"] < 8 * 2e-3, r


def test_sharded_ensemble_sampler(tmp_path):
    res = _run("ensemble", tmp_path)
    assert res["world"] == 2 and res["walkers"] == 64 and res["ranks_agree"] and res["halves_differ"]
    mu, cov = np.array(res["mu"]), np.array(res["true_cov"])
    sd = np.sqrt(np.diag(cov))
    assert np.all(np.abs(np.array(res["mean"]) - mu) < 0.15 * sd), (res["mean"], mu)
    assert np.all(np.abs(np.sqrt(np.diag(np.array(res["cov"]))) / sd - 1) < 0.12)


def test_sharded_hmc_chains(tmp_path):
    res = _run("hmc", tmp_path)
    assert res["world"] == 2 and res["shape"] == [60, 2048, 2] and 0.5 < res["acc"] <= 1.0
    # pooled moments of the two ranks' chains == the single-GPU moments of tests/test_gpu_api.py (oracle quadrature there)
    a, b = np.array(res["first_half_mean"]), np.array(res["second_half_mean"])
    assert np.all(np.abs(a - b) < 0.05) and not np.allclose(a, b)
    from linna_b200 import arch
    from oracle.oracle import Oracle
    from tests.helpers import fixture_problem, load_golden
    o = Oracle(fixture_problem(load_golden("fixture")), arch)
    ax = np.linspace(-4, 4, 161)
    U1, U2 = np.meshgrid(ax, ax, indexing="ij")
    grid = np.stack([U1.ravel(), U2.ravel()], 1)
    w = np.exp(o.lnp(grid, np.float64)["lnp"])
    w /= w.sum()
    mean = (grid * w[:, None]).sum(0)
    std = np.sqrt(((grid - mean) ** 2 * w[:, None]).sum(0))
    assert np.all(np.abs(np.array(res["mean"]) - mean) < 0.03) and np.all(np.abs(np.array(res["std"]) / std - 1) < 0.05)

"""The chain file the reference ships with its fixture, read by this package's own HDF5 reader: the numbers the
reference's test_reading asserts (tests/test_main.py:46-52 of chto/linna) must come out -- that pins the reader, the
restated emcee autocorrelation time and the cut of read_chain_and_cut (linna/util.py:68-94)."""
import os
import shutil

import numpy as np

from tests.helpers import GOLDEN

H5 = os.path.join(GOLDEN, "ref_chain", "chemcee_256.h5")


def test_h5_reader_sees_the_emcee_backend_layout():
    from linna_b200.h5read import H5File
    f = H5File(H5)
    assert f.keys("/") == ["mcmc"]
    assert f.keys("/mcmc") == ["accepted", "chain", "chain_transformed", "log_prob"]
    at = f.attrs("/mcmc")
    assert int(at["iteration"]) == 200 and int(at["nwalkers"]) == 4 and int(at["ndim"]) == 2
    chain = f.dataset("/mcmc/chain")
    assert chain.shape[1:] == (4, 2) and chain.dtype == np.float64
    acc = f.dataset("/mcmc/accepted")
    assert acc.shape == (4,) and np.all(acc > 100) and np.all(acc <= 200)
    assert np.all(chain[200:210] == 0)          # emcee grows ahead of what it has written


def test_read_chain_and_cut_reproduces_the_reference_test_reading(tmp_path):
    from linna_b200.sampler import ChainStore, integrated_time, read_chain_and_cut
    d = tmp_path / "iter_0"
    d.mkdir()
    shutil.copy(H5, d / "chemcee_256.h5")
    store = ChainStore(str(d / "chemcee_256.h5"))
    assert store.exists() and store.iteration == 200 and store.chain.shape == (200, 4, 2)
    tau = integrated_time(store.chain)
    np.testing.assert_allclose(tau, [17.38388933, 11.33458716], rtol=1e-7)      # emcee.autocorr.integrated_time, c = 5
    chain, lp, _ = read_chain_and_cut(str(d / "chemcee_256.h5"), 1, 2, method="emcee")     # nkeepArr = [1], ntimesArr = [2]
    assert chain.shape == (56, 2) and lp.shape == (14, 4)
    # the reference's own assertion, to its own 5 decimals (and in fact to the last bit)
    np.testing.assert_almost_equal(np.mean(chain), 0.15151080063411168, decimal=5)
    np.testing.assert_almost_equal(np.std(chain), 0.9633211647095377, decimal=5)
    assert abs(np.mean(chain) - 0.15151080063411168) < 1e-14 and abs(np.std(chain) - 0.9633211647095377) < 1e-14
    # the stored chain_transformed is Transform(chain) (linna/sampler.py:356): flat priors on [-2, 2]
    import linna.util as U
    tr = U.Transform([dict(param="x%d" % i, dist="flat", arg1=-2.0, arg2=2.0) for i in range(2)])
    np.testing.assert_allclose(np.asarray(tr(store.chain[-1].astype(np.float64))), store.chain_transformed[-1], atol=1e-6)

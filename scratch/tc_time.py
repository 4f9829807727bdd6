import sys, os, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from linna_b200 import engine, synthetic, arch
mode = sys.argv[1] if len(sys.argv) > 1 else "lnp"
ns = [int(a) for a in sys.argv[2:]] or [100000]
p = synthetic.make_problem(30, 500, seed=0)
e = engine.engine_from_problem(p, with_likelihood=False)
m0 = e.predict(np.asarray(p.theta0, np.float32)[None, :], engine.LINNA_OUT_M)[0]
p.set_data_from_prediction(m0)
e.set_likelihood(p.priors, np.asarray(p.data, np.float32), p.inv_cov, 1.0)
e.set_path("tc")
for n in ns:
    u = torch.from_numpy(synthetic.walkers(n, 30, scale=0.3, seed=1)).cuda()
    call = e.lnp if mode == "lnp" else e.lnp_grad
    for _ in range(3): call(u)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(5): call(u)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 5
    print("SEG", os.environ.get("LINNA_TC_SEG_KC"), mode, "n", n, "ms %.3f  evals/s %.4g" % (ms, n / ms * 1e3), flush=True)
if os.environ.get("LINNA_TC_DEBUG"):
    cnt = e.tc_counters()
    if len(cnt):
        m_ = cnt.mean(axis=0)
        print("producer: total %.0f  wait_empty %.0f (%.0f%%)  wait_ready %.0f (%.0f%%)" % (m_[0], m_[1], 100*m_[1]/m_[0], m_[2], 100*m_[2]/m_[0]))
        lead = cnt[cnt[:, 3] > 0]
        m_ = np.concatenate([m_[:3], lead.mean(axis=0)[3:6], m_[6:]])
        for gname, o in (("epi g0", 6), ("epi g1", 10)):
            print("%s  : total %.0f  wait_pfull %.0f (%.0f%%)  drain %.0f (%.0f%%)  chunk-epilogue %.0f (%.0f%%)" % (gname, m_[o], m_[o+1], 100*m_[o+1]/m_[o], m_[o+2], 100*m_[o+2]/m_[o], m_[o+3], 100*m_[o+3]/m_[o]) + "  of which wait_sfree %.0f" % m_[14 if o == 6 else 15])
        print("mma wait_full in the first %d stages of a layer pass: %.0f" % (4, lead.mean(axis=0)[14]))
        print("mma     : total %.0f  wait_full %.0f (%.0f%%)  wait_pempty %.0f (%.0f%%)" % (m_[3], m_[4], 100*m_[4]/m_[3], m_[5], 100*m_[5]/m_[3]))
        ns_ = 24
        f = lambda a: " ".join("%d" % v for v in a)
        print("per-step cycles, MMA warp (leader)       :", f(lead.mean(axis=0)[16:16+ns_]))
        print("per-step cycles, epilogue group 0        :", f(cnt.mean(axis=0)[40:40+ns_]))
        print("   of which waiting for accumulators     :", f(cnt.mean(axis=0)[64:64+ns_]))
        print("   of which chunk epilogues              :", f(cnt.mean(axis=0)[88:88+ns_]))

import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch, time
from linna_b200 import sampler
rng=np.random.default_rng(0)
x=rng.standard_normal((60,100000,8)).astype(np.float32)
t=time.time(); a=sampler.checkmeanstd(x,0.2,0.15); t1=time.time()-t
xd=torch.from_numpy(x).cuda(); torch.cuda.synchronize()
t=time.time(); b=sampler.checkmeanstd(xd,0.2,0.15); torch.cuda.synchronize(); t2=time.time()-t
print(a,b,"host-array call %.3f s, device-tensor call %.4f s"%(t1,t2))

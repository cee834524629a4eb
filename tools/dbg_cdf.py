import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from tempest_b200.steps import Kernels
dev=torch.device('cuda:0'); k=Kernels(dev)
n=1<<25
g=torch.Generator(device=dev).manual_seed(1)
chi=(torch.randn((n,10),dtype=torch.float64,device=dev,generator=g)**2).sum(1)
w=torch.exp(-0.5*chi*4.0*0.37); w/=w.sum()
cdf=k.cdf(w,n)
torch.cuda.synchronize()
ws=k.ws._buf['cdf_ws']
nt=(n+1023)//1024
al=lambda x:(x+255)//256*256
off=al(8*nt)*3
E=ws[off:off+4*nt].view(torch.int32).cpu().numpy()
hard=(E==-2**31)
print("nt",nt,"hard tiles",hard.sum(),"first hard idx",np.nonzero(hard)[0][:40], "distinct E", len(set(E[~hard])))
runs=np.sum(np.diff(E)!=0)
print("E changes",runs)
ref=np.cumsum(w.cpu().numpy()); print("exact", np.array_equal(ref, cdf.cpu().numpy()))

"""How far apart are two fp32 evaluations of the SAME DeepFM training steps?

Runs the numpy fp32 oracle (oracle/oracle.py, the restatement the GPU parity tests check against) twice on the
Criteo-shaped workload: once as is, once with the batch rows permuted (same samples, same arithmetic, another
summation order in the batch reductions).  With 256-wide batch-normed towers the two runs separate after one or two
updates - a ReLU mask flips on a pre-activation within an ulp of zero, one sample's gradient changes discretely, and
the six-layer batch-norm chain amplifies it - which is why tests/test_gpu_deepfm_parity.py gates multi-step parity at
the config's real shape per step on identical weights.  Recorded run (this container, B=2048):

  256,128,64: step 0 logits 2.4e-06 rows 5.7e-09 | step 1 logits 2.7e-06 rows 5.2e-05 | step 2 logits 4.9e-03 rows 2.7e-04
  64,32     : step 0 logits 1.3e-06 rows 2.1e-09 | step 1 logits 1.0e-06 rows 3.2e-09 | step 2 logits 1.2e-06 rows 9.6e-06

usage: python tools/fp32_order_sensitivity.py 2048 256,128,64
"""
import sys, copy, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import oracle as O
from easyrec_b200 import workloads
import test_gpu_deepfm_parity as P
F,D=39,16
def init(V,dnn,final):
  rng=np.random.default_rng(0)
  def mk(sizes, din):
    out=[]
    for u in sizes:
      lim=np.sqrt(6/(din+u)); out.append(dict(W=rng.uniform(-lim,lim,(din,u)).astype(np.float32), b=np.zeros(u,np.float32), gamma=np.ones(u,np.float32), beta=np.zeros(u,np.float32), mean=np.zeros(u,np.float32), var=np.ones(u,np.float32))); din=u
    return out
  n=V+13
  return {'params':{'dnn':mk(dnn,F*D),'final':mk(final,1+D+dnn[-1]),'out_W':rng.uniform(-.1,.1,(final[-1],1)).astype(np.float32),'out_b':np.zeros(1,np.float32)},'acc':{},
     't16':(rng.standard_normal((n,16))*0.0025).astype(np.float32),'a16':np.full((n,16),0.1,np.float32),'t1':(rng.standard_normal((n,1))*0.01).astype(np.float32),'a1':np.full((n,1),0.1,np.float32)}
B,V=int(sys.argv[1]),100003
dnn=final=tuple(int(x) for x in sys.argv[2].split(','))
s0=init(V,dnn,final); s1=copy.deepcopy(s0)
for step in range(3):
  ids,dense,labels=workloads.criteo_batch(B,50+step)
  perm=np.random.default_rng(7).permutation(B)
  l0,_,_=P._oracle_step(s0,ids,dense,labels,B,V)
  ids2=ids.reshape(26,B)[:,perm].reshape(-1)
  l1,_,_=P._oracle_step(s1,ids2,dense[perm],labels[perm],B,V)
  pd={}
  for tag in ('dnn','final'):
    for i,(a,b) in enumerate(zip(s0['params'][tag],s1['params'][tag])):
      for k in ('W','gamma','beta'): pd['%s%d.%s'%(tag,i,k)]=float(np.abs(a[k]-b[k]).max())
  worst=sorted(pd.items(),key=lambda kv:-kv[1])[:3]
  print('step',step,'logits maxdiff %.2e'%np.abs(l0[perm]-l1).max(),'tables %.2e'%np.abs(s0['t16']-s1['t16']).max(), worst, flush=True)

import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, time
from oracle.oracle import Oracle
from fhe_linformer_b200 import Engine
P=dict(logN=int(sys.argv[1]) if len(sys.argv)>1 else 12, L=6, dnum=3)
o=Oracle(**P); e=Engine(sparse_h=64, **P)
assert (o.moduli==e.moduli).all() and (o.roots==e.roots).all() and (o.sf==e.sf).all()
N=o.N; rng=np.random.default_rng(1); q=[int(x) for x in o.moduli]
def rnd(midx): return np.stack([rng.integers(0,q[m],N,dtype=np.uint64) for m in midx])
allm=list(range(o.L+o.K))
a=rnd(allm); d=e.to_dev(a)
A=e.ntt(d,allm).download(); assert (A==o.ntt(a,allm)).all(), "ntt"
assert (e.intt(d,allm).download()==a).all(), "intt"
b=rnd(allm); db=e.to_dev(b); d=e.to_dev(a)
for nm in ['add','sub','mul']:
    assert (getattr(e,nm)(d,db,allm).download()==getattr(o,nm)(a,b,allm)).all(), nm
g=o.galois(3); assert (e.automorph(d,g).download()==o.automorph_eval(a,g)).all(),"auto"
for l in [6,5,3,2]:
    x=rnd(range(l)); dx=e.to_dev(x)
    assert (e.rescale(dx).download()==o.rescale(x)).all(), "rescale"
    for dg in range((l+o.alpha-1)//o.alpha):
        assert (e.modup(dx,dg).download()==o.modup(x,dg)).all(), ("modup",l,dg)
    xe=rnd(list(range(l))+[o.L+k for k in range(o.K)]); assert (e.moddown(e.to_dev(xe)).download()==o.moddown(xe)).all(), "moddown"
sk=o.gen_sk(7,h=64); g=o.galois(1); evk=o.gen_galois_key(11,sk,g); rk=o.gen_relin_key(12,sk); devk=e.to_dev(evk); drk=e.to_dev(rk)
for l in [6,4,3,1]:
    ct=np.stack([rnd(range(l)),rnd(range(l))]); ct2=np.stack([rnd(range(l)),rnd(range(l))])
    k0,k1=o.keyswitch(ct[1],evk); ks=e.keyswitch(e.to_dev(ct[1]),devk).download(); assert (ks[0]==k0).all() and (ks[1]==k1).all(), ("ks",l)
    assert (e.rotate(e.to_dev(ct),g,devk).download()==o.rotate(ct,g,evk)).all(), ("rot",l)
    assert (e.mul_relin(e.to_dev(ct),e.to_dev(ct2),drk).download()==o.mul_relin(ct,ct2,rk)).all(), ("mulrelin",l)
    assert (e.host_rotate(ct,g,devk)==o.rotate(ct,g,evk)).all()
print("ALL OK logN",P['logN'])

"""Independent numpy restatement of clipper.cpp:172-323 (findDenseClique), used once to cross-check
oracle/clipper_oracle.c over 200 random u0 of the reference KAT: identical node sets and scores."""
import numpy as np, sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from oracle import pyoracle as O
import clipper_kats as K
def solve_np(Mu, u0, eps=1e-9, beta=0.25, tol_u=1e-8, tol_F=1e-9):
    M = Mu + Mu.T; C = (M != 0).astype(float); n=len(M); ones=np.ones(n)
    u = M@u0 + u0; u/=np.linalg.norm(u)
    d=0
    Cbu = ones*u.sum() - C@u - u
    idx=(Cbu>eps)&(u>eps)
    if idx.sum()>0:
        d=((M@u+u)[idx]/Cbu[idx]).mean()
    for i in range(1000):
        g=(1+d)*u - d*ones*u.sum() + M@u + C@u*d
        F=u@g
        for j in range(200):
            alpha=1
            for k in range(99):
                un=np.maximum(u+alpha*g,0); un/=np.linalg.norm(un)
                gn=(1+d)*un - d*ones*un.sum() + M@un + C@un*d
                Fn=un@gn; dF=Fn-F
                if dF < -eps: alpha*=beta
                else: break
            du=np.linalg.norm(un-u); F=Fn; u=un; g=gn
            if du<tol_u or abs(dF)<tol_F: break
        Cbu = ones*u.sum() - C@u - u
        idx=(Cbu>eps)&(u>eps)
        if idx.sum()>0:
            d+=np.abs((M@u+u)[idx]/Cbu[idx]).mean()
        else: break
    return u,F,i
model,data=K.kat_clouds(); p=O.clipper_params(); A,Mu=O.clipper_score_pairwise(p,model,data)
cnt={}
for seed in range(200):
    u0=np.random.default_rng(seed).uniform(0,1,len(A))
    sol=O.clipper_find_dense_clique(p,Mu,u0)
    u,F,i=solve_np(Mu,u0)
    k=int(round(F)); nodes=sorted(np.argsort(-u)[:k].tolist())
    key=(tuple(sorted(sol['nodes'].tolist())), tuple(nodes))
    cnt[key]=cnt.get(key,0)+1
    if abs(F-sol['score'])>1e-9: print('MISMATCH',seed,F,sol['score'])
for k,v in sorted(cnt.items(), key=lambda kv:-kv[1]): print(v,k)

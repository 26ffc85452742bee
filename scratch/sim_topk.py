import numpy as np
rng=np.random.default_rng(0)
N,D,k=20000,256,32
X=rng.standard_normal((N,D)).astype(np.float32)
def key(v):
    b=v.view(np.uint32)
    return np.where(b&0x80000000, ~b, b|0x80000000).astype(np.uint64)
def steps_bitwise(x):
    ky=key(x); thr=0; s=0
    for bit in range(31,-1,-1):
        cand=thr|(1<<bit); c=int((ky>=cand).sum()); s+=1
        if c>=k:
            thr=cand
            if c==k: break
    return s
def steps_interp(x, maxit=12, mode="secant"):
    # lanes: element e of lane l: column j*128+l*4+i  -> reshape
    xx=x.reshape(2,32,4).transpose(1,0,2).reshape(32,8)
    lm=xx.max(1)
    lo=lm.min(); clo=int((x>=lo).sum())
    if clo==k: return 1
    hi=lm.max(); chi=int((x>=hi).sum())   # counts: at hi typically 1
    s=2
    if chi>=k: return s  # ties at max.. ignore
    side=0
    flo=clo-k+0.5; fhi=chi-k+0.5   # flo>0, fhi<0 want root of f(T)=count(>=T)-k+0.5
    for it in range(maxit):
        if mode=="bisect": t=0.5*(lo+hi)
        else:
            t=lo+(hi-lo)*flo/(flo-fhi)
        t=np.float32(t)
        if not (lo<t<hi): t=np.float32(0.5*(lo+hi))
        if not (lo<t<hi): return s+8   # no room: tie case
        c=int((x>=t).sum()); s+=1
        if c==k: return s
        f=c-k+0.5
        if f>0:
            lo=t; flo=f
            if mode=="illinois" and side==1: fhi*=0.5
            side=1
        else:
            hi=t; fhi=f
            if mode=="illinois" and side==-1: flo*=0.5
            side=-1
    return s+10
for name,fn in [("bitwise",steps_bitwise),("secant",lambda x:steps_interp(x,mode="secant")),("illinois",lambda x:steps_interp(x,mode="illinois")),("bisect",lambda x:steps_interp(x,mode="bisect"))]:
    st=np.array([fn(X[i]) for i in range(3000)])
    print(name, st.mean(), np.percentile(st,[50,90,99]), st.max())

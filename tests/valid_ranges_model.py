import numpy as np
F=np.float32
def lin_depth(dmin,dmax,step,j,D):
    return F(step*F(j)+dmin) if j < D//2 else F(F(-step)*F(D-1-j)+dmax)   # (fma rounding differs slightly; irrelevant for the property)
def make_valid(T,q,dmin,dmax,dstep,D):
    def valid(j):
        d=lin_depth(dmin,dmax,dstep,j,D)
        p=[F(q[0]*d),F(q[1]*d),F(q[2]*d)]
        x=[F(F(F(F(p[0]*T[4*r])+F(p[1]*T[4*r+1]))+F(p[2]*T[4*r+2]))+T[4*r+3]) for r in range(3)]
        n=np.sqrt(F(x[2]*x[2]+F(x[1]*x[1]+F(x[0]*x[0]))))
        return bool(n<F(1.0))
    return valid
def exact_hull(valid,D):
    v=[j for j in range(D) if valid(j)]
    return (v[0],v[-1]-v[0]+1) if v else (0,0)
def fast_hull(valid,T,q,dmin,dmax,dstep,D,stats):
    w0,w1=0,D-1; none=False
    ux=F(T[0]*q[0]+T[1]*q[1]+T[2]*q[2]); uy=F(T[4]*q[0]+T[5]*q[1]+T[6]*q[2]); uz=F(T[8]*q[0]+T[9]*q[1]+T[10]*q[2])
    aa=F(ux*ux+uy*uy+uz*uz); bb=F(ux*T[3]+uy*T[7]+uz*T[11]); cc=F(T[3]*T[3]+T[7]*T[7]+T[11]*T[11]-F(1))
    if aa>1e-20 and aa<1e20 and dstep>1e-12 and abs(dmin)<1e4 and abs(dmax)<1e4 and abs(bb)<1e20 and abs(cc)<1e20:
        inv=F(1)/aa
        if F(cc-bb*bb*inv)>F(0.02): none=True
        else:
            sq=F(np.sqrt(max(F(bb*bb-aa*cc),F(0)))*inv); dc=F(-bb*inv)
            flo=F((dc-sq-dmin)/dstep); fhi=F((dc+sq-dmin)/dstep)
            if flo==flo and fhi==fhi:
                if fhi<-2 or flo>D+1: none=True
                else:
                    w0=max(0,int(np.floor(max(flo,F(-4))))-2); w1=min(D-1,int(np.ceil(min(fhi,F(D+4))))+2)
    lo=-1;hi=-1;tests=0
    if not none:
        for j in range(w0,w1+1):
            tests+=1
            if valid(j): lo=j;break
        if lo<0 and (w0>0 or w1<D-1):
            stats['full']+=1
            w1=D-1
            for j in range(D):
                tests+=1
                if valid(j): lo=j;break
        if lo>=0:
            while lo>0:
                tests+=1
                if valid(lo-1): lo-=1
                else: break
            hi=lo
            j=w1
            while j>lo:
                tests+=1
                if valid(j): hi=j;break
                j-=1
            while hi<D-1:
                tests+=1
                if valid(hi+1): hi+=1
                else: break
    stats['tests']+=tests
    return (lo,hi-lo+1) if lo>=0 else (0,0)
def run(n,seed,D=50):
    rng=np.random.default_rng(seed); stats=dict(full=0,tests=0); bad=0; nonempty=0
    for i in range(n):
        s=F(rng.uniform(0.5,3.0)); th=rng.uniform(-np.pi,np.pi)
        R=np.array([[np.cos(th),0,np.sin(th)],[0,1,0],[-np.sin(th),0,np.cos(th)]])@np.diag([1,-1,-1])
        if i%3==0:
            A=rng.normal(size=(3,3)); R,_=np.linalg.qr(A)
        tco=np.array([rng.uniform(-5,5),rng.uniform(-1,2),rng.uniform(3,25)])
        Tco=np.eye(4); Tco[:3,:3]=s*R; Tco[:3,3]=tco
        Toc=np.linalg.inv(Tco)
        T=Toc[:3,:].astype(F).reshape(-1)
        dmin=F(tco[2]-s); dmax=F(tco[2]+s)
        if i%7==0: dmin=F(dmin+rng.uniform(-3,3)); dmax=F(dmin+abs(rng.uniform(0.1,6)))
        dstep=F((dmax-dmin)/F(D-1))
        # ray: towards a point near the object (so that many rays graze the ball)
        target=tco+s*rng.normal(size=3)*rng.choice([0.3,0.9,1.0,1.05,1.5])
        q=(target/target[2]).astype(F)
        valid=make_valid(T,q,dmin,dmax,dstep,D)
        e=exact_hull(valid,D); f=fast_hull(valid,T,q,dmin,dmax,dstep,D,stats)
        nonempty+= e[1]>0
        if e!=f: bad+=1; print("MISMATCH",i,e,f)
    return bad,nonempty,stats
if __name__=="__main__":
    import sys
    bad,ne,st=run(int(sys.argv[1]) if len(sys.argv)>1 else 3000, 1)
    print("bad",bad,"nonempty",ne,st)

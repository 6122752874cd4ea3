import numpy as np, sys
sys.path.insert(0,'.')
from oracle import oracle as O
LD=np.longdouble

def mats(degree, nq, colloc, skew):
    n=degree+1
    nodes,_=O.gauss_lobatto(n)
    xq,w = (O.gauss_lobatto(nq) if colloc else O.gauss_legendre(nq))
    S=O.lagrange_eval(nodes,xq); D=O.lagrange_deriv(xq,xq)
    f0=O.lagrange_eval(xq,[0.0])[0]; f1=O.lagrange_eval(xq,[1.0])[0]
    W=np.diag(w); Wi=np.diag(1/w)
    Sinv = O._inv_ld(S) if nq==n else O._inv_ld(S.T@W@S)@S.T@W
    s=LD(skew)
    V = Sinv@(-s*D + (1-s)*Wi@D.T@W)@S
    l0 = Sinv@Wi@f0; l1=Sinv@Wi@f1
    return V,l0,l1

def build(degree,nq,colloc,skew,a,h):
    # returns C (n x n), L0, L1 (n) for direction with speed a, size h
    V,l0,l1=mats(degree,nq,colloc,skew)
    n=degree+1
    a=LD(a);h=LD(h);s=LD(skew)
    C=(a/h)*V
    out=[]
    for f,(l,e,nf) in enumerate(((l0,0,-1),(l1,n-1,+1))):
        an=a*nf
        alpha=-(an/2+abs(a)/2-s*an)/h
        beta=-(an-abs(a))/2/h
        C[:,e]+=alpha*l
        out.append(beta*l)
    return C.astype(np.float64),out[0].astype(np.float64),out[1].astype(np.float64)

def apply_kron(mesh,degree,nq,colloc,skew,vel,src):
    dim=mesh.dim;n=degree+1
    u=src.reshape(tuple(reversed(mesh.n_cells[:dim]))+(n,)*dim)
    out=np.zeros_like(u)
    for d in range(dim):
        C,L0,L1=build(degree,nq,colloc,skew,vel[d],mesh.h[d])
        axn=2*dim-1-d; axc=dim-1-d
        out+=np.moveaxis(np.tensordot(C,u,axes=([1],[axn])),0,axn)
        # neighbour traces
        lo=np.take(np.roll(u,1,axis=axc),n-1,axis=axn)  # lower neighbour's upper layer
        hi=np.take(np.roll(u,-1,axis=axc),0,axis=axn)
        sh=[1]*(2*dim); sh[axn]=n
        out+=L0.reshape(sh)*np.expand_dims(lo,axn)+L1.reshape(sh)*np.expand_dims(hi,axn)
    return out.reshape(-1)

rng=np.random.default_rng(1)
for (dx,dv,nc,deg,nq,col,skew) in [(1,1,(4,3),3,4,False,0.0),(2,2,(3,2,4,2),3,4,False,0.5),(2,2,(3,2,4,2),3,5,False,0.5),(2,2,(2,2,2,2),3,4,True,0.3),(2,1,(2,3,2),2,3,False,1.0),(3,3,(2,2,2,2,2,2),3,4,False,0.5)]:
    dim=dx+dv
    mesh=O.Mesh(dx,dv,nc,tuple(-1+0.1*d for d in range(dim)),tuple(1+0.2*d for d in range(dim)),(True,)*dim)
    vel=np.array([1,0.15,-0.05,0.1,-0.15,0.5])[:dim]
    orc=O.Oracle(mesh,deg,nq=nq,collocation=col,skew=skew,velocity=vel,nthreads=4)
    src=rng.standard_normal(orc.ndofs)
    ref=orc.apply(src)
    mine=apply_kron(mesh,deg,nq,col,skew,vel,src)
    print(dx,dv,deg,nq,col,skew, np.max(np.abs(ref-mine))/np.max(np.abs(ref)))

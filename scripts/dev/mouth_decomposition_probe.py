"""Developer probe (CPU only) behind DESIGN.md section 9, row f4: the reference's pose converters decompose the mouth shape with
scipy.optimize.minimize (L-BFGS-B, numerical gradients) on the non-smooth objective ||d M - p||_2 + 0.01 ||d||_1 over [0,1]^4
(src/tha4/mocap/ifacialmocap_pose_converter_25.py:574-579, mediapipe_face_pose_converter_00.py:627-632).  This script solves the same
convex program exactly (iteratively re-weighted box-QPs, each solved by enumerating the 81 active sets) and compares: on 400 random
mouth points the two agree to 1e-3 in 97.75 % of the cases; in the rest the reference's optimiser stops early at the kink of the norm
(objective up to 0.089 above the optimum, coefficients up to 0.24 away).  The reference's result is therefore defined by that
solver's iterates, not by the program it minimises: a batched GPU solver cannot be pinned to it."""
import numpy as np, scipy.optimize, itertools, time
M = np.array([[1.0,1.0,0.0,0.0],[0.0,1.0,0.0,0.0],[0.5,0.3,0.25,0.75],[1.0,0.5,0.5,0.4]])
LAM=0.01
def loss(d,p): return np.linalg.norm(d@M-p)+LAM*np.linalg.norm(d,ord=1)
def ref(p):
    r=scipy.optimize.minimize(lambda d: loss(d,p), np.array([0,0,0,0]), bounds=[(0.0,1.0)]*4)
    return r["x"]
# exact box-QP: min 0.5||M^T d - p||^2 + c * sum(d), 0<=d<=1   via active-set enumeration
A=M@M.T   # 4x4 (rows of M are basis points): (d@M - p) -> gradient = M (M^T d - p) = A d - M p
def boxqp(p,c):
    b=M@p - c          # gradient = A d - b
    best=None;bestv=None
    for pat in itertools.product((0,1,2),repeat=4):   # 0: at 0, 1: at 1, 2: free
        free=[i for i in range(4) if pat[i]==2]
        d=np.array([1.0 if pat[i]==1 else 0.0 for i in range(4)])
        if free:
            Aff=A[np.ix_(free,free)]
            rhs=b[free]-A[np.ix_(free,[i for i in range(4) if pat[i]!=2])]@d[[i for i in range(4) if pat[i]!=2]]
            try: x=np.linalg.solve(Aff,rhs)
            except np.linalg.LinAlgError: continue
            if np.any(x< -1e-12) or np.any(x>1+1e-12): continue
            d[free]=np.clip(x,0,1)
        g=A@d-b
        ok=True
        for i in range(4):
            if pat[i]==0 and g[i]< -1e-10: ok=False
            if pat[i]==1 and g[i]> 1e-10: ok=False
        if not ok: continue
        v=0.5*np.sum((d@M-p)**2)+c*np.sum(d)
        if bestv is None or v<bestv: best,bestv=d,v
    return best
def exact(p,iters=60):
    d=boxqp(p,0.0)
    w=np.linalg.norm(d@M-p)
    for _ in range(iters):
        if w<1e-14: break
        d2=boxqp(p,LAM*w)
        w2=np.linalg.norm(d2@M-p)
        if abs(w2-w)<1e-15: d=d2; break
        d,w=d2,w2
    return d
rng=np.random.default_rng(0)
worst=0;gapmax=0;neg=0
T=400
t0=time.time()
res=[]
for k in range(T):
    mo=rng.uniform(0.01,1.0); ml=min(1.0,rng.uniform(0,1.2)); mf=rng.uniform(0,1.0); mp=rng.uniform(0,1.0)
    if k%4==0: mf=mp=0.0
    p=np.array([mo,ml,mf,mp])
    a=ref(p); b=exact(p)
    fa,fb=loss(a,p),loss(b,p)
    res.append((np.abs(a-b).max(), fa-fb))
res=np.array(res)
print('time',time.time()-t0)
print('max |d_scipy - d_exact|', res[:,0].max(), 'mean', res[:,0].mean(), 'p99', np.quantile(res[:,0],0.99))
print('objective gap scipy-exact: min', res[:,1].min(), 'max', res[:,1].max())
print('frac with |dd|>1e-3', (res[:,0]>1e-3).mean(), '>1e-2', (res[:,0]>1e-2).mean())

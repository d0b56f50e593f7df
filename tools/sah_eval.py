import sys, numpy as np, time
sys.path.insert(0,'/root/repo')
from vanrijn_b200 import scenes
sys.setrecursionlimit(100000)
pos,nrm,faces = scenes.bunny_proxy(5)   # 20480 triangles (same shape as the 81920 one, for speed)
V,_ = scenes.mesh_arrays(pos,nrm,faces)
tri = V.reshape(-1,3,3)
lo = tri.min(1); hi = tri.max(1); cen=(lo+hi)/2
n=len(tri)
def build(idx, mode):
    blo=lo[idx].min(0); bhi=hi[idx].max(0)
    if len(idx)==1: return (blo,bhi,int(idx[0]))
    ext=bhi-blo
    if mode=='median':
        ax=int(np.argmax(ext)); order=idx[np.argsort(cen[idx,ax],kind='stable')]; piv=len(idx)//2
    else:
        best=(np.inf,None,None)
        for ax in range(3):
            order=idx[np.argsort(cen[idx,ax],kind='stable')]
            l_lo=np.minimum.accumulate(lo[order],0); l_hi=np.maximum.accumulate(hi[order],0)
            r_lo=np.minimum.accumulate(lo[order][::-1],0)[::-1]; r_hi=np.maximum.accumulate(hi[order][::-1],0)[::-1]
            def area(a,b):
                d=b-a; return 2*(d[:,0]*d[:,1]+d[:,1]*d[:,2]+d[:,0]*d[:,2])
            k=np.arange(1,len(order))
            cost=area(l_lo[:-1],l_hi[:-1])*k + area(r_lo[1:],r_hi[1:])*(len(order)-k)
            j=int(np.argmin(cost))
            if cost[j]<best[0]: best=(cost[j],order,j+1)
        order,piv=best[1],best[2]
    return (blo,bhi,build(order[:piv],mode),build(order[piv:],mode))
def slab(o,inv,blo,bhi):
    t1=(blo-o)*inv; t2=(bhi-o)*inv
    tn=np.minimum(t1,t2).max(); tf=np.maximum(t1,t2).min()
    return tn,tf
def tri_hit(o,d,t):
    v0,v1,v2=tri[t]; e1=v1-v0; e2=v2-v0; p=np.cross(d,e2); det=e1@p
    if abs(det)<1e-14: return np.inf
    tv=o-v0; u=(tv@p)/det
    if u<0 or u>1: return np.inf
    q=np.cross(tv,e1); v=(d@q)/det
    if v<0 or u+v>1: return np.inf
    tt=(e2@q)/det
    return tt if tt>0 else np.inf
def trace(root,o,d):
    inv=1.0/d; best=np.inf; boxes=0; tris=0
    stack=[root]
    while stack:
        nd=stack.pop()
        if len(nd)==3:
            tris+=1; best=min(best,tri_hit(o,d,nd[2])); continue
        hits=[]
        for c in (nd[2],nd[3]):
            boxes+=1
            tn,tf=slab(o,inv,c[0],c[1])
            if tn<=tf and tf>=0 and tn<=best: hits.append((tn,c))
        hits.sort(key=lambda x:-x[0])
        for h in hits: stack.append(h[1])
    return boxes,tris
rng=np.random.default_rng(1)
# rays: camera rays toward the mesh + random bounce-like rays from points on a sphere around it
cam=np.array([-2.0,1.0,-5.0]); ctr=np.array([0,-0.5,0.0])
rays=[]
for i in range(1500):
    tgt=ctr+rng.normal(size=3)*0.8; d=tgt-cam; rays.append((cam,d/np.linalg.norm(d)))
for i in range(1500):
    u=rng.normal(size=3); u/=np.linalg.norm(u); o=ctr+2.2*u
    d=rng.normal(size=3); d/=np.linalg.norm(d)
    if d@(ctr-o)<0: d=-d
    rays.append((o,d))
for mode in ('median','sah'):
    t0=time.time(); root=build(np.arange(n),mode); tb=time.time()-t0
    B=T=0
    for o,d in rays:
        b,t=trace(root,o,d); B+=b; T+=t
    print(mode,"build %.1fs"%tb,"box tests/ray %.1f"%(B/len(rays)),"tri tests/ray %.2f"%(T/len(rays)))

"""CPU check that the two block-level culls of the composite kernels never reject a splat a pixel would accept
(csrc/composite_common.cuh::splat_hits_block: the box of the a*G >= 1/255 region, and the ellipse-vs-block test behind
it).  An INDEPENDENT numpy restatement of both tests (fp32, same expressions) is run over every (8x4 pixel block, tile
list entry) pair of a scene and compared with the exact per-pixel criterion (the reference's fp64 Gaussian,
kernels.h:195-224): a "needed" pair is one where some pixel of the block has a*G >= 1/255.

    python tools/cull_conservativeness_check.py [cfg N reso svec_scale]     (default: c3 30000 320 2.0)

Written while diagnosing the drop-in difference of round 2 (which turned out to be tie order, not a cull); result for
the default scene: 309 072 needed pairs, 0 box false negatives, 0 ellipse false negatives."""
import sys, ctypes, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from gsgen_b200.scenes import make_scene
from tests.util import ocam_of, fp
args = sys.argv[1:]
cfg_name, N_, reso_, scale_ = (args + ["c3", "30000", "320", "2.0"][len(args):])[:4]
sc = make_scene(cfg_name, N=int(N_), reso=int(reso_)); sc.svec = (sc.svec * float(scale_)).contiguous()
cam, c2w = sc.cams[0], sc.c2ws[0]
ocam = ocam_of(cam)
normals, pts = oracle.get_frustum(ocam, c2w)
mask = oracle.cull_bsphere(sc.mean, sc.svec, normals, pts, 6.0)
m2, cov, _, dp = oracle.project_gaussians(sc.mean[mask], sc.qvec[mask], sc.svec[mask], c2w, True)
D, tl, br = oracle.tile_culling_aabb_count(m2, cov, 16, ocam, 6.0)
th, tw = cam.n_tiles
ids, start, end = oracle.tile_culling_aabb_start_end(tl, br, dp, th, tw, D)
print("visible", int(mask.sum()), "D", D)

m2n = m2.detach().numpy().astype(np.float32); covn = cov.detach().reshape(-1,4).numpy().astype(np.float32)
al = sc.alpha[mask].numpy().astype(np.float32)
# make_splat port (fp64 as in the header)
c0,c1,c2,c3 = [covn[:,i].astype(np.float64) for i in range(4)]
det = c0*c3-c1*c2; b=0.5*(c1+c2)
a = np.minimum(al, np.float32(0.99)).astype(np.float32)
A=c3/det; B=-b/det; Cc=c0/det
l00=np.sqrt(A); l01=B/l00; l11sq=Cc-l01*l01
kChol=0.84932180028801907
p0=(kChol*l00).astype(np.float32); p1=(kChol*l01).astype(np.float32); p2=(kChol*np.sqrt(l11sq)).astype(np.float32)
a255=255.0*a.astype(np.float64)
qmax=2.0*np.log(a255)*(1+1e-6)+1e-6
dq=A*Cc-B*B
hx=((np.sqrt(qmax*Cc/dq)*(1+1e-4)).astype(np.float32)+np.float32(1e-6)).astype(np.float32)
hy=((np.sqrt(qmax*A/dq)*(1+1e-4)).astype(np.float32)+np.float32(1e-6)).astype(np.float32)
hx[~(a255>1)]=-1; hy[~(a255>1)]=-1
mx=m2n[:,0]; my=m2n[:,1]
psx=np.float32(1.0/cam.fx); psy=np.float32(1.0/cam.fy); tlx=np.float32(-cam.cx/cam.fx); tly=np.float32(-cam.cy/cam.fy)
f32=np.float32
ids_n=ids.numpy(); st=start.numpy(); en=end.numpy()
bad_box=bad_ell=0; need_total=0; examples=[]
for t in range(th*tw):
    if st[t]<0: continue
    g=ids_n[st[t]:en[t]].astype(np.int64)
    ty,tx=divmod(t,tw)
    for w in range(8):
        bx0=tx*16+(w&1)*8; by0=ty*16+(w>>1)*4
        X0=f32(bx0)*psx+tlx; X1=f32(bx0+7)*psx+tlx; Y0=f32(by0)*psy+tly; Y1=f32(by0+3)*psy+tly
        # exact need: any pixel in block with a*G>=1/255 (fp64 reference formula)
        xs=(np.arange(bx0,bx0+8,dtype=np.float32)*psx+tlx).astype(np.float64); ys=(np.arange(by0,by0+4,dtype=np.float32)*psy+tly).astype(np.float64)
        dx=xs[None,:,None]-mx[g].astype(np.float64)[:,None,None]; dy=ys[None,None,:]-my[g].astype(np.float64)[:,None,None]
        q=((dx*c3[g][:,None,None]-dy*c2[g][:,None,None])*dx+(-dx*c1[g][:,None,None]+dy*c0[g][:,None,None])*dy)/det[g][:,None,None]
        G=np.exp(-0.5*q).astype(np.float32)
        need=((a[g][:,None,None]*G)>=f32(0.00392156862745098)).any(axis=(1,2))
        # box test
        box=(mx[g]-hx[g]<=X1)&(mx[g]+hx[g]>=X0)&(my[g]-hy[g]<=Y1)&(my[g]+hy[g]>=Y0)
        # ellipse test (fp32)
        dx0=(X0-mx[g]).astype(f32); dx1=(X1-mx[g]).astype(f32); dy0=(Y0-my[g]).astype(f32); dy1=(Y1-my[g]).astype(f32)
        inside=(dx0<=0)&(dx1>=0)&(dy0<=0)&(dy1>=0)
        L=(np.log2(f32(255.0)*a[g]).astype(f32)*f32(1.001)+f32(1e-4)).astype(f32)
        P0,P1,P2=p0[g],p1[g],p2[g]
        ua=P0*dx0; ub=P0*dx1; o0=P1*dy0; o1=P1*dy1; v0=P2*dy0; v1=P2*dy1
        cc0=np.minimum(np.maximum(0,ua+o0),ub+o0); cc1=np.minimum(np.maximum(0,ua+o1),ub+o1)
        best=np.minimum(cc0*cc0+v0*v0, cc1*cc1+v1*v1)
        inv=1.0/(P1*P1+P2*P2)
        t0=np.minimum(np.maximum(-ua*P1*inv,dy0),dy1); t1=np.minimum(np.maximum(-ub*P1*inv,dy0),dy1)
        e0u=P1*t0+ua; e0v=P2*t0; e1u=P1*t1+ub; e1v=P2*t1
        best=np.minimum(best,np.minimum(e0u*e0u+e0v*e0v,e1u*e1u+e1v*e1v))
        ell=inside|(best<=L)
        need_total+=int(need.sum())
        bb=need&~box; be=need&box&~ell
        bad_box+=int(bb.sum()); bad_ell+=int(be.sum())
        if be.any() and len(examples)<3:
            k=np.nonzero(be)[0][0]; examples.append((t,w,int(g[k]),float(best[k]),float(L[k]),float(a[g[k]]),float(P0[k]),float(P1[k]),float(P2[k]),float(dx0[k]),float(dx1[k]),float(dy0[k]),float(dy1[k])))
        if bb.any() and len(examples)<6:
            k=np.nonzero(bb)[0][0]; examples.append(('box',t,w,int(g[k]),float(hx[g[k]]),float(hy[g[k]]),float(mx[g[k]]),float(my[g[k]]),float(X0),float(X1),float(Y0),float(Y1)))
print("needed (block,entry) pairs", need_total, "box false negatives", bad_box, "ellipse false negatives", bad_ell)
for e in examples: print(e)

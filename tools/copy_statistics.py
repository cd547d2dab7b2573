"""CPU-side counting on random grid positions: projected copies, (copy, ray-in-pitch-band) pairs of the pitch-run walk, and
in-box (ray, heading) candidates per position, by copy size ("what was measured this round")."""
import numpy as np, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from invertsy_b200.env.world import Seville2009
from invertsy_b200.render import fibonacci_lattice
pol, col = Seville2009.load_world()
pitch, yaw = fibonacci_lattice(1000)
H=36; heads = (np.arange(H)*2*np.pi/H + np.pi)%(2*np.pi)-np.pi
rng=np.random.RandomState(0)
tot = dict(E=0,pairs_old=0,cand=0,inside=0)
hist_c=[]; 
for it in range(12):
    p = np.array([rng.uniform(0,10), rng.uniform(0,10), 0.01])
    v = pol.astype(np.float64)-p
    d = np.linalg.norm(v,axis=2).min(1)
    vis = d<=2.0
    v=v[vis]
    ph = np.arctan2(v[:,:,1],v[:,:,0]); th=np.arctan2(np.hypot(v[:,:,0],v[:,:,1]),v[:,:,2])-np.pi/2
    copies=[]
    for t in range(v.shape[0]):
        if ph[t].max()-ph[t].min()>=np.pi:
            p1=np.where(ph[t]<0,ph[t]+2*np.pi,ph[t]); p2=np.where(ph[t]>0,ph[t]-2*np.pi,ph[t])
            copies.append((th[t],p1)); copies.append((th[t],p2))
        else: copies.append((th[t],ph[t]))
    E=len(copies); tot['E']+=E
    for (t3,p3) in copies:
        tlo,thi,plo,phi=t3.min(),t3.max(),p3.min(),p3.max()
        band = (pitch>=tlo)&(pitch<=thi)
        nb=band.sum(); tot['pairs_old']+=nb
        # candidates: (ray,heading) with wrap(yaw+h) in [plo,phi] (clipped to (-pi,pi])
        yb=yaw[band]
        q=(yb[:,None]+heads[None,:]+np.pi)%(2*np.pi)-np.pi
        inb=((q>=max(plo,-np.pi))&(q<=min(phi,np.pi)))
        c=inb.sum(); tot['cand']+=c; hist_c.append((nb,c,thi-tlo,min(phi,np.pi)-max(plo,-np.pi)))
print({k:v/12 for k,v in tot.items()})
h=np.array(hist_c)
print("copies per pos", len(h)/12)
for lo,hi in ((0,8),(8,32),(32,128),(128,512),(512,99999)):
    m=(h[:,1]>=lo)&(h[:,1]<hi)
    print("cand in [%d,%d): copies %.1f/pos  cand share %.3f  old-pairs share %.3f  mean tall %.1f deg wide %.1f deg"%(lo,hi,m.sum()/12,h[m,1].sum()/h[:,1].sum(),h[m,0].sum()/h[:,0].sum(),np.degrees(h[m,2].mean()),np.degrees(h[m,3].mean())))

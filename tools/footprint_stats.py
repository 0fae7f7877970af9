"""Developer tool (CPU): exact culling / footprint statistics of config 2 from the oracle state
(how many tile instances can reach a tile, an 8x4 / 8x8 / 16x8 pixel patch; pairs passing the alpha test; pairs blended)."""
import numpy as np, torch, sys
sys.path.insert(0,'/root/repo')
from saro_gs_b200 import synthetic
from oracle import oracle
scene, cam = synthetic.config2_scene()
r = oracle.forward_scene(scene, cam, torch.zeros(3), precision="f32")
W,H = cam.width, cam.height
tx = (W+15)//16
m2 = r.means2D; co = r.conic_opacity
rng = np.random.default_rng(0)
tiles = rng.choice(len(r.ranges), 400, replace=False)
tot = dict(inst=0, tile_keep=0, p8x4=0, p8x8=0, p16x8=0, p4x4=0, pairs_alpha=0, pairs_contrib=0)
ncon = r.n_contrib
for t in tiles:
    s,e = r.ranges[t]
    if e<=s: continue
    ids = r.point_list[s:e]
    x0 = (t % tx)*16; y0=(t//tx)*16
    px = np.arange(x0,x0+16)[None,None,:].astype(np.float32); py=np.arange(y0,y0+16)[None,:,None].astype(np.float32)
    dx = m2[ids,0][:,None,None]-px; dy = m2[ids,1][:,None,None]-py
    A=co[ids,0][:,None,None]; B=co[ids,1][:,None,None]; C=co[ids,2][:,None,None]; o=co[ids,3][:,None,None]
    power = -0.5*(A*dx*dx + C*dy*dy) - B*dx*dy
    alpha = np.minimum(0.99, o*np.exp(power))
    ok = (power<=0)&(alpha>=1/255)
    inside = (px<W)&(py<H)
    ok = ok & inside
    # list pos < n_contrib
    yy = np.clip(np.arange(y0,y0+16),0,H-1); xx=np.clip(np.arange(x0,x0+16),0,W-1)
    nc = ncon[np.ix_(yy,xx)][None]
    pos = np.arange(e-s)[:,None,None]
    contrib = ok & (pos < nc)
    tot['inst'] += e-s
    tot['tile_keep'] += ok.any(axis=(1,2)).sum()
    tot['pairs_alpha'] += ok.sum()
    tot['pairs_contrib'] += contrib.sum()
    def patches(ph,pw,m):
        n = m.shape[0]
        return m.reshape(n,16//ph,ph,16//pw,pw).any(axis=(2,4)).sum()
    tot['p8x4'] += patches(4,8,ok); tot['p8x8']+=patches(8,8,ok); tot['p16x8']+=patches(8,16,ok); tot['p4x4']+=patches(4,4,ok)
    tot.setdefault('c8x4',0); tot['c8x4'] += patches(4,8,contrib)
    tot.setdefault('c8x8',0); tot['c8x8'] += patches(8,8,contrib)
    tot.setdefault('c16x8',0); tot['c16x8'] += patches(8,16,contrib)
print(tot)
n=tot['inst']
for k,v in tot.items(): print(k, v/n)

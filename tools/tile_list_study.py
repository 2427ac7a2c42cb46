"""CPU study (oracle buffers, cfg3 view 0): share of tile-list entries whose alpha extent misses their tile, chunks the
compositor walks would visit without them, and the distribution of tile-list lengths.  Analysis infrastructure: imports
oracle/, never imported by the product.  python tools/tile_list_study.py"""
import sys, numpy as np, torch
sys.path[:0] = ['/root/repo', '/root/repo/hair-gs_b200', '/root/repo/tests']
import common
from oracle import pyoracle
d = common.strand_inputs(10000, 100, 1024, 1024, 'cpu', seed=0, view=0, n_views=16)
f = pyoracle.Forward({k: (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k, v in d.items()})
m2, co, pl = f.array("means2D"), f.array("conic_opacity"), f.array("point_list")
ranges = f.array("ranges"); ncontrib = f.array("n_contrib").reshape(1024,1024)
A,B,C,o = [co[:, i].astype(np.float64) for i in range(4)]
det = A*C-B*B
t = 2*np.log(255*np.maximum(o,1/255))*1.002+1e-3
with np.errstate(all='ignore'):
    hx = np.sqrt(t*C/det)*1.001+1e-3; hy = np.sqrt(t*A/det)*1.001+1e-3
x,y,ex,ey = m2[pl,0].astype(np.float64), m2[pl,1].astype(np.float64), hx[pl], hy[pl]
rs=np.random.default_rng(0); ne=np.nonzero(ranges[:,1]>ranges[:,0])[0]; sel=rs.choice(ne,400,replace=False)
ch_all=ch_live=0; tot_len=0; live_len=0
lens=(ranges[:,1]-ranges[:,0]); 
for tl in sel:
    r0,r1=ranges[tl]; X0,Y0=(tl%64)*16,(tl//64)*16
    xs,ys,exs,eys=x[r0:r1],y[r0:r1],ex[r0:r1],ey[r0:r1]
    live=~((xs-exs>X0+15)|(xs+exs<X0)|(ys-eys>Y0+15)|(ys+eys<Y0))
    tot_len+=len(xs); live_len+=live.sum()
    cl=np.cumsum(live)
    for w in range(8):
        bx,by=X0+(w&1)*8,Y0+(w>>1)*4
        L=int(ncontrib[by:by+4,bx:bx+8].max())
        if L==0: continue
        ch_all+=(L+31)//32
        ch_live+=(int(cl[L-1])+31)//32
print('instances live in tile', live_len/tot_len, 'chunks visited all', ch_all, 'with dead removed', ch_live, ch_live/ch_all)
print('tile list length percentiles (non-empty):', np.percentile(lens[lens>0],[50,90,99,100]), 'share of instances in lists >4096:', lens[lens>4096].sum()/lens.sum(), '>1024:', lens[lens>1024].sum()/lens.sum())

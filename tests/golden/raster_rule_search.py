"""Scores 64 variants of Pillow's scanline polygon fill against the reference's recorded robot episodes (poses from the oracle replay,
pixels from the gif frames) to find the rule set of the Pillow the recordings were made with; see tests/test_gif_episodes.py."""
import sys, itertools, math
import os; HERE = os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np, boxlcd_b200 as blcd
from oracle import oracle
import test_gif_episodes as T
F=np.float32
def round_up(f): return int(math.floor(f+0.5)) if f>=0 else -int(math.floor(abs(f)+0.5))
def round_down(f): return int(math.ceil(f-0.5)) if f>=0 else -int(math.ceil(abs(f)-0.5))
def hline(img,x0,y,x1):
    H,W=img.shape
    if y<0 or y>=H: return
    if x0>x1: x0,x1=x1,x0
    x0=max(x0,0); x1=min(x1,W-1)
    if x0<=x1: img[y,x0:x1+1]=1
def poly(img,P,opt):
    A,B,C,D,E,G=opt
    H,W=img.shape
    n=len(P)
    edges=[(P[i],P[i+1]) for i in range(n-1)]
    if P[-1]!=P[0]: edges.append((P[-1],P[0]))
    tab=[]; ylo=None; yhi=None
    for (x0,y0),(x1,y1) in edges:
        ymin,ymax=min(y0,y1),max(y0,y1)
        if C or y0!=y1:
            ylo=ymin if ylo is None else min(ylo,ymin); yhi=ymax if yhi is None else max(yhi,ymax)
        if y0==y1:
            if A: hline(img,x0,y0,x1)
            continue
        dx=F(x1-x0)/F(y1-y0)
        tab.append((x0,y0,x1,y1,ymin,ymax,dx))
    if ylo is None: return
    Ymin=max(ylo,0); Ymax=min(yhi,H if not E else H-1)
    for y in range(Ymin,Ymax+1):
        xx=[]
        for (x0,y0,x1,y1,ymin,ymax,dx) in tab:
            if G:   # half-open rule: ymin <= y < ymax
                if ymin<=y<ymax or (y==ymax and y==Ymax and False): xx.append(float(F(F(y-y0)*dx)+F(x0)))
                continue
            if ymin<=y<=ymax:
                v=float(F(F(y-y0)*dx)+F(x0)); xx.append(v)
                if B and y==ymax and y<Ymax: xx.append(v)
        xx.sort()
        pos=None
        for k in range(1,len(xx),2):
            xs,xe=round_up(xx[k-1]),round_down(xx[k])
            if D:
                if pos is not None:
                    if xe<pos: continue
                    if xs<pos: xs=pos
                if xe<xs: continue
                hline(img,xs,y,xe); pos=xe+1
            else: hline(img,xs,y,xe)
def ellipse_from_oracle(shapes,poses,env,sp):
    return None
def collect():
    data=[]
    for name in ['Urchin','Luxo','UrchinCube','UrchinBall','LuxoBall']:
        lcd, init, actions = T.load_robot(name)
        env = blcd.env_map[name](); sp=env.layout.spec
        bodies=np.zeros((1,sp.n_bodies,6),np.float32); bodies[0,:,:3]=init
        ow=oracle.OracleWorlds(sp,1); ow.set_bodies(bodies)
        shapes=ow.lcd_shapes(0)
        K=T.ROBOT_GIFS[name][0]
        for t in range(K):
            ow.step(actions[t].astype(np.float32)[None])
            poses,_=ow.get_poses()
            polys=[]; circ_shapes=shapes.copy()
            for b in range(sp.n_bodies):
                x,y,s,c=[F(v) for v in poses[0,b]]
                if shapes[b]['kind']==0: continue
                nv=int(shapes[b]['n'])
                pts=[]
                for i in range(nv):
                    vx,vy=F(shapes[b]['verts'][i][0]),F(shapes[b]['verts'][i][1])
                    wx=F(F(c*vx)-F(s*vy))+x; wy=F(F(s*vx)+F(c*vy))+y
                    pts.append((int(float(wx)/env.WIDTH*sp.lcd_w), int(float(wy)/env.WIDTH*sp.lcd_w)))
                polys.append(pts)
            # circles rendered by the oracle alone: make polygons degenerate by zero-vertex trick -> render with only circles
            only=shapes.copy()
            for b in range(sp.n_bodies):
                if only[b]['kind']!=0: only[b]['n']=0
            base=~oracle.unpack_bits(oracle.lcd_render(only, poses, env.WIDTH, sp.lcd_w, sp.lcd_h, 1), sp.lcd_w)[0]
            data.append((name,t,polys,base[::-1].astype(np.uint8),(~lcd[t])[::-1].astype(np.uint8)))
    return data
data=collect()
print('frames',len(data))
best=[]
for opt in itertools.product([0,1],repeat=6):
    exact=0; px=0
    for name,t,polys,base,want in data:
        img=base.copy()
        for P in polys: poly(img,P,opt)
        d=int((img!=want).sum()); exact+=d==0; px+=d
    best.append((exact,-px,opt))
best.sort(reverse=True)
for b in best[:12]: print(b)

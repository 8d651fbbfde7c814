import sys, numpy as np
for p in ("oracle","tests","real-time-path-tracing-voxel-blocks_b200/python"): sys.path.insert(0,p)
import common, oracle as O, vpt, vpt_scenes as S
O.build()
W=H=256
inp=common.scene_inputs((2,1,2))
g=common.setup(vpt.Vpt(W,H),inp,spp=1,total=3,diffuse=1); o=common.setup(O.Oracle(W,H),inp,spp=1,total=3,diffuse=1)
p=S.default_denoising_params(); cam=common.scene_camera(W,H)
for f in range(3):
    g.render(cam,cam,f); o.render(cam,cam,f); g.denoise(p,cam,cam,f,f+1); o.denoise(p,cam,cam,f,f+1)
    a,b=g.read("HistoryLength"),o.read("HistoryLength")
    ys,xs=np.nonzero(a!=b)
    print(f, "mismatch", len(ys))
    if len(ys):
        d=o.read("Depth"); nr=o.read("NormalRoughness"); ph=o.read("PrimaryHits")
        for y,x in list(zip(ys,xs))[:12]:
            print("  px",x,y,"gpu",a[y,x],"orc",b[y,x],"depth",d[y,x],"n",nr[y,x,:3],"hit",ph[y,x], "nbr depths", d[max(y-1,0):y+2,max(x-1,0):x+2].ravel())

import sys, numpy as np
sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/hybrid-rendering_b200')
import oracle as O, pyhr
W,H=192,112
sc = pyhr.SynthScene(pyhr.SCENE_SHADOWS_TEST); ss=O.ShadingScene(sc, brute=True); bn=pyhr.blue_noise()
ctx=pyhr.Context(0); ctx.set_bluenoise(*bn); ctx.build_scene(sc); ctx.gbuffer_create(W,H)
mn,mx=sc.bounds()
dd=pyhr.DDGIPass(ctx,W,H,0); dd.params.probe_distance=4.0; dd.params.normal_bias=1.0
for k in range(3): dd.params.sky_color[k]=(0.3,0.4,0.6)[k]
odd=O.DDGIOracle(W,H,0,dd.params,mn,mx)
f=pyhr.make_frame((0,14,34),(0,3,0),W,H); g=pyhr.write_gbuffer(sc,f,W,H); ctx.gbuffer_upload(0,g); cur=O.GBufMips(g)
rot=pyhr.rotation_matrix(0.7,(0.3,1.0,-0.5))
dd.render(f,rot); odd.render(ss,cur,f,rot)
a=dd.download(3).astype(np.float32); b=O.h2f(odd.cur_dep)
d=np.abs(a-b); print("max",d.max(), "count>0.05", (d>0.05).sum(), "of", d.size)
ys,xs,cs=np.where(d>0.05)
for y,x,c in list(zip(ys,xs,cs))[:12]:
    print((y,x,c), a[y,x], b[y,x], "probe-local", (x-1)%18, (y-1)%18)
ra=dd.download(0).astype(np.float32); rb=O.h2f(odd.radiance).reshape(ra.shape); dr=np.abs(ra-rb)
print("radiance max", dr.max(), "count>0.01", (dr>0.01).sum())

import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
import conftest  # noqa
import numpy as np, torch
import pyhr
import test_gpu_multi as T

world, sh_scale, ao_scale = 2, 0, 1
W, H = T.W, T.H
sc = pyhr.SynthScene(pyhr.SCENE_SHADOWS_TEST)
ranks = [T.make_rank(0, sc, sh_scale, ao_scale, r, world) for r in range(world)]
for r in range(world):
    for q in range(world):
        if q != r:
            ranks[r][1].link_local(q, ranks[q][1]); ranks[r][2].link_local(q, ranks[q][2])
streams = [torch.cuda.Stream() for _ in range(world)]
mode = sys.argv[1] if len(sys.argv) > 1 else "both"
for f in T.frames(3, pan_from=2, vertical=0.35):
    g = pyhr.write_gbuffer(sc, f, W, H)
    for (c, sh, ao), st in zip(ranks, streams): c.gbuffer_upload(f.ping_pong, g, st.cuda_stream)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    try:
        for (c, sh, ao), st in zip(ranks, streams):
            ta = time.perf_counter(); sh.render(f, st.cuda_stream); print(f"  sh.render host {1e3*(time.perf_counter()-ta):.2f} ms")
        if mode == "both":
            for (c, sh, ao), st in zip(ranks, streams):
                ta = time.perf_counter(); ao.render(f, st.cuda_stream); print(f"  ao.render host {1e3*(time.perf_counter()-ta):.2f} ms")
    except Exception as e:
        print("EXC", e)
    torch.cuda.synchronize()
    print(f"frame {f.num_frames}: {1e3*(time.perf_counter()-t0):.1f} ms total")

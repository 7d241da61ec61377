"""Streaming host frames (hr_gbuffer_stage_upload / hr_gbuffer_commit_staged / hr_pass_download_async): the staged surface is
swapped into the ping-pong slots without a copy while the next frame's PCIe transfer is already running.  The frames must
come out bit-identical to the plain hr_gbuffer_upload path (same kernels, same inputs — only the plumbing differs)."""
import numpy as np
import pytest
import torch

import pyhr

pytestmark = pytest.mark.gpu


def _frames(sc, W, H, n):
    out, f = [], None
    for i in range(n):
        cam = (0.5 * i, 14.0 + 0.2 * i, 34.0 - 0.3 * i)  # moving camera: non-trivial motion vectors, history taps move
        f = pyhr.make_frame(cam, (0.0, 3.0, 0.0), W, H, prev=f, num_frames=i)
        out.append((f, pyhr.write_gbuffer(sc, f, W, H, pinned=True)))
    return out


def _run(sc, frames, W, H, streamed):
    ctx = pyhr.Context(0)
    outs = []
    try:
        ctx.set_bluenoise(*pyhr.blue_noise())
        ctx.build_scene(sc)
        ctx.gbuffer_create(W, H)
        sh, ao = pyhr.Pass(ctx, "shadows", W, H, 0), pyhr.Pass(ctx, "ao", W, H, 1)
        stream = torch.cuda.current_stream().cuda_stream
        if streamed:
            o_sh = [torch.empty((H, W, 2), dtype=torch.float16).pin_memory().numpy() for _ in frames]
            o_ao = [torch.empty((H, W), dtype=torch.float16).pin_memory().numpy() for _ in frames]
            ctx.gbuffer_stage_upload(frames[0][1])
            for i, (f, _) in enumerate(frames):
                ctx.gbuffer_commit_staged(f.ping_pong, stream)
                if i + 1 < len(frames):
                    ctx.gbuffer_stage_upload(frames[i + 1][1])  # overlaps this frame's render
                sh.render(f, stream)
                ao.render(f, stream)
                sh.download_async(100, o_sh[i], stream)
                ao.download_async(100, o_ao[i], stream)
            torch.cuda.synchronize()
            outs = [(a.copy(), b.copy()) for a, b in zip(o_sh, o_ao)]
        else:
            for f, g in frames:
                ctx.gbuffer_upload(f.ping_pong, g, stream)
                sh.render(f, stream)
                ao.render(f, stream)
                outs.append((sh.download(100, stream).copy(), ao.download(100, stream).copy()))
        sh.destroy()
        ao.destroy()
    finally:
        ctx.close()
    return outs


def test_streamed_frames_match_plain_upload():
    W, H = 320, 176
    sc = pyhr.SynthScene(pyhr.SCENE_SHADOWS_TEST, 0)
    frames = _frames(sc, W, H, 6)
    a = _run(sc, frames, W, H, streamed=False)
    b = _run(sc, frames, W, H, streamed=True)
    for i, ((sa, aa), (sb, ab)) in enumerate(zip(a, b)):
        assert np.array_equal(sa.view(np.uint16), sb.view(np.uint16)), f"frame {i}: shadows output differs between upload paths"
        assert np.array_equal(aa.view(np.uint16), ab.view(np.uint16)), f"frame {i}: AO output differs between upload paths"
    assert a[-1][0].view(np.uint16).any()


def test_stage_protocol_errors():
    W, H = 64, 48
    sc = pyhr.SynthScene(pyhr.SCENE_SHADOWS_TEST, 0)
    f = pyhr.make_frame((0.0, 14.0, 34.0), (0.0, 3.0, 0.0), W, H)
    g = pyhr.write_gbuffer(sc, f, W, H, pinned=True)
    ctx = pyhr.Context(0)
    try:
        ctx.gbuffer_create(W, H)
        with pytest.raises(RuntimeError):
            ctx.gbuffer_commit_staged(0)  # nothing staged
        ctx.gbuffer_stage_upload(g)
        with pytest.raises(RuntimeError):
            ctx.gbuffer_stage_upload(g)  # one staged frame at a time
        ctx.gbuffer_commit_staged(0)
        torch.cuda.synchronize()
        got = ctx.gbuffer_download(0, 0, 0, W, H)
        assert np.array_equal(got, g.depth)
    finally:
        ctx.close()

#!/usr/bin/env python
"""bench.py — denoised frames/s of the ray-trace + SVGF hot path at 4K on N B200s (one process per GPU).

A "step" is one frame of the hot path: shadows (full-res: K1 ray trace -> K3 temporal -> 4x K5 a-trous) and ambient
occlusion (reference default half-res: K7 -> K9 -> 2x K10 blur -> K11 upsample) over a synthetic 3840x2160 G-buffer of
the 262 144-triangle arcade scene, 1 ray / pixel / effect, static camera in steady state (history saturated at 32
frames; the blue-noise sample index advances every frame so the traced rays change every frame).

  value   frames/s with the G-buffer already resident in HBM (hr_gbuffer_bind_device + both passes per step)
  e2e     frames/s through the C ABI with HOST buffers, every step: pinned-host G-buffer over PCIe (streamed:
          hr_gbuffer_stage_upload of frame N+1 overlaps the render of frame N, hr_gbuffer_commit_staged swaps it in),
          both passes, device -> host copy of both denoised outputs.  e2e.serial_value is the same work with plain
          hr_gbuffer_upload / hr_pass_download (no overlap)
  roofline  the shadows a-trous kernel (K5): algorithmic 24 B/px/iteration / its CUDA-event duration
  cpu_baseline / --impl reference   the CPU oracle (a port; the reference has no CPU path and cannot be built here)
          timed on this box's host cores on a 1/16-area (960x540) render of the same scene, extrapolated x16.

Timing: W>=3 warm-up steps, K timed steps bracketed by barrier + cuda synchronize, CUDA events on the launch stream,
max over ranks.  Inputs (G-buffer 199 MB + history/intermediate images > 300 MB per frame) exceed the 126 MB L2, so
no explicit flush is needed ("inputs_larger_than_l2").
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "hybrid-rendering_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

CAM_POS, CAM_TGT = (0.0, 9.0, -4.0), (2.0, 7.0, 60.0)
LIGHT_ROT_X = 25.0
ATROUS_BYTES_PER_PX = 24.0  # RG16F in 4 + GB2 8 + GB3 8 + RG16F out 4
# measured DRAM bytes per a-trous launch at 4K on this workload (average of the four iterations: 75.4 / 76.0 / 76.9 / 80.4 MB),
# one `ncu --set full` capture, profiles/r1o_ncu_full_summary.csv — below the 199 MB algorithmic figure (tile skipping + L2)
ATROUS_TRAFFIC_BYTES = 77.2e6


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def oracle_fps(width, height, n_tris, frames, extrapolate):
    """CPU oracle (OpenMP over all host threads) on a (width x height) render; returns (frames/s at full size, threads, seconds)."""
    import oracle as O
    import pyhr
    sc = pyhr.SynthScene(pyhr.SCENE_ARCADE, n_tris)
    tri, _ = sc.world_triangles()
    osc = O.Scene(tri)
    bn = pyhr.blue_noise()
    light = pyhr.default_light(rot_x_deg=LIGHT_ROT_X)
    sh, ao = O.ShadowsOracle(width, height, 0), O.AOOracle(width, height, 1)
    f = pyhr.make_frame(CAM_POS, CAM_TGT, width, height, light=light)
    f = pyhr.make_frame(CAM_POS, CAM_TGT, width, height, prev=f, num_frames=1, light=light)
    g = O.GBufMips(pyhr.write_gbuffer(sc, f, width, height))
    sh.render(osc, g, g, f, bn)  # warm-up (also leaves valid history)
    ao.render(osc, g, g, f, bn)
    t0 = time.perf_counter()
    for i in range(frames):
        f = pyhr.make_frame(CAM_POS, CAM_TGT, width, height, prev=f, num_frames=2 + i, light=light)
        sh.render(osc, g, g, f, bn)
        ao.render(osc, g, g, f, bn)
    dt = (time.perf_counter() - t0) / frames
    return 1.0 / (dt * extrapolate), O.lib().orc_num_threads(), dt * frames


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--tris", type=int, default=262144)
    ap.add_argument("--ao-scale", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full", action="store_true", help="skip the shadows+AO+DDGI+reflections leg")
    ap.add_argument("--full-sharded", action="store_true", help="run the full-pipeline leg on sharded runs too (default: single GPU only)")
    args = ap.parse_args()
    W, H = args.width, args.height
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = f"{W}x{H} 1spp shadows(full-res)+AO(scale {args.ao_scale}) full SVGF, arcade {args.tris} tris, static camera steady state"
    config = {"workload": workload, "width": W, "height": H, "triangles": args.tris, "passes": ["shadows", "ao"], "spp": 1,
              "parallelism": f"row-band x{world}", "l2_policy": "inputs_larger_than_l2"}

    if args.impl == "reference":
        # the reference's own CPU implementation does not exist and the reference cannot be built here;
        # this arm times the oracle port with every host thread on a bounded sample of the same workload.
        if rank != 0:
            return
        sw, shh = W // 4, H // 4
        steps = max(1, min(args.steps, 3))
        fps, threads, secs = oracle_fps(sw, shh, args.tris, steps, (W * H) / (sw * shh))
        line = {"metric": "denoised frames/s @4K (shadows+AO, 1 spp, full SVGF)", "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps,
                "warmup": 1, "ms_per_step": 1000.0 / fps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (fp16 storage)",
                "data": "synthetic", "config": config, "impl": "reference",
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                                 "sample": f"{sw}x{shh} (1/16 area) render of the same scene, {steps} frames in {secs:.1f} s, extrapolated x16"},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import pyhr

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs (host, synthetic) ----------------------------------------------------------------------------
    light = pyhr.default_light(rot_x_deg=LIGHT_ROT_X)
    sc = pyhr.SynthScene(pyhr.SCENE_ARCADE, args.tris)
    f0 = pyhr.make_frame(CAM_POS, CAM_TGT, W, H, light=light)
    f1 = pyhr.make_frame(CAM_POS, CAM_TGT, W, H, prev=f0, num_frames=1, light=light)  # static camera: zero motion vectors
    g_host = pyhr.write_gbuffer(sc, f1, W, H, pinned=True)

    ctx = pyhr.Context(local_rank)
    if os.environ.get("HR_ATROUS_IMPL"):  # A/B switch for kernel experiments (0 naive, 1 tiled, 2 chain = default)
        ctx.lib.hr_debug_set(1, int(os.environ["HR_ATROUS_IMPL"]))
    if os.environ.get("HR_TRACE_IMPL"):  # 0 one warp per 8x4 block (default), 1 persistent threads + compaction
        ctx.lib.hr_debug_set(2, int(os.environ["HR_TRACE_IMPL"]))
    if os.environ.get("HR_FORCE_SHARED_RT"):  # single GPU: run the cooperative (multi-GPU) ray-trace kernel, for overhead A/B
        ctx.lib.hr_debug_set(4, int(os.environ["HR_FORCE_SHARED_RT"]))
    if os.environ.get("HR_ATROUS_ROWS"):  # 1 (default) row-interleaved tiles for a-trous steps 4 and 8, 0 dense tiles
        ctx.lib.hr_debug_set(5, int(os.environ["HR_ATROUS_ROWS"]))
    if os.environ.get("HR_BVH_QUALITY"):  # 0 Karras radix tree, 1 PLOC (default); must be set before the scene build
        ctx.lib.hr_debug_set(3, int(os.environ["HR_BVH_QUALITY"]))
    ctx.set_bluenoise(*pyhr.blue_noise())
    if world > 1:
        # row-band sharding with the library's own NCCL exchange: rank 0 creates the ncclUniqueId, torch.distributed ships it
        uid = [pyhr.shard_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.shard_init(rank, world, uid[0])
        # the denoised frame stays distributed: every rank keeps (and, in the e2e leg, downloads) its own band; the temporal
        # history is read through peer mappings, so no per-frame collective is left on the data path
        ctx.shard_set_gather(False)
    t_b = time.perf_counter()
    scene_h = ctx.build_scene(sc)
    torch.cuda.synchronize()
    print(f"[bench] scene build (upload + BVH): {(time.perf_counter() - t_b) * 1e3:.2f} ms", file=sys.stderr)
    ctx.gbuffer_create(W, H)
    sh = pyhr.Pass(ctx, "shadows", W, H, 0)
    ao = pyhr.Pass(ctx, "ao", W, H, args.ao_scale)
    stream = torch.cuda.current_stream().cuda_stream

    # device-resident copy of the G-buffer for the `value` leg
    d_gb1 = torch.from_numpy(g_host.gb1).cuda()
    d_gb2 = torch.from_numpy(g_host.gb2.view(np.int16)).cuda()
    d_gb3 = torch.from_numpy(g_host.gb3.view(np.int16)).cuda()
    d_depth = torch.from_numpy(g_host.depth).cuda()
    dev_desc = pyhr.hr_gbuffer_desc(W, H, d_gb1.data_ptr(), d_gb2.data_ptr(), d_gb3.data_ptr(), d_depth.data_ptr())

    band0, band1 = pyhr.shard_rows(H, rank, world)  # this rank's rows of the full-resolution outputs
    out_sh = torch.empty((band1 - band0, W, 2), dtype=torch.float16).pin_memory().numpy()
    out_ao = torch.empty((band1 - band0, W), dtype=torch.float16).pin_memory().numpy()

    state = {"f": f1, "n": 2}

    def next_frame():
        state["f"] = pyhr.make_frame(CAM_POS, CAM_TGT, W, H, prev=state["f"], num_frames=state["n"], light=light)
        state["n"] += 1
        return state["f"]

    # Sharded runs put the two (independent) passes on two streams: with 1/N of the rows per launch every kernel is short,
    # and the ramp / tail / peer-wait bubbles of one pass are filled by the other pass's kernels.  The single-GPU run keeps
    # one stream (the GPU is already busy and the per-kernel roofline timings stay undisturbed).
    overlap = world > 1 and not os.environ.get("HR_NO_PASS_OVERLAP")
    s_ao = torch.cuda.Stream() if overlap else None
    ev_bound, ev_ao_done = torch.cuda.Event(), torch.cuda.Event()
    config["pass_streams"] = 2 if overlap else 1

    def step_resident():
        f = next_frame()
        if not overlap:
            ctx.gbuffer_bind_device(f.ping_pong, dev_desc, stream)
            sh.render(f, stream)
            ao.render(f, stream)
            return
        cur = torch.cuda.current_stream()
        cur.wait_event(ev_ao_done)  # the slot re-bound now (its mips are rebuilt) was read as "previous" by last frame's AO pass
        ctx.gbuffer_bind_device(f.ping_pong, dev_desc, stream)
        ev_bound.record(cur)
        sh.render(f, stream)
        s_ao.wait_event(ev_bound)
        ao.render(f, s_ao.cuda_stream)
        ev_ao_done.record(s_ao)

    def join_passes():
        if overlap:
            torch.cuda.current_stream().wait_event(ev_ao_done)

    def step_e2e_serial():  # upload -> render -> download, one after the other on one stream
        f = next_frame()
        ctx.gbuffer_upload(f.ping_pong, g_host, stream)
        sh.render(f, stream)
        ao.render(f, stream)
        sh.download_rows_async(100, band0, band1, out_sh, stream)
        ao.download_rows_async(100, band0, band1, out_ao, stream)
        torch.cuda.current_stream().synchronize()

    def step_e2e():
        # streaming host frames: this frame's G-buffer was staged (PCIe copy on the library's upload stream) while the
        # previous frame rendered; commit it, start the next frame's copy, render, read the results back
        f = next_frame()
        ctx.gbuffer_commit_staged(f.ping_pong, stream)
        ctx.gbuffer_stage_upload(g_host)
        sh.render(f, stream)
        ao.render(f, stream)
        sh.download_rows_async(100, band0, band1, out_sh, stream)
        ao.download_rows_async(100, band0, band1, out_ao, stream)

    # history warm-up to steady state (both slots bound, history length saturates at 32)
    ctx.gbuffer_bind_device(0, dev_desc, stream)
    ctx.gbuffer_bind_device(1, dev_desc, stream)
    for _ in range(max(args.warmup, 3) + 30):
        step_resident()
    torch.cuda.synchronize()

    # ---- value: device-resident inputs ---------------------------------------------------------------------------
    ctx.set_profiling(True)
    sampler = ClockSampler(local_rank)
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    ev0.record()
    for _ in range(args.steps):
        step_resident()
    join_passes()
    ev1.record()
    barrier()
    clocks = sampler.result()
    ms_total = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - launches0
    sh_stages = sh.stage_times()
    ao_stages = ao.stage_times()
    ctx.set_profiling(False)

    # ---- e2e: host buffers through the C ABI ------------------------------------------------------------------------
    for _ in range(3):
        step_e2e_serial()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_e2e_serial()
    e1.record()
    barrier()
    ms_e2e_serial = e0.elapsed_time(e1)
    # streamed: every timed step commits one staged frame, starts the upload of the next, renders and downloads; the
    # frame staged before the clock starts is paid back by the one staged in the last step and never rendered
    ctx.gbuffer_stage_upload(g_host)
    for _ in range(3):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()  # synchronises the device: includes the upload stream's last copy
    ms_e2e = e0.elapsed_time(e1)

    # ---- full hybrid pipeline (BASELINE config 4 pass set at 1 spp): shadows + AO + DDGI (4096 probes x 256 rays) + reflections -----
    full = None
    # the informational full-pipeline leg runs on the single-GPU line only: the sharded lines are the scaling measurement of
    # the headline metric and stay free of the (NCCL-exchanged) reflections / DDGI passes unless asked for
    if not args.no_full and (world == 1 or args.full_sharded):
        dd = pyhr.DDGIPass(ctx, W, H, 0)
        rf = pyhr.ReflectionsPass(ctx, W, H, 1)
        dd.params.probe_distance, dd.params.normal_bias = 4.2, 0.5  # arcade bounds 60 x 28.6 x 128 => 16 x 8 x 32 = 4096 probes
        for P in (dd.params, rf.params):
            P.sky_color[0], P.sky_color[1], P.sky_color[2] = 0.3, 0.4, 0.6
        rng = np.random.default_rng(1234)

        def step_full():
            f = next_frame()
            ctx.gbuffer_bind_device(f.ping_pong, dev_desc, stream)
            sh.render(f, stream)
            ao.render(f, stream)
            ax = rng.uniform(-1, 1, 3)
            dd.render(f, pyhr.rotation_matrix(float(rng.uniform(0, 2 * np.pi)), ax / np.linalg.norm(ax)), stream)
            rf.render(f, dd, stream)

        for _ in range(12):
            step_full()
        barrier()
        ctx.set_profiling(True)
        for p_ in (sh, ao):
            p_.stage_times()
        f0e, f1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0e.record()
        for _ in range(args.steps):
            step_full()
        f1e.record()
        barrier()
        ms_full = f0e.elapsed_time(f1e)
        dd_st, rf_st = dict(dd.stage_times()), dict(rf.stage_times())
        ctx.set_profiling(False)
        ra = [v for k_, v in rf_st.items() if k_.startswith("A-Trous")]
        ra_ms = float(np.mean(ra)) if ra else None
        rb0, rb1 = pyhr.shard_rows(H // 2, 0, world)
        r_rows = (min(rb1 + 16, H // 2) - max(rb0 - 16, 0)) if world > 1 else H // 2
        r_bytes = 36.0 * (W // 2) * r_rows  # RGBA16F 8 + GB2 8 + GB3 8 + depth 4 read, RGBA16F 8 written
        n_probes = 1
        u_ = dd.uniforms()
        n_probes = u_.probe_counts[0] * u_.probe_counts[1] * u_.probe_counts[2]
        full = {"ms_total": ms_full, "ddgi": dd_st, "reflections": rf_st, "refl_atrous_ms": ra_ms, "refl_atrous_bytes": r_bytes, "probes": n_probes,
                "rays_per_probe": int(u_.rays_per_probe)}
        dd.destroy()
        rf.destroy()

    t = torch.tensor([ms_total, ms_e2e, full["ms_total"] if full else 0.0], dtype=torch.float64, device="cuda")
    # per-rank sum of the profiled stage times of one frame: shows the load imbalance between the row bands
    busy = torch.tensor([sum(ms for _, ms in sh_stages), sum(ms for _, ms in ao_stages), dict(sh_stages).get("Ray Trace", 0.0) + dict(ao_stages).get("Ray Trace", 0.0)],
                        dtype=torch.float64, device="cuda")
    busy_all = [busy.clone() for _ in range(world)]
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_gather(busy_all, busy)
    ms_total, ms_e2e, ms_full_max = float(t[0]), float(t[1]), float(t[2])

    if rank == 0:
        fps = args.steps / (ms_total / 1e3)
        fps_e2e = args.steps / (ms_e2e / 1e3)
        peak, peak_src = read_peaks()
        atrous = [ms for name, ms in sh_stages if name.startswith("A-Trous")]
        at_ms = float(np.mean(atrous)) if atrous else None
        px = W * H
        # rows rank 0's a-trous launches cover: its band +- 16 halo rows when sharded, the whole image otherwise
        b0, b1 = pyhr.shard_rows(H, 0, world)
        at_rows = (min(b1 + 16, H) - max(b0 - 16, 0)) if world > 1 else H
        at_px = W * at_rows
        achieved = (ATROUS_BYTES_PER_PX * at_px / 1e9) / (at_ms / 1e3) if at_ms else None
        rays_per_frame = px + (W >> args.ao_scale) * (H >> args.ao_scale)  # upper bound: one ray per non-sky pixel per effect
        rt_ms = dict(sh_stages).get("Ray Trace", 0.0) + dict(ao_stages).get("Ray Trace", 0.0)
        line = {
            "metric": "denoised frames/s @4K (shadows+AO, 1 spp, full SVGF)", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3) + 30, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 (fp16 storage)", "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": int(g_host.nbytes()), "d2h_bytes_per_step": int(W * H * 6),
                    "mode": "streamed: hr_gbuffer_stage_upload of frame N+1 overlaps the render of frame N (pinned host buffers)",
                    "serial_value": args.steps / (ms_e2e_serial / 1e3)},
            "rank_stage_sums_ms": [{"shadows": round(float(b[0]), 4), "ao": round(float(b[1]), 4), "ray_trace": round(float(b[2]), 4)} for b in busy_all],
            "roofline": {"kernel": "k_atrous_v3 (shadows a-trous, K5)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": ATROUS_TRAFFIC_BYTES if world == 1 else None,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, profiles/r1o_ncu_full_summary.csv",
                         "peak_source": peak_src,
                         "avg_launch_ms": at_ms, "algorithmic_bytes_per_launch": ATROUS_BYTES_PER_PX * at_px},
            "stages_ms": {"shadows": dict(sh_stages), "ao": dict(ao_stages)},
            "mrays_per_s": {"primary_rays_per_frame_upper_bound": rays_per_frame, "trace_kernels_ms": rt_ms,
                            "value": (rays_per_frame / 1e6) / (rt_ms / 1e3) if rt_ms else None},
        }
        if full:
            ra_ach = (full["refl_atrous_bytes"] / 1e9) / (full["refl_atrous_ms"] / 1e3) if full["refl_atrous_ms"] else None
            n_gi_rays = full["probes"] * full["rays_per_probe"]
            line["full_pipeline"] = {
                "passes": ["shadows(full)", "ao(half)", f"ddgi({full['probes']} probes x {full['rays_per_probe']} rays, full-res sample)", "reflections(half)"],
                "value": args.steps / (ms_full_max / 1e3), "unit": "frames/s", "ms_per_step": ms_full_max / args.steps,
                "stages_ms": {"ddgi": full["ddgi"], "reflections": full["reflections"]},
                "roofline_reflections_atrous": {"kernel": "k_refl_atrous (K16)", "bound": "hbm", "achieved": ra_ach, "peak": peak, "unit": "GB/s",
                                                "frac": (ra_ach / peak) if ra_ach else None, "avg_launch_ms": full["refl_atrous_ms"],
                                                "algorithmic_bytes_per_launch": full["refl_atrous_bytes"]},
                "mrays_per_s": {"ddgi_primary": (n_gi_rays / 1e6) / (full["ddgi"].get("Ray Trace", 0.0) / 1e3) if full["ddgi"].get("Ray Trace") else None,
                                "reflections_primary_upper_bound": ((W // 2) * (H // 2) / 1e6) / (full["reflections"].get("Ray Trace", 0.0) / 1e3)
                                if full["reflections"].get("Ray Trace") else None}}
        if not args.no_cpu_baseline and world == 1:
            sw, shh = W // 4, H // 4
            cfps, threads, secs = oracle_fps(sw, shh, args.tris, 2, (W * H) / (sw * shh))
            line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": threads, "kind": "port",
                                    "sample": f"{sw}x{shh} (1/16 area) render of the same scene, 2 frames in {secs:.1f} s, extrapolated x16"}
        print(json.dumps(line))
    sh.destroy()
    ao.destroy()
    ctx.lib.hr_scene_destroy(scene_h)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json configs[1], "C2"):
synthetic 1,001,112-triangle mesh, 1920x1080 primary + shadow ray casting through the CUDA layer.

    python bench.py --gpus 1 --steps 20 --warmup 3                   # our arm (sm_100a kernels behind the C ABI)
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1   # the reference's own CPU code on the host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one ray-casting pass over the whole screen: K1 eye rays -> K2 closest hit -> shadow rays to a point light ->
K2s any hit (hc_raycast_pass).  value = (primary + shadow rays) / device time, inputs resident in HBM; e2e = the same pass
through the reference-facing call sequence with HOST buffers: upload of the EngineGlobals blob (what RenderDriverRTE::Draw
does before every pass, RenderDriverRTE.cpp:1723-1725) + pass + read-back of the hit and visibility buffers to pinned host
memory.  At N > 1 every rank casts its own full frame (one process per GPU, replicated scene, no data-path collective:
the reference's process-per-GPU mode, README.md:99-103) — weak scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT = 1920, 1080
METRIC, UNIT = "Mrays/s", "Mrays/s"
WORKLOAD = "C2: synthetic 1,001,112-triangle mesh, 1920x1080 primary + shadow ray casting"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms; only samples inside [t_load, t_end] (GPU under our load) are kept."""

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self, t_load=None, t_end=None):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), [nme for k, nme in enumerate(names) if f[4 + k].lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.path)
        inside = [r for r in rows if (t_load is None or r[0] >= t_load) and (t_end is None or r[0] <= t_end + 0.02)]
        use = inside if inside else rows
        if use:
            out["sm_mhz"] = float(np.median([r[1] for r in use]))
            out["sm_max_mhz"] = float(max(r[2] for r in use))
            out["reasons"] = sorted({x for r in use for x in r[3]})
            out["samples"] = len(inside)
        return out


def _dist():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


# ------------------------------------------------------------------------------------------------------------ reference arm
def reference_rays(scn, ref):
    """Primary rays of the frame (zero jitter) from the reference's own MakeRandEyeRay."""
    from tests import refapi, scenes
    W, H = scn.width, scn.height
    pd = ref.make_rand_eye_rays(scn.globals_blob, W, H, scenes.pixel_grid(W, H), np.zeros((W*H, 4), np.float32))
    return refapi.rays_from_pos_dir(pd)


def reference_step(scn, ref, rays):
    """One pass of the reference CPU path over `rays`: closest hit (BVH4InstTraverse via IntegratorCommon::rayTrace semantics),
    shadow rays to the point light, shadow query (IntegratorCommon::shadowTrace semantics).  Returns rays traced."""
    from hydracore_b200.scene import C2_LIGHT_POS
    nodes, tris = scn.bvh["nodes"], scn.bvh["tris"]
    hits = ref.trace_closest(nodes, tris, rays)
    hit = hits["primId"] >= 0
    pos = rays[:, 0:3] + rays[:, 4:7]*hits["t"][:, None]
    L = np.array(C2_LIGHT_POS, np.float32)
    d = L - pos[hit]
    dist = np.sqrt((d*d).sum(1))
    sd = d/dist[:, None]
    eps = np.maximum(np.abs(pos[hit]).max(1), 1.0)*np.float32(1e-4)
    sp = pos[hit] + sd*eps[:, None]
    sh = np.zeros((sp.shape[0], 8), np.float32)
    sh[:, 0:3] = sp
    sh[:, 4:7] = sd
    sh[:, 7] = np.sqrt(((sp - L)**2).sum(1))*np.float32(0.995)
    ref.trace_shadow(nodes, tris, sh)
    return rays.shape[0] + sh.shape[0]


def cpu_baseline(scn, budget_s=12.0):
    """The reference's own CPU implementation (oracle/_ref) timed on this box's host cores over a bounded sample of the workload.
    Also returns the traversal work counters (quads / leaves / triangles per primary ray) from the oracle restatement,
    which define the algorithmic bytes of the roofline."""
    from tests import refapi
    ref = refapi.Ref.try_load()
    kind = "reference"
    if ref is None:
        raise RuntimeError("oracle/_ref/libhydra_ref.so is missing (built by __graft_entry__.build() where /root/reference exists)")
    cores = len(os.sched_getaffinity(0))
    cores = min(cores, 32) if False else cores
    rays = reference_rays(scn, ref)
    n = rays.shape[0]
    # calibrate on 1/32 of the frame (strided rows keep the sample representative), then size the sample for ~budget seconds
    probe = rays[::32]
    t0 = time.perf_counter()
    traced = reference_step(scn, ref, probe)
    dt = time.perf_counter() - t0
    rate = traced/dt
    frac = min(1.0, budget_s*rate/(2.0*n))
    stride = max(1, int(round(1.0/frac)))
    sample = rays[::stride]
    t0 = time.perf_counter()
    traced = reference_step(scn, ref, sample)
    dt = time.perf_counter() - t0
    orc = refapi.Oracle()
    sub = rays[::61]
    _h, cnt = orc.trace_closest(scn.bvh["nodes"], scn.bvh["tris"], sub, count=True)
    qlt = [float(c)/sub.shape[0] for c in cnt]
    return {"value": traced/dt/1e6, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "every %d-th pixel of the 1080p frame: %d primary + shadow rays in %.2f s; closest hit = reference BVH4InstTraverse "
                      "(Embree unavailable), OpenMP over all host threads" % (stride, traced, dt)}, qlt


def run_reference(args):
    world, rank, _local = _dist()
    if rank != 0:
        return 0
    import __graft_entry__ as g  # noqa: F401  (libraries are prebuilt; the bvh builder lives in the product .so)
    from hydracore_b200 import scene as S
    from tests import refapi
    ref = refapi.Ref.try_load()
    if ref is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libhydra_ref.so missing"}))
        return 0
    scn = S.scene_c2(WIDTH, HEIGHT)
    rays = reference_rays(scn, ref)
    cores = len(os.sched_getaffinity(0))
    # bounded sample per step: calibrate so that (steps + warmup) x sample stays within ~2 minutes
    t0 = time.perf_counter()
    traced = reference_step(scn, ref, rays[::64])
    rate = traced/(time.perf_counter() - t0)
    per_step_s = min(20.0, 120.0/max(1, args.steps + args.warmup))
    stride = max(1, int(round(2.0*rays.shape[0]/(per_step_s*rate))))
    sample = rays[::stride]
    for _ in range(args.warmup):
        reference_step(scn, ref, sample)
    t0 = time.perf_counter()
    traced = 0
    for _ in range(args.steps):
        traced += reference_step(scn, ref, sample)
    dt = time.perf_counter() - t0
    v = traced/dt/1e6
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3*dt/args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOAD, "resolution": [WIDTH, HEIGHT], "triangles": 1001112,
                                            "sample": "every %d-th pixel per step" % stride},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference",
                             "sample": "every %d-th pixel of the frame per step, %d rays per step" % (stride, traced//max(1, args.steps))},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import hydracore_b200 as hc
    from hydracore_b200 import scene as S
    from hydracore_b200._lib import HC_HOST, HC_DEVICE
    world, rank, local = _dist()
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # NCCL writes its version / debug lines to stdout by default: keep stdout = the one JSON line
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    scn = S.scene_c2(WIDTH, HEIGHT)
    lay = hc.CudaLayer(device=local)
    lay.LoadScene(scn)
    n = WIDTH*HEIGHT
    light = S.C2_LIGHT_POS

    # device-resident result buffers owned by torch (so that NCCL / torch can see them) and pinned host mirrors for e2e
    hits_d = torch.empty(n*4, dtype=torch.int32, device=dev)
    vis_d = torch.empty(n, dtype=torch.uint8, device=dev)
    hits_h = torch.empty(n*4, dtype=torch.int32, pin_memory=True)
    vis_h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    flush = torch.empty(256*1024*1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        lay.RaycastPass(light, hits_d.data_ptr(), vis_d.data_ptr(), HC_DEVICE)
        return lay.last_trace_ms()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    # untimed pre-roll of the same step (>= 0.4 s): nvidia-smi needs ~0.1 s to come up and a 20-step timed region lasts only tens of ms;
    # clock samples are kept from here to the end of the timed region, i.e. only while the GPU runs this workload
    t_load = time.time()
    while time.time() - t_load < (0.0 if args.profile else 0.4):
        step_device()
    lay.ResetPerfCounters()
    barrier()
    ms = []
    for _ in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (outside the per-step CUDA events)
        torch.cuda.synchronize()
        ms.append(step_device())            # device time of the 4 kernels of this step, CUDA events on the launching stream
    barrier()
    clocks = sampler.stop(t_load, time.time())
    stats = lay.GetRaysStat()
    shadow_rays = int(vis_d.numel())        # one shadow ray slot per pixel (missed pixels carry t_far = 0 and are not traced)
    n_hit = int((hits_d.view(-1, 4)[:, 1] >= 0).sum().item())
    rays_per_step = n + n_hit
    t_step = float(np.sum(ms))/1e3
    if dist is not None:
        t = torch.tensor([t_step], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_step = float(t.item())
        tot = torch.tensor([float(rays_per_step)], device=dev, dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        rays_all = float(tot.item())
    else:
        rays_all = float(rays_per_step)
    value = rays_all*args.steps/t_step/1e6
    ms_per_step = 1e3*t_step/args.steps

    # ---- measured denominators beside MEASURED_PEAKS.json: L2 read bandwidth (48 MiB working set, 20 sweeps) and HBM read bandwidth (2 GiB, 1 sweep)
    mem_peaks = None
    if rank == 0:
        mem_peaks = {"l2_read_gbs": lay.measure_read_bandwidth(48 << 20, 20), "hbm_read_gbs": lay.measure_read_bandwidth(2 << 30, 1),
                     "how": "hc_measure_read_bandwidth: 128-bit ld.global.cg streaming reads, 148 x 8 CTAs, best of 5"}

    # ---- e2e: globals upload + pass + read-back to pinned host memory, wall clock around synchronous API calls
    blob = torch.from_numpy(scn.globals_blob.copy()).pin_memory()
    blob_np = blob.numpy()

    def step_e2e():
        lay.PrepareEngineGlobals(blob_np)
        lay.RaycastPass(light, hits_h.data_ptr(), vis_h.data_ptr(), HC_HOST)

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    te = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([te], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        te = float(t.item())
    e2e_value = rays_all*args.steps/te/1e6
    # parity guard: the host copy must agree with the device-resident result
    assert torch.equal(hits_h, hits_d.cpu()) and torch.equal(vis_h, vis_d.cpu())

    # ---- C3 / C4 (reported next to the headline): MISPT on the 1M-triangle terrain with mixed materials + 2 area lights, and on the
    # 20M-triangle instanced scene (fully incoherent diffuse secondary rays, compaction + material sort in the loop).  Image plane in
    # interleaved 32x32 tiles across the ranks, HDR sums combined by one NCCL reduce: strong scaling of one frame.
    extras = {}
    if not args.no_c3:
        from hydracore_b200 import multigpu as MG
        lay.close()
        for key, label, build in (("c1", "C1: hydra_app/tests/test_42 (25,612 triangles, Lambert / Phong blend / emissive, rect area light, DOF), unidirectional PT, 512x512",
                                   lambda: __import__("hydracore_b200.hydra_scene", fromlist=["x"]).build_scene(
                                       __import__("hydracore_b200.hydra_scene", fromlist=["x"]).load_fixture(os.path.join(ROOT, "tests", "golden", "hydra_scenes.npz"), "test_42"), 512, 512)),
                                  ("c3", "C3: MISPT trace_depth 8, Lambert/GGX/glass/blend + 2 area lights, 1,001,116 triangles, 1080p, 32x32 interleaved tiles",
                                   lambda: S.scene_c3(WIDTH, HEIGHT)),
                                  ("c4", "C4: MISPT trace_depth 5 on 200 instances x 100,352 triangles = 20,070,400 instanced triangles, Lambert, 1080p, 32x32 interleaved tiles",
                                   lambda: S.scene_c4(WIDTH, HEIGHT)),
                                  ("c5", "C5: MISPT-QMC (Sobol-Niederreiter screen + lens dimensions) on the C3 scene at 3840x2160; rank g takes the sample indices "
                                         "i = g (mod G) of every pass into a full-size SUM buffer, one NCCL reduce of 8,294,400 x float4",
                                   lambda: S.scene_c3(3840, 2160))):
            if args.profile and key != "c3":
                continue                               # profiler runs: the C2 steps and the C3 passes only (keeps the ncu launch list short and stable)
            scn3 = build()
            lay = hc.CudaLayer(device=local)
            lay.LoadScene(scn3)
            integ = {"c1": 0, "c5": 3}.get(key, 2)     # C1: unidirectional PT (INTEGRATOR_PT = 0); C3 / C4: MISPT (= 2); C5: MISPT-QMC (= 3)
            lay.SetTiles(32, rank, world)
            lay.InitPathTracing(777)
            lay.TracingPass(integ, 2)                  # warm-up passes
            lay.ResetPerfCounters()
            barrier()
            passes = 64 if key == "c1" else 4
            t0 = time.perf_counter()
            lay.TracingPass(integ, passes)
            barrier()
            t_pass = time.perf_counter() - t0
            st3 = lay.GetRaysStat()
            ev_ms = st3["msClosest"] + st3["msShadow"] + st3["msShade"] + st3["msOther"]
            t0 = time.perf_counter()
            fb = MG.reduce_framebuffer(lay, dist, dev, dst=0)
            t_red = time.perf_counter() - t0
            vals = torch.tensor([t_pass, ev_ms, t_red, float(st3["paths"]), float(st3["raysClosest"]), float(st3["raysShadow"]),
                                 st3["msClosest"], st3["msShadow"], st3["msShade"], st3["msOther"]], device=dev, dtype=torch.float64)
            mx, sm = vals.clone(), vals.clone()
            if dist is not None:
                dist.all_reduce(mx, op=dist.ReduceOp.MAX)
                dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            mx, sm = mx.tolist(), sm.tolist()
            mean_img = float((fb.view(-1, 4)[:, :3].sum()/(scn3.width*scn3.height*3*(passes + 2))).item()) if rank == 0 else 0.0
            extras[key] = {"workload": label, "passes": passes, "ms_per_pass_wall_max": 1e3*mx[0]/passes, "ms_per_pass_device_max": mx[1]/passes,
                           "paths_per_s": sm[3]/mx[0], "mrays_per_s": (sm[4] + sm[5])/mx[0]/1e6,
                           "rays_closest_per_pass": sm[4]/passes, "rays_shadow_per_pass": sm[5]/passes,
                           "stage_ms_per_pass_max": {"closest": mx[6]/passes, "shadow_added": mx[7]/passes, "shade": mx[8]/passes, "raygen_sort": mx[9]/passes},
                           "stage_note": "closest-hit and any-hit launches of a bounce overlap on two streams: shadow_added = time from the end of the closest-hit launch to the join",
                           "reduce_ms": 1e3*mx[2], "reduce_bytes": scn3.width*scn3.height*16 if world > 1 else 0, "mean_radiance": mean_img,
                           "scaling": "strong (one frame split over the ranks)"}
            if key == "c5":
                extras[key]["spp_per_s"] = passes/mx[0]            # one pass of all ranks together = 1 sample per pixel of the 4K frame
                extras[key]["partition"] = "Sobol sample index i = g (mod G); samples land on arbitrary pixels, so every rank keeps a full-size SUM buffer"
            if key == "c1" and rank == 0 and world == 1 and not args.profile and not args.no_cpu_baseline:
                # the reference's own CPU integrator (IntegratorStupidPT compiled in place, oracle/_ref) on the same scene, host cores
                from tests import refapi
                rf = refapi.Ref.try_load()
                if rf is not None:
                    rs = rf.scene(scn3)
                    rs.render(0, 777, 1)
                    t0 = time.perf_counter()
                    _img, npass = rs.render(0, 777, 2)
                    dtc = time.perf_counter() - t0
                    rs.close()
                    extras[key]["cpu_reference"] = {"paths_per_s": scn3.width*scn3.height*2/dtc, "cores": len(os.sched_getaffinity(0)), "kind": "reference",
                                                    "sample": "2 passes of 512x512 by IntegratorStupidPT (oracle/_ref, OpenMP, BVH4InstTraverse instead of Embree)"}
            lay.close()
        lay = None

    if rank != 0:
        if lay is not None:
            lay.close()
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (K2 closest hit) + CPU baseline (N = 1 only)
    peak, peak_src = _peaks()
    qlt_path = os.path.join(ROOT, "profiles", "c2_algorithmic_bytes.json")
    cpu = None
    if world == 1 and not args.no_cpu_baseline and not args.profile:
        cpu, qlt = cpu_baseline(scn)
        try:
            json.dump({"quads_per_ray": qlt[0], "leaves_per_ray": qlt[1], "tris_per_ray": qlt[2],
                       "formula": "B_ray = 32 + 4 + 16 + 128*Q + 16*L + 48*T (SURVEY.md 8d)",
                       "source": "oracle/hydra_oracle.cpp traversal counters on every 61st primary ray of the C2 frame (bench.py, N=1)"},
                      open(qlt_path, "w"), indent=1)
        except OSError:
            pass
    else:
        d = json.load(open(qlt_path))
        qlt = [d["quads_per_ray"], d["leaves_per_ray"], d["tris_per_ray"]]
    bytes_per_ray = 52.0 + 128.0*qlt[0] + 16.0*qlt[1] + 48.0*qlt[2]
    ms_closest = stats["msClosest"]/args.steps
    achieved = bytes_per_ray*n/(ms_closest*1e-3)/1e9
    traffic, ncu_extra = None, None
    tp = os.path.join(ROOT, "profiles", "c2_ncu_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic = tj.get("k_trace_closest_dram_bytes_per_launch")
        ncu_extra = {k: tj[k] for k in ("issue_active_pct", "lsu_wavefronts_pct_of_peak", "active_threads_per_warp_instruction", "l1_hit_pct", "l2_hit_pct",
                                        "l2_to_l1_bytes_per_launch", "l1_load_bytes_per_launch", "source") if k in tj}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "resolution": [WIDTH, HEIGHT], "triangles": 1001112, "rays_per_step_per_gpu": rays_per_step,
                       "l2": "flushed between steps (256 MiB memset outside the per-step CUDA events)",
                       "parallelism": "replicated scene, one full frame per GPU, no collective" if world > 1 else "single GPU"},
            "mrays_primary": n/(stats["msClosest"]/args.steps)/1e3, "mrays_shadow": n_hit/(stats["msShadow"]/args.steps)/1e3,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(blob_np.nbytes), "d2h_bytes_per_step": int(n*16 + n),
                    "ms_per_step": 1e3*te/args.steps},
            "gpu_launches": int(stats["kernelLaunches"]),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved/peak, "traffic": traffic,
                         "kernel": "k_trace<closest>", "bytes_per_ray": bytes_per_ray, "quads_leaves_tris_per_ray": qlt,
                         "ms_per_launch": ms_closest, "peak_source": peak_src,
                         "note": "algorithmic bytes are BVH/triangle fetches that L1 (84 % hit) and the 126 MB L2 serve, so frac can exceed 1; "
                                 "the kernel is bound by instruction issue and the L1/LSU wavefront rate (fields below, from the ncu capture in profiles/)",
                         "measured_memory_peaks": mem_peaks, "ncu": ncu_extra,
                         "frac_vs_measured_l2_read": (achieved/mem_peaks["l2_read_gbs"]) if mem_peaks else None}}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    line.update(extras)
    print(json.dumps(line))
    if lay is not None:
        lay.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="profiling run (ncu): no clock pre-roll, no CPU baseline")
    ap.add_argument("--no-c3", action="store_true", help="skip the C3 / C4 (MISPT path tracing) sections")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

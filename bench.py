#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json configs[1], "C2"):
synthetic 1,001,112-triangle mesh, 1920x1080 primary + shadow ray casting through the CUDA layer.

    python bench.py --gpus 1 --steps 20 --warmup 3                   # our arm (sm_100a kernels behind the C ABI / the C++ IHWLayer)
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1   # the reference's own CPU code on the host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one ray-casting pass over the whole screen: K1 eye rays -> K2 closest hit -> shadow rays to a point light ->
K2s any hit (hc_raycast_pass).  value = (primary + shadow rays of the frame) / device time, inputs resident in HBM; e2e = the same pass
driven through the reference's plugin interface - the C++ GPUCUDALayer : IHWLayer (PrepareEngineGlobals / Tables as RenderDriverRTE::Draw
does before every pass, RenderDriverRTE.cpp:1723-1725, then BeginTracingPass / EndTracingPass) - with HOST result buffers.
At N > 1 the SAME frame is split over the GPUs in interleaved 32x32 tiles (one process per GPU, scene replicated): every rank casts the
rays of its tiles and the hit / visibility records are gathered on rank 0 over NVLink by the library (NCCL send / recv of the owned
pixels, hc_comm.cu) inside the timed step - strong scaling of one frame.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT = 1920, 1080
METRIC, UNIT = "Mrays/s", "Mrays/s"
WORKLOAD = "C2: synthetic 1,001,112-triangle mesh, 1920x1080 primary + shadow ray casting"
TILE = 32


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


_T0 = time.time()


def _log(msg):
    """progress marker on stderr (all ranks), so that a stuck multi-GPU run shows where it stopped"""
    sys.stderr.write("[bench %6.1fs rank %s] %s\n" % (time.time() - _T0, os.environ.get("RANK", "0"), msg))
    sys.stderr.flush()


def _dist():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


# ------------------------------------------------------------------------------------------------------------ reference arm
def _ref_threads(ref):
    """All host threads the reference code can use: the affinity mask, capped at the reference's own INTEGRATOR_MAX_THREADS_NUM = 32
    (CPUExp_Integrators.h:267).  Set explicitly: torchrun exports OMP_NUM_THREADS=1, which libgomp would otherwise obey."""
    return ref.omp_threads(min(len(os.sched_getaffinity(0)), 32))


def reference_rays(scn, ref):
    """Primary rays of the frame (zero jitter) from the reference's own MakeRandEyeRay."""
    from tests import refapi, scenes
    W, H = scn.width, scn.height
    pd = ref.make_rand_eye_rays(scn.globals_blob, W, H, scenes.pixel_grid(W, H), np.zeros((W*H, 4), np.float32))
    return refapi.rays_from_pos_dir(pd)


def reference_step(scn, ref, rays):
    """One pass of the reference CPU path over `rays`, entirely inside the reference library (oracle/ref_driver.cpp ref_raycast_step): closest
    hit (BVH4InstTraverse, IntegratorCommon::rayTrace semantics), one shadow ray per hit to the point light, shadow query
    (IntegratorCommon::shadowTrace semantics).  Returns rays traced."""
    from hydracore_b200.scene import C2_LIGHT_POS
    _h, _v, traced = ref.raycast_step(scn.bvh["nodes"], scn.bvh["tris"], rays, C2_LIGHT_POS)
    return traced


def cpu_baseline(scn, budget_s=12.0):
    """The reference's own CPU implementation (oracle/_ref) timed on this box's host cores over a bounded sample of the workload.
    Also returns the traversal work counters (quads / leaves / triangles per primary ray) from the oracle restatement,
    which define the algorithmic bytes of the roofline."""
    from tests import refapi
    ref = refapi.Ref.try_load()
    if ref is None:
        raise RuntimeError("oracle/_ref/libhydra_ref.so is missing (built by __graft_entry__.build() where /root/reference exists)")
    threads = _ref_threads(ref)
    rays = reference_rays(scn, ref)
    n = rays.shape[0]
    # calibrate on 1/32 of the frame (strided pixels keep the sample representative), then size the sample for ~budget seconds
    probe = np.ascontiguousarray(rays[::32])
    t0 = time.perf_counter()
    traced = reference_step(scn, ref, probe)
    rate = traced/(time.perf_counter() - t0)
    stride = max(1, int(round(2.0*n/(budget_s*rate))))
    sample = np.ascontiguousarray(rays[::stride])
    t0 = time.perf_counter()
    traced, reps = 0, 0
    while reps == 0 or (time.perf_counter() - t0 < 0.5*budget_s and reps < 64):      # the whole frame takes well under the budget on a many-core host: repeat it
        traced += reference_step(scn, ref, sample)
        reps += 1
    dt = time.perf_counter() - t0
    orc = refapi.Oracle()
    sub = np.ascontiguousarray(rays[::61])
    _h, cnt = orc.trace_closest(scn.bvh["nodes"], scn.bvh["tris"], sub, count=True)
    qlt = [float(c)/sub.shape[0] for c in cnt]
    return {"value": traced/dt/1e6, "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": "every %d-th pixel of the 1080p frame, %d time(s): %d primary + shadow rays in %.2f s; closest hit = reference BVH4InstTraverse "
                      "(Embree unavailable), OpenMP on %d threads (omp_get_max_threads)" % (stride, reps, traced, dt, threads)}, qlt


def cpu_baseline_pt(scn, kind, label, window):
    """The reference's own CPU integrator (oracle/_ref) on a pixel window of the same scene: paths/s on the host cores."""
    from tests import refapi
    ref = refapi.Ref.try_load()
    if ref is None:
        return None
    threads = _ref_threads(ref)
    rs = ref.scene(scn)
    x0, y0, x1, y1 = window
    if kind == 3:
        W = scn.width
        first, count = y0*W, (y1 - y0)*W                       # QMC passes take a range of sample indices (they land on arbitrary pixels)
        r = ref.L.ref_render_create(rs.h, 3, 777)
        rs._renders.append(r)
        ref.L.ref_render_pass_qmc_range(r, first, min(count, 4096))
        t0 = time.perf_counter()
        ref.L.ref_render_pass_qmc_range(r, first, count)
        dt = time.perf_counter() - t0
        paths = count
        what = "%d consecutive sample indices of one pass" % count
    else:
        rs.render(kind, 777, 1, window=(x0, y0, min(x1, x0 + 64), min(y1, y0 + 16)))        # warm-up
        npass = 8 if (x1 - x0)*(y1 - y0) <= 512*512 else 2
        t0 = time.perf_counter()
        rs.render(kind, 777, npass, window=window)
        dt = time.perf_counter() - t0
        paths = (x1 - x0)*(y1 - y0)*npass
        what = "pixel window [%d, %d) x [%d, %d), %d passes" % (x0, x1, y0, y1, npass)
    rs.close()
    return {"paths_per_s": paths/dt, "cores": threads, "kind": "reference",
            "sample": "%s by %s (oracle/_ref, OpenMP on %d threads, BVH4InstTraverse instead of Embree), %.2f s" % (what, label, threads, dt)}


def run_reference(args):
    world, rank, _local = _dist()
    if rank != 0:
        return 0
    import __graft_entry__ as g  # noqa: F401  (libraries are prebuilt; the bvh builder lives in the product .so)
    from hydracore_b200 import scene as S
    from tests import refapi
    ref = refapi.Ref.try_load()
    if ref is None:
        _emit({"impl": "reference", "unavailable": "oracle/_ref/libhydra_ref.so missing"})
        return 0
    threads = _ref_threads(ref)
    scn = S.scene_c2(WIDTH, HEIGHT)
    rays = reference_rays(scn, ref)
    # bounded sample per step: calibrate so that (steps + warmup) x sample stays within ~2 minutes
    probe = np.ascontiguousarray(rays[::64])
    t0 = time.perf_counter()
    traced = reference_step(scn, ref, probe)
    rate = traced/(time.perf_counter() - t0)
    per_step_s = min(20.0, 120.0/max(1, args.steps + args.warmup))
    stride = max(1, int(round(2.0*rays.shape[0]/(per_step_s*rate))))
    sample = np.ascontiguousarray(rays[::stride])
    for _ in range(args.warmup):
        reference_step(scn, ref, sample)
    t0 = time.perf_counter()
    traced = 0
    for _ in range(args.steps):
        traced += reference_step(scn, ref, sample)
    dt = time.perf_counter() - t0
    v = traced/dt/1e6
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3*dt/args.steps, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOAD, "resolution": [WIDTH, HEIGHT], "triangles": 1001112,
                                            "sample": "every %d-th pixel per step" % stride, "omp_threads": threads},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "reference",
                             "sample": "every %d-th pixel of the frame per step, %d rays per step, OpenMP on %d threads" % (stride, traced//max(1, args.steps), threads)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)
    return 0


# ------------------------------------------------------------------------------------------------------------ our arm
def _incoherent_rays(torch, dev, rays, hits, n):
    """Cosine-distributed secondary rays from the primary hit points: fully incoherent directions, origins in pixel order."""
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    h = hits.view(-1, 4)
    hit = h[:, 1] >= 0
    t = h[:, 0].view(torch.float32)
    r8 = rays.view(-1, 8)
    pos = r8[:, 0:3] + r8[:, 4:7]*t[:, None]
    nh = int(hit.sum().item())
    u = torch.rand(nh, 2, device=dev, generator=g)
    rr = torch.sqrt(u[:, 0])
    phi = 2*np.pi*u[:, 1]
    d = torch.stack([rr*torch.cos(phi), torch.sqrt(1 - u[:, 0]).clamp_min(1e-3), rr*torch.sin(phi)], 1)
    inc = torch.zeros(nh, 8, device=dev)
    inc[:, 0:3] = pos[hit] + torch.tensor([0, 1e-3, 0], device=dev)
    inc[:, 4:7] = d/d.norm(dim=1, keepdim=True)
    inc[:, 7] = 3.0e38
    return inc.contiguous(), nh


def run_ours(args):
    import torch
    import hydracore_b200 as hc
    from hydracore_b200 import scene as S
    from hydracore_b200 import multigpu as MG
    from hydracore_b200._lib import HC_DEVICE
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

    _log("process group up, building the C2 scene")
    scn = S.scene_c2(WIDTH, HEIGHT)
    lay = hc.CudaLayer(device=local)
    lay.LoadScene(scn)
    n = WIDTH*HEIGHT
    light = S.C2_LIGHT_POS
    if world > 1:
        lay.SetTiles(TILE, rank, world)              # this rank's interleaved 32x32 tiles of the ONE frame
        _log("scene loaded, joining the library's communicator")
        MG.join_communicator(lay, dist, dev)         # NCCL communicator inside the library (unique id carried by torch.distributed)
        _log("communicator joined")

    # device-resident result buffers owned by torch and pinned host mirrors for e2e (complete on rank 0)
    hits_d = torch.zeros(n*4, dtype=torch.int32, device=dev)
    vis_d = torch.zeros(n, dtype=torch.uint8, device=dev)
    hits_h = torch.empty(n*4, dtype=torch.int32, pin_memory=True)
    vis_h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    flush = torch.empty(256*1024*1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        lay.RaycastPass(light, hits_d.data_ptr(), vis_d.data_ptr(), HC_DEVICE)      # split frame: includes the gather of the records on rank 0
        return lay.last_trace_ms()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    _log("warm-up steps done")
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
        barrier()
        ms.append(step_device())            # device time of this step's launches (+ gather), CUDA events on the launching stream
    barrier()
    clocks = sampler.stop(t_load, time.time())
    _log("timed steps done")
    stats = lay.GetRaysStat()
    n_hit = int((hits_d.view(-1, 4)[:, 1] >= 0).sum().item())       # rank 0 holds the whole frame
    t_step = float(np.sum(ms))/1e3
    launches = float(stats["kernelLaunches"])
    if dist is not None:
        t = torch.tensor([t_step], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_step = float(t.item())
        c = torch.tensor([float(n_hit) if rank == 0 else 0.0, launches, stats["msClosest"], stats["msShadow"], stats["msOther"]], device=dev, dtype=torch.float64)
        mx = c.clone()
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        n_hit, launches = int(c[0].item()), float(c[1].item())
        ms_closest_max, ms_shadow_max, ms_other_max = float(mx[2].item()), float(mx[3].item()), float(mx[4].item())
    else:
        ms_closest_max, ms_shadow_max, ms_other_max = stats["msClosest"], stats["msShadow"], stats["msOther"]
    rays_per_step = n + n_hit                # one frame: every pixel's primary ray + one shadow ray per hit
    value = rays_per_step*args.steps/t_step/1e6
    ms_per_step = 1e3*t_step/args.steps

    # ---- incoherent secondary rays (N = 1): cosine-distributed about the up axis from the primary hit points, closest hit, L2 flushed
    mrays_incoherent = None
    if world == 1:
        rays = torch.empty(n*8, dtype=torch.float32, device=dev)
        lay.make_eye_rays_device(WIDTH, HEIGHT, rays.data_ptr())
        inc, nh = _incoherent_rays(torch, dev, rays, hits_d, n)
        hits2 = torch.empty(nh*4, dtype=torch.int32, device=dev)
        tms = []
        for k in range(3 + 10):
            flush.zero_()
            torch.cuda.synchronize()
            lay.trace_closest_device(inc.data_ptr(), nh, hits2.data_ptr())
            if k >= 3:
                tms.append(lay.last_trace_ms())
        mrays_incoherent = {"value": nh/float(np.median(tms))/1e3, "unit": "Mrays/s", "rays": nh, "ms": float(np.median(tms)),
                            "hit_fraction": float((hits2.view(-1, 4)[:, 1] >= 0).float().mean().item()),
                            "what": "closest hit of cosine-distributed secondary rays leaving the primary hit points of the C2 frame (pixel order), rays resident in HBM, L2 flushed, median of 10"}
        del rays, inc, hits2

    # ---- measured denominators beside MEASURED_PEAKS.json: L2 read bandwidth (48 MiB working set, 20 sweeps) and HBM read bandwidth (2 GiB, 1 sweep)
    mem_peaks = None
    if rank == 0:
        mem_peaks = {"l2_read_gbs": lay.measure_read_bandwidth(48 << 20, 20), "hbm_read_gbs": lay.measure_read_bandwidth(2 << 30, 1),
                     "how": "hc_measure_read_bandwidth: 128-bit ld.global.cg streaming reads, 148 x 8 CTAs, best of 5 (its ncu record: profiles/r02_final_l2_microbench_ncu.md)"}

    # ---- e2e through the reference's plugin interface: GPUCUDALayer : IHWLayer (C++), host result buffers
    e2e = None
    from tests import layerapi
    if layerapi.CppLayer.available():
        consts = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_consts.json")))
        _log("e2e: loading the scene through the C++ IHWLayer")
        cpp = layerapi.CppLayer(WIDTH, HEIGHT, 0, local)
        layerapi.load_scene_like_render_driver(cpp, scn, consts)
        varsI, varsF, flags = cpp.GetAllFlagsAndVars()
        cpp.SetAllFlagsAndVars(varsI, varsF, flags & ~consts["HRT_UNIFIED_IMAGE_SAMPLING"])      # ray-casting mode of the layer (BeginTracingPass without unified sampling)
        cpp.CallNamedFunc("raycast_light", "%r %r %r" % tuple(float(v) for v in light))
        if world > 1:
            cpp.CallNamedFunc("tiles", "%d %d %d" % (TILE, rank, world))
            idt = torch.zeros(256, dtype=torch.uint8, device=dev)
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(cpp.CommIdHex().encode()), dtype=torch.uint8))
            dist.broadcast(idt, src=0)
            cpp.CallNamedFunc("comm", "%d %d %s" % (rank, world, bytes(idt.cpu().numpy().tobytes()).decode()))
        cpp.CallNamedFunc("raycast_results", "%d %d" % ((hits_h.data_ptr(), vis_h.data_ptr()) if rank == 0 else (0, 0)))

        def step_e2e():
            cpp.PrepareEngineGlobalsAndTables()      # RenderDriverRTE::Draw re-assembles and re-uploads the globals before every pass
            cpp.TracingPasses(1)                     # BeginTracingPass + EndTracingPass

        _log("e2e: layer ready")
        for _ in range(3):
            step_e2e()
        barrier()
        _log("e2e: warm-up done")
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        te = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([te], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        if rank == 0:
            # parity guard: the host copy from the plugin path must agree with the device-resident result of the C-ABI path
            assert torch.equal(hits_h, hits_d.cpu()) and torch.equal(vis_h, vis_d.cpu()), "IHWLayer ray-casting results differ from the C-ABI path"
        e2e = {"value": rays_per_step*args.steps/te/1e6, "unit": UNIT, "h2d_bytes_per_step": int(scn.globals_blob.nbytes),
               "d2h_bytes_per_step": int(n*16 + n), "ms_per_step": 1e3*te/args.steps,
               "path": "GPUCUDALayer : IHWLayer (libhydra_cuda_layer.so): PrepareEngineGlobals + PrepareEngineTables + BeginTracingPass + EndTracingPass, results in host memory (rank 0)"}
        cpp.close()
    else:
        raise RuntimeError("hydracore_b200/cpp/_build/libhydra_cuda_layer.so is missing: the e2e leg goes through the C++ IHWLayer (build() makes it where /root/reference exists)")

    # ---- C1 / C3 / C4 / C5 (reported next to the headline): path tracing.  Image plane in interleaved 32x32 tiles across the ranks (C5: sample
    # indices modulo the world size), framebuffers combined on rank 0 by hc_fb_reduce INSIDE the timed frame: strong scaling of one frame.
    extras = {}
    if not args.no_c3:
        lay.close()
        HS = __import__("hydracore_b200.hydra_scene", fromlist=["x"])
        c1_build = lambda: HS.build_scene(HS.load_fixture(os.path.join(ROOT, "tests", "golden", "hydra_scenes.npz"), "test_42"), 512, 512)
        for key, label, build, window in (
                ("c1", "C1: hydra_app/tests/test_42 (25,612 triangles, Lambert / Phong blend / emissive, rect area light, DOF), unidirectional PT, 512x512",
                 c1_build, (0, 0, 512, 512)),
                ("c1_streams8", "C1 with 8 sample streams per pixel (hc_pt_set_sample_streams): 8 passes of the 512x512 frame share one wavefront of 2M paths",
                 c1_build, (0, 0, 512, 512)),
                ("c3", "C3: MISPT trace_depth 8, Lambert/GGX/glass/blend + 2 area lights, 1,001,116 triangles, 1080p, 32x32 interleaved tiles",
                 lambda: S.scene_c3(WIDTH, HEIGHT), (0, 270, 1920, 810)),
                ("c4", "C4: MISPT trace_depth 5 on 200 instances x 100,352 triangles = 20,070,400 instanced triangles, 6 materials (material sort on), 1080p, 32x32 interleaved tiles",
                 lambda: S.scene_c4(WIDTH, HEIGHT), (0, 270, 1920, 810)),
                ("c5", "C5: MISPT-QMC (Sobol-Niederreiter screen + lens dimensions) on the C3 scene at 3840x2160; rank g takes the sample indices "
                       "i = g (mod G) of every pass into a full-size SUM buffer, ncclReduce of 8,294,400 x float4 inside the frame",
                 lambda: S.scene_c3(3840, 2160), (0, 1000, 3840, 1256))):
            if args.profile and key != "c3":
                continue                               # profiler runs: the C2 steps and the C3 passes only (keeps the ncu launch list short and stable)
            _log("extras: building " + key)
            scn3 = build()
            lay = hc.CudaLayer(device=local)
            # sample streams: S generators per pixel, pass p draws from stream p mod S, so a rank that owns 1/G of the tiles keeps up to S passes in
            # flight as one wavefront (up to 8M paths by default: four 1080p passes on one GPU, all eight from two GPUs on).  The image depends on
            # (seed, S) only, so S is the SAME at every N.  C1 keeps the single-generator rule its 64 spp parity test is stated on; c1_streams8 shows the other.
            streams = {"c1": 1}.get(key, 8)
            lay.SetSampleStreams(streams)
            lay.LoadScene(scn3)
            integ = {"c1": 0, "c1_streams8": 0, "c5": 3}.get(key, 2)     # C1: unidirectional PT (INTEGRATOR_PT = 0); C3 / C4: MISPT (= 2); C5: MISPT-QMC (= 3)
            mode = 1 if key == "c5" else 0
            lay.SetTiles(TILE, rank, world)
            if world > 1:
                MG.join_communicator(lay, dist, dev)
            lay.InitPathTracing(777)
            warm = 8 if streams > 1 else 2
            lay.TracingPass(integ, warm)               # warm-up passes (one full cycle of the streams)
            _log("extras: " + key + " warm-up passes done, first reduce")
            lay.ReduceFramebuffer(0, mode)             # warm-up of the exchange (NCCL connects lazily)
            _log("extras: " + key + " reduce done")
            passes = 64 if key.startswith("c1") else (8 if streams > 1 else 4)
            # several frames, the MEDIAN frame is reported (frame time = max over ranks, reduce and the closing barrier included): with 8 processes on
            # the 32 vCPUs of the box single frames show host-side stalls of 1-5 ms on one rank or another (profiles/r02_streams_ranks_n8.log)
            frames = 1 if args.profile else 5
            rows = []
            for _f in range(frames):
                lay.ResetPerfCounters()
                barrier()
                t0 = time.perf_counter()
                lay.TracingPass(integ, passes)
                t_pass = time.perf_counter() - t0
                red_ms = lay.ReduceFramebuffer(0, mode)    # the frame is not complete before rank 0 holds it
                barrier()
                t_frame = time.perf_counter() - t0
                st3 = lay.GetRaysStat()
                ev_ms = st3["msClosest"] + st3["msShadow"] + st3["msShade"] + st3["msOther"]
                vals = torch.tensor([t_frame, ev_ms, red_ms, float(st3["paths"]), float(st3["raysClosest"]), float(st3["raysShadow"]),
                                     st3["msClosest"], st3["msShadow"], st3["msShade"], st3["msOther"], t_pass], device=dev, dtype=torch.float64)
                mx, sm, mn = vals.clone(), vals.clone(), vals.clone()
                if dist is not None:
                    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
                    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
                    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
                rows.append((mx.tolist(), sm.tolist(), mn.tolist()))
            rows.sort(key=lambda r: r[0][0])
            frame_ms = [1e3*r[0][0] for r in rows]
            mx, sm, mn = rows[len(rows)//2]
            mean_img = float(lay.GetSumImage()[..., :3].sum()/(scn3.width*scn3.height*3*(passes*frames + warm))) if rank == 0 else 0.0
            extras[key] = {"workload": label, "passes": passes, "frames_timed": frames, "frame_ms_min_median_max": [frame_ms[0], frame_ms[len(frame_ms)//2], frame_ms[-1]],
                           "sample_streams": streams, "passes_per_wavefront": lay.GroupPasses(), "ms_per_pass_wall_max": 1e3*mx[0]/passes, "ms_per_pass_device_max": (mx[1] + mx[2])/passes,
                           "ms_per_pass_without_reduce_wall_max": 1e3*mx[10]/passes,
                           "paths_per_s": sm[3]/mx[0], "mrays_per_s": (sm[4] + sm[5])/mx[0]/1e6,
                           "rays_closest_per_pass": sm[4]/passes, "rays_shadow_per_pass": sm[5]/passes,
                           "stage_ms_per_pass_max": {"closest": mx[6]/passes, "shadow_added": mx[7]/passes, "shade": mx[8]/passes, "raygen_sort": mx[9]/passes},
                           "stage_note": "closest-hit and any-hit launches of a bounce overlap on two streams: shadow_added = time from the end of the closest-hit launch to the join",
                           "reduce_ms": mn[2], "reduce_ms_incl_wait_for_slowest_rank": mx[2],
                           "reduce_note": "device time of hc_fb_reduce per rank: the minimum over ranks is the exchange itself (the last rank to arrive does not wait), the maximum includes the load imbalance of the frame",
                           "reduce": "hc_fb_reduce inside the frame: " + ("ncclReduce(sum) of the full-size buffers" if mode == 1 else "NCCL send / recv of the owned tiles only (1/G of the image per rank)"),
                           "reduce_bytes_per_rank": (scn3.width*scn3.height*16 if mode == 1 else scn3.width*scn3.height*16//world) if world > 1 else 0,
                           "mean_radiance": mean_img, "scaling": "strong (one frame split over the ranks, reduce included)"}
            if key == "c5":
                extras[key]["spp_per_s"] = passes/mx[0]            # one pass of all ranks together = 1 sample per pixel of the 4K frame
                extras[key]["partition"] = "Sobol sample index i = g (mod G); samples land on arbitrary pixels, so every rank keeps a full-size SUM buffer"
            if rank == 0 and world == 1 and not args.profile and not args.no_cpu_baseline and key != "c1_streams8":
                # the reference's own CPU integrator (oracle/_ref: the reference sources compiled in place) on a pixel window of the same scene
                names = {0: "IntegratorStupidPT", 2: "IntegratorMISPTLoop2", 3: "IntegratorMISPT_QMC"}
                cb = cpu_baseline_pt(scn3, integ, names[integ], window)
                if cb is not None:
                    extras[key]["cpu_reference"] = cb
                    extras[key]["vs_cpu_reference_paths"] = extras[key]["paths_per_s"]/cb["paths_per_s"]
            lay.close()
        lay = None

    if rank != 0:
        if lay is not None:
            lay.close()
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (K2 closest hit) + CPU baseline (N = 1 only)
    peak_hbm, peak_src = _peaks()
    qlt_path = os.path.join(ROOT, "profiles", "c2_algorithmic_bytes.json")
    cpu = None
    if world == 1 and not args.no_cpu_baseline and not args.profile:
        cpu, qlt = cpu_baseline(scn)
        qlt_src = "oracle traversal counters on every 61st primary ray of the C2 frame, this run"
    else:
        d = json.load(open(qlt_path))
        qlt = [d["quads_per_ray"], d["leaves_per_ray"], d["tris_per_ray"]]
        qlt_src = "profiles/c2_algorithmic_bytes.json (committed; the same counters from an N = 1 run)"
    bytes_per_ray = 52.0 + 128.0*qlt[0] + 16.0*qlt[1] + 48.0*qlt[2]
    ms_closest = ms_closest_max/args.steps
    rays_per_launch = n/world                    # this kernel's share of the frame on the slowest rank
    achieved = bytes_per_ray*rays_per_launch/(ms_closest*1e-3)/1e9
    ncu = None
    tp = os.path.join(ROOT, "profiles", "c2_ncu_traffic.json")
    if os.path.exists(tp):
        ncu = json.load(open(tp))
    l2_peak = mem_peaks["l2_read_gbs"] if mem_peaks else None
    traffic = ncu.get("k_trace_closest_dram_bytes_per_launch") if ncu else None
    roofline = {
        # the algorithmic bytes are BVH-node / triangle fetches that L1 and the 126 MB L2 serve; ncu shows the kernel bound by instruction issue and
        # L1-hit latency, not by a memory ceiling.  frac is therefore taken against the measured L2 read bandwidth; HBM is reported beside it
        "bound": "issue",
        "achieved": achieved, "peak": l2_peak, "unit": "GB/s", "frac": (achieved/l2_peak) if l2_peak else None,
        "peak_source": "L2 streaming-read bandwidth measured in this run (hc_measure_read_bandwidth); HBM peak beside it: " + peak_src,
        "hbm_peak": peak_hbm, "frac_vs_hbm_peak": achieved/peak_hbm,
        "traffic": traffic, "traffic_source": (ncu.get("source", "") + " - from profiles/, NOT measured in this run") if ncu else None,
        "hbm_frac_of_dram_traffic": (traffic/(ms_closest*1e-3)/1e9/peak_hbm) if traffic else None,
        "kernel": "k_trace<closest>", "bytes_per_ray": bytes_per_ray, "quads_leaves_tris_per_ray": qlt, "quads_leaves_tris_source": qlt_src,
        "rays_per_launch": rays_per_launch, "ms_per_launch": ms_closest,
        "ncu": {k: ncu[k] for k in ("issue_active_pct", "lsu_wavefronts_pct_of_peak", "active_threads_per_warp_instruction", "l1_hit_pct", "l2_hit_pct",
                                    "l2_to_l1_bytes_per_launch", "l1_load_bytes_per_launch", "source") if ncu and k in ncu},
        "measured_memory_peaks": mem_peaks}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "resolution": [WIDTH, HEIGHT], "triangles": 1001112, "rays_per_step": rays_per_step,
                       "l2": "flushed between steps (256 MiB memset outside the per-step CUDA events)",
                       "parallelism": ("one frame in interleaved %dx%d tiles over %d GPUs (replicated scene), hit / visibility records gathered on rank 0 by NCCL send / recv "
                                       "inside the step" % (TILE, TILE, world)) if world > 1 else "single GPU"},
            "mrays_primary": (n/world)/(ms_closest_max/args.steps)/1e3*world, "mrays_shadow": n_hit/(ms_shadow_max/args.steps)/1e3,
            "mrays_incoherent": mrays_incoherent["value"] if mrays_incoherent else None, "incoherent": mrays_incoherent,
            "gather_ms_per_step": (ms_other_max/args.steps) if world > 1 else 0.0,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline}
    if cpu is not None:
        line["cpu_baseline"] = cpu
    line.update(extras)
    _emit(line)
    if lay is not None:
        lay.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def _emit(line):
    """The ONE JSON line goes to the process's real stdout; everything else written to fd 1 meanwhile (NCCL prints its version line there from C)
    was redirected to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def _watchdog(limit_s):
    """A multi-GPU run that is stuck (a rank died, a collective never completes) must not hang the caller for ever: after limit_s seconds the
    process prints where it was and exits with status 3.  Never triggers on a healthy run (N = 8 takes about a minute)."""
    import threading

    def fire():
        _log("watchdog: no JSON line after %d s - giving up; stacks follow" % limit_s)
        try:
            import faulthandler
            faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
        except Exception:
            pass
        os._exit(3)
    t = threading.Timer(limit_s, fire)
    t.daemon = True
    t.start()
    return t


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="profiling run (ncu): no clock pre-roll, no CPU baseline")
    ap.add_argument("--no-c3", action="store_true", help="skip the C1 / C3 / C4 / C5 (path tracing) sections")
    args = ap.parse_args()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        _watchdog(900)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

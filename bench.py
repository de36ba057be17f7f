#!/usr/bin/env python
"""bench.py -- the headline benchmark: Mrays/s of the staircase render path (BASELINE.json config 3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload staircase|rtiow|raybatch]

A "step" is one frame: runRenderer(ns) over a scene that initRenderer has already put in HBM (`value`), or the whole
reference-facing call sequence from HOST buffers -- initRenderer + runRenderer + reading the frame -- (`e2e`).
Workload at N = 1: the procedural staircase-class mesh (~312 k triangles, BVH_00.04 layout, 17 levels x 5 per leaf),
1200x800, 100 spp, depth 64, the reference's own RNG seeding -- the frame is bit-identical to the reference kernel's,
so both arms trace exactly the same rays.  configs[1] (RTIOW spheres) is NOT the default because the reference's HEAD
contains no sphere renderer to put in the reference arm; it is available as --workload rtiow.
At N > 1 every rank renders `spp` samples of every pixel on its own RNG stream (weak scaling: N x spp in total) and
the un-normalised float4 sums are combined with ONE NCCL reduce to rank 0 at frame end.

The reference arm (--impl reference) is the reference's own CUDA kernel (oracle/_ref/libref.so, compiled unmodified
from /root/reference for sm_100a) driven through its three entry points by oracle/_ref/ref_driver on the same GPU (at N > 2
a step is a bounded 200-spp sample of the N x spp job: its frame time is linear in the sample count, the rate is reported):
the reference has no CPU path; a CPU restatement (oracle/cpu_oracle.cpp) is reported as `cpu_baseline`.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracing-optimized_b200", "python"))

FLT_MAX = 3.4028234663852886e38
EXTEND_BYTES_PER_RAY = 56   # queue entry 4 + {origin,rng} 16 + {dir,flags} 16 read; {t,u,v,id} 16 + shade-queue entry 4 written
SHADOW_BYTES_PER_RAY = 85   # entry 4 + {origin} 16 + {dir,dist} 16 + {contribution,flags} 16 read; colour 16 read + 16 written; pending 1
RESUME_BYTES = 64           # a parked ray: state 8 + hit 16 + entry 4 written, then ray 32 + the same 28 read again
BATCH_BYTES_PER_RAY = 52    # 32 in, 16 + 4 out
# Flop model: the order-exact walk costs 36 flops per dual-node visit (two slab tests); a visit of an 8-wide node of
# the renderer's own tree costs 8 x (6 fused multiply-adds + 4 min/max + 1 subtraction) + 12 of per-node setup = 148; a triangle
# test 48; 3 per ray for the reciprocal direction.
FLOPS_PER_DUAL_VISIT, FLOPS_PER_WIDE_VISIT, FLOPS_PER_TRI_TEST, FLOPS_PER_RAY = 36.0, 148.0, 48.0, 3.0
# sphere scenes (20 flops per sphere test): slab test of one box 6 sub + 6 mul + 6 min/max = 18; shading ~150 per ray
FLOPS_PER_BOX_TEST, FLOPS_PER_SPHERE_TEST, SPHERE_SHADE_FLOPS = 18, 20, 150
NCU_FILE = os.path.join(ROOT, "profiles", "r02", "trace_ncu.json")  # DRAM bytes of the dominant kernel's bulk launch (ncu --set full)


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE bulk launch of `kernel`, from the ncu capture kept under profiles/ (the
    capture is a separate run of the same frame: bench.py itself never runs under a profiler)."""
    if os.path.exists(NCU_FILE):
        d = json.load(open(NCU_FILE)).get(kernel)
        if d:
            return d
    return None


def shard_samples(ns_total, world):
    """Samples per rank: ns_total split as evenly as possible, larger shares first."""
    base, extra = divmod(ns_total, world)
    return [base + (1 if r < extra else 0) for r in range(world)]


def reduce_and_finalize(acc, ns_total, rank):
    """The frame-end exchange: one sum-reduce of the per-rank un-normalised radiance sums to rank 0, then / ns_total.
    `acc` is a torch tensor (..., 4). Used with NCCL on device buffers by this file and with gloo by tests/test_multi_rank.py."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(acc, dst=0, op=dist.ReduceOp.SUM)
    return acc / float(ns_total) if rank == 0 else None


def peaks():
    p = dict(hbm_gbs=6650.0, src="fallback (B200_PROFILING.md)", sm_max_mhz=1965.0)
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        m = json.load(open(path))
        p = dict(hbm_gbs=float(m["hbm_gbs"]), src="MEASURED_PEAKS.json", sm_max_mhz=float(m.get("sm_max_mhz", 1965.0)))
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                                          str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["unavailable"])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 7 and r[3 + k].lower().startswith("active") for r in self.rows)]
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=float(self.rows[0][1]), reasons=reasons, samples=len(sm))


def workload_params(args):
    if args.workload == "staircase":
        return dict(workload="staircase mesh %dx%d %dspp/GPU depth %d (BASELINE config 3)" % (args.nx, args.ny, args.spp, args.depth),
                    scene="procedural staircase detail %.2f, %d^2 textures, 5 prims/leaf%s" %
                          (args.detail, args.tex, ", BVH built with SAH splits (same BVH_00.04 layout)" if args.bvh == "sah" else ""))
    if args.workload == "rtiow":
        return dict(workload="RTIOW 488 spheres %dx%d %dspp/GPU depth 50 (BASELINE config 2)" % (args.nx, args.ny, args.spp),
                    scene="host LCG seed 1, spheres in __constant__")
    return dict(workload="ray batch %d rays vs staircase BVH (BASELINE config 5)" % args.rays,
                scene="procedural staircase detail %.2f" % args.detail)


RAYCOUNT_FILE = os.path.join(ROOT, "profiles", "raycounts.json")


def raycount_key(args, ns):
    if args.workload == "rtiow":
        return "rtiow:seed1:%dx%dx%d:d50" % (args.nx, args.ny, ns)
    # (the caller's tree does not change which rays are traced -- frames are identical for both build modes -- but the key says
    # which file the count was taken on)
    return "staircase:%.3f:%d:%dx%dx%d:d%d%s" % (args.detail, args.tex, args.nx, args.ny, ns, args.depth, ":sah" if args.bvh == "sah" else "")


def known_raycount(args, ns):
    if os.path.exists(RAYCOUNT_FILE):
        return json.load(open(RAYCOUNT_FILE)).get(raycount_key(args, ns))
    return None


# ------------------------------------------------------------------------------------------------ reference arm --
def run_reference(args, rank, world):
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    base = dict(impl="reference", n_gpus=world, steps=args.steps, warmup=args.warmup, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic")
    if not oracle.have_ref():
        print(json.dumps(dict(base, unavailable="oracle/_ref (the reference's CUDA build) is not present on this box")))
        return
    # The other arm's whole job is spp x world samples per pixel; the reference has one GPU to do it on. Its frame time is
    # linear in the sample count (one thread per pixel, samples in sequence), so at world > 2 a step is a BOUNDED SAMPLE of
    # that job -- 200 spp -- and the rate (Mrays/s) is what is reported; the run stays within a minute or two at every N.
    job_ns = args.spp * world
    ns = min(job_ns, max(2 * args.spp, args.spp))
    cfg = workload_params(args)  # the same keys and values as the other arm's `config`; what is specific to this arm is in `arm`
    arm = dict(gpus_used=1, total_spp=job_ns, sampled_spp=ns, l2="state + textures (>230 MB) exceed L2; rewritten every step")
    if args.workload == "staircase":
        spec = args.detail
        if args.bvh == "sah":  # the reference reads the SAH-built tree from a BVH_00.04 file, like any scene of its own
            import tempfile
            spec = os.path.join(tempfile.mkdtemp(), "staircase_sah.bvh")
            assert oracle.crt.Scene.staircase(args.detail, args.tex, 5, sah=True).save_bvh(spec) == 0
        _, info = oracle.ref_render(spec, args.tex, 5, args.nx, args.ny, ns, args.depth, "-", warmup=args.warmup, steps=args.steps)
        e2e = oracle._run([os.path.join(oracle.REF_DIR, "ref_driver"), "render_e2e", str(spec), str(args.tex), "5", str(args.nx),
                           str(args.ny), str(ns), str(args.depth), "1", str(min(args.steps, 3)), "-"])
        kind = "reference"
    elif args.workload == "rtiow":
        _, info = oracle.ref_spheres(1, args.nx, args.ny, ns, 50, "-", warmup=args.warmup, steps=args.steps)
        e2e = info
        kind = "reference-derived (oracle/ref_spheres.cu: HEAD has no sphere renderer)"
    else:
        print(json.dumps(dict(base, unavailable="ray-batch reference timing is part of tools/parity_report.py")))
        return
    ms = sum(info["ms"]) / len(info["ms"])
    rays = known_raycount(args, ns)
    est = False
    if rays is None:  # estimate from a row-strided CPU sample (the reference cannot count: its STATS build does not compile on Linux)
        scene = oracle.crt.Scene.staircase(args.detail, args.tex, 5) if args.workload == "staircase" else None
        stride = 40
        if scene is not None:
            _, cnt = oracle.render(scene, args.nx, args.ny, min(ns, 8), args.depth, count=True, row_stride=stride)
            per_sample = (cnt["primary"] + cnt["secondary"] + cnt["shadow"]) / max(cnt["primary"], 1)
        else:
            _, cnt = oracle.render_spheres(oracle.crt.rtiow_scene(1), args.nx, args.ny, min(ns, 8), 50, count=True, row_stride=stride)
            per_sample = (cnt["primary"] + cnt["secondary"]) / max(cnt["primary"], 1)
        rays = int(per_sample * args.nx * args.ny * ns)
        est = True
    value = rays / (ms * 1e3)
    e2e_ms = sum(e2e["ms"]) / len(e2e["ms"])
    line = dict(base, metric="Mrays/s", value=value, unit="Mrays/s", ms_per_step=ms, config=cfg, arm=arm, rays_per_step=rays,
                rays_estimated=est, msamples_per_s=args.nx * args.ny * ns / (ms * 1e3),
                cpu_baseline=dict(value=value, unit="Mrays/s", cores=0, kind=kind,
                                  sample="%d of the job's %d spp, full frame, on the GPU: the reference's render path is a CUDA kernel, it has "
                                         "no CPU implementation" % (ns, job_ns)),
                e2e=dict(value=rays / (e2e_ms * 1e3), unit="Mrays/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0, ms_per_step=e2e_ms,
                         note="initRenderer + runRenderer + frame read per step, through the same 3 entry points"),
                gpu_launches=args.steps)
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------ our arm --
def scene_bytes(scene):
    ks = scene.ks
    tex = sum(ks.textures[i].width * ks.textures[i].height * 12 for i in range(ks.numTextures))
    return scene.num_slots * 64 + scene.num_nodes * 24 + ks.numMaterials * 24 + tex


def run_ours(args, rank, world, local_rank):
    import torch
    import crt_b200 as crt
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's banner must not share stdout with the JSON line
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    L = crt.device_lib()  # raises when the CUDA library is missing: no fallback
    pk = peaks()
    spheres = args.workload == "rtiow"
    depth = 50 if spheres else args.depth
    scene = crt.rtiow_scene(1) if spheres else crt.Scene.staircase(args.detail, args.tex, 5, sah=args.bvh == "sah")
    nx, ny, ns = args.nx, args.ny, args.spp
    ns_total = ns * world
    h2d = (488 * 40) if spheres else scene_bytes(scene)
    d2h = nx * ny * 12

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def allsum(x):
        if not dist:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(t.item())

    def allmax(x):
        if not dist:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    acc = torch.zeros(ny * nx, 4, device="cuda") if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def open_frame():
        crt.set_options(device=local_rank, sample_stream=rank, defer_finalize=1 if world > 1 else 0, slots_per_pixel=args.slots)
        fr = crt.Frame(scene, nx, ny, depth)
        if world > 1:
            L.setRendererAccumDevice(acc.data_ptr())
        return fr

    def step(fr):
        """One frame. Returns (device ms incl. the reduce, rays, launches)."""
        L.runRenderer(ns, 8, 8)
        st = crt.stats()
        ms = st.msTotal
        launches = st.kernelLaunches
        if world > 1:
            ev0.record()
            dist.reduce(acc, dst=0, op=dist.ReduceOp.SUM)
            ev1.record()
            torch.cuda.synchronize()
            ms += ev0.elapsed_time(ev1)
            if rank == 0:
                L.finalizeFrame(ns_total)
                launches += 1
        return ms, st.raysExtend + st.raysShadow, launches

    fr = open_frame()
    for _ in range(args.warmup):
        step(fr)
    barrier()
    dev_ms, rays_step, launches = 0.0, 0, 0
    with ClockSampler(local_rank) as clk:
        t0 = time.perf_counter()
        for _ in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            ms, rays_step, l = step(fr)
            dev_ms += ms
            launches += l
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
    ms_per_step = allmax(dev_ms / args.steps)
    total_rays = allsum(rays_step)
    launches = int(allsum(launches))
    value = total_rays / (ms_per_step * 1e3)
    clocks = clk.summary()

    # ---- roofline of the dominant kernel (the trace kernel): one extra step in the timed configuration with CUDA events around every
    # trace / shade launch (setRendererProfiling(2): same kernels, same chaser beside them, launched one by one instead of as a
    # graph), and one counting step for the flop side; both untimed
    roof, extra = None, {}
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    fp32_peak = sms * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
    if rank == 0 and not spheres:
        import ctypes as C
        wide = crt.wide_info()
        L.setRendererProfiling(2)
        L.runRenderer(ns, 8, 8)
        L.setRendererProfiling(0)
        ps = crt.stats()
        ce, cs, cn, ct = C.c_ulonglong(), C.c_ulonglong(), C.c_ulonglong(), C.c_ulonglong()
        L.getRendererChaserCounts(C.byref(ce), C.byref(cs), C.byref(cn), C.byref(ct))
        wave_extend, wave_shadow = ps.raysExtend - ce.value, ps.raysShadow - cs.value  # the rays the trace kernel traced
        L.setRendererCounting(1)
        L.runRenderer(ns, 8, 8)
        L.setRendererCounting(0)
        nv, tt = C.c_ulonglong(), C.c_ulonglong()
        L.getRendererTraversalCounts(C.byref(nv), C.byref(tt))
        L.getRendererChaserCounts(C.byref(ce), C.byref(cs), C.byref(cn), C.byref(ct))
        iters = max(int(ps.iterations), 1)
        avg_ms = ps.msTrace / iters
        rays_all = ps.raysExtend + ps.raysShadow
        kernel = "wideTraceKernel<false,*,true>" if wide.active else "traceKernel<false,*>"
        per_visit = FLOPS_PER_WIDE_VISIT if wide.active else FLOPS_PER_DUAL_VISIT
        bytes_per_launch = (EXTEND_BYTES_PER_RAY * wave_extend + SHADOW_BYTES_PER_RAY * wave_shadow + RESUME_BYTES * ps.resumes) / iters
        hbm_achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9
        flops_wave = FLOPS_PER_RAY * (wave_extend + wave_shadow) + per_visit * (nv.value - cn.value) + FLOPS_PER_TRI_TEST * (tt.value - ct.value)
        flops_all = FLOPS_PER_RAY * rays_all + per_visit * nv.value + FLOPS_PER_TRI_TEST * tt.value
        fp32_achieved = flops_wave / (ps.msTrace * 1e-3) / 1e12
        hbm_frac, fp32_frac = hbm_achieved / pk["hbm_gbs"], fp32_achieved / fp32_peak
        ncu = measured_traffic("wideTraceKernel" if wide.active else "traceKernel")
        roof = dict(bound="fp32" if fp32_frac >= hbm_frac else "hbm", kernel=kernel,
                    achieved=fp32_achieved if fp32_frac >= hbm_frac else hbm_achieved, peak=fp32_peak if fp32_frac >= hbm_frac else pk["hbm_gbs"],
                    unit="TFLOP/s" if fp32_frac >= hbm_frac else "GB/s", frac=max(fp32_frac, hbm_frac),
                    traffic=ncu["dram_bytes_per_launch"] if ncu else None,
                    traffic_source=(ncu["source"] if ncu else None),
                    traffic_over_algorithmic=(ncu["dram_bytes_per_launch"] / (ncu["algorithmic_bytes_per_launch"]) if ncu else None),
                    peak_source="SMs x 128 lanes x 2 x sm_max_mhz (cudaGetDeviceProperties, MEASURED_PEAKS.json)", avg_launch_ms=avg_ms, launches=iters,
                    binding_resource="instruction issue, ALU pipe first (profiles/r02: issue active 70 %, ALU pipe 60 %, 18 of 32 lanes per "
                                     "instruction); the scene is L2-resident by design, so the HBM term is small",
                    hbm=dict(achieved=hbm_achieved, peak=pk["hbm_gbs"], unit="GB/s", frac=hbm_frac, peak_source=pk["src"],
                             bytes_per_ray=dict(extend=EXTEND_BYTES_PER_RAY, shadow=SHADOW_BYTES_PER_RAY, resumed=RESUME_BYTES),
                             algorithmic_bytes_per_launch=bytes_per_launch),
                    fp32=dict(achieved=fp32_achieved, peak=fp32_peak, unit="TFLOP/s", frac=fp32_frac,
                              kernels="the trace kernel's own rays over its own launch time",
                              whole_frame_frac=flops_all / (ms_per_step * 1e-3) / 1e12 / fp32_peak,
                              node_visits_per_ray=nv.value / max(rays_all, 1), tri_tests_per_ray=tt.value / max(rays_all, 1),
                              formula="%g*rays + %g*node visits + %g*triangle tests (%s)" %
                                      (FLOPS_PER_RAY, per_visit, FLOPS_PER_TRI_TEST, "8-wide nodes of the renderer's own tree" if wide.active else "dual-node visits, SURVEY.md 8d")))
        extra = dict(kernel_ms_profiled=dict(trace=ps.msTrace, shade=ps.msShade,
                                             note="CUDA events around every trace / shade launch of one frame in the timed configuration (chaser waves run beside them)"),
                     wavefront_iterations=iters, resumed_rays=int(ps.resumes), deferred_shades=int(ps.deferred),
                     chaser=dict(rays=int(ce.value + cs.value), share_of_rays=(ce.value + cs.value) / max(rays_all, 1)),
                     acceleration_structure=dict(own_tree=bool(wide.active), wide_nodes=int(wide.numNodes), depth=int(wide.depth),
                                                 build_ms_inside_initRenderer=float(wide.buildMs), built_on="device" if wide.buildThreads == 0 else "host",
                                                 rays_retraced_in_reference_order=int(wide.lastFrameRedo)))
    elif rank == 0:
        # sphere scenes: ONE persistent launch per frame (spheresMegaKernel: a lane owns a pixel, the path lives in registers, the
        # 488 spheres and their BVH are L1-resident), so the kernel's launch time is the step time and the byte side is the frame
        # it writes; the flop side counts the box and sphere tests of one extra counting step (untimed)
        import ctypes as C
        L.setRendererCounting(1)
        L.runRenderer(ns, 8, 8)
        L.setRendererCounting(0)
        nb, nt = C.c_ulonglong(), C.c_ulonglong()
        L.getRendererTraversalCounts(C.byref(nb), C.byref(nt))
        rays_mine = max(rays_step, 1)
        flops = (3.0 + SPHERE_SHADE_FLOPS) * rays_mine + FLOPS_PER_BOX_TEST * nb.value + FLOPS_PER_SPHERE_TEST * nt.value
        fp32_achieved = flops / (ms_per_step * 1e-3) / 1e12
        bytes_frame = 16.0 * nx * ny + 32.0 * nx * ny  # float4 sums written once, finalize reads them and writes the frame
        hbm_achieved = bytes_frame / (ms_per_step * 1e-3) / 1e9
        roof = dict(bound="fp32", kernel="spheresMegaKernel<false> (whole frame: one launch)", achieved=fp32_achieved, peak=fp32_peak, unit="TFLOP/s",
                    frac=fp32_achieved / fp32_peak, traffic=None, peak_source="SMs x 128 lanes x 2 x sm_max_mhz (cudaGetDeviceProperties, MEASURED_PEAKS.json)",
                    binding_resource="instruction issue at low lane use (profiles/r02/sph_mega_ncu_summary.txt: issue active 80 %, 9-10 of 32 lanes per "
                                     "instruction): walks of different length inside a warp; no queue traffic at all, the scene is L1-resident",
                    hbm=dict(achieved=hbm_achieved, peak=pk["hbm_gbs"], unit="GB/s", frac=hbm_achieved / pk["hbm_gbs"], peak_source=pk["src"],
                             algorithmic_bytes_per_launch=bytes_frame),
                    fp32=dict(achieved=fp32_achieved, peak=fp32_peak, unit="TFLOP/s", frac=fp32_achieved / fp32_peak,
                              box_tests_per_ray=nb.value / rays_mine, sphere_tests_per_ray=nt.value / rays_mine,
                              formula="(3 + %d)*rays + %d*box tests + %d*sphere tests" % (SPHERE_SHADE_FLOPS, FLOPS_PER_BOX_TEST, FLOPS_PER_SPHERE_TEST)))
    fr.close()

    # ---- e2e: host buffers -> frame on the host, through the reference-facing entry points, every step
    e2e_steps = max(1, min(args.steps, 3))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        fr = open_frame()                 # H2D: triangles, nodes, materials, textures (caller-owned pageable memory, as the ABI hands it over)
        step(fr)
        if rank == 0:
            host = fr.frame(copy=True)    # D2H: the managed frame buffer read on the host
            if extra.get("acceleration_structure"):  # the tree build of a WARM frame (the first frame of a process also allocates)
                extra["acceleration_structure"]["build_ms_inside_initRenderer"] = float(crt.wide_info().buildMs)
        fr.close()
    barrier()
    e2e_ms = allmax((time.perf_counter() - t0) * 1e3 / e2e_steps)
    e2e_value = total_rays / (e2e_ms * 1e3)

    # ---- BASELINE config 4 (strong scaling): ONE frame of 3840x2160 at 1024 spp IN TOTAL, the samples split over the ranks
    # (1024 / N each, rank g on RNG stream g), the un-normalised sums reduced once to rank 0. Not part of `value`; no warm-up.
    c4, c4_tp = None, None
    if args.c4 and not spheres:
        nx4, ny4, total4 = 3840, 2160, 1024
        frames4 = {}

        def render_c4(slots):
            mine = shard_samples(total4, world)[rank]
            acc4 = torch.zeros(ny4 * nx4, 4, device="cuda") if world > 1 else None
            crt.set_options(device=local_rank, sample_stream=rank, defer_finalize=1 if world > 1 else 0, slots_per_pixel=slots)
            fr4 = crt.Frame(scene, nx4, ny4, depth)
            if world > 1:
                L.setRendererAccumDevice(acc4.data_ptr())
            barrier()
            t0 = time.perf_counter()
            L.runRenderer(mine, 8, 8)
            st4 = crt.stats()
            ms4 = st4.msTotal
            if world > 1:
                ev0.record()
                dist.reduce(acc4, dst=0, op=dist.ReduceOp.SUM)
                ev1.record()
                torch.cuda.synchronize()
                ms4 += ev0.elapsed_time(ev1)
                if rank == 0:
                    L.finalizeFrame(total4)
            barrier()
            wall4 = (time.perf_counter() - t0) * 1e3
            ms4 = allmax(ms4)
            rays4 = allsum(st4.raysExtend + st4.raysShadow)
            if rank == 0:
                frames4[slots] = fr4.frame(copy=True)
            fr4.close()
            del acc4
            return dict(ms=ms4, wall_ms=wall4, mrays_per_s=rays4 / (ms4 * 1e3), msamples_per_s=nx4 * ny4 * total4 / (ms4 * 1e3), rays=int(rays4),
                        spp_per_gpu=shard_samples(total4, world), frame="3840x2160", total_spp=total4, scaling="strong", slots_per_pixel=max(slots, 1),
                        note="device time of runRenderer (max over ranks) + the NCCL reduce of %d MB per rank + finalize on rank 0" % (nx4 * ny4 * 16 >> 20))

        c4 = render_c4(args.slots)
        if args.c4_throughput and args.slots <= 1 and min(shard_samples(total4, world)) % 4 == 0:
            # the same frame in throughput mode (4 path slots per pixel, each its own RNG stream: no per-pixel serial chain): other
            # noise, same expectation -- the two frames are compared below
            c4_tp = render_c4(4)
            if rank == 0:
                import numpy as np
                a, b = frames4[args.slots].astype(np.float64), frames4[4].astype(np.float64)
                c4_tp["vs_parity_mode"] = dict(speedup=c4["ms"] / c4_tp["ms"], rmse=float(np.sqrt(((a - b) ** 2).mean())), mean_parity=float(a.mean()),
                                               mean_throughput=float(b.mean()), relative_mean_difference=float(abs(a.mean() - b.mean()) / a.mean()),
                                               note="two independent 1024-spp estimates of the same frame: rmse is the noise of their difference")
        frames4.clear()

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---- CPU baseline: the oracle port on the host cores, row-strided sample of the same frame
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    cpu = None
    if world == 1:
        stride, cpu_ns = 4, min(ns, 25)  # ~10-20 s of work on 16 host threads
        t0 = time.perf_counter()
        if spheres:
            _, cnt = oracle.render_spheres(scene, nx, ny, cpu_ns, depth, count=True, row_stride=stride)
            cr = cnt["primary"] + cnt["secondary"]
        else:
            _, cnt = oracle.render(scene, nx, ny, cpu_ns, depth, count=True, row_stride=stride)
            cr = cnt["primary"] + cnt["secondary"] + cnt["shadow"]
        sec = time.perf_counter() - t0
        cpu = dict(value=cr / sec / 1e6, unit="Mrays/s", cores=oracle.lib().oracleNumThreads(), kind="port",
                   sample="every %dth row of the %dx%d frame at %d spp (%d rays, %.1f s), OpenMP over rows, -O3 -ffp-contract=off" %
                          (stride, nx, ny, cpu_ns, cr, sec))

    cfg = workload_params(args)  # the same keys and values as the reference arm's `config`; what is specific to this arm is in `arm`
    arm = dict(parallelism="sample-sharded x%d, 1 NCCL reduce/frame" % world if world > 1 else "single GPU",
               total_spp=ns_total, rng="reference seeding (stream = rank), %d slot(s)/pixel" % max(args.slots, 1),
               l2="256 MB written between timed steps; path state + textures (>230 MB) exceed the 126 MB L2")
    line = dict(metric="Mrays/s", value=value, unit="Mrays/s", n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_per_step,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", config=cfg, arm=arm,
                rays_per_step=int(total_rays), msamples_per_s=nx * ny * ns_total / (ms_per_step * 1e3), wall_ms_per_step=wall_ms / args.steps,
                clocks=clocks, e2e=dict(value=e2e_value, unit="Mrays/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                                        ms_per_step=e2e_ms, steps=e2e_steps,
                                        note="initRenderer (scene upload from caller-owned host memory) + runRenderer + frame read + cleanupRenderer, per step"),
                gpu_launches=launches, roofline=roof, cpu_baseline=cpu, c4_strong=c4, c4_strong_throughput=c4_tp, **extra)
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


def run_raybatch(args, rank, world, local_rank):
    """BASELINE config 5: closest-hit queries on a device-resident ray batch (weak scaling: every rank its own batch)."""
    import torch
    import crt_b200 as crt
    torch.cuda.set_device(local_rank)
    L = crt.device_lib()
    pk = peaks()
    n = args.rays
    scene = crt.Scene.staircase(args.detail, 64, 5)
    crt.set_options(device=local_rank)
    with crt.Frame(scene, 64, 64, 1):
        dO, dD, dH, dM = (L.rendererDeviceAlloc(16 * n), L.rendererDeviceAlloc(16 * n), L.rendererDeviceAlloc(16 * n), L.rendererDeviceAlloc(4 * n))
        L.generateRayBatchDevice(dO, dD, n, 8192, 4096, 0.01, FLT_MAX)
        for _ in range(args.warmup):
            L.intersectBatchDevice(dO, dD, n, dH, dM)
        with ClockSampler(local_rank) as clk:
            ms = [L.intersectBatchDevice(dO, dD, n, dH, dM) for _ in range(args.steps)]
        L.setRendererCounting(1)
        L.intersectBatchDevice(dO, dD, n, dH, dM)
        L.setRendererCounting(0)
        import ctypes as C
        nv, tt = C.c_ulonglong(), C.c_ulonglong()
        L.getRendererTraversalCounts(C.byref(nv), C.byref(tt))
        wide_active, redo = bool(crt.wide_info().active), int(crt.wide_info().lastBatchRedo)
        # e2e: host rays in, host hits out
        ro, rd = np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32)
        L.rendererCopyToHost(ro.ctypes.data, dO, 16 * n)
        L.rendererCopyToHost(rd.ctypes.data, dD, 16 * n)
        hit, mesh = np.zeros((n, 4), np.float32), np.zeros(n, np.int32)
        t0 = time.perf_counter()
        L.rendererCopyToDevice(dO, ro.ctypes.data, 16 * n)
        L.rendererCopyToDevice(dD, rd.ctypes.data, 16 * n)
        L.intersectBatchDevice(dO, dD, n, dH, dM)
        L.rendererCopyToHost(hit.ctypes.data, dH, 16 * n)
        L.rendererCopyToHost(mesh.ctypes.data, dM, 4 * n)
        e2e_ms = (time.perf_counter() - t0) * 1e3
        for p in (dO, dD, dH, dM):
            L.rendererDeviceFree(p)
    if rank != 0:
        return
    avg = sum(ms) / len(ms)
    achieved = BATCH_BYTES_PER_RAY * n / (avg * 1e-3) / 1e9
    flops = FLOPS_PER_RAY * n + (FLOPS_PER_WIDE_VISIT if wide_active else FLOPS_PER_DUAL_VISIT) * nv.value + FLOPS_PER_TRI_TEST * tt.value
    sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
    fp32_peak = sms * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
    line = dict(metric="Mrays/s", value=n * world / (avg * 1e3), unit="Mrays/s", n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=avg,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload_params(args), l2="ray batch (%d MB) exceeds L2" % (52 * n >> 20)), clocks=clk.summary(),
                e2e=dict(value=n / (e2e_ms * 1e3), unit="Mrays/s", h2d_bytes_per_step=32 * n, d2h_bytes_per_step=20 * n, ms_per_step=e2e_ms),
                gpu_launches=args.steps,
                roofline=dict(bound="fp32" if flops / (avg * 1e-3) / 1e12 / fp32_peak >= achieved / pk["hbm_gbs"] else "hbm",
                              kernel="wideIntersectBatchKernel<false,true> (+ intersectBatchKernel for %d uncertified rays)" % redo if wide_active else "intersectBatchKernel<false>",
                              achieved=max(flops / (avg * 1e-3) / 1e12, 0.0) if flops / (avg * 1e-3) / 1e12 / fp32_peak >= achieved / pk["hbm_gbs"] else achieved,
                              peak=fp32_peak if flops / (avg * 1e-3) / 1e12 / fp32_peak >= achieved / pk["hbm_gbs"] else pk["hbm_gbs"],
                              unit="TFLOP/s" if flops / (avg * 1e-3) / 1e12 / fp32_peak >= achieved / pk["hbm_gbs"] else "GB/s",
                              frac=max(flops / (avg * 1e-3) / 1e12 / fp32_peak, achieved / pk["hbm_gbs"]),
                              traffic=(measured_traffic("wideIntersectBatchKernel") or {}).get("dram_bytes_per_launch"),
                              hbm=dict(achieved=achieved, peak=pk["hbm_gbs"], unit="GB/s", frac=achieved / pk["hbm_gbs"], peak_source=pk["src"], bytes_per_ray=BATCH_BYTES_PER_RAY),
                              fp32=dict(achieved=flops / (avg * 1e-3) / 1e12, peak=fp32_peak, unit="TFLOP/s", frac=flops / (avg * 1e-3) / 1e12 / fp32_peak,
                                        node_visits_per_ray=nv.value / n, tri_tests_per_ray=tt.value / n)),
                cpu_baseline=None)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="staircase", choices=["staircase", "rtiow", "raybatch"])
    ap.add_argument("--nx", type=int, default=1200)
    ap.add_argument("--ny", type=int, default=800)
    ap.add_argument("--spp", type=int, default=100)
    ap.add_argument("--depth", type=int, default=64)
    ap.add_argument("--detail", type=float, default=1.0)
    ap.add_argument("--tex", type=int, default=1024)
    ap.add_argument("--rays", type=int, default=1 << 26)
    ap.add_argument("--c4-throughput", dest="c4_throughput", action="store_true",
                    help="render config 4 a second time with 4 path slots per pixel and compare the two frames (profiles/r02/bench_c3.json: 23.9 s vs 23.1 s "
                         "in parity mode, means equal to 4e-7: not faster any more, so parity mode stays the default)")
    ap.add_argument("--slots", type=int, default=0, help="path slots per pixel (0/1 = reference RNG streams)")
    ap.add_argument("--no-c4", dest="c4", action="store_false", help="skip the BASELINE config 4 frame (3840x2160, 1024 spp in total: ~30 s on one GPU)")
    ap.add_argument("--bvh", default="median", choices=["median", "sah"],
                    help="how the host builds the scene's BVH_00.04 tree: the reference author's median split (default) or the "
                         "surface-area heuristic inside the same layout (both arms get the same file)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:  # python bench.py --gpus N without torchrun: re-launch under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29400 + os.getpid() % 500)] + sys.argv
        sys.exit(subprocess.call(cmd))
    # Anything a library prints on stdout (NCCL's version banner, for one) must not mix with the ONE JSON line: file
    # descriptor 1 points at stderr while the benchmark runs, print() goes to the saved original stdout.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload == "raybatch":
        run_raybatch(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)
    real_stdout.flush()


if __name__ == "__main__":
    main()

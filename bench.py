#!/usr/bin/env python3
"""Headline benchmark: samples/s (and Mrays/s) of the hair path-tracing hot path.

    python bench.py --gpus N --steps K --warmup W          # CUDA arm
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm

Default workload = BASELINE.json configs[2], the one `metric` is quoted on: synthetic
curly hair, 1M line segments (4 lines / Bezier), blonde sigma_a = {0.06, 0.10,
0.20}, 1280 x 1280, path integrator, 8 bounces, under a 2048 x 1024 sky texture
(the size of the reference's tests/curly-hair/textures/sky.hdr).  `--workload`
selects the other BASELINE configurations: c1 / c1sq (sphere + hair block, 720 x 405
and 720 x 720), c2 (250k straight segments), c4 (the 18-point BSDF sweep), c5 (5M
segments, 1920 x 1920).  A "step" is one pass of `--spp-step` samples per pixel over
the whole image (= spp_step consecutive trace_samples calls of the reference).

One JSON line is printed by rank 0 (see the contract in the task statement):
  value     whole-job samples/s, scene and state resident in HBM, device timed
  e2e       same metric through the C ABI from HOST buffers: scene upload + BVH
            build + light tables + state creation + K x (trace + [N > 1: NCCL reduce of
            the accumulation buffers to rank 0 + resolve] + download of the COMPLETE
            render image into pinned host memory on rank 0)
  roofline  traversal kernel (k_intersect): algorithmic bytes / measured time
  cpu_baseline  the reference's own CPU code on a bounded sample of the workload,
            preceded by a hit-record spot check of the scene that was timed
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "yocto-hair_b200"))

SKY_SIZE = (2048, 1024)  # tests/curly-hair/textures/sky.hdr is 2048 x 1024

WORKLOADS = {
    # name: (generator, segments, resolution, config text)
    "c1": ("block", 200_000, 720, "sphere + synthetic hair block (200k segments), default hair material beta_m=beta_n=0.3 alpha=2 eta=1.55, 720x405 (the scene's 16:9 camera), path, 8 bounces"),
    "c1sq": ("block_sq", 200_000, 720, "sphere + synthetic hair block (200k segments), default hair material, 720x720, path, 8 bounces"),
    "c2": ("straight", 250_000, 720, "synthetic straight hair, 250k segments (2 lines/Bezier), eumelanin=1.3, 720x720, path, 8 bounces"),
    "c3": ("curly", 1_000_000, 1280, "synthetic curly hair, 1M segments (4 lines/Bezier), blonde sigma_a={0.06,0.10,0.20}, 1280x1280, path, 8 bounces"),
    "c4": ("sweep", 200_000, 720, "BSDF parameter sweep beta_m in {0.1,0.25,0.6} x beta_n in {0.3,0.6,0.9} x alpha in {0,2} on the synthetic hair block (200k segments), 720x720, path, 8 bounces: 18 scenes, aggregate"),
    "c5": ("curly", 5_000_000, 1920, "synthetic head of hair, 5M segments, 1920x1920, path, 8 bounces"),
    "tiny": ("curly", 40_000, 256, "tiny curly hair (40k segments, 256x256) for harness checks"),
}
SWEEP = [(bm, bn, al) for bm in (0.1, 0.25, 0.6) for bn in (0.3, 0.6, 0.9) for al in (0.0, 2.0)]


def build_scene(kind: str, segments: int, sky: bool, hair=None):
    from yocto_hair_b200 import synth
    if kind == "curly":
        return synth.curly_hair_scene(segments, sky=sky, sky_size=SKY_SIZE)
    if kind == "straight":
        return synth.straight_hair_scene(segments, sky=sky, sky_size=SKY_SIZE)
    if kind == "block":
        return synth.hair_block_scene(segments, **(hair or {}))
    if kind in ("block_sq", "sweep"):
        return synth.hair_block_scene(segments, aspect=1.0, **(hair or {}))
    raise ValueError(kind)


def build_workload(name: str, sky: bool, hair=None):
    kind, segs, res, text = WORKLOADS[name]
    scene, cam = build_scene(kind, segs, sky, hair)
    return scene, cam, res, text


def scene_bytes(scene) -> int:
    n = 0
    for s in scene.shapes:
        n += s.positions.nbytes + s.normals.nbytes + s.radius.nbytes + s.lines.nbytes + s.triangles.nbytes + s.points.nbytes
    for t in scene.textures:
        n += t.colorf.nbytes
    return n


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v == "Active":
                    reasons.add(name)
        busy = [x for x in sm if x > 0.5 * mx] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traversal_stats_probe(workload: str, sky: bool, spp: int, max_split: int, hair=None) -> dict:
    """Runs the -DYH_TRAVERSAL_STATS build once (outside any timed region) and
    returns nodes / primitives fetched per ray for this workload."""
    env = dict(os.environ, YH_B200_LIB=str(ROOT / "yocto-hair_b200" / "csrc" / "libyocto_hair_b200_stats.so"))
    code = (
        "import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import bench\n"
        "from yocto_hair_b200.pathtrace import DeviceScene\nfrom yocto_hair_b200.scene import make_params\n"
        "sc, cam, res, _ = bench.build_workload(%r, %r, %r)\n"
        "ds = DeviceScene(sc).init_bvh(%d).init_lights()\n"
        "st = ds.init_state(cam, make_params(resolution=res)); s = st.trace_samples(%d)\n"
        "print(json.dumps(dict(rays=s.rays, nodes=s.nodes_visited, prims=s.prims_tested)))\n"
    ) % (str(ROOT), str(ROOT / "yocto-hair_b200"), workload, sky, hair, max_split, spp)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    if out.returncode != 0:
        raise RuntimeError("stats probe failed: " + out.stderr[-2000:])
    return json.loads(out.stdout.strip().splitlines()[-1])


def sweep_hair(point):
    bm, bn, al = point
    return dict(beta_m=bm, beta_n=bn, alpha=al)


def run_reference(args):
    """CPU arm: the reference's own code (oracle/_ref) or the port, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle
    from yocto_hair_b200.scene import make_params
    orc = oracle.best_oracle()
    points = [sweep_hair(p) for p in SWEEP] if args.workload == "c4" else [None]
    spp = args.ref_spp_step
    total = build_s = 0.0
    samples = 0
    for hair in points:
        scene, cam, res, text = build_workload(args.workload, args.sky, hair)
        t0 = time.time()
        rs = orc.scene(scene)
        build_s += time.time() - t0
        st = rs.state(cam, make_params(resolution=res))
        for _ in range(args.warmup):
            st.trace_samples(1)
        steps = args.steps if hair is None else 1  # the sweep is 18 scenes: one step each
        total += sum(st.trace_samples(spp) for _ in range(steps))
        samples += st.width * st.height * spp * steps
        st.close(); rs.close()
    value = samples / total
    cores = orc.hardware_concurrency()
    sample = f"{args.steps} steps x {spp} spp of the workload ({samples/1e6:.1f} M samples), all {cores} host threads"
    print(json.dumps({
        "impl": "reference", "metric": "samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": text, "spp_per_step": spp, "sky_texture": args.sky, "bvh_build_s": round(build_s, 3)},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": orc.kind, "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline(args, scene, cam, res, device_scene):
    """Rank 0, N = 1: first the checker's job - a hit-record spot check of the scene that was just timed (bit-exact, or
    the bench line is not printed) - then the reference's own tracer on a bounded sample of the workload."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle
    from yocto_hair_b200 import synth
    from yocto_hair_b200.scene import make_params
    orc = oracle.best_oracle()
    rs = orc.scene(scene)
    rays = np.concatenate([synth.random_rays(scene, 12_000), synth.camera_rays(cam, 256, 256, 8_000)])
    hr = rs.intersect(rays.view(np.float32).reshape(-1, 8), threads=orc.hardware_concurrency())
    hd = device_scene.intersect(rays)
    h = hr["hit"] == 1
    same = (np.array_equal(hr["hit"], hd["hit"]) and np.array_equal(hr["object"][h], hd["object"][h])
            and np.array_equal(hr["element"][h], hd["element"][h])
            and np.array_equal(hr["distance"][h].view(np.uint32), hd["distance"][h].view(np.uint32))
            and np.array_equal(hr["uv"][h].view(np.uint32), hd["uv"][h].view(np.uint32)))
    if not same:
        raise SystemExit("bench: hit records of the timed scene differ from the %s oracle - no number is reported" % orc.kind)
    st = rs.state(cam, make_params(resolution=res))
    st.trace_samples(1)
    # bounded sample: aim at ~15 s of CPU work
    t1 = st.trace_samples(1)
    spp = int(max(1, min(64, round(15.0 / max(t1, 1e-3)))))
    secs = st.trace_samples(spp)
    samples = st.width * st.height * spp
    cores = orc.hardware_concurrency()
    st.close(); rs.close()
    return {"value": samples / secs, "unit": "samples/s", "cores": cores, "kind": orc.kind,
            "sample": f"same scene and camera, {spp} spp of {res}-pixel-wide image ({samples/1e6:.1f} M samples, {secs:.1f} s), "
                      f"std::thread::hardware_concurrency() = {cores} threads",
            "hit_spot_check": f"{len(rays)} rays ({int(h.sum())} hits): object / element / uv / distance bit-exact vs the {orc.kind} oracle"}


def run_cuda_point(args, hair, ctx):
    """One scene of the workload on the CUDA path; returns the pieces of the bench line."""
    import torch
    import torch.distributed as dist
    from yocto_hair_b200 import parallel
    from yocto_hair_b200.pathtrace import DeviceScene
    from yocto_hair_b200.scene import make_params

    world, rank, local = ctx["world"], ctx["rank"], ctx["local"]
    scene, cam, res, text = build_workload(args.workload, args.sky, hair)
    params = make_params(resolution=res)
    spp = args.spp_step
    steps, warmup = (args.steps, args.warmup) if hair is None else (1, 1)  # the sweep: 18 scenes, one step each
    # sharding: image tiles keep the reference's one-stream-per-pixel RNG (bit-identical images) but a pixel's
    # samples are sequential, so the wavefront narrows with 1/N; sample-range sharding keeps it full width
    # (counter-mode streams).  auto = tiles on one GPU, samples on several.
    shard = args.shard if args.shard != "auto" else ("tiles" if world == 1 else "samples")
    if shard == "samples" and spp % world:
        raise SystemExit(f"--spp-step {spp} must be a multiple of the number of GPUs for sample sharding")
    spp_rank = spp // world if shard == "samples" else spp
    dev = f"cuda:{local}"

    def make_state(ds, mode=shard):
        st = ds.init_state(cam, params)
        if mode == "samples":
            st.set_sample_shard(rank, world)
            if args.slots != 1:
                st.set_sample_slots(args.slots)
        elif world > 1:
            st.set_shard(rank, world)
        return st

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- resident arm: scene + state in HBM, device-timed -------------
    ds = DeviceScene(scene).init_bvh(args.max_split).init_lights()
    bvh = ds.bvh_stats()
    st = make_state(ds)
    width, height = st.width, st.height
    for _ in range(warmup):
        st.trace_samples(spp_rank)
    if world > 1 and warmup > 0:  # the collective's own warm-up (NCCL sets its channels up on the first call: 25-50 ms once)
        parallel.reduce_image(*parallel.state_tensors(st, dev), 0)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_ms = 0.0
    rays = light_rays = launches = iters = 0
    t_wall = time.perf_counter()
    for _ in range(steps):
        s = st.trace_samples(spp_rank)
        dev_ms += s.device_ms
        rays += s.rays; light_rays += s.light_rays; launches += s.launches; iters += s.iterations
    reduce_ms, reduce_ok = 0.0, None
    if world > 1:  # one NCCL reduce of the fp32 accumulation buffers over NVLink, at the end
        accum, nsamp = parallel.state_tensors(st, dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        total_a, total_s = parallel.reduce_image(accum, nsamp, 0)
        e1.record()
        torch.cuda.synchronize()
        reduce_ms = e0.elapsed_time(e1)
        if rank == 0:  # after the reduce rank 0 holds the whole image: every pixel has all its samples
            want = spp * (warmup + steps)
            reduce_ok = bool((total_s == want).all().item()) and bool(torch.isfinite(total_a).all().item())
            if not reduce_ok:
                raise SystemExit("bench: the %d-rank reduce did not produce a complete image - no number is reported" % world)
        del total_a, total_s
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall)
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel timing pass (outside the timed region): with profiling on, the wavefront groups run on ONE
    # stream and every traversal / shading launch is bracketed by CUDA events on it
    st.set_profiling(True)
    prof_steps = max(1, min(2, steps))
    trav_ms = shade_ms = prof_dev_ms = 0.0
    prof_rays = prof_iters = 0
    for _ in range(prof_steps):
        s = st.trace_samples(spp_rank)
        trav_ms += s.traverse_ms; shade_ms += s.shade_ms; prof_dev_ms += s.device_ms
        prof_rays += s.rays; prof_iters += s.iterations
    st.set_profiling(False)
    t = torch.tensor([dev_ms + reduce_ms, wall_ms, trav_ms, shade_ms, prof_dev_ms], dtype=torch.float64, device=dev)
    c = torch.tensor([rays, light_rays, launches, iters], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    total_ms, wall_max_ms, trav_max_ms, shade_max_ms, prof_max_ms = t.tolist()
    rays_all, light_rays_all, launches_all, iters_all = c.tolist()
    samples = width * height * spp * steps

    # N = 1 also reports the counter-mode rate (the mode the N > 1 default runs in), so that the scaling curve has
    # a baseline in its own RNG / workload mode (outside the timed region)
    counter_value = None
    if world == 1 and hair is None and not args.no_counter_baseline:
        stc = make_state(ds, "samples")
        stc.trace_samples(spp)
        ms = sum(stc.trace_samples(spp).device_ms for _ in range(2))
        counter_value = width * height * spp * 2 / (ms * 1e-3)
        stc.close()

    cpu = None
    if world == 1 and hair is None and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, scene, cam, res, ds)

    # ---------------- e2e arm: host buffers through the C ABI ----------------------
    st.close(); ds.close()
    render_host = torch.empty((height, width, 4), dtype=torch.float32).pin_memory()
    render_np = render_host.numpy()
    render_dev = torch.empty((height, width, 4), dtype=torch.float32, device=dev) if world > 1 else None
    scratch = None
    barrier()
    t0 = time.perf_counter()
    ds2 = DeviceScene(scene).init_bvh(args.max_split).init_lights()
    st2 = make_state(ds2)
    setup_ms = 1e3 * (time.perf_counter() - t0)
    e2e_reduce_ms = 0.0
    for _ in range(steps):
        st2.trace_samples(spp_rank)
        if world == 1:
            st2.download_into(render_np)
        else:  # the complete image: NCCL reduce to rank 0 + resolve + download, every step
            tr = time.perf_counter()
            accum, nsamp = parallel.state_tensors(st2, dev)
            if scratch is None:
                scratch = (torch.empty_like(accum), torch.empty_like(nsamp))
            total_a, total_s = parallel.reduce_image(accum, nsamp, 0, out=scratch)
            if rank == 0:
                parallel.resolve_dev(total_a, total_s, render_dev)
                render_host.copy_(render_dev)
            torch.cuda.synchronize()
            e2e_reduce_ms += 1e3 * (time.perf_counter() - tr)
    barrier()
    e2e_ms = torch.tensor([1e3 * (time.perf_counter() - t0), setup_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_total_ms, setup_max_ms = e2e_ms.tolist()
    checksum = float(np.asarray(render_np[..., :3], dtype=np.float64).mean()) if rank == 0 else 0.0
    if rank == 0 and world > 1:
        if not bool((scratch[1] == spp * steps).all().item()) or not np.isfinite(render_np).all():
            raise SystemExit("bench: the e2e image on rank 0 is incomplete - no number is reported")

    # the reference's own calling contract: ONE sample per trace_samples call, state->render valid (downloaded) after
    # each (host/yocto_pathtrace_b200.cpp); bounded to a few calls, N = 1 only
    dropin = None
    if world == 1 and hair is None:
        n_calls = 8
        st2.trace_samples(1); st2.download_into(render_np)
        torch.cuda.synchronize()
        tc = time.perf_counter()
        for _ in range(n_calls):
            st2.trace_samples(1)
            st2.download_into(render_np)
        dropin = {"value": width * height * n_calls / (time.perf_counter() - tc), "unit": "samples/s", "calls": n_calls,
                  "note": "1 spp per yh_trace_samples call + yh_state_download of the render after every call"}
    st2.close(); ds2.close()

    return dict(text=text, shard=shard, world=world, spp=spp, steps=steps, warmup=warmup, width=width, height=height, bvh=bvh,
                samples=samples, total_ms=total_ms, wall_max_ms=wall_max_ms, trav_max_ms=trav_max_ms, shade_max_ms=shade_max_ms,
                prof_max_ms=prof_max_ms, prof_steps=prof_steps, rays_all=rays_all, light_rays_all=light_rays_all,
                launches_all=launches_all, iters_all=iters_all, local_rays=prof_rays, local_iters=prof_iters, local_trav_ms=trav_ms,
                reduce_ms=reduce_ms, reduce_ok=reduce_ok, clocks=clocks, checksum=checksum, e2e_total_ms=e2e_total_ms,
                setup_ms=setup_max_ms, e2e_reduce_ms=e2e_reduce_ms, h2d=scene_bytes(scene), counter_value=counter_value, cpu=cpu,
                dropin=dropin, hair=hair)


def run_cuda(args):
    import torch
    import torch.distributed as dist
    from yocto_hair_b200 import _capi as capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # libraries (NCCL's version banner) write to fd 1: keep the real stdout for the ONE JSON line
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = capi.load()
    capi.check(lib.yh_set_device(local))
    ctx = dict(world=world, rank=rank, local=local)

    points = [sweep_hair(p) for p in SWEEP] if args.workload == "c4" else [None]
    runs = [run_cuda_point(args, hair, ctx) for hair in points]

    if rank == 0:
        r0 = runs[0]
        tot = lambda k: sum(r[k] for r in runs)  # noqa: E731
        samples, total_ms = tot("samples"), tot("total_ms")
        value = samples / (total_ms * 1e-3)
        e2e_value = samples / (tot("e2e_total_ms") * 1e-3)
        steps_all = tot("steps")
        peak, peak_src = measured_peak()
        roof = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
                "peak_source": peak_src, "kernel": "k_intersect"}
        try:
            probe = traversal_stats_probe(args.workload, args.sky, 2, args.max_split, r0["hair"])
            nodes_per_ray = probe["nodes"] / probe["rays"]
            prims_per_ray = probe["prims"] / probe["rays"]
            # one 32-byte slot per node step and per primitive test (yh_scene.cuh), + ray in (32 B) and hit out (24 B)
            bytes_per_ray = 32.0 * nodes_per_ray + 32.0 * prims_per_ray + 56.0
            # rank 0's own launches: one k_intersect launch per wavefront iteration
            local_rays, local_iters, local_trav_ms = tot("local_rays"), tot("local_iters"), tot("local_trav_ms")
            bytes_per_launch = bytes_per_ray * local_rays / max(1, local_iters)
            ms_per_launch = local_trav_ms / max(1, local_iters)
            roof.update(achieved=bytes_per_launch / (ms_per_launch * 1e-3) / 1e9, launches=local_iters,
                        nodes_per_ray=nodes_per_ray, prims_per_ray=prims_per_ray, bytes_per_ray=bytes_per_ray,
                        bytes_per_launch=bytes_per_launch, ms_per_launch=ms_per_launch,
                        kernel_share_of_step=tot("trav_max_ms") / tot("prof_max_ms"),
                        timing="CUDA events around every k_intersect launch in a serialised pass of %d step(s) after the timed region "
                               "(in the timed region the wavefront groups overlap on separate streams)" % r0["prof_steps"])
            roof["frac"] = roof["achieved"] / peak
            prof = ROOT / "profiles" / "traffic_k_intersect.json"
            if prof.exists():
                table = json.loads(prof.read_text())
                entry = table.get(args.workload) if isinstance(table.get(args.workload), dict) else (table if args.workload == "c3" and "dram_bytes_per_launch" in table else None)
                if entry:
                    roof["traffic"] = entry.get("dram_bytes_per_launch")
                    roof["l2_gbs"] = entry.get("l2_gbs")
                    # DRAM sees far less than the algorithmic bytes: the tree is L2 resident and the kernel sits at no
                    # memory roof - the HBM peak stays the stated denominator, the label says what bounds it
                    if roof["traffic"] and roof["traffic"] < 0.25 * bytes_per_launch:
                        roof["bound"] = "l2-latency (tree L2-resident: DRAM traffic << algorithmic bytes; frac is still quoted against the HBM peak)"
        except Exception as exc:  # the bench line must still print
            roof["error"] = str(exc)[:300]
        world, shard, spp = r0["world"], r0["shard"], r0["spp"]
        bvh = r0["bvh"]
        config = {"workload": r0["text"], "spp_per_step": spp, "sky_texture": args.sky,
                  "sky_size": list(SKY_SIZE) if args.sky and WORKLOADS[args.workload][0] in ("curly", "straight") else None,
                  "parallelism": f"{shard}{world}",
                  "rng": ("counter (per-sample PCG32 streams; statistically equivalent image); slots per pixel: %s" % (args.slots or "auto")) if shard == "samples"
                         else "reference (one PCG32 stream per pixel; image bit-identical to the CPU reference)",
                  "l2_policy": "working set larger than L2 is not guaranteed; rays are regenerated every step "
                               "(no input is re-read), see DESIGN.md",
                  "bvh": {"prims": bvh.num_primitives, "refs": bvh.num_references, "nodes": bvh.num_nodes,
                          "node_bytes": bvh.node_bytes, "prim_bytes": bvh.prim_bytes, "max_split": args.max_split}}
        if len(runs) > 1:
            config["points"] = [dict(r["hair"], msamples_per_sec=round(r["samples"] / r["total_ms"] / 1e3, 2)) for r in runs]
        line = {
            "metric": "samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": steps_all,
            "warmup": tot("warmup"), "ms_per_step": total_ms / steps_all, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "mrays_per_sec": tot("rays_all") / (total_ms * 1e-3) / 1e6, "light_mrays_per_sec": tot("light_rays_all") / (total_ms * 1e-3) / 1e6,
            "rays_per_sample": tot("rays_all") / samples, "wall_ms_per_step": tot("wall_max_ms") / steps_all,
            "traverse_ms": tot("trav_max_ms"), "shade_ms": tot("shade_max_ms"), "profiled_pass_ms": tot("prof_max_ms"),
            "reduce_ms": tot("reduce_ms"), "reduce_complete": r0["reduce_ok"],
            "gpu_launches": int(tot("launches_all")), "clocks": r0["clocks"], "render_mean": r0["checksum"],
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": tot("h2d") // steps_all,
                    "d2h_bytes_per_step": r0["width"] * r0["height"] * 16, "setup_ms": tot("setup_ms"),
                    "reduce_resolve_download_ms": tot("e2e_reduce_ms") if world > 1 else None,
                    "per_sample_call": r0["dropin"],
                    "note": "timed region = yh_scene_create (H2D of all scene arrays) + init_bvh + init_lights + yh_state_create + "
                            "K x (yh_trace_samples + " + ("NCCL reduce of accumulation and counts to rank 0 + yh_resolve_dev + download of the "
                            "COMPLETE image into pinned host memory on rank 0)" if world > 1 else "yh_state_download into pinned host memory)")},
            "roofline": roof, "cpu_baseline": r0["cpu"],
        }
        if r0["counter_value"] is not None:
            line["n1_counter_mode_value"] = r0["counter_value"]  # the N > 1 default's mode on this one GPU
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


def run_group(args):
    """--group: ONE process drives all GPUs through the C ABI's yh_group_* (no torch.distributed): the path a C++
    caller takes (yscenetrace_b200 --gpus N).  Same metric, same JSON shape; the collective is the fused
    peer-memory reduce + resolve kernel."""
    import torch
    from yocto_hair_b200.pathtrace import DeviceGroup
    from yocto_hair_b200.scene import make_params
    n = args.gpus
    scene, cam, res, text = build_workload(args.workload, args.sky)
    shard = args.shard if args.shard != "auto" else ("tiles" if n == 1 else "samples")
    spp = args.spp_step
    params = make_params(resolution=res)
    grp = DeviceGroup(scene, num_devices=n, max_split=args.max_split).init_state(cam, params, shard)
    for _ in range(args.warmup):
        grp.trace_samples(spp)
    sampler = ClockSampler(0)
    sampler.start()
    dev_ms, rays, launches = 0.0, 0, 0
    for _ in range(args.steps):
        s = grp.trace_samples(spp)
        dev_ms += s.device_ms; rays += s.rays; launches += s.launches
    _, a, cnt = grp.download()
    _, reduce_ms = grp.stats()
    clocks = sampler.stop()
    complete = bool((cnt == spp * (args.warmup + args.steps)).all()) and bool(np.isfinite(a).all())
    if not complete:
        raise SystemExit("bench --group: the reduced image is incomplete - no number is reported")
    samples = grp.width * grp.height * spp * args.steps
    value = samples / ((dev_ms + reduce_ms) * 1e-3)
    grp.close()
    render = torch.empty((res, res, 4), dtype=torch.float32).pin_memory().numpy()
    t0 = time.perf_counter()
    grp = DeviceGroup(scene, num_devices=n, max_split=args.max_split).init_state(cam, params, shard)
    setup_ms = 1e3 * (time.perf_counter() - t0)
    if render.shape[:2] != (grp.height, grp.width):
        render = np.zeros((grp.height, grp.width, 4), np.float32)
    for _ in range(args.steps):
        grp.trace_samples(spp)
        grp.download_into(render)
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    _, reduce_ms2 = grp.stats()
    line = {"metric": "samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": (dev_ms + reduce_ms) / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": text, "spp_per_step": spp, "sky_texture": args.sky, "parallelism": f"{shard}{n}, one process (yh_group_*)",
                       "collective": "fused peer-memory reduce + resolve kernel on GPU 0 (loads over NVLink)"},
            "mrays_per_sec": rays / (dev_ms * 1e-3) / 1e6, "reduce_ms": reduce_ms, "reduce_complete": complete, "gpu_launches": int(launches),
            "clocks": clocks, "render_mean": float(render[..., :3].astype(np.float64).mean()),
            "e2e": {"value": samples / (e2e_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": scene_bytes(scene) * n // args.steps,
                    "d2h_bytes_per_step": grp.width * grp.height * 16, "setup_ms": setup_ms, "reduce_resolve_ms_last": reduce_ms2,
                    "note": "timed region = yh_group_create (scene upload + BVH + lights on every GPU) + yh_group_init_state + K x "
                            "(yh_group_trace_samples + yh_group_download of the complete image into pinned host memory)"}}
    grp.close()
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--spp-step", type=int, default=128, help="samples per pixel per step (CUDA arm): one yh_trace_samples call per rank")
    ap.add_argument("--ref-spp-step", type=int, default=2, help="samples per pixel per step (reference arm)")
    ap.add_argument("--max-split", type=int, default=0)
    ap.add_argument("--shard", default="auto", choices=["auto", "tiles", "samples"],
                    help="multi-GPU sharding: image tiles (reference RNG streams) or sample ranges (counter-mode streams)")
    ap.add_argument("--slots", type=int, default=0, help="counter mode: concurrent path slots per pixel (0 = auto: fill the device; 1 = one slot per pixel)")
    ap.add_argument("--group", action="store_true", help="one process drives all --gpus through yh_group_* (no torchrun)")
    ap.add_argument("--no-sky", dest="sky", action="store_false", help="uniform environment instead of the sky texture")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-counter-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.group:
        run_group(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()

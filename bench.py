#!/usr/bin/env python3
"""Headline benchmark: samples/s (and Mrays/s) of the hair path-tracing hot path.

    python bench.py --gpus N --steps K --warmup W          # CUDA arm
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU reference arm

Workload (BASELINE.json configs[2], the one `metric` is quoted on): synthetic
curly hair, 1M line segments (4 lines / Bezier), blonde sigma_a = {0.06, 0.10,
0.20}, 1280 x 1280, path integrator, 8 bounces, tile-sharded over N GPUs.
A "step" is one pass of `--spp-step` samples per pixel over the whole image
(= spp_step consecutive trace_samples calls of the reference).

One JSON line is printed by rank 0 (see the contract in the task statement):
  value     whole-job samples/s, scene and state resident in HBM, device timed
  e2e       same metric through the C ABI from HOST buffers: scene upload + BVH
            build + light tables + state creation + K x (trace + download of the
            render image into pinned host memory)
  roofline  traversal kernel (k_intersect): algorithmic bytes / measured time
  cpu_baseline  the reference's own CPU code on a bounded sample of the workload
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

WORKLOADS = {
    # name: (generator, segments, resolution, config text)
    "c3": ("curly", 1_000_000, 1280, "synthetic curly hair, 1M segments (4 lines/Bezier), blonde sigma_a={0.06,0.10,0.20}, 1280x1280, path, 8 bounces"),
    "c2": ("straight", 250_000, 720, "synthetic straight hair, 250k segments (2 lines/Bezier), eumelanin=1.3, 720x720, path, 8 bounces"),
    "c5": ("curly", 5_000_000, 1920, "synthetic head of hair, 5M segments, 1920x1920, path, 8 bounces"),
    "tiny": ("curly", 40_000, 256, "tiny curly hair (40k segments, 256x256) for harness checks"),
}


def build_workload(name: str, sky: bool):
    from yocto_hair_b200 import synth
    kind, segs, res, text = WORKLOADS[name]
    if kind == "curly":
        scene, cam = synth.curly_hair_scene(segs, sky=sky)
    else:
        scene, cam = synth.straight_hair_scene(segs, sky=sky)
    return scene, cam, res, text


def scene_bytes(scene) -> int:
    n = 0
    for s in scene.shapes:
        n += s.positions.nbytes + s.normals.nbytes + s.radius.nbytes + s.lines.nbytes + s.triangles.nbytes
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


def traversal_stats_probe(workload: str, sky: bool, spp: int, max_split: int) -> dict:
    """Runs the -DYH_TRAVERSAL_STATS build once (outside any timed region) and
    returns nodes / primitives fetched per ray for this workload."""
    env = dict(os.environ, YH_B200_LIB=str(ROOT / "yocto-hair_b200" / "csrc" / "libyocto_hair_b200_stats.so"))
    code = (
        "import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import bench\n"
        "from yocto_hair_b200.pathtrace import DeviceScene\nfrom yocto_hair_b200.scene import make_params\n"
        "sc, cam, res, _ = bench.build_workload(%r, %r)\n"
        "ds = DeviceScene(sc).init_bvh(%d).init_lights()\n"
        "st = ds.init_state(cam, make_params(resolution=res)); s = st.trace_samples(%d)\n"
        "print(json.dumps(dict(rays=s.rays, nodes=s.nodes_visited, prims=s.prims_tested)))\n"
    ) % (str(ROOT), str(ROOT / "yocto-hair_b200"), workload, sky, max_split, spp)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    if out.returncode != 0:
        raise RuntimeError("stats probe failed: " + out.stderr[-2000:])
    return json.loads(out.stdout.strip().splitlines()[-1])


def run_reference(args):
    """CPU arm: the reference's own code (oracle/_ref) or the port, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle
    from yocto_hair_b200.scene import make_params
    orc = oracle.best_oracle()
    scene, cam, res, text = build_workload(args.workload, args.sky)
    t0 = time.time()
    rs = orc.scene(scene)
    build_s = time.time() - t0
    st = rs.state(cam, make_params(resolution=res))
    spp = args.ref_spp_step
    for _ in range(args.warmup):
        st.trace_samples(1)
    secs = [st.trace_samples(spp) for _ in range(args.steps)]
    total = sum(secs)
    samples = st.width * st.height * spp * args.steps
    value = samples / total
    cores = orc.hardware_concurrency()
    sample = f"{args.steps} steps x {spp} spp of the {res}x{res} workload ({samples/1e6:.1f} M samples), all {cores} host threads"
    print(json.dumps({
        "impl": "reference", "metric": "samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": text, "spp_per_step": spp, "sky_texture": args.sky, "bvh_build_s": round(build_s, 3)},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": orc.kind, "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline(args, scene, cam, res):
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle
    from yocto_hair_b200.scene import make_params
    orc = oracle.best_oracle()
    rs = orc.scene(scene)
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
            "sample": f"same scene and camera, {spp} spp of {res}x{res} ({samples/1e6:.1f} M samples, {secs:.1f} s), "
                      f"std::thread::hardware_concurrency() = {cores} threads"}


def run_cuda(args):
    import torch
    import torch.distributed as dist
    from yocto_hair_b200 import _capi as capi
    from yocto_hair_b200 import parallel
    from yocto_hair_b200.pathtrace import DeviceScene
    from yocto_hair_b200.scene import make_params

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

    scene, cam, res, text = build_workload(args.workload, args.sky)
    params = make_params(resolution=res)
    spp = args.spp_step
    # sharding: image tiles keep the reference's one-stream-per-pixel RNG (bit-identical images) but a pixel's
    # samples are sequential, so the wavefront narrows with 1/N; sample-range sharding keeps it full width
    # (counter-mode streams).  auto = tiles on one GPU, samples on several.
    shard = args.shard if args.shard != "auto" else ("tiles" if world == 1 else "samples")
    if shard == "samples" and spp % world:
        raise SystemExit(f"--spp-step {spp} must be a multiple of the number of GPUs for sample sharding")
    spp_rank = spp // world if shard == "samples" else spp

    def make_state(ds):
        st = ds.init_state(cam, params)
        if shard == "samples":
            st.set_sample_shard(rank, world)
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
    for _ in range(args.warmup):
        st.trace_samples(spp_rank)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_ms = 0.0
    rays = light_rays = launches = iters = 0
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        s = st.trace_samples(spp_rank)
        dev_ms += s.device_ms
        rays += s.rays; light_rays += s.light_rays; launches += s.launches; iters += s.iterations
    reduce_ms, reduce_ok = 0.0, None
    if world > 1:  # one NCCL reduce of the fp32 accumulation buffers over NVLink, at the end
        accum, nsamp = parallel.state_tensors(st, f"cuda:{local}")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        parallel.reduce_image(accum, nsamp, 0)
        e1.record()
        torch.cuda.synchronize()
        reduce_ms = e0.elapsed_time(e1)
        if rank == 0:  # after the reduce rank 0 holds the whole image: every pixel has all its samples
            want = spp * (args.warmup + args.steps)
            reduce_ok = bool((nsamp == want).all().item()) and bool(torch.isfinite(accum).all().item())
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall)
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel timing pass (outside the timed region): with profiling on, the wavefront groups run on ONE
    # stream and every traversal / shading launch is bracketed by CUDA events on it
    st.set_profiling(True)
    prof_steps = max(1, min(2, args.steps))
    trav_ms = shade_ms = prof_dev_ms = 0.0
    prof_rays = prof_iters = 0
    for _ in range(prof_steps):
        s = st.trace_samples(spp_rank)
        trav_ms += s.traverse_ms; shade_ms += s.shade_ms; prof_dev_ms += s.device_ms
        prof_rays += s.rays; prof_iters += s.iterations
    st.set_profiling(False)
    local_rays, local_iters, local_trav_ms = prof_rays, prof_iters, trav_ms
    t = torch.tensor([dev_ms + reduce_ms, wall_ms, trav_ms, shade_ms, prof_dev_ms], dtype=torch.float64, device=f"cuda:{local}")
    c = torch.tensor([rays, light_rays, launches, iters], dtype=torch.int64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    total_ms, wall_max_ms, trav_max_ms, shade_max_ms, prof_max_ms = t.tolist()
    rays_all, light_rays_all, launches_all, iters_all = c.tolist()
    samples = width * height * spp * args.steps
    value = samples / (total_ms * 1e-3)

    # ---------------- e2e arm: host buffers through the C ABI ----------------------
    st.close(); ds.close()
    render_host = torch.empty((height, width, 4), dtype=torch.float32).pin_memory().numpy()
    barrier()
    t0 = time.perf_counter()
    ds2 = DeviceScene(scene).init_bvh(args.max_split).init_lights()
    st2 = make_state(ds2)
    setup_ms = 1e3 * (time.perf_counter() - t0)
    for _ in range(args.steps):
        st2.trace_samples(spp_rank)
        st2.download_into(render_host)
    barrier()
    e2e_ms = torch.tensor([1e3 * (time.perf_counter() - t0)], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = samples / (e2e_ms.item() * 1e-3)
    checksum = float(np.asarray(render_host[..., :3], dtype=np.float64).mean())
    st2.close(); ds2.close()

    if rank == 0:
        peak, peak_src = measured_peak()
        roof = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
                "peak_source": peak_src, "kernel": "k_intersect"}
        try:
            probe = traversal_stats_probe(args.workload, args.sky, 2, args.max_split)
            nodes_per_ray = probe["nodes"] / probe["rays"]
            prims_per_ray = probe["prims"] / probe["rays"]
            # one 32-byte slot per node step and per primitive test (yh_scene.cuh), + ray in (32 B) and hit out (24 B)
            bytes_per_ray = 32.0 * nodes_per_ray + 32.0 * prims_per_ray + 56.0
            # rank 0's own launches: one k_intersect launch per wavefront iteration
            bytes_per_launch = bytes_per_ray * local_rays / max(1, local_iters)
            ms_per_launch = local_trav_ms / max(1, local_iters)
            roof.update(achieved=bytes_per_launch / (ms_per_launch * 1e-3) / 1e9, launches=local_iters,
                        nodes_per_ray=nodes_per_ray, prims_per_ray=prims_per_ray, bytes_per_ray=bytes_per_ray,
                        bytes_per_launch=bytes_per_launch, ms_per_launch=ms_per_launch,
                        kernel_share_of_step=trav_max_ms / prof_max_ms,
                        timing="CUDA events around every k_intersect launch in a serialised pass of %d step(s) after the timed region "
                               "(in the timed region the wavefront groups overlap on separate streams)" % prof_steps)
            roof["frac"] = roof["achieved"] / peak
            prof = ROOT / "profiles" / "traffic_k_intersect.json"
            if prof.exists():
                roof["traffic"] = json.loads(prof.read_text()).get("dram_bytes_per_launch")
        except Exception as exc:  # the bench line must still print
            roof["error"] = str(exc)[:300]
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(args, scene, cam, res)
        line = {
            "metric": "samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": text, "spp_per_step": spp, "sky_texture": args.sky, "parallelism": f"{shard}{world}",
                       "rng": "counter (per-sample PCG32 streams)" if shard == "samples" else "reference (one PCG32 stream per pixel)",
                       "l2_policy": "working set larger than L2 is not guaranteed; rays are regenerated every step "
                                    "(no input is re-read), see DESIGN.md",
                       "bvh": {"prims": bvh.num_primitives, "refs": bvh.num_references, "nodes": bvh.num_nodes,
                               "node_bytes": bvh.node_bytes, "prim_bytes": bvh.prim_bytes, "max_split": args.max_split}},
            "mrays_per_sec": rays_all / (total_ms * 1e-3) / 1e6, "light_mrays_per_sec": light_rays_all / (total_ms * 1e-3) / 1e6,
            "rays_per_sample": rays_all / samples, "wall_ms_per_step": wall_max_ms / args.steps,
            "traverse_ms": trav_max_ms, "shade_ms": shade_max_ms, "profiled_pass_ms": prof_max_ms, "reduce_ms": reduce_ms, "reduce_complete": reduce_ok,
            "gpu_launches": int(launches_all), "clocks": clocks, "render_mean": checksum,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": scene_bytes(scene) // args.steps,
                    "d2h_bytes_per_step": width * height * 16, "setup_ms": setup_ms,
                    "note": "timed region = yh_scene_create (H2D of all scene arrays) + init_bvh + init_lights + "
                            "yh_state_create + K x (yh_trace_samples + yh_state_download into pinned host memory)"},
            "roofline": roof, "cpu_baseline": cpu,
        }
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--spp-step", type=int, default=64, help="samples per pixel per step (CUDA arm): one yh_trace_samples call per rank")
    ap.add_argument("--ref-spp-step", type=int, default=2, help="samples per pixel per step (reference arm)")
    ap.add_argument("--max-split", type=int, default=0)
    ap.add_argument("--shard", default="auto", choices=["auto", "tiles", "samples"],
                    help="multi-GPU sharding: image tiles (reference RNG streams) or sample ranges (counter-mode streams)")
    ap.add_argument("--no-sky", dest="sky", action="store_false", help="uniform environment instead of the sky texture")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()

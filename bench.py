#!/usr/bin/env python
"""bench.py — the hot path's headline metric (BASELINE.json): Mrays/s and s/frame of a 1920x1080, 64 spp
path-traced frame of a ~1M-triangle synthetic mesh, with the HBM roofline fraction of the dominant kernel
and the reference's own CPU path timed on the same box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload north_star|north_star_motion|config2|small]

A step = one frame.  Own arm: the scene is built through libfjscene (the host mirror of fj_scene_interface) from
`.scn` text and rendered by libfjgpu (sm_100a).  `value` is timed with the scene resident in HBM and the frame left
on the device; `e2e` is the same frame through SiRenderScene with host buffers (scene arrays re-sent from pinned
host memory every step, pixels copied back to the host framebuffer).  N > 1: tiles are sharded round-robin over the
ranks (no data-path collective while tracing), one NCCL all-gather of the packed tile blocks ends the frame.
--impl reference: the UNMODIFIED reference (oracle/_ref, built from /root/reference by oracle/Makefile.ref) renders a
bounded render_region of the same frame on all host cores.
"""
import argparse
import ctypes as C
import gc
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
import __graft_entry__ as entry  # noqa: E402

WORKLOADS = {
    # name: (scene builder, kwargs, description)
    "north_star": ("pathtracing_blob", dict(n=707, res=(1920, 1080), rate=8, depth=3),
                   "S-blob(707)=999698 tris + emissive shell, pathtracing_shader depth 3, 1920x1080, 8x8=64spp, tile 32, filter 2"),
    "north_star_motion": ("pathtracing_blob", dict(n=707, res=(1920, 1080), rate=8, depth=3, motion=True),
                          "north-star scene with motion blur: the blob's rotate and the camera's translate carry two time samples (SURVEY.md 8f row 4)"),
    "config2": ("plastic_blob", dict(n=187, res=(1280, 720), rate=4),
                "S-blob(187)=69938 tris, plastic_shader + 1 point light, 1280x720, 4x4=16spp"),
    "config3": ("pathtracing_blob", dict(n=1871, res=(1920, 1080), rate=8, depth=3),
                "S-blob(1871)=7.0M tris + emissive shell (dragon stand-in), pathtracing_shader depth 3, 1920x1080, 64spp"),
    "config4": ("instanced_blobs", dict(n=740, res=(1920, 1080), rate=8),
                "16 instances of S-blob(740)=1.09M tris + floor, plastic_shader, GridLight 16 samples, 1920x1080, 64spp"),
    "config5": ("pathtracing_soup", dict(ntris=10_000_000, res=(3840, 2160), rate=16, depth=8),
                "S-random 10M-triangle soup + emissive shell, pathtracing_shader depth 8, 3840x2160, 256spp"),
    "profile": ("pathtracing_blob", dict(n=707, res=(480, 270), rate=8, depth=3),
                "north-star scene at 480x270 (1/16 of the frame) for ncu --set full captures"),
    "small": ("pathtracing_blob", dict(n=64, res=(320, 180), rate=4, depth=3),
              "S-blob(64)=8192 tris + shell, pathtracing_shader, 320x180, 16spp (plumbing check)"),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workdir():
    d = os.environ.get("FJ_BENCH_DIR") or os.path.join(tempfile.gettempdir(), "fj_bench_%d" % os.getuid())
    os.makedirs(d, exist_ok=True)
    return d


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).

    nvidia-smi is started during the warm-up and polls about three times per timed region: measured on a B200 (profiles/
    r2b_clock_sampler_period.txt) a sampler started right at the timed region with `-lms 200` cost the 20-step north-star run
    10-175 ms per 351-ms step (NVML start-up and every poll stall the launching thread), 22 ms at 1000 ms, 0.4 ms at 5000 ms;
    started before the warm-up with a 1-s period the same run still lost 1-16 ms per step.  (What remains with a single poll
    per region, 3-17 ms per step, was not the sampler: two or three of the twenty steps took 20-70 ms longer on the host side
    with identical kernel times — `step_wall_ms` in the JSON line shows them.  It was libfjgpu's cudaMemGetInfo per frame, which
    queues behind whoever holds the kernel driver's lock; with the batch budget cached the gap is 0.3-0.4 ms.)
    Only the rows that arrive between mark() and stop() are reported; a region shorter than one period falls back to the last
    rows of the warm-up (same load) and says so."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0, period_ms=1000):
        self.rows, self.proc, self.index = [], None, index
        self.period_ms = int(os.environ.get("FJ_CLOCK_LMS", period_ms))
        self.t_mark = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", str(max(50, self.period_ms))], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def wait_first_row(self, timeout_s=6.0):
        """Blocks until nvidia-smi has delivered its first row (NVML is up), so that its start-up stays out of the timed region."""
        t_end = time.perf_counter() + timeout_s
        while self.proc and not self.rows and time.perf_counter() < t_end and self.proc.poll() is None:
            time.sleep(0.02)

    def mark(self):
        """Start of the timed region."""
        self.t_mark = time.perf_counter()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.perf_counter()
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        t0 = self.t_mark if self.t_mark is not None else 0.0
        inside = [r for t, r in self.rows if t0 <= t <= t_end]
        note = None
        if not inside:                       # a timed region shorter than one polling period
            inside = [r for t, r in self.rows][-2:]
            note = "timed region shorter than the polling period: last rows of the warm-up (same load)"
        sm, mx, reasons = [], [], set()
        for r in inside:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm), "period_ms": self.period_ms}
        if note:
            out["note"] = note
        return out


def algorithmic_bytes(stats, ninst):
    """SURVEY.md §8(d) single-path model (FP32 layout figures): per ray that hits, 228 B (ray 32 + instance matrices 96
    + indices 12 + positions 36 + normals 36 + result 16) + 64 B per BVH level on one root-to-leaf path of the hit
    mesh and of the instance tree; a miss reads the ray and one root node (96 B); +32 B per camera sample for the
    sample write + filter read."""
    l_tlas = max(0, math.ceil(math.log2(max(ninst, 1))))
    rays = stats.rays
    hit = stats.rays_hit
    return hit * (228 + 64 * l_tlas) + 64 * stats.hit_mesh_levels + (rays - hit) * 96 + 32 * stats.camera_samples


# ---------------------------------------------------------------------------------------------- CPU arms
# Both CPU legs (`--impl reference` and the own arm's `cpu_baseline`) time the UNMODIFIED reference (oracle/_ref, built from
# /root/reference by oracle/Makefile.ref) on a bounded sample of the SAME frame: tile-row strips drawn from a seeded
# permutation of ALL tile rows of the frame (a step = one strip; over the steps the sample covers the frame top to bottom,
# background and object alike — round 1 sampled centre tiles only).  Mrays/s = rays of exactly those tiles / seconds; the
# reference has no ray counter, so the rays of the sampled tiles are counted after the timed region (libfjgpu's per-type
# counters for the same tiles when a GPU is present, else the oracle port) — a unit conversion, never part of the timing.
STRIP_SEED = 20261017


def ref_paths():
    ref = os.path.join(REPO, "oracle", "_ref")
    probe = os.path.join(ref, "bin", "ref_probe")
    return ref, probe


def run_reference(scene_text, regions, timeout=3000):
    """Runs the unmodified reference: one RenderScene per entry of `regions` (render_region xmin ymin xmax ymax) in ONE process.
    Returns the per-frame seconds measured between its frame-start and frame-done callbacks."""
    ref, probe = ref_paths()
    txt = scene_text + "".join("SetProperty4 ren1 render_region %d %d %d %d\nRenderScene ren1\n" % tuple(r) for r in regions)
    scn = os.path.join(workdir(), "ref_%d.scn" % os.getpid())
    with open(scn, "w") as f:
        f.write(txt)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ref, "lib") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    res = subprocess.run([probe, "run", scn], env=env, capture_output=True, text=True, timeout=timeout)
    if res.returncode != 0:
        raise RuntimeError("reference failed: " + res.stdout[-1500:] + res.stderr[-1500:])
    secs = [float(l.split()[1]) for l in res.stdout.split("\n") if l.startswith("FJ_FRAME_SECONDS")]
    if len(secs) != len(regions):
        raise RuntimeError("reference reported %d frames for %d regions" % (len(secs), len(regions)))
    return secs


def strip_regions(res, n, tiles_per_strip, tile=32):
    """`n` tile-row strips of `tiles_per_strip` tiles each: rows from a seeded permutation of all tile rows (cycled), the
    strip's first column seeded too when it is narrower than the frame."""
    import random
    rng = random.Random(STRIP_SEED)
    tx, ty = -(-res[0] // tile), -(-res[1] // tile)
    rows = list(range(ty))
    rng.shuffle(rows)
    w = max(1, min(tx, tiles_per_strip))
    out = []
    for k in range(n):
        row = rows[k % ty]
        c0 = rng.randrange(0, tx - w + 1)
        out.append((c0 * tile, row * tile, min((c0 + w) * tile, res[0]), min((row + 1) * tile, res[1])))
    return out


def tiles_of_regions(res, regions, tile=32):
    """The frame's tiles (full-frame ids) covered by tile-aligned `regions`, one list per region."""
    from fujiyama_renderer_b200 import sharding
    all_tiles = sharding.make_tiles(res[0], res[1], tile)
    per_row = -(-res[0] // tile)
    out = []
    for x0, y0, x1, y1 in regions:
        out.append([all_tiles[(y0 // tile) * per_row + c] for c in range(x0 // tile, -(-x1 // tile))])
    return out


def count_rays(builder, kw, tile_lists, prefer_gpu=True):
    """Rays (all types) and camera samples of the given tiles of the workload's frame, per list.  Returns (counts, how)."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import workloads
    import scenekit as sk
    desc = workloads.desc_for(builder, kw)
    st = desc.to_structs()
    how = None
    out = []
    try:
        import torch
        has_gpu = prefer_gpu and torch.cuda.is_available() and os.environ.get("FJ_REF_COUNT", "gpu") != "oracle"
    except Exception:
        has_gpu = False
    if has_gpu:
        from fujiyama_renderer_b200 import device
        dev = device.Device(int(os.environ.get("LOCAL_RANK", "0")))
        try:
            dev.load_structs(st)
            for tl in tile_lists:
                s_ = dev.render_resident(st["params"], tl)
                out.append((s_.rays, s_.camera_samples))
        finally:
            dev.close()
        how = "libfjgpu per-type ray counters for the same tiles"
    else:
        for tl in tile_lists:
            _, s_ = sk.oracle_render(desc, rng_mode=0, threads=min(os.cpu_count() or 1, 64), st=st, tiles=tl)
            out.append((s_.rays, s_.camera_samples))
        how = "oracle port (counter RNG) on the same tiles"
    return out, how


def cpu_sample(builder, kw, nframes, budget_s):
    """Times the unmodified reference on `nframes` strips sized so that all of them take about `budget_s` seconds.
    Returns dict(secs, regions, tile_lists, threads, calibration)."""
    from fujiyama_renderer_b200 import scenes
    ref, probe = ref_paths()
    threads = min(os.cpu_count() or 1, 64)
    text = getattr(scenes, builder)(workdir(), os.path.join(ref, "lib"), threads=threads, **kw)
    res = kw["res"]
    tx = -(-res[0] // 32)
    # calibrate on one strip of `threads` tiles in the first row of the permutation's middle (>= one tile per worker thread)
    cal = strip_regions(res, 1, min(tx, max(threads, 4)))
    t_cal = run_reference(text, cal)[0]
    cal_tiles = len(tiles_of_regions(res, cal)[0])
    per_tile = max(t_cal, 1e-3) / cal_tiles
    want = int(budget_s / max(nframes, 1) / per_tile)
    width = max(min(tx, max(threads, 4)), min(tx, want))
    regions = strip_regions(res, nframes, width)
    secs = run_reference(text, regions)
    return {"secs": secs, "regions": regions, "tile_lists": tiles_of_regions(res, regions), "threads": threads,
            "calibration": {"seconds": t_cal, "tiles": cal_tiles}, "strip_tiles": width, "frame_tiles": [tx, -(-res[1] // 32)]}


def shared_config(args, desc, world):
    """`config` is identical in both arms (the driver compares them): the workload, not the measurement."""
    return {"workload": args.workload, "scene": desc,
            "l2": "GPU arm: 256 MiB flush between steps; scene + 2.3 GB/frame sample stream exceed L2 (not applicable to the CPU arm)",
            "parallelism": "GPU arm: tiles round-robin over %d rank(s)%s; CPU arm: the reference's worker threads over the tiles of a strip" % (
                world, ", one NCCL all_gather of tile blocks" if world > 1 else "")}


def reference_arm(args, builder, kw, desc):
    ref, probe = ref_paths()
    base = {"impl": "reference", "metric": "Mrays/s", "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": shared_config(args, desc, args.gpus)}
    if not os.path.exists(probe):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not built (needs /root/reference at build time)"}))
        return
    n = args.warmup + args.steps
    smp = cpu_sample(builder, kw, n, float(os.environ.get("FJ_REF_BUDGET_S", "150")))
    secs = smp["secs"][args.warmup:]
    counts, how = count_rays(builder, kw, smp["tile_lists"][args.warmup:])
    rays = sum(c[0] for c in counts)
    samples = sum(c[1] for c in counts)
    t = sum(secs)
    mrays = rays / t / 1e6
    sample = ("%d strips of %d tiles (rows %s of %d, seeded permutation %d) = %d camera samples, %d rays (%s), %d threads" % (
        len(secs), smp["strip_tiles"], [r[1] // 32 for r in smp["regions"][args.warmup:]], smp["frame_tiles"][1], STRIP_SEED,
        samples, rays, how, smp["threads"]))
    out = dict(base, value=mrays, ms_per_step=1e3 * t / len(secs),
               cpu_baseline={"value": mrays, "unit": "Mrays/s", "cores": smp["threads"], "kind": "reference", "sample": sample,
                             "camera_samples_per_s": samples / t, "rays_per_sample": rays / max(samples, 1),
                             "calibration": smp["calibration"]},
               e2e={"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(out))


def cpu_baseline_leg(builder, kw):
    """Bounded sample of the same workload on the host cores (about 15 s of CPU work) for the own arm's line."""
    ref, probe = ref_paths()
    if not os.path.exists(probe):
        return None
    nframes = 2
    smp = cpu_sample(builder, kw, nframes, float(os.environ.get("FJ_CPU_BASELINE_S", "16")))
    counts, how = count_rays(builder, kw, smp["tile_lists"])
    rays, samples, t = sum(c[0] for c in counts), sum(c[1] for c in counts), sum(smp["secs"])
    return {"value": rays / t / 1e6, "unit": "Mrays/s", "cores": smp["threads"], "kind": "reference",
            "sample": "unmodified reference (oracle/_ref), %d strips of %d tiles (rows %s of %d, seeded) = %d camera samples, %d rays (%s), %.1f s" % (
                nframes, smp["strip_tiles"], [r[1] // 32 for r in smp["regions"]], smp["frame_tiles"][1], samples, rays, how, t),
            "camera_samples_per_s": samples / t, "rays_per_sample": rays / max(samples, 1)}


# ---------------------------------------------------------------------------------------------- parity leg
def parity_leg(builder, kw, frame, ntiles=6):
    """Holds the frame the bench just timed to the oracle: `ntiles` tiles drawn (seeded) from ALL tiles of the frame, rendered by
    the oracle port with the same counter RNG and the tiles' full-frame ids, compared with the same pixels of the timed GPU
    frame; the same tiles rendered once more through the C-ABI give the per-type ray counts to compare.  The oracle is the
    checker here, never the thing measured."""
    import random
    import numpy as np
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import workloads
    import scenekit as sk
    from fujiyama_renderer_b200 import device
    desc = workloads.desc_for(builder, kw)
    st = desc.to_structs()
    all_tiles = desc.tiles()
    rng = random.Random(STRIP_SEED + 1)
    pick = sorted(rng.sample(range(len(all_tiles)), min(ntiles, len(all_tiles))))
    tiles = [all_tiles[i] for i in pick]
    ref, rstats = sk.oracle_render(desc, rng_mode=0, threads=min(os.cpu_count() or 1, 64), st=st, tiles=tiles)
    dev = device.Device(int(os.environ.get("LOCAL_RANK", "0")))
    try:
        dev.load_structs(st)
        sub, gstats = dev.render(st["params"], tiles)
    finally:
        dev.close()
    mask = np.zeros(ref.shape[:2], bool)
    for _, x0, y0, x1, y1 in tiles:
        mask[y0:y1, x0:x1] = True
    d = frame[mask].astype(np.float64) - ref[mask].astype(np.float64)
    rmse = float(np.sqrt((d * d).mean(0)).max())
    keys = ("rays_camera", "rays_shadow", "rays_diffuse", "rays_reflect", "rays_refract", "camera_samples")
    return {"rmse": rmse, "max_abs": float(np.abs(d).max()), "bar": 1e-4, "tiles": [t[0] for t in tiles], "pixels": int(mask.sum()),
            "rays_equal": all(getattr(gstats, k) == getattr(rstats, k) for k in keys),
            "timed_frame_equals_subset_render": bool(np.array_equal(frame[mask], sub[mask])),
            "rays_checked": int(rstats.rays), "against": "oracle port (counter RNG, full-frame tile ids), pinned on the reference's golden vectors"}


# ---------------------------------------------------------------------------------------------- own arm
def own_arm(args, builder, kw, desc):
    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the renderer has no CPU fallback")
    torch.cuda.set_device(local)
    # everything but the final JSON line is kept off stdout: libfjscene prints the reference's progress lines there and NCCL
    # its version banner
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    sys.stdout.flush()
    os.dup2(devnull, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    entry.load_package()
    from fujiyama_renderer_b200 import fujiyama, scenes, abi, sharding

    wd = workdir()
    if rank == 0:
        text = getattr(scenes, builder)(wd, "/opt/fujiyama/lib", **kw)
    if world > 1:
        dist.barrier()
    text = getattr(scenes, builder)(wd, "/opt/fujiyama/lib", **kw)
    res, rate = kw["res"], kw["rate"]
    s = fujiyama.Session(echo=False, device=local, rank=rank, world_size=world)

    def frame():
        s.run("RenderScene ren1\n")
        return s.stats()

    try:
        s.run(text)
        tiles = sharding.make_tiles(res[0], res[1], 32)
        ntiles = len(tiles)
        my_tiles = len(sharding.rank_tiles(tiles, rank, world))
        max_tiles = sharding.blocks_per_rank(ntiles, world)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        blocks = gathered = None
        if world > 1:
            blocks = torch.zeros((max_tiles, 32, 32, 4), dtype=torch.float32, device="cuda")
            s.set_device_blocks(blocks.data_ptr(), 32, 32)
        else:
            s.set_resident(True)

        def step():
            st = frame()
            if world > 1:
                nonlocal gathered
                gathered = sharding.all_gather_blocks(blocks, world, dist)
            return st

        # The clock sampler starts after the first warm-up step, with a period of a third of the expected timed region (three
        # rows per run: every nvidia-smi poll stalls the launching thread for tens of ms, see ClockSampler), and the warm-up
        # goes on only when its first row has arrived.
        clocks = None
        for w in range(max(args.warmup, 1)):
            tw = time.perf_counter()
            st, info, upload_s = step()
            flush.zero_()
            if w == 0:
                torch.cuda.synchronize()
                step_ms = (time.perf_counter() - tw) * 1e3
                clocks = ClockSampler(local, period_ms=int(min(5000, max(500, args.steps * step_ms / 3))))
                if rank == 0:
                    clocks.start()
                    clocks.wait_first_row()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        gc.collect()
        gc.disable()                        # no collector pause between two frames of the timed region (re-enabled below)
        clocks.mark()
        t0 = time.perf_counter()
        ev0.record()
        tot = abi.Stats()
        ms_trace = ms_resolve = ms_shade = 0.0
        launches = trace_launches = 0
        step_wall = []
        for _ in range(args.steps):
            tw = time.perf_counter()
            st, info, _ = step()
            flush.zero_()                   # L2 flush between timed iterations (256 MiB > 126 MB L2)
            step_wall.append((time.perf_counter() - tw) * 1e3)
            for f, _t in abi.Stats._fields_:
                if f.startswith("rays_") or f in ("camera_samples", "hit_mesh_levels", "node_steps", "tri_tests", "leaf_phases", "leaf_rounds"):
                    setattr(tot, f, getattr(tot, f) + getattr(st, f))
            ms_trace += st.ms_trace
            ms_resolve += st.ms_resolve
            ms_shade += st.ms_shade
            launches += st.kernel_launches
            trace_launches += st.trace_launches
        ev1.record()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        gc.enable()
        dev_ms = ev0.elapsed_time(ev1)
        if world > 1:
            dist.barrier()
        clk = clocks.stop() if rank == 0 else None
        ms = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
        cnt = torch.tensor([tot.rays, tot.camera_samples, launches], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dev_ms, wall_ms = float(ms[0]), float(ms[1])
        total_rays, total_samples, total_launches = float(cnt[0]), float(cnt[1]), int(cnt[2])
        ms_per_step = dev_ms / args.steps
        value = total_rays / (dev_ms * 1e-3) / 1e6

        # roofline of the dominant kernel (k_extend, the closest-hit kernel) on this rank: CUDA events around every one of
        # its launches on the launching stream (fjgpu_stats.ms_trace / trace_launches)
        n_trace_launches = max(1, trace_launches)
        algo = algorithmic_bytes(tot, int(info.instances))
        peaks = {}
        try:
            with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        peak, peak_src = (float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs") if "hbm_gbs" in peaks else (6650.0, "fallback 6.65 TB/s")
        achieved = algo / (ms_trace * 1e-3) / 1e9 if ms_trace > 0 else 0.0
        traffic = None
        tpath = os.path.join(REPO, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(args.workload)
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "kernel": "k_extend", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": algo / n_trace_launches, "launches": n_trace_launches,
                    "avg_launch_ms": ms_trace / n_trace_launches, "bytes_per_ray": algo / max(tot.rays, 1),
                    "share_of_step": ms_trace / (dev_ms if world == 1 else max(dev_ms, 1e-9)),
                    "frac_of_8TBs": achieved / 8000.0}

        # e2e: the same frame through SiRenderScene with HOST buffers, the same way at every N — every rank re-sends its copy
        # of the scene arrays host -> device from pinned memory each step (what a host that edited or re-loaded the scene
        # pays), renders its tiles, and the frame lands in host memory: N = 1 fjgpu_render_tiles into the host framebuffer;
        # N > 1 tile blocks -> one NCCL all-gather -> rank 0 un-permutes on the device and copies the frame ONCE into pinned
        # host memory (fjgpu_assemble_frame)
        host_frame = None
        s.set_resend(True)
        if world == 1:
            s.set_resident(False)
            frame()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e_rays = 0
            for _ in range(args.steps):
                st, _, _ = frame()
                e_rays += st.rays
                fb = s.framebuffer("fb1")
            t_e2e = time.perf_counter() - t0
            host_frame = fb
            e2e = {"value": e_rays / t_e2e / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": s.resend_bytes(),
                   "d2h_bytes_per_step": int(fb.nbytes), "ms_per_step": 1e3 * t_e2e / args.steps,
                   "api": "libfjscene SiRenderScene -> fjgpu_scene_resend + fjgpu_render_tiles (host framebuffer)"}
        else:
            pinned = torch.empty((res[1], res[0], 4), dtype=torch.float32, pin_memory=True) if rank == 0 else None

            def e2e_step():
                st_ = step()
                if rank == 0:
                    s.assemble_gathered(gathered.data_ptr(), world, 32, 32, pinned.data_ptr())
                return st_

            e2e_step()
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            e_rays = 0
            for _ in range(args.steps):
                st, _, _ = e2e_step()
                e_rays += st.rays
            torch.cuda.synchronize()
            dist.barrier()
            t_e2e = time.perf_counter() - t0
            er = torch.tensor([float(e_rays), float(s.resend_bytes())], dtype=torch.float64, device="cuda")
            dist.all_reduce(er, op=dist.ReduceOp.SUM)
            if rank == 0:
                host_frame = pinned.numpy()
            e2e = {"value": float(er[0]) / t_e2e / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(er[1]),
                   "d2h_bytes_per_step": int(res[0] * res[1] * 16), "ms_per_step": 1e3 * t_e2e / args.steps,
                   "api": "libfjscene SiRenderScene per rank (fjgpu_scene_resend + fjgpu_render_tiles_device) -> NCCL all_gather of tile "
                          "blocks -> fjgpu_assemble_frame on rank 0 (one copy into pinned host memory)"}
        s.set_resend(False)
    finally:
        os.dup2(saved, 1)
        os.close(devnull)

    parity = None
    if rank == 0 and not args.no_parity and host_frame is not None:
        parity = parity_leg(builder, kw, host_frame, ntiles=int(os.environ.get("FJ_PARITY_TILES", "6")))
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline_leg(builder, kw)
        except Exception as e:      # the baseline is reported, never required for the GPU number
            cpu = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "reference", "sample": "failed: %s" % e}
    if rank == 0 and parity is not None and not (parity["rmse"] < parity["bar"] and parity["rays_equal"]):
        print(json.dumps({"error": "parity check failed", "parity": parity}))
        raise SystemExit("bench.py: the timed frame differs from the oracle: %r" % (parity,))
    if rank == 0:
        out = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_per_step, "s_per_frame": ms_per_step * 1e-3, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": shared_config(args, desc, world),
               "frame": {"rays_per_frame": total_rays / args.steps, "camera_samples_per_frame": total_samples / args.steps,
                         "scene_hbm_bytes": int(info.hbm_bytes), "bvh_build_s": info.build_seconds, "scene_upload_s": upload_s,
                         "blas_nodes": int(info.blas_nodes), "blas_max_depth": int(info.blas_max_depth),
                         "node_steps_per_ray": tot.node_steps / max(tot.rays, 1), "tri_tests_per_ray": tot.tri_tests / max(tot.rays, 1),
                         "tri_pairs_per_leaf_phase": tot.tri_tests / max(tot.leaf_phases, 1), "rounds_per_leaf_phase": tot.leaf_rounds / max(tot.leaf_phases, 1)},
               "wall_ms_per_step": wall_ms / args.steps, "step_wall_ms": {"min": min(step_wall), "median": sorted(step_wall)[len(step_wall) // 2], "max": max(step_wall)},
               "gpu_launches": total_launches, "clocks": clk,
               "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
               "kernel_ms_per_step": {"k_extend": ms_trace / args.steps, "k_generate+k_shade": ms_shade / args.steps,
                                      "k_resolve_tiles": ms_resolve / args.steps}}
        print(json.dumps(out))
    s.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="north_star", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    builder, kw, desc = WORKLOADS[args.workload]
    if os.environ.get("FJ_BENCH_N"):          # experiment knob: mesh resolution of the blob (not used by the driver)
        kw = dict(kw, n=int(os.environ["FJ_BENCH_N"]))
        desc += " [n=%d]" % kw["n"]
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        entry.load_package()
        reference_arm(args, builder, kw, desc)
        return
    own_arm(args, builder, kw, desc)


if __name__ == "__main__":
    main()

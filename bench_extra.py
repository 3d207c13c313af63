"""Secondary workloads of bench.py (BASELINE.json configs[2..4]); same launch contract, one JSON line on rank 0.

  --workload batch10k   configs[3]: a batch of independent 10k-point registrations sharded by index over the ranks
  --workload map1cm     configs[4]: full-res Kinect v1 frames integrated into a 600x600x500 1 cm grid, z-slab sharded,
                        one all-gather of the lifted points per frame
  --workload trajectory configs[2]: Kinect v2 trajectory: per-frame back-projection, ICP on a <=10k subsample
                        against the previous frame, full-res ray integration into 300x300x250 at 2 cm
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "icp-slam-prototype_b200", "python"))

ITERS = 20


def _setup(local, world):
    import torch
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return torch


def _hbm_peak():
    """Measured copy bandwidth of this pool's B200s (driver-written MEASURED_PEAKS.json), else the recipe's fallback."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(path))["hbm_gbs"] if os.path.exists(path) else 6650.0


def _cpu_port_rate(fn, units, what):
    """cpu_baseline for a secondary workload: the oracle port (oracle/, one thread) on a bounded sample, units / s."""
    try:
        from oracle import oracle as orc
        orc.build()
        t0 = time.perf_counter()
        fn(orc)
        dt = time.perf_counter() - t0
        return {"value": units / dt, "cores": 1, "kind": "port", "sample": f"{what} ({dt:.2f} s)"}
    except Exception as e:  # test infrastructure: the GPU number stands without it
        return {"unavailable": str(e)}


def _barrier(torch, world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(torch, world, local, value):
    if world == 1:
        return value
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _fp32_nominal():
    return 148 * 128 * 2 * 1.965e9 / 1e12


def run_batch10k(args, rank, world, local, emit=True):
    """configs[3]: args.batch independent 10k-point registrations, contiguous blocks of the batch per rank, no
    data-path collective; the poses are gathered at the end.  Strong scaling.  Returns the JSON line on rank 0."""
    import icpb200
    from icpb200 import dist as D
    from icpb200 import synth
    torch = _setup(local, world)
    ctx = icpb200.Context(local)
    K = icpb200.reference_intrinsics_v1()
    cam = np.array([5, 5, 5], np.float32)
    total = args.batch
    # the job's plumbing lives in libicpb200 (icpb_comm: NCCL): block ownership and the final gather of the poses;
    # torch.distributed only carries the 128-byte NCCL id
    comm = icpb200.Comm.from_torch(ctx) if world > 1 else None
    lo, hi = comm.shard_range(total) if comm is not None else (0, total)
    assert (lo, hi) == D.shard_range(total, rank, world)
    # 8 distinct rendered frame pairs; registration i uses pair i % 8 with its own seeded 10k subsample
    pairs = []
    full = ctx.cloud(640 * 480)
    for s in range(8):
        d0, d1, col, _ = synth.frame_pair(seed=synth.MASTER_SEED + 31 * s)
        full.from_depth(d0, col, K); full.transform(None, cam); t = full.download()
        full.from_depth(d1, col, K); full.transform(None, cam); d = full.download()
        pairs.append((d, t))
    full.close()
    datas, targets, pristine, host = [], [], [], []
    for i in range(lo, hi):
        d, t = pairs[i % 8]
        dp = synth.subsample_exact(d, 10000, 1000 + i)
        tp = synth.subsample_exact(t, 10000, 5000 + i)
        host.append((dp, tp))
        datas.append(ctx.cloud_from_points(dp)); targets.append(ctx.cloud_from_points(tp)); pristine.append(ctx.cloud_from_points(dp))
    chunk = args.batch_chunk if getattr(args, "batch_chunk", 0) else 128
    # ICPB_NN_AUTO: the library's own choice -- for a batch of clouds this size the exact cooperative cell-grid search
    # (same associations as the scan); ICPB_BATCH_NN=brute times the scan as the headline instead
    nn_mode = {"brute": icpb200.NN_BRUTE, "grid": icpb200.NN_GRID, "auto": icpb200.NN_AUTO}[os.environ.get("ICPB_BATCH_NN", "auto")]

    def step(mode=None):
        ctx.timer_start()
        res = []
        for b in range(0, len(datas), chunk):
            for dcl, pcl in zip(datas[b:b + chunk], pristine[b:b + chunk]):
                dcl.copy_from(pcl)
            res += ctx.icp_register_batch(datas[b:b + chunk], targets[b:b + chunk], ITERS, 0.0, 0.75, icpb200.SOLVE_REFERENCE,
                                          nn_mode=nn_mode if mode is None else mode)
        return ctx.timer_stop(), res

    for _ in range(args.warmup):
        step()
    _barrier(torch, world)
    l0 = ctx.launch_count()
    ms_all = []
    for _ in range(args.steps):
        ms, res = step()
        ms_all.append(ms)
    l1 = ctx.launch_count()
    _barrier(torch, world)
    tot = _max_over_ranks(torch, world, local, float(np.sum(ms_all)))
    # roofline of the dominant kernel: the same steps again with CUDA events around every search launch (one set of
    # events and one pair counter per batch call: every registration of a call reports the call's figures)
    used_grid = res[0]["nn_mode_used"] == icpb200.NN_GRID
    ctx.set_profiling(True)
    nn_ms, nn_launches, qpt, splits, filt, pairs = 0.0, 0, 0, 0, 0, 0
    for _ in range(max(1, min(args.steps, 2))):
        _, rp = step()
        for b in range(0, len(rp), chunk):
            nn_ms += rp[b]["nn_partial_ms"]; nn_launches += rp[b]["nn_partial_launches"]; pairs += rp[b]["grid_pairs"]
        qpt, splits, filt = rp[0]["nn_qpt"], rp[0]["nn_splits"], rp[0]["nn_filter_used"]
    ctx.set_profiling(False)
    # the brute-force scan on the same batch (the north star's kernel), for the record
    brute = None
    if used_grid:
        for _ in range(2):
            step(icpb200.NN_BRUTE)
        b_ms = [step(icpb200.NN_BRUTE)[0] for _ in range(2)]
        ctx.set_profiling(True)
        _, rb = step(icpb200.NN_BRUTE)
        ctx.set_profiling(False)
        b_nn = sum(rb[b]["nn_partial_ms"] for b in range(0, len(rb), chunk))
        b_launches = sum(rb[b]["nn_partial_launches"] for b in range(0, len(rb), chunk))
        b_tot = _max_over_ranks(torch, world, local, float(np.mean(b_ms)))
        b_ach = 8.0 * 1e4 * 1e4 * min(chunk, len(datas)) / max(b_nn / max(b_launches, 1) * 1e-3, 1e-12) / 1e12
        brute = {"registrations_per_s": total * 1000.0 / b_tot, "ms_per_step": b_tot,
                 "same_poses": bool(all(np.array_equal(x["pose_R"], y["pose_R"]) and np.array_equal(x["pose_t"], y["pose_t"])
                                        for x, y in zip(rb, res))),
                 "roofline": {"bound": "fp32", "kernel": f"nn_partial_warp_kernel<{rb[0]['nn_qpt']}>, {rb[0]['nn_splits']} splits",
                              "achieved": b_ach, "peak": _fp32_nominal(), "unit": "TFLOP/s", "frac": b_ach / _fp32_nominal()}}
    # e2e: host point lists in, poses out -- uploads and the result blocks inside the timed region
    e2e_ms = []
    for s in range(2):
        ctx.sync()
        t0 = time.perf_counter()
        for b in range(0, len(datas), chunk):
            for (dp, tp), dcl, tcl in zip(host[b:b + chunk], datas[b:b + chunk], targets[b:b + chunk]):
                dcl.upload(dp); tcl.upload(tp)
            ctx.icp_register_batch(datas[b:b + chunk], targets[b:b + chunk], ITERS, 0.0, 0.75, icpb200.SOLVE_REFERENCE,
                                   nn_mode=nn_mode)
        ctx.sync()
        e2e_ms.append((time.perf_counter() - t0) * 1e3)
    e2e_tot = _max_over_ranks(torch, world, local, float(e2e_ms[-1]))
    rows = np.array([np.concatenate([r["pose_R"].ravel(), r["pose_t"]]) for r in res], dtype=np.float64).reshape(-1, 12)
    if comm is not None:   # equal-sized blocks for the all-gather: pad to the largest share, strip after
        cap = -(-total // world)
        padded = np.zeros((cap, 12), np.float64)
        padded[: len(rows)] = rows
        gathered = comm.allgather_host(padded)
        allrows = np.concatenate([gathered[r][: D.shard_range(total, r, world)[1] - D.shard_range(total, r, world)[0]]
                                  for r in range(world)])
    else:
        allrows = rows
    line = None
    if rank == 0:
        ms_per_step = tot / args.steps
        digest = hashlib.sha256(np.ascontiguousarray(allrows).tobytes()).hexdigest()
        # oracle check + CPU baseline in one: registration 0 of the batch through the oracle port on one thread
        cpu, oracle_ok = None, None
        try:
            from oracle import oracle as orc
            orc.build()
            dp, tp = host[0]
            t0 = time.perf_counter()
            ref, rout, _, _ = orc.icp(dp, tp, ITERS, 0.0, 0.75, orc.SOLVE_REFERENCE, n_threads=1)
            cdt = time.perf_counter() - t0
            oracle_ok = bool(np.array_equal(ref["pose_R"], res[0]["pose_R"]) and np.array_equal(ref["pose_t"], res[0]["pose_t"])
                             and ref["n_assoc"] == res[0]["n_assoc"])
            cpu = {"value": 1.0 / cdt, "unit": "registrations/s", "cores": 1, "kind": "port",
                   "sample": f"registration 0 of the batch (10k x 10k, {ITERS + 1} passes) through the oracle port, one thread ({cdt:.2f} s)"}
        except ImportError as e:
            cpu = {"unavailable": str(e)}
        assert oracle_ok is not False, "batch registration 0 differs from the oracle"
        n_loc = len(datas)
        if used_grid:   # the search evaluates the pairs it stages, not n x m: 8 flop for each of THOSE (counted on the device)
            flop_per_launch = 8.0 * pairs / max(nn_launches, 1)
        else:
            flop_per_launch = 8.0 * 1e4 * 1e4 * min(chunk, n_loc)
        avg_s = nn_ms / max(nn_launches, 1) * 1e-3
        ach = flop_per_launch / max(avg_s, 1e-12) / 1e12
        line = {"metric": "icp_registrations_per_s", "value": total * 1000.0 / ms_per_step, "unit": "registrations/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"configs[3]: batch of {total} independent Kinect v1 10k-point registrations, "
                                       "20 iterations, sharded by index", "per_call_batch": chunk,
                           "pose_sha256": digest, "registration_0_equals_oracle": oracle_ok,
                           "l2": "working set (batch clouds + partials) exceeds L2"},
                "roofline": {"bound": "fp32",
                             "kernel": ("nn_grid_coop_kernel + nn_grid_heavy_kernel (exact cooperative cell-grid search; flops = 8 x the "
                                        "pairs it evaluates, counted on the device)") if used_grid else
                                       f"nn_partial ({'warp' if filt == icpb200.FILTER_WARP else 'centred'}-filter)<{qpt}>, {splits} splits",
                             "achieved": ach, "peak": _fp32_nominal(), "unit": "TFLOP/s", "frac": ach / _fp32_nominal(),
                             "traffic": None, "flop_per_launch": flop_per_launch, "avg_launch_ms": avg_s * 1e3,
                             "launches_timed": nn_launches,
                             "peak_source": "nominal 148 SM x 128 lanes x 2 x 1.965 GHz (rank 0's launches)"},
                "cpu_baseline": cpu,
                "e2e": {"value": total * 1000.0 / e2e_tot, "unit": "registrations/s",
                        "h2d_bytes_per_step": int(total * 2 * 10000 * 16), "d2h_bytes_per_step": int(total * 304)},
                "extra": {"nn_mode": "ICPB_NN_AUTO -> " + ("cell-grid search" if used_grid else "brute-force scan"),
                          "brute_force_scan": brute},
                "gpu_launches": int(l1 - l0)}
        if emit:
            print(json.dumps(line), flush=True)
    for c in datas + targets + pristine:
        c.close()
    if comm is not None:
        _barrier(torch, world)
        comm.close()
    ctx.close()
    return line


def run_map1cm(args, rank, world, local, emit=True):
    """configs[4]: full-resolution Kinect v1 frames into the 600x600x500 1 cm grid, z-slab sharded.  Strong scaling.
    Everything multi-GPU happens inside libicpb200 (icpb_comm / icpb_slabmap: NCCL all-gather of the lifted row bands,
    several frames per exchange, double buffered); torch.distributed only carries the 128-byte NCCL id and the hashes.
    The concatenated slabs are compared with the ORACLE's grid of the same sequence and with the unsharded GPU grid."""
    import icpb200
    from icpb200 import dist as D
    from icpb200 import synth
    torch = _setup(local, world)
    ctx = icpb200.Context(local)
    K = icpb200.reference_intrinsics_v1()
    dims, cell = (600, 600, 500), 0.01
    frames = args.frames or 32
    fpe = int(os.environ.get("ICPB_FRAMES_PER_EXCHANGE", "8"))
    poses = synth.trajectory(frames, step_deg=0.8, step_m=0.02)
    depths = [synth.render_depth(R, t, synth.KINECT_V1, seed=f) for f, (R, t) in enumerate(poses)]
    Rs = np.stack([np.asarray(R, np.float32) for R, _ in poses])
    ts = np.stack([np.asarray(t, np.float32) for _, t in poses])
    # slab boundaries balanced on the MEASURED work of the ray walk: three frames of the sequence go through a scratch
    # whole map twice (the second time the bricks their endpoints occupy are there) with the per-layer work histogram
    # switched on; equal shares of it are equal shares of the walk.  Every rank computes the same integers: no exchange.
    bounds = None
    if world > 1:
        probe = ctx.cloud(640 * 480)
        scratch = ctx.map(dims, cell)
        work = np.zeros(dims[2], np.uint64)
        for rep in range(2):
            for f in (0, frames // 2, frames - 1):
                R, t = poses[f]
                probe.from_depth(depths[f], None, K)
                probe.transform(np.asarray(R, np.float32), np.asarray(t, np.float32))
                if rep == 0:
                    scratch.integrate_rays(probe, tuple(float(x) for x in t), 25, 25, False)
                else:
                    scratch.integrate_rays_profiled(probe, tuple(float(x) for x in t), work, 25, 25)
        bounds = icpb200.slab_bounds_from_work(work, world)
        probe.close(); scratch.close()
    comm = icpb200.Comm.from_torch(ctx) if world > 1 else None

    # depth frames resident in HBM before the timed region (the metric's definition); int16 view of the u16 bits
    host_depths = torch.from_numpy(np.stack(depths).astype(np.uint16).view(np.int16)).pin_memory()
    d_depths = host_depths.to(torch.device("cuda", local))
    torch.cuda.synchronize()
    frame_bytes = 640 * 480 * 2

    # The histogram counts instructions, not the memory latency behind the steps that read voxels: the boundaries are
    # refined (untimed, at most three rounds) with the ray-walk time every rank MEASURES for its slab -- the histogram
    # inside each slab is rescaled to the slab's measured time and the equal shares are taken again.
    sm = icpb200.SlabMapC(ctx, comm, dims, cell, 640, 480, bounds)
    rebalance = []
    if world > 1:
        model = work.astype(np.float64) + 1e-3
        for it in range(3):
            for rep in range(2):
                ctx.set_profiling(rep == 1)
                sm.integrate_sequence_device(d_depths.data_ptr(), frames, K, Rs, ts, 25, 25, fpe)
                ctx.sync()
            t_rays, _ = ctx.profile_read(icpb200.PROF_MAP_RAYS)
            ctx.profile_read(icpb200.PROF_MAP_ENDPOINTS)
            ctx.set_profiling(False)
            times = comm.allgather_host(np.array([t_rays], np.float64)).ravel()
            rebalance.append({"bounds": [int(b) for b in bounds], "rays_ms_per_rank": [round(float(x), 3) for x in times]})
            if times.max() <= 1.08 * times.mean():
                break
            for g in range(world):
                seg = slice(bounds[g], bounds[g + 1])
                model[seg] *= times[g] / max(model[seg].sum(), 1e-9)
            new_bounds = icpb200.slab_bounds_from_work(np.maximum(model * 1e6, 1).astype(np.uint64), world)
            if new_bounds == list(bounds):
                break
            bounds = new_bounds
            sm.close()
            sm = icpb200.SlabMapC(ctx, comm, dims, cell, 640, 480, bounds)

    def step(h2d=False):
        ctx.timer_start()
        if h2d:   # e2e leg: every frame comes from pinned host memory (torch's stream; the host waits for the copy)
            d_depths.copy_(host_depths, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        sm.integrate_sequence_device(d_depths.data_ptr(), frames, K, Rs, ts, 25, 25, fpe)
        return ctx.timer_stop()

    for _ in range(args.warmup):
        step()
    sm.map.clear()
    step()                                  # the grid that is hashed: one pass over the sequence from an empty map
    slab = sm.download()
    _barrier(torch, world)
    ms_all = [step() for _ in range(args.steps)]
    _barrier(torch, world)
    tot = _max_over_ranks(torch, world, local, float(np.sum(ms_all)))
    # the ray-walk kernel's own time (roofline), same sequence with CUDA events around every launch
    ctx.set_profiling(True)
    step()
    rays_ms, rays_launches = ctx.profile_read(icpb200.PROF_MAP_RAYS)
    ends_ms, _ = ctx.profile_read(icpb200.PROF_MAP_ENDPOINTS)
    ctx.set_profiling(False)
    rays_ms_rank = rays_ms
    rays_ms = _max_over_ranks(torch, world, local, rays_ms)
    e2e = [step(h2d=True) for _ in range(2)]
    e2e_tot = _max_over_ranks(torch, world, local, float(e2e[-1]))
    h = hashlib.sha256(slab.tobytes()).hexdigest()
    if world > 1:
        import torch.distributed as dist
        hs = [None] * world
        dist.all_gather_object(hs, (sm.z_lo, sm.z_hi, h, int((slab > 0).sum())))
        per_rank_rays = [None] * world
        dist.all_gather_object(per_rank_rays, round(float(rays_ms_rank), 3))
    else:
        hs = [(sm.z_lo, sm.z_hi, h, int((slab > 0).sum()))]
        per_rank_rays = [round(float(rays_ms_rank), 3)]
    del slab
    line = None
    if rank == 0:
        # T7: the slabs must be the single-GPU grid, byte for byte -- integrated here unsharded through the
        # host-synchronising calls, which also count the rays and the voxels their walks pass
        one = ctx.map(dims, cell)
        cl = ctx.cloud(640 * 480)
        npts = visited = 0
        for (R, t), dpt in zip(poses, depths):
            cl.from_depth(dpt, None, K)
            cl.transform(np.asarray(R, np.float32), np.asarray(t, np.float32))
            npts += cl.n
            visited += one.integrate_rays(cl, tuple(float(x) for x in t), 25, 25, True)
        g1 = one.download()
        unsharded_ok = all(hashlib.sha256(np.ascontiguousarray(g1[:, :, lo:hi]).tobytes()).hexdigest() == hh
                           for (lo, hi, hh, _) in hs)
        one.close(); cl.close()
        del g1
        assert unsharded_ok, "z-slab result differs from the single-GPU grid"
        ms_per_step = tot / args.steps
        updates = visited + npts  # voxels visited by rays + endpoint updates, per pass over the sequence
        alg_bytes = 2.0 * visited + 12.0 * npts + 14.0 * npts
        rays_bytes = 2.0 * visited + 12.0 * npts        # 2 B per visited voxel + 12 B per ray
        peak = _hbm_peak() * world  # whole-job bytes over the slowest rank's time: against the N GPUs' aggregate bandwidth
        rays_gbs = rays_bytes / max(rays_ms * 1e-3, 1e-12) / 1e9
        # the whole sequence through the oracle: every slab must equal the oracle grid's slab
        cb, oracle_ok = None, None
        try:
            from oracle import oracle as orc
            orc.build()
            grid = np.zeros(dims, np.uint8)
            t0 = time.perf_counter()
            ov = 0
            for (R, t), dpt in zip(poses, depths):
                pts, _, _ = orc.backproject(dpt, None, orc.kinect_v1())
                pts = orc.translate(orc.rotate(pts, np.asarray(R, np.float32)), np.asarray(t, np.float32))
                ov += orc.map_integrate_rays(grid, dims, np.float32(cell), pts, tuple(float(x) for x in t), 25, 25)
            cdt = time.perf_counter() - t0
            oracle_ok = bool(ov == visited and all(
                hashlib.sha256(np.ascontiguousarray(grid[:, :, lo:hi]).tobytes()).hexdigest() == hh for (lo, hi, hh, _) in hs))
            cb = {"value": (ov + npts) / cdt, "unit": "voxel updates/s", "cores": 1, "kind": "port",
                  "sample": f"the whole {frames}-frame sequence (lift + ray integration) through the oracle port, one thread ({cdt:.1f} s)"}
            del grid
        except ImportError as e:
            cb = {"unavailable": str(e)}
        assert oracle_ok is not False, "z-slab result differs from the oracle's grid"
        line = {"metric": "voxel_updates_per_s", "value": updates / (ms_per_step * 1e-3), "unit": "voxel updates/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": f"configs[4]: {frames} full-res Kinect v1 frames into a 600x600x500 1 cm uint8 grid "
                                       "(180 MB), z-slab sharded (boundaries at equal shares of the measured per-layer walk work); every rank lifts its row "
                                       f"band, one NCCL all-gather per {fpe} frames inside libicpb200 (double buffered, three "
                                       "streams), bands walked in place; depth frames resident in HBM, no host "
                                       "synchronisation inside the sequence",
                           "frames_per_exchange": fpe, "rebalance_rounds": rebalance,
                           "rays_per_pass": npts, "voxels_visited_per_pass": visited, "slabs": hs,
                           "slabs_equal_single_gpu_grid": unsharded_ok, "slabs_equal_oracle_grid": oracle_ok,
                           "l2": "grid (180 MB) exceeds L2 at 1 GPU"},
                "extra": {"frames_per_s": frames / (ms_per_step * 1e-3),
                          "algorithmic_GBps": alg_bytes / (ms_per_step * 1e-3) / 1e9,
                          "stage_ms_per_pass": {"map_rays_max_over_ranks": rays_ms, "map_rays_per_rank": per_rank_rays,
                                                "map_endpoints_rank0": ends_ms}},
                "roofline": {"bound": "hbm", "kernel": "map_rays_brick_kernel", "achieved": rays_gbs, "peak": peak,
                             "unit": "GB/s", "frac": rays_gbs / peak,
                             # dram__bytes_read + write of one launch at N=1, profiles/r02_ncu_map_rays_brick.txt
                             "traffic": 22297088 if world == 1 else None,
                             "bytes_per_launch": rays_bytes / max(rays_launches, 1),
                             "avg_launch_ms": rays_ms / max(rays_launches, 1), "launches_timed": rays_launches,
                             "note": "algorithmic bytes = 2 B per voxel the walk passes + 12 B per ray (SURVEY.md 8d), whole "
                                     "job; time = the slowest rank's summed launches; peak = n_gpus x the measured copy "
                                     "bandwidth of one GPU.  The brick-skipping walk does not touch "
                                     "the voxels of empty bricks, so the figure can exceed what the memory system moves"},
                "cpu_baseline": cb,
                "e2e": {"value": updates / (e2e_tot * 1e-3), "unit": "voxel updates/s",
                        "h2d_bytes_per_step": int(frames * frame_bytes), "d2h_bytes_per_step": 0}}
        if emit:
            print(json.dumps(line), flush=True)
    if world > 1:
        _barrier(torch, world)
    sm.close()
    if comm is not None:
        comm.close()
    ctx.close()
    return line


def run_backproject(args, rank, world, local, emit=True):
    """HBM-bound stage in isolation: a batch of resident Kinect v1 frames -> XYZ clouds in one sync-free launch."""
    import icpb200
    from icpb200 import synth
    torch = _setup(local, world)
    stream = torch.cuda.current_stream().cuda_stream
    ctx = icpb200.Context(local, stream=stream)
    K = icpb200.reference_intrinsics_v1()
    nf = args.frames or 256
    base = [synth.render_depth(R, t, synth.KINECT_V1, seed=f) for f, (R, t) in enumerate(synth.trajectory(8, step_deg=1.0))]
    w, h = synth.KINECT_V1["w"], synth.KINECT_V1["h"]
    dev = torch.device("cuda", local)
    d_depth = torch.from_numpy(np.stack([base[f % 8] for f in range(nf)]).astype(np.int16)).to(dev)
    d_pts = torch.empty((nf, w * h, 4), dtype=torch.float32, device=dev)
    d_cnt = torch.zeros(nf, dtype=torch.int32, device=dev)

    def step():
        ctx.timer_start()
        ctx.backproject_batch_device(d_depth.data_ptr(), None, nf, w, h, K, d_pts.data_ptr(), w * h, d_cnt.data_ptr())
        return ctx.timer_stop()

    for _ in range(args.warmup):
        step()
    _barrier(torch, world)
    ms_all = [step() for _ in range(args.steps)]
    _barrier(torch, world)
    tot = _max_over_ranks(torch, world, local, float(np.sum(ms_all)))
    npts = int(d_cnt.sum().item())
    if rank == 0:
        ms_per_step = tot / args.steps
        alg = 2.0 * nf * w * h + 16.0 * npts           # depth read + points written
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
            os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        gbs = alg / (ms_per_step * 1e-3) / 1e9
        line = {"metric": "backprojected_frames_per_s", "value": world * nf / (ms_per_step * 1e-3), "unit": "frames/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{nf} resident Kinect v1 640x480 depth frames -> XYZ (pointcloud.cpp:109-165), one launch",
                           "points": npts, "l2": f"inputs+outputs ({alg / 1e6:.0f} MB) exceed L2"},
                "roofline": {"bound": "hbm", "kernel": "backproject_kernel", "achieved": gbs, "peak": peak, "unit": "GB/s",
                             "frac": gbs / peak, "traffic": None, "bytes_per_launch": alg}}
        cb = _cpu_port_rate(lambda orc: [orc.backproject(base[f], None, orc.kinect_v1()) for f in range(8)], 8,
                            "oracle back-projection of 8 of the frames, one thread")
        cb["unit"] = "frames/s"
        line["cpu_baseline"] = cb
        if emit:
            print(json.dumps(line), flush=True)
    ctx.close()
    return line if rank == 0 else None


def run_trajectory(args, rank, world, local, emit=True):
    """configs[2]: every rank runs the same trajectory (replicas; a single sequence does not shard)."""
    import icpb200
    from icpb200 import synth
    torch = _setup(local, world)
    ctx = icpb200.Context(local)
    K = icpb200.reference_intrinsics_v2()
    frames = args.frames or 300
    dims, cell = (300, 300, 250), 0.02
    poses = synth.trajectory(frames)
    t_r = time.perf_counter()
    depths = [synth.render_depth(R, t, synth.KINECT_V2, seed=f) for f, (R, t) in enumerate(poses)]
    render_s = time.perf_counter() - t_r
    W, Hh = synth.KINECT_V2["w"], synth.KINECT_V2["h"]
    sub, prev_sub = ctx.cloud(W * Hh), ctx.cloud(W * Hh)
    # map stage through the sync-free frame path on a SECOND context (its own stream): depth frames resident in HBM; the
    # full-resolution lift, the ray walk and the endpoint update of frame f need only the pose, so they overlap frame
    # f+1's registration on the device as well as on the host
    ctx_map = icpb200.Context(local)
    m = ctx_map.map(dims, cell)
    dev = torch.device("cuda", local)
    d_depths = torch.from_numpy(np.stack(depths).astype(np.uint16).view(np.int16)).to(dev)
    band = torch.zeros((W * Hh + 1, 4), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    nvalid = [int((d != 0).sum()) for d in depths]

    def run():
        m.clear()
        ctx_map.sync()
        R, t = poses[0][0].astype(np.float64), poses[0][1].astype(np.float64)
        stage = {"backproject": 0.0, "icp": 0.0, "map": 0.0}
        ctx.timer_start()
        for f in range(frames):
            t0 = time.perf_counter()
            stride = max(1, -(-nvalid[f] // 10000))
            sub.from_depth(depths[f], None, K, icpb200.SUB_STRIDE, stride)
            t1 = time.perf_counter()
            if f > 0:
                sub.transform(R.astype(np.float32), t.astype(np.float32))       # initial guess: previous pose
                res, _, _ = ctx.icp_register(sub, prev_sub, ITERS, 0.0, 0.75, icpb200.SOLVE_KABSCH)
                R, t = res["pose_R"] @ R, res["pose_R"] @ t + res["pose_t"]
            else:
                sub.transform(R.astype(np.float32), t.astype(np.float32))
            prev_sub.copy_from(sub)
            t2 = time.perf_counter()
            ctx_map.frame_lift_band_device(d_depths.data_ptr() + f * W * Hh * 2, W, Hh, 0, Hh, K, R.astype(np.float32),
                                       t.astype(np.float32), band.data_ptr(), W * Hh)
            m.integrate_bands_device(band.data_ptr(), 1, W * Hh, tuple(float(x) for x in t), 25, 25)
            t3 = time.perf_counter()
            stage["backproject"] += t1 - t0; stage["icp"] += t2 - t1; stage["map"] += t3 - t2
        ctx_map.sync()  # the map stream's tail is inside the timed region
        return ctx.timer_stop(), stage, (R, t)

    run()  # warm-up pass over the whole sequence
    _barrier(torch, world)
    ms, stage, (R, t) = run()
    _barrier(torch, world)
    grid = m.download()
    tot = _max_over_ranks(torch, world, local, ms)
    if rank == 0:
        err_t = float(np.linalg.norm(t - poses[-1][1]))
        line = {"metric": "slam_frames_per_s", "value": world * frames / (tot * 1e-3), "unit": "frames/s", "n_gpus": world,
                "steps": 1, "warmup": 1, "ms_per_step": tot, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"configs[2]: {frames}-frame Kinect v2 512x424 trajectory, ICP (<=10k subsample, "
                                       "Kabsch, 20 iterations) + full-res ray integration into 300x300x250 at 2 cm",
                           "grid_sha256": hashlib.sha256(grid.tobytes()).hexdigest(), "occupied_voxels": int((grid > 0).sum()),
                           "final_position_error_m": err_t, "per_rank": "replica of the same sequence"},
                "extra": {"stage_host_wall_s": stage, "render_s_untimed": render_s,
                          "note": "the map stage is enqueued without a host synchronisation (sync-free frame path): its "
                                  "host wall time is launch cost only, the device work overlaps the next frame"}}
        if emit:
            print(json.dumps(line), flush=True)
    ctx_map.close()
    ctx.close()
    return line if rank == 0 else None


def run_live(args, rank, world, local, emit=True):
    """the per-frame loop exactly as the reference runs it (icp.cpp:28-285, key-point association
    against the growing map cloud, rule-C map update), frames/s through the C-ABI; beside it the reference's OWN
    icp::getTransformation (oracle/_ref, single thread like the reference) on the same frames and key-points.
    cv::FAST is out of scope: key-points are seeded pixels with depth in every frame."""
    import icpb200
    from icpb200 import synth
    torch = _setup(local, world)
    ctx = icpb200.Context(local)
    K = icpb200.reference_intrinsics_v1()
    frames = args.frames or 40
    n_kp = 1500
    poses = synth.trajectory(frames, step_deg=0.5, step_m=0.01)
    depths = [synth.render_depth(R, t, synth.KINECT_V1, seed=f) for f, (R, t) in enumerate(poses)]
    h, w = depths[0].shape
    bgr = np.random.default_rng(3).integers(0, 255, (h, w, 3), dtype=np.uint8)
    ok = np.ones((h, w), bool)
    for d in depths:
        ok &= d != 0
    ys, xs = np.nonzero(ok[8:-8, 8:-8])
    sel = np.random.default_rng(7).choice(len(xs), n_kp, replace=False)
    kxy = np.stack([xs[sel] + 8, ys[sel] + 8], 1).astype(np.float32)
    # the 1-in-40 subsample decisions (rand() % 40 in the reference, pointcloud.cpp:28), drawn before the timed region
    decs = [(np.random.default_rng(100 + f).integers(0, 40, int((d != 0).sum())) == 0).astype(np.uint8)
            for f, d in enumerate(depths)]
    dims, cell = (300, 300, 300), float(np.float32(10.0) / np.float32(300.0))     # map.hpp:9-10,17
    FX, CX = np.float32(468.60), np.float32(318.27)

    def lift(depth):   # pointcloud.cpp:64-97, host scalar code in the reference as well
        x, y = kxy[:, 0].astype(np.int64), kxy[:, 1].astype(np.int64)
        pz = depth[y, x].astype(np.float32) / np.float32(5000.0)
        out = np.zeros(len(x), icpb200.POINT_DTYPE)
        out["x"] = (x.astype(np.float32) - CX) * pz / FX; out["y"] = (y.astype(np.float32) - CX) * pz / FX; out["z"] = pz
        out["c0"], out["c1"], out["c2"] = bgr[y, x, 0], bgr[y, x, 1], bgr[y, x, 2]
        return out

    def run():
        m = ctx.map(dims, cell)
        map_kp = ctx.cloud(1 << 18)
        pts, kps, non = ctx.cloud(w * h), ctx.cloud(n_kp), ctx.cloud(17 * n_kp)
        camR, camP, last_t = np.eye(3, dtype=np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32)
        started = False
        ctx.sync()
        t0 = time.perf_counter()
        for f in range(1, frames):
            cur, prev = depths[f], depths[f - 1]
            dec = decs[f]
            if not started:                                                             # icp.cpp:47-68
                camR, camP = np.eye(3, dtype=np.float32), np.array([5, 5, 5], np.float32)
                pk = ctx.cloud_from_points(lift(prev))
                pk.transform(camR, camP)
                m.update_tracked(pk, icpb200.TRACK_INIT, 180, 180, map_kp)
                pk.close()
                started = True
            pts.from_depth(cur, bgr, K, icpb200.SUB_STREAM, 40, 0, dec)
            kps.upload(lift(cur))
            pts.transform(camR, camP); kps.transform(camR, camP)                        # :70-71
            res = ctx.icp_register_keypoints(kps, pts, map_kp, 16, 1e-4, 0.1, last_translation=tuple(last_t),
                                             non_associations=non)
            camR = (camR @ res["cam_rotation"]).astype(np.float32)
            camP = (camP + res["cam_position"]).astype(np.float32)
            last_t = (-res["offset"]).astype(np.float32)
            if res["n_assoc"] > 0 and non.n > 0:
                m.update_tracked(non, icpb200.TRACK_NONASSOC, 25, 180, map_kp)          # :271
        ctx.sync()
        dt = time.perf_counter() - t0
        out = (dt, camR, camP, map_kp.n)
        for c in (m, map_kp, pts, kps, non):
            c.close()
        return out

    run()
    _barrier(torch, world)
    dt, camR, camP, n_map = run()
    _barrier(torch, world)
    tot = _max_over_ranks(torch, world, local, dt * 1e3)
    if rank == 0:
        cpu = None
        try:
            from oracle import ref
            if ref.available():
                ref.map_reset()
                k = min(frames, 9)
                t0 = time.perf_counter()
                for f in range(1, k):
                    _, rR, rP = ref.get_transformation(depths[f], depths[f - 1], bgr, kxy, 16, 1e-4, 1 if f == 1 else 12345 + f)
                cdt = time.perf_counter() - t0
                cpu = {"value": (k - 1) / cdt, "unit": "frames/s", "cores": 1, "kind": "reference",
                       "sample": f"the reference's own icp::getTransformation (oracle/_ref) on the first {k - 1} frame pairs "
                                 f"of the same sequence, {n_kp} key-points ({cdt:.2f} s)"}
        except Exception as e:  # the library is test infrastructure; the GPU number stands without it
            cpu = {"unavailable": str(e)}
        # the same frames through the C++ drop-in layer (icp::getTransformation of include/icpb200/icp.hpp, the call of
        # SLAM.cpp:277), both association modes: tests/cpp/bench_compat.cpp
        compat = None
        try:
            import struct
            import subprocess
            import tempfile
            binp = os.path.join(ROOT, "icp-slam-prototype_b200", "lib", "bench_compat")
            kf = min(frames, 12)
            with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
                f.write(struct.pack("iiii", w, h, kf, len(kxy)))
                for d in depths[:kf]:
                    f.write(np.ascontiguousarray(d, np.uint16).tobytes())
                f.write(bgr.tobytes()); f.write(np.ascontiguousarray(kxy, np.float32).tobytes())
                tmp = f.name
            out = subprocess.run([binp, tmp], capture_output=True, text=True, timeout=300)
            # A/B: the same frames with one rand() call per pixel for the subsample draws (pointcloud.cpp:22-28) instead
            # of the bulk draws on the generator's state (host/compat.cpp, draw_keep): same stream, same results
            out_ab = subprocess.run([binp, tmp], capture_output=True, text=True, timeout=300,
                                    env=dict(os.environ, ICPB_COMPAT_RAND_CALLS="1"))
            os.unlink(tmp)
            kv = dict(tok.split("=") for tok in out.stdout.split() if "=" in tok)
            kv_ab = dict(tok.split("=") for tok in out_ab.stdout.split() if "=" in tok)
            compat = {"all_points_ms_per_frame": float(kv["all_points_ms"]), "keypoints_ms_per_frame": float(kv["keypoints_ms"]),
                      "frames": int(kv["frames"]),
                      "with_one_rand_call_per_pixel": {"all_points_ms_per_frame": float(kv_ab["all_points_ms"]),
                                                       "keypoints_ms_per_frame": float(kv_ab["keypoints_ms"])},
                      "what": "icp::getTransformation through the C++ drop-in headers (depth, colour and key-point "
                              "images in, 4x4 matrix out, as the reference's signature demands; the clouds stay on the "
                              "device, the subsample draws are made in bulk on the C library's generator state), per frame"}
        except Exception as e:  # the harness is optional: the C-ABI number stands without it
            compat = {"unavailable": str(e)}
        line = {"metric": "live_loop_frames_per_s", "value": world * (frames - 1) / (tot * 1e-3), "unit": "frames/s",
                "n_gpus": world, "steps": 1, "warmup": 1, "ms_per_step": tot, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"8f-2: {frames - 1} Kinect v1 frame pairs through the reference's live loop "
                                       f"(1-in-40 subsampled cloud, {n_kp} key-points against the growing map cloud, "
                                       "16 iterations max, threshold 1e-4, rule-C map update on the 300^3 grid); includes "
                                       "host->device copies of every frame",
                           "map_keypoints_at_end": int(n_map), "per_rank": "replica of the same sequence"},
                "cpu_baseline": cpu, "extra": {"cpp_drop_in": compat}}
        if emit:
            print(json.dumps(line), flush=True)
    ctx.close()
    return line if rank == 0 else None


def run_normals(args, rank, world, local, emit=True):
    """P3 in isolation (SLAM.cpp:412-430): a batch of resident Kinect v1 frames -> per-pixel normals, one launch."""
    import icpb200
    from icpb200 import synth
    torch = _setup(local, world)
    stream = torch.cuda.current_stream().cuda_stream
    ctx = icpb200.Context(local, stream=stream)
    nf = args.frames or 256
    base = [synth.render_depth(R, t, synth.KINECT_V1, seed=f) for f, (R, t) in enumerate(synth.trajectory(8, step_deg=1.0))]
    w, h = synth.KINECT_V1["w"], synth.KINECT_V1["h"]
    dev = torch.device("cuda", local)
    d_depth = torch.from_numpy(np.stack([base[f % 8] for f in range(nf)]).astype(np.uint16).view(np.int16)).to(dev)
    d_out = torch.empty((nf, h, w, 3), dtype=torch.float32, device=dev)

    def step():
        ctx.timer_start()
        ctx.normals_batch_device(d_depth.data_ptr(), nf, w, h, d_out.data_ptr())
        return ctx.timer_stop()

    for _ in range(args.warmup):
        step()
    _barrier(torch, world)
    ms_all = [step() for _ in range(args.steps)]
    _barrier(torch, world)
    tot = _max_over_ranks(torch, world, local, float(np.sum(ms_all)))
    if rank == 0:
        ms = tot / args.steps
        bytes_per_launch = float(nf) * w * h * 14.0   # 2 B depth read + 12 B normal written per pixel
        peak = _hbm_peak()
        line = {"metric": "normal_maps_per_s", "value": world * nf / (ms * 1e-3), "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{nf} resident Kinect v1 640x480 depth frames -> normals (SLAM.cpp:412-430), one launch",
                           "l2": f"inputs+outputs ({bytes_per_launch / 1e6:.0f} MB) exceed L2"},
                "roofline": {"bound": "hbm", "kernel": "normals4_kernel", "achieved": bytes_per_launch / (ms * 1e-3) / 1e9,
                             "peak": peak, "unit": "GB/s", "frac": bytes_per_launch / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                             "bytes_per_launch": bytes_per_launch,
                             "note": "issue-bound before HBM binds (71 % issue utilisation, profiles/r01_ncu_normals_batch.txt): "
                                     "the reference normalises in double (cv::normalize), one IEEE FP64 sqrt and one FP64 "
                                     "reciprocal per pixel"}}
        cb = _cpu_port_rate(lambda orc: [orc.normals(base[f]) for f in range(8)], 8,
                            "oracle normals of 8 of the frames, one thread")
        cb["unit"] = "frames/s"
        line["cpu_baseline"] = cb
        if emit:
            print(json.dumps(line), flush=True)
    ctx.close()
    return line if rank == 0 else None

#!/usr/bin/env python
"""Benchmark of the rasterizer hot path: fwd+bwd views/s at 3 M Gaussians, 1959x1090, all channels (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config tnt-3m] [--views-per-rank V]

A *step* is one view-sharded batch: every rank renders (forward + backward, gradients accumulated in place into the nine
raw parameter-gradient groups) its V views of the N*V-view batch and, for N > 1, those parameter gradients are summed with
NCCL all-reduce, Gaussian range by Gaussian range under the remaining backward work.  Per-GPU work is fixed as N grows (weak
scaling).  Prints ONE JSON line (see DESIGN.md "Measurement"); at N > 1 it carries `dp_check`: the all-reduced gradients of
one step against the same rank running the whole batch sequentially.

`--impl reference` times the unmodified reference CUDA rasterizer (compiled into oracle/_ref by oracle/build_ref.py)
through its own Python binding on the same workload (rank 0 only; the reference has no multi-GPU path); when that
build is absent it falls back to the CPU oracle port on a bounded sample.
"""
import argparse
import gc as pygc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for sub in ("gs-2m_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, sub))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import synthetic_scenes as syn  # noqa: E402

METRIC = "fwd+bwd views/sec @3M Gaussians 1959x1090"
UNIT = "views/s"


# ----------------------------------------------------------------------------------------------------------------
def algorithmic_bytes(P, V, R, N, F, M, tiles):
    """Compulsory HBM traffic per view, SURVEY.md section 8(d) (every input read once, every output written once,
    every intermediate written once and read once per consuming stage)."""
    b = {}
    b["preprocess_fwd"] = 20 * P + V * (12 * M + 32 + 67)
    b["scan"] = 8 * P
    b["duplicate"] = 20 * V + 12 * R
    b["sort"] = 24 * R
    b["ranges"] = 8 * R + 8 * tiles
    b["blend_fwd"] = R * (40 + 4 * F) + N * (4 * (3 + F) + 8) + 4 * P
    b["blend_bwd"] = R * (40 + 4 * F) + N * (4 * (3 + F) + 8) + 4 * V * (12 + F)
    b["preprocess_bwd"] = V * (12 * M + 115) + 4 * P * (34 + 3 * M)
    b["total"] = sum(b.values())
    return b


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed regions run.

    The poller is started BEFORE the warm-up steps and only terminated when the whole run is over: starting or ending an
    nvidia-smi process next to a running CUDA job stalls that job for tens of milliseconds (seen as a single 40-60 ms step, or a
    whole leg at a third of its speed, in about one run out of eight when the poller was started right in front of the first
    timed step and terminated right in front of another leg).  `summary()` uses the samples that fall inside the timed
    windows (`window()` context manager)."""

    def __init__(self, index):
        self.index, self.samples, self.proc, self.windows = index, [], None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    class _Window:
        def __init__(self, owner):
            self.owner = owner

        def __enter__(self):
            self.t0 = time.perf_counter()

        def __exit__(self, *exc):
            self.owner.windows.append((self.t0, time.perf_counter()))

    def window(self):
        return ClockSampler._Window(self)

    def summary(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        inside = [s for t, s in list(self.samples) if any(a - 0.2 <= t <= b + 0.2 for a, b in self.windows)]
        use = inside or [s for _t, s in list(self.samples)]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in use:
            try:
                sm.append(float(s[0])); mx.append(float(s[1]))
            except Exception:
                continue
            for n, v in zip(names, s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            self.proc = None


def build_workload(cfg, n_views_total, my_views, device):
    """Replicated scene + this rank's cameras/features, all resident on `device`."""
    scene = syn.scene_to(syn.make_scene(cfg["P"], shell_fraction=cfg["shell"], cluster=cfg.get("cluster"), opacity_cap=cfg.get("opacity_cap")), device)
    cams_cpu = syn.make_cameras(n_views_total, cfg["W"], cfg["H"], radius=cfg["cam_radius"])
    cams = {v: syn.camera_to(cams_cpu[v], device) for v in my_views}
    feats = {v: syn.pack_features(scene, cams[v], cfg["F"]) for v in my_views}
    gc, gb = syn.make_upstream_grads(cfg["W"], cfg["H"], cfg["F"])
    return scene, cams_cpu, cams, feats, gc.to(device), gb.to(device)


# ----------------------------------------------------------------------------------------------------------------
def run_ours(args, cfg, rank, world, device):
    import diff_gaussian_rasterization as dgr
    from diff_gaussian_rasterization import _native
    from diff_gaussian_rasterization.packing import activate_and_pack
    import view_parallel as vp
    lib = _native.load()
    P, W, H, F, M = cfg["P"], cfg["W"], cfg["H"], cfg["F"], 16
    V_per = args.views_per_rank
    n_views = world * V_per
    my_views = list(vp.shard_views(n_views, world, rank))       # (re-assigned by cost below when N > 1)
    scene = syn.scene_to(syn.make_scene(P, shell_fraction=cfg["shell"], cluster=cfg.get("cluster"), opacity_cap=cfg.get("opacity_cap")), device)
    cams_cpu = syn.make_cameras(n_views, W, H, radius=cfg["cam_radius"])
    cams = {v: syn.camera_to(cams_cpu[v], device) for v in range(n_views)}       # all views: dp_check replays the whole batch
    settings = {v: syn.raster_settings_for(cams[v], F, dgr.GaussianRasterizationSettings) for v in range(n_views)}
    gc, gb = (t.to(device) for t in syn.make_upstream_grads(W, H, F))
    raw = syn.raw_parameters(scene)
    order = ("xyz", "scaling", "rotation", "opacity", "albedo", "roughness", "metallic")
    blend_metallic = F in (2, 6, 10)
    stats = {}
    holder = {}

    # ---- which rank renders which view (N > 1): a step waits for its slowest rank at every collective, and the views differ in
    # cost (instance count: 6.9 .. 7.9 M on the ring of cameras), so the batch is dealt out by cost — each rank measures the
    # instance counts of its contiguous share once (a training loop knows them from the camera's previous visit), the counts
    # are all-gathered and every rank derives the same longest-first table.  Same batch, same views per rank, same sum.
    assignment = None
    if world > 1 and not args.contiguous_views:
        mine_R = []
        for v in my_views:
            st = settings[v]
            with torch.no_grad():
                s_, q_, o_, f_ = activate_and_pack(*[raw[k] for k in order], st.viewmatrix, st.campos, blend_metallic=blend_metallic)
            mine_R.append(float(dgr.forward_raw(raw["xyz"], scene.shs, None, o_, s_, q_, None, f_, st, for_backward=False)[4].num_rendered))
        gathered = [None] * world
        dist.all_gather_object(gathered, (my_views, mine_R))
        costs = [0.0] * n_views
        for vs_, rs_ in gathered:
            for v, r_ in zip(vs_, rs_):
                costs[v] = r_
        assignment = vp.balance_views(costs, world)
        my_views = assignment[rank]
        stats["view_costs"] = costs

    # ---- the training-faithful chain: the features and the activated scale / rotation / opacity depend on the
    # camera, so every view runs  fused activation+packing forward -> rasterizer forward -> reverse blend  as soon as it is
    # scheduled, and the per-Gaussian backward + fused packing backward (+= into the nine RAW parameter-gradient groups, 64
    # floats per Gaussian) afterwards, Gaussian range by Gaussian range, each finished range all-reduced under the next one.
    # Forward calls of the timed legs use ONE instance capacity for the binning arena, fixed after the warm-up (1.25 x the largest
    # count seen + 64 Ki, the wrapper's own rule): the wrapper's automatic hint moves with every view, and an arena whose size is
    # new to torch's caching allocator costs a cudaMalloc of ~1 GB in the middle of a step.  By default the forward still waits for
    # its count (and raises if the capacity did not suffice).  With --no-wait-forward it does not (`no_wait`): the host queues a
    # step ahead (bounded by max_steps_ahead), and whether the capacity sufficed is a bit in each view's device-side control
    # block; the bits are OR-ed into one word on the device and read once per timed region.
    holder["cap"], holder["maxR"] = None, 0
    holder["ovf"] = torch.zeros(1, dtype=torch.int32, device=device)

    def forward(*a):
        cap = holder["cap"]
        if cap and args.no_wait_forward:
            out = dgr.forward_raw(*a, capacity=cap, no_wait=True)
            holder["ovf"].bitwise_or_(dgr.control_block(a[-1], out[4])[2:3])
        elif cap:
            out = dgr.forward_raw(*a, capacity=cap)
        else:
            out = dgr.forward_raw(*a)
            holder["maxR"] = max(holder["maxR"], int(out[4].num_rendered))
        return out

    def check_capacity(what):
        flags = int(holder["ovf"])
        holder["ovf"].zero_()
        if flags:
            raise RuntimeError("%s: a no_wait forward reported control-block flags %d (1 = instance capacity exceeded)" % (what, flags))

    def begin_view(v, grad_color=None, st=None):
        st = st or settings[v]
        with torch.no_grad():
            s_, q_, o_, f_ = activate_and_pack(*[raw[k] for k in order], st.viewmatrix, st.campos, blend_metallic=blend_metallic)
        color, radii, observe, buffer, state = forward(raw["xyz"], scene.shs, None, o_, s_, q_, None, f_, st)
        g_c = grad_color(color) if grad_color is not None else gc
        dgr.backward_raw(g_c, gb, raw["xyz"], scene.shs, None, s_, q_, None, f_, radii, st, state,
                         grads=holder["step"].buckets.raster, phase="blend")
        stats["R"], stats["radii"] = state.num_rendered, radii
        return {"v": v, "st": st, "s": s_, "q": q_, "f": f_, "radii": radii, "observe": observe, "state": state}

    def finish_view(h, buckets, accumulate, rows):
        st = h["st"]
        # (the packing chain of the view's camera runs inside the per-Gaussian kernel: gs2m_backward_args::chain)
        dgr.backward_raw(gc, gb, raw["xyz"], scene.shs, None, h["s"], h["q"], None, h["f"], h["radii"], st, h["state"],
                         grads=buckets.raster, accumulate=2 if accumulate else 0, phase="gaussians", rows=rows,
                         densify_stats=holder["step"].stats.backward_args(),
                         chain=buckets.chain_spec(raw, blend_metallic=blend_metallic))

    def finish_views(handles, buckets, rows, accumulate=False):
        # all of the rank's views for one Gaussian range in ONE pass (gs2m_rasterize_backward_views): each thread owns a
        # Gaussian, sums the views that see it on chip and writes the 64 raw-gradient floats once
        chain = buckets.chain_spec(raw, blend_metallic=blend_metallic)
        dgr.backward_views_raw([dict(grad_color=gc, grad_buffer=gb, means3D=raw["xyz"], shs=scene.shs, scales=h["s"], rotations=h["q"],
                                     features=h["f"], radii=h["radii"], raster_settings=h["st"], state=h["state"],
                                     grads=buckets.raster, densify_stats=holder["step"].stats.backward_args(), chain=chain)
                                for h in handles], rows=rows, accumulate=accumulate)

    def make_step(begin, world_=world, rank_=rank, n_streams=args.streams, buckets=None, in_flight=None):
        st_ = vp.ViewShardedStep(P, M, device, world=world_, rank=rank_, n_streams=n_streams, buckets_cls=vp.ParameterBuckets,
                                 begin_view=begin, finish_view=finish_view, n_chunks=args.chunks, buckets=buckets,
                                 finish_views=None if args.per_view_finish else finish_views,
                                 assignment=assignment if world_ == world else None, max_views_in_flight=in_flight,
                                 max_steps_ahead=1 if args.no_wait_forward else None)
        st_.buckets.fused_chain = True
        return st_

    step = holder["step"] = make_step(begin_view)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def timed(step_, steps):
        holder["step"] = step_
        pygc.collect()      # a full collection of the interpreter's heap takes 0.1-0.3 s (torch alone is ~1 M objects); if the
        pygc.freeze()       # cyclic GC chose to run one inside a timed region the launch queue would drain and the step times
        barrier()           # would measure the host.  Collect now, park the survivors, keep the GC enabled.
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        t0 = time.perf_counter()
        e0.record()
        for k in range(steps):
            step_.run(n_views)
            marks[k].record()
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms = torch.tensor([max(e0.elapsed_time(e1), 0.0)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        stats["step_ms"] = [round(a.elapsed_time(b), 3) for a, b in zip([e0] + marks[:-1], marks)]    # this rank's steps, one by one
        return float(ms[0]), wall

    clocks = holder["clocks"] = ClockSampler(device.index)
    if rank == 0 and not args.no_clock_sampler:
        clocks.start()          # (before the warm-up, terminated at the very end of the run: see ClockSampler)
    for _ in range(args.warmup):
        step.run(n_views)
    barrier()
    V_vis = int((stats["radii"] > 0).sum())
    R = int(stats["R"])
    holder["cap"] = int(holder["maxR"] * 1.25) + 65536          # one arena size for all timed views from here on
    for _ in range(3):                                          # (warm-up at that size: the allocator's pool reaches its
        step.run(n_views)                                       # final state before anything is timed)
    barrier()
    check_capacity("warm-up")

    # ---- timed region 1 (headline): device-resident inputs ----
    launches0 = lib.gs2m_launch_count()
    with clocks.window():
        total_ms, _ = timed(step, args.steps)
    check_capacity("headline")
    headline_step_ms = list(stats["step_ms"])
    launches = torch.tensor([lib.gs2m_launch_count() - launches0], device=device, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    value = n_views * args.steps / (total_ms * 1e-3)

    # ---- per-kernel device times: the same step with the library's cudaEvent brackets switched on (they sit on the
    # launching stream around every stage; kept out of the region above so that event bookkeeping cannot perturb it;
    # single stream: with several views in flight the brackets of one stream would also span the other stream's kernels)
    step_prof = holder["step"] = make_step(begin_view, world_=1, rank_=0, n_streams=1, buckets=step.buckets)
    prof_views = my_views[:min(len(my_views), 4)]
    lib.gs2m_profile_enable(1)
    _native.profile_read()
    step_prof._run_deferred(prof_views, reduce=False)
    torch.cuda.synchronize(device)
    stage = _native.profile_read()
    lib.gs2m_profile_enable(0)
    barrier()

    # ---- timed region 2: end to end through the public API with HOST buffers ----
    # What lives on the host in GS-2M's training loop is the camera and its ground-truth image; every view therefore
    # copies (pinned host -> device) its camera matrices and a GT image, renders, turns the GT image into dL/dcolor
    # on the device (L2 photometric loss, the caller's job), runs the backward, and reads the loss back to the host.
    # dL/dbuffer (regularisers on the feature planes) is produced on the device in the real loop and stays resident.
    # The next view's host->device copies are issued on a side stream so they overlap the current view's kernels.
    pin = lambda t: t.cpu().pin_memory()  # noqa: E731
    h_cam = {v: (pin(cams_cpu[v].world_view_transform), pin(cams_cpu[v].full_proj_transform),
                 pin(cams_cpu[v].camera_center)) for v in my_views}
    gen = torch.Generator().manual_seed(7)
    h_gt = torch.rand((3, H, W), generator=gen).pin_memory()
    h_loss = torch.zeros(len(my_views), dtype=torch.float32).pin_memory()
    bg = torch.zeros(3, device=device)
    copy_stream = torch.cuda.Stream(device=device)
    # camera matrices (tiny) are read again by the view's deferred per-Gaussian stage: one slot per view, double-buffered over
    # steps so that the next step's copies never wait for this step's tail; the ground-truth images (25.6 MB each) are only
    # needed until the view's loss gradient is formed: a ring of one slot per local view
    n_ahead = min(4, len(my_views))       # a view's copies are queued this many views ahead of its forward
    cam_slots = [[dict(wvt=torch.empty((4, 4), device=device), full=torch.empty((4, 4), device=device),
                       cpos=torch.empty(3, device=device)) for _ in my_views] for _ in range(2)]
    gt_slots = [dict(gt=torch.empty((3, H, W), device=device), ready=torch.cuda.Event(), free=torch.cuda.Event())
                for _ in range(len(my_views))]
    h2d = sum(t.numel() * 4 for t in h_cam[my_views[0]]) + h_gt.numel() * 4
    d2h = 4
    seq = {"k": 0}          # views begun so far, over all steps: view k of the run is view k % n of step k // n

    def prefetch(k, pos, parity):
        sl, cs, v = gt_slots[k % len(gt_slots)], cam_slots[parity][pos], my_views[pos]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(sl["free"])
            sl["gt"].copy_(h_gt, non_blocking=True)
            cs["wvt"].copy_(h_cam[v][0], non_blocking=True)
            cs["full"].copy_(h_cam[v][1], non_blocking=True)
            cs["cpos"].copy_(h_cam[v][2], non_blocking=True)
            sl["ready"].record(copy_stream)

    def begin_view_e2e(v):
        pos, k, n = my_views.index(v), seq["k"], len(my_views)
        parity = (k // n) & 1
        if k == 0:
            for j in range(n_ahead):
                prefetch(j, j % n, (j // n) & 1)
        kk = k + n_ahead                              # global index of the view prefetched now (may belong to the next step)
        prefetch(kk, kk % n, (kk // n) & 1)
        sl, cs = gt_slots[k % len(gt_slots)], cam_slots[parity][pos]
        torch.cuda.current_stream(device).wait_event(sl["ready"])
        st = dgr.GaussianRasterizationSettings(H, W, cams_cpu[v].tanfovx, cams_cpu[v].tanfovy, bg, 1.0, cs["wvt"],
                                               cs["full"], 3, cs["cpos"], False, F)

        def grad_color(color):
            diff = color - sl["gt"]
            h_loss[pos].copy_((diff * diff).mean(), non_blocking=True)
            g_c = diff * (2.0 / diff.numel())
            sl["free"].record(torch.cuda.current_stream(device))
            return g_c
        seq["k"] = k + 1
        return begin_view(v, grad_color=grad_color, st=st)

    step_e2e = make_step(begin_view_e2e, buckets=step.buckets)

    def run_e2e(steps):
        holder["step"] = step_e2e
        pygc.collect(); pygc.freeze()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        t0 = time.perf_counter()
        e0.record()
        for k in range(steps):
            step_e2e.run(n_views)
            marks[k].record()
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms = torch.tensor([max(e0.elapsed_time(e1), 0.0)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        stats["e2e_step_ms"] = [round(a.elapsed_time(b), 3) for a, b in zip([e0] + marks[:-1], marks)]
        return float(ms[0]), wall

    run_e2e(2)
    with clocks.window():
        e2e_ms, wall_ms = run_e2e(args.steps)
    e2e_value = n_views * args.steps / (e2e_ms * 1e-3)
    check_capacity("e2e")
    clock_info = clocks.summary() if rank == 0 else None

    # ---- data-parallel correctness, outside the timed regions (N > 1): after one step every rank must hold the gradient of
    # the WHOLE batch, i.e. what a single rank gets by running all world*V views one after the other, and all ranks must
    # hold identical bits ----
    dp_check = None
    if world > 1:
        holder["step"] = step
        step.run(n_views)
        torch.cuda.synchronize(device)
        got = {k: t.clone() for k, t in step.buckets.tensors.items()}
        # (the whole batch on one rank: V_per views between their two phases at a time, like the ranks themselves — 64 views
        # of a 6 M scene would otherwise hold 64 x 3.5 GB of arenas)
        seq_step = holder["step"] = make_step(begin_view, world_=1, rank_=0, n_streams=1, buckets=step.buckets, in_flight=V_per)
        cap_timed, holder["cap"] = holder["cap"], None            # other ranks' views: exact instance counts again
        seq_step.run(n_views, reduce=False)
        torch.cuda.synchronize(device)
        seq = {k: t.clone() for k, t in step.buckets.tensors.items()}
        seq_step.run(n_views, reduce=False)          # the same sequential sum once more: this rank's own run-to-run noise
        torch.cuda.synchronize(device)

        def rel(a, b):
            return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
        ill = ("scaling", "rotation")   # fed by the ill-conditioned conic backward: order-of-summation noise of ~1e-3
        errs = {k: rel(got[k], seq[k]) for k in vp.ParameterBuckets.names}
        noise = {k: rel(step.buckets.tensors[k], seq[k]) for k in vp.ParameterBuckets.names}
        checksum = torch.stack([t.view(torch.int32).to(torch.int64).sum() for t in got.values()]).sum().reshape(1)
        sums = [torch.zeros_like(checksum) for _ in range(world)]
        dist.all_gather(sums, checksum)
        worst = torch.tensor([max(v for k, v in errs.items() if k not in ill), max(errs[k] for k in ill),
                              max(noise[k] for k in ill), max(v for k, v in noise.items() if k not in ill)],
                             device=device, dtype=torch.float64)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        identical = all(int(x) == int(sums[0]) for x in sums)
        well, ill_err, ill_noise, well_noise = (float(x) for x in worst)
        dp_check = {"views": n_views, "max_rel_err_vs_sequential_sum": well,
                    "max_rel_err_ill_conditioned(scaling,rotation)": ill_err,
                    "sequential_run_to_run_noise": {"well_conditioned": well_noise, "scaling,rotation": ill_noise},
                    "ranks_bit_identical": identical,
                    "pass": bool(well <= 1e-4 and ill_err <= max(2e-3, 4.0 * ill_noise) and identical),
                    "what": "all-reduced raw-parameter gradients of one step on every rank vs the same rank running all %d views "
                            "sequentially (max|d|/max|ref| per group, max over groups and ranks).  Gate: 1e-4; for the "
                            "scaling/rotation pair, which the ill-conditioned conic backward feeds, max(2e-3, 4 x the noise "
                            "between two sequential runs on the same rank)" % n_views}
        del seq
        del got
        holder["cap"] = cap_timed
        if not dp_check["pass"] and rank == 0:
            print("dp_check FAILED: %s" % json.dumps(dp_check), file=sys.stderr)

    # ---- where the N > 1 step time goes: the same step with the exchange switched off, slowest and fastest rank ----
    comm = None
    if world > 1:
        holder["step"] = step
        pygc.collect(); pygc.freeze()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step.run(n_views, reduce=False)
        e1.record()
        torch.cuda.synchronize(device)
        own = torch.tensor([e0.elapsed_time(e1) / args.steps], device=device)
        lo, hi = own.clone(), own.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        comm = {"ms_per_step_without_allreduce": {"slowest_rank": round(float(hi[0]), 3), "fastest_rank": round(float(lo[0]), 3)},
                "views": "cost-balanced assignment (instance counts, longest first)" if assignment is not None else "contiguous split",
                "what": "the headline step with the all-reduces left out, timed on every rank: the slowest rank bounds the "
                        "step from below, the rest of ms_per_step is exposed exchange"}
        if assignment is not None:
            c = stats["view_costs"]
            comm["instances_per_rank"] = [int(sum(c[v] for v in vs_)) for vs_ in assignment]
        barrier()

    # ---- side leg: the rasterizer alone on given inputs (the reference arm's workload; round 1's headline): forward +
    # backward per view into the gradients of the rasterizer's own inputs (73 floats per Gaussian), same deferred step ----
    param_bytes = step.buckets.nbytes_reduced()
    step.bucket_sets = step_e2e.bucket_sets = step_prof.bucket_sets = None
    step.buckets = step_e2e.buckets = step_prof.buckets = None
    holder["step"] = None
    torch.cuda.empty_cache()
    feats = {v: syn.pack_features(scene, cams[v], F) for v in my_views}

    def begin_raster(v):
        st = settings[v]
        color, radii, observe, buffer, state = forward(scene.means3D, scene.shs, None, scene.opacities, scene.scales,
                                                       scene.rotations, None, feats[v], st)
        dgr.backward_raw(gc, gb, scene.means3D, scene.shs, None, scene.scales, scene.rotations, None, feats[v], radii, st,
                         state, grads=holder["step"].buckets.tensors, phase="blend")
        return {"v": v, "radii": radii, "observe": observe, "state": state}

    def finish_raster(h, buckets, accumulate, rows):
        v = h["v"]
        dgr.backward_raw(gc, gb, scene.means3D, scene.shs, None, scene.scales, scene.rotations, None, feats[v], h["radii"],
                         settings[v], h["state"], grads=buckets.tensors, accumulate=accumulate, phase="gaussians", rows=rows,
                         densify_stats=holder["step"].stats.backward_args())

    step_r = vp.ViewShardedStep(P, M, device, world=world, rank=rank, n_streams=args.streams, begin_view=begin_raster,
                                finish_view=finish_raster, n_chunks=args.chunks, assignment=assignment,
                                max_steps_ahead=1 if args.no_wait_forward else None)
    holder["step"] = step_r
    for _ in range(4):
        step_r.run(n_views)
    raster_ms, _ = timed(step_r, args.steps)
    check_capacity("raster_only")
    raster_value = n_views * args.steps / (raster_ms * 1e-3)
    raster_bytes = step_r.buckets.nbytes_reduced()

    # ---- side leg (N = 1): what an unmodified GS-2M iteration sees — the drop-in autograd API on ONE stream, two views per
    # iteration with the gradients accumulating in the leaves' .grad (train.py:95, utils/loss_utils.py:253) ----
    dropin = None
    if world == 1:
        step_r.bucket_sets = step_r.buckets = None
        torch.cuda.empty_cache()
        leaves = dict(means3D=scene.means3D.clone().requires_grad_(True), opacities=scene.opacities.clone().requires_grad_(True),
                      shs=scene.shs.clone().requires_grad_(True), scales=scene.scales.clone().requires_grad_(True),
                      rotations=scene.rotations.clone().requires_grad_(True))
        m2d = torch.zeros(P, 4, device=device, requires_grad=True)
        fl = {v: feats[v].clone().requires_grad_(True) for v in my_views[:2]}

        def iteration():
            for t in list(leaves.values()) + list(fl.values()) + [m2d]:
                t.grad = None                                   # optimizer.zero_grad(set_to_none=True), train.py:259
            for v in my_views[:2]:
                color, radii, observe, buffer = dgr.GaussianRasterizer(settings[v])(
                    means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], shs=leaves["shs"], colors_precomp=None,
                    scales=leaves["scales"], rotations=leaves["rotations"], cov3D_precomp=None, features=fl[v])
                torch.autograd.backward([color, buffer], [gc, gb])
        for _ in range(2):
            iteration()
        torch.cuda.synchronize(device)
        n_it = max(2, args.steps * V_per // 4)
        pygc.collect(); pygc.freeze()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n_it):
            iteration()
        e1.record()
        torch.cuda.synchronize(device)
        d_ms, d_wall = e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3
        n_v = 2 * n_it
        dropin = {"value": round(n_v / (d_ms * 1e-3), 3), "unit": UNIT, "ms_per_view": round(d_ms / n_v, 4),
                  "wall_ms_per_view": round(d_wall / n_v, 4),
                  "what": "GaussianRasterizer(settings)(...) + torch.autograd.backward on torch's current stream, 2 views per "
                          "iteration, .grad set to None per iteration and accumulated by autograd (rasterizer only, like the "
                          "reference arm)"}

    if rank != 0:
        return None
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    ab = algorithmic_bytes(P, V_vis, R, W * H, F, M, tiles)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    n_prof = max(len(prof_views), 1)
    bwd_ms, bwd_calls = stage["blend_bwd"]
    bwd_avg_ms = bwd_ms / max(bwd_calls, 1)
    achieved = ab["blend_bwd"] / (bwd_avg_ms * 1e-3) / 1e9 if bwd_avg_ms > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("blend_backward_kernel")
    except Exception:
        pass
    per_view_ms = total_ms / (args.steps * V_per)
    raster_view_ms = raster_ms / (args.steps * V_per)
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: %d Gaussians, %dx%d, SH deg 3, feature_count %d (RGB+alpha+depth+normal+albedo+"
                               "roughness+metallic), %d views per rank per step; per view: fused activation+packing fwd, rasterizer "
                               "fwd+bwd, fused packing bwd into the 9 raw parameter-gradient groups (64 floats/Gaussian); "
                               "view-sharded DP, NCCL all-reduce of the parameter gradients overlapped range by range"
                               % (args.config, P, W, H, F, V_per),
                   "views_per_step": n_views, "streams_per_rank": step.n_streams, "gaussian_ranges": len(step.chunks),
                   "allreduce_bytes": param_bytes, "visible_gaussians": V_vis, "instances_R": R,
                   "l2_policy": "working set per view (%.1f GB algorithmic) exceeds the 126 MB L2; no flush needed"
                                % (ab["total"] / 1e9)},
        "ms_per_view": round(per_view_ms, 4),
        "step_ms": {"value": headline_step_ms, "e2e": stats.get("e2e_step_ms"),
                    "what": "rank 0's timed steps one by one (CUDA events between the steps; the headline uses the bracket around all of them)"},
        "gpu_launches": int(launches[0]),
        "raster_only": {"value": round(raster_value, 3), "unit": UNIT, "ms_per_view": round(raster_view_ms, 4),
                        "allreduce_bytes": raster_bytes,
                        "what": "rasterizer forward+backward alone on given inputs (the reference arm's workload): gradients of "
                                "the rasterizer's own inputs, 73 floats/Gaussian, same deferred view-sharded step"},
        "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": h2d * V_per, "d2h_bytes_per_step": d2h * V_per,
                "wall_ms_per_step": round(wall_ms / args.steps, 4)},
        "roofline": {"bound": "hbm", "kernel": "blend_backward_kernel<%d>" % F, "achieved": round(achieved, 2),
                     "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 5), "traffic": traffic,
                     "algorithmic_bytes_per_launch": ab["blend_bwd"], "avg_launch_ms": round(bwd_avg_ms, 4),
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                     "note": "blend kernels are FP32-issue/shared-memory bound, not HBM bound (DESIGN.md)"},
        "roofline_view": {"algorithmic_bytes_per_view": ab["total"],
                          "achieved": round(ab["total"] / (raster_view_ms * 1e-3) / 1e9, 2),
                          "frac": round(ab["total"] / (raster_view_ms * 1e-3) / 1e9 / peak, 5), "unit": "GB/s",
                          "note": "rasterizer-only leg (the byte model of SURVEY 8d covers the rasterizer's stages)"},
        "stage_ms_per_view": {k: round(v[0] / n_prof, 4) for k, v in stage.items()},
        "clocks": clock_info,
    }
    clocks.stop()            # every GPU leg is over
    if dp_check is not None:
        line["dp_check"] = dp_check
        line["scaling_breakdown"] = comm
    if dropin is not None:
        line["dropin"] = dropin
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(cfg, args)
    return line


# ----------------------------------------------------------------------------------------------------------------
def cpu_baseline(cfg, args, budget_tiles=None):
    """The CPU oracle (pure-PyTorch port of the same math) timed on the host cores on a bounded sample of the same
    workload: the full per-Gaussian stage + key sort for one view, and the blend forward+backward on an evenly spread
    subset of tiles, scaled to the whole image (stated in `sample`; never a silent extrapolation)."""
    import cpu_rasterizer as cr
    import golden_io
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    P, W, H, F = cfg["P"], cfg["W"], cfg["H"], cfg["F"]
    scene = syn.make_scene(P, shell_fraction=cfg["shell"], cluster=cfg.get("cluster"), opacity_cap=cfg.get("opacity_cap"))
    cam = syn.make_cameras(1, W, H, radius=cfg["cam_radius"])[0]
    feats = syn.pack_features(scene, cam, F)
    gc, gb = syn.make_upstream_grads(W, H, F)
    settings = syn.raster_settings_for(cam, F, golden_io.Settings)
    tiles_total = ((W + 15) // 16) * ((H + 15) // 16)
    n_sample = budget_tiles or args.cpu_tiles
    sample = list(range(0, tiles_total, max(1, tiles_total // n_sample)))[:n_sample]
    orc = cr.CpuRasterizer(torch.float32)
    t0 = time.perf_counter()
    orc.forward(scene.means3D, scene.shs, None, scene.opacities, scene.scales, scene.rotations, None, feats, settings,
                tiles=[])                       # per-Gaussian stage + duplication + sort, no blending
    t_geom = time.perf_counter() - t0
    t0 = time.perf_counter()
    cr.blend_forward(orc.pre, orc.lists, orc.inp["features"], orc.bg, W, H, F, tiles=sample)
    t_fwd = time.perf_counter() - t0
    # backward needs the forward's per-pixel state on the sampled tiles
    color, buffer, final_T, n_contrib, _ = cr.blend_forward(orc.pre, orc.lists, orc.inp["features"], orc.bg, W, H, F,
                                                            tiles=sample)
    orc.final_T, orc.n_contrib = final_T, n_contrib
    t0 = time.perf_counter()
    orc.backward(gc, gb, tiles=sample)          # blend backward on the sample + full per-Gaussian backward (autograd)
    t_bwd = time.perf_counter() - t0
    scale = tiles_total / len(sample)
    # the per-Gaussian backward inside orc.backward is not tile-proportional; time it apart to scale only the blend
    t0 = time.perf_counter()
    orc.backward(gc, gb, tiles=[])
    t_pre_bwd = time.perf_counter() - t0
    t_view = t_geom + t_fwd * scale + (t_bwd - t_pre_bwd) * scale + t_pre_bwd
    return {"value": round(1.0 / t_view, 5), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "oracle/cpu_rasterizer.py fp32, 1 view of the same workload: per-Gaussian fwd+bwd and key sort in "
                      "full (%.1f s + %.1f s), blend fwd+bwd on %d of %d tiles (%.1f s) scaled x%.1f"
                      % (t_geom, t_pre_bwd, len(sample), tiles_total, t_fwd + t_bwd - t_pre_bwd, scale),
            "seconds_per_view_estimated": round(t_view, 2)}


# ----------------------------------------------------------------------------------------------------------------
def run_reference(args, cfg, rank, world, device):
    """Reference arm: the unmodified reference rasterizer through its own binding (GaussianRasterizer + autograd)."""
    if rank != 0:
        return None
    import build_ref
    P, W, H, F = cfg["P"], cfg["W"], cfg["H"], cfg["F"]
    V_per = args.views_per_rank
    if not (build_ref.available() and torch.cuda.is_available()):
        cb = cpu_baseline(cfg, args)
        cb["kind"] = "port"
        return {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * V_per / cb["value"], 2),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "%s (CPU oracle port: %s)" % (
                    args.config, "no CUDA device" if not torch.cuda.is_available() else "compiled reference oracle/_ref not available")},
                "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    ref = build_ref.load()
    my_views = list(range(V_per))
    scene, cams_cpu, cams, feats, gc, gb = build_workload(cfg, V_per, my_views, device)
    leaves = dict(means3D=scene.means3D.clone().requires_grad_(True), opacities=scene.opacities.clone().requires_grad_(True),
                  shs=scene.shs.clone().requires_grad_(True), scales=scene.scales.clone().requires_grad_(True),
                  rotations=scene.rotations.clone().requires_grad_(True))
    settings = {v: syn.raster_settings_for(cams[v], F, ref.GaussianRasterizationSettings) for v in my_views}

    def step():
        for v in my_views:
            f = feats[v].clone().requires_grad_(True)
            m2d = torch.zeros(P, 4, device=device, requires_grad=True)
            color, radii, observe, buffer = ref.GaussianRasterizer(settings[v])(
                means3D=leaves["means3D"], means2D=m2d, opacities=leaves["opacities"], shs=leaves["shs"],
                colors_precomp=None, scales=leaves["scales"], rotations=leaves["rotations"], cov3D_precomp=None, features=f)
            torch.autograd.backward([color, buffer], [gc, gb])   # gradients accumulate in .grad across the views

    clocks = ClockSampler(device.index)
    clocks.start()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize(device)
    pygc.collect(); pygc.freeze()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with clocks.window():
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize(device)
    total_ms = e0.elapsed_time(e1)
    clock_info = clocks.summary()
    clocks.stop()
    value = V_per * args.steps / (total_ms * 1e-3)
    return {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_ms / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s: %d Gaussians, %dx%d, feature_count %d, fwd+bwd, %d views per step on ONE GPU "
                                   "(the reference has no multi-GPU path; rank 0 only)" % (args.config, P, W, H, F, V_per)},
            "ms_per_view": round(total_ms / (args.steps * V_per), 4),
            "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": 0, "kind": "reference",
                             "sample": "compiled reference CUDA rasterizer (oracle/_ref, sm_100 build of the unmodified "
                                       "sources) on the GPU through its own Python binding; full workload, no sampling"},
            "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "clocks": clock_info}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="tnt-3m", choices=sorted(syn.CONFIGS))
    ap.add_argument("--views-per-rank", type=int, default=8)
    ap.add_argument("--streams", type=int, default=2, help="views in flight per rank (CUDA streams)")
    ap.add_argument("--chunks", type=int, default=4, help="Gaussian ranges of the deferred per-Gaussian backward / all-reduce")
    ap.add_argument("--no-wait-forward", action="store_true",
                    help="timed legs: forwards never read the instance count back (no_wait), host bounded to one step ahead")
    ap.add_argument("--no-clock-sampler", action="store_true", help="diagnostic: do not poll nvidia-smi during the timed regions")
    ap.add_argument("--nccl-normal-priority", action="store_true", help="N > 1: NCCL on a normal-priority stream (default: high)")
    ap.add_argument("--contiguous-views", action="store_true",
                    help="N > 1: contiguous split of the batch over the ranks instead of the cost-balanced assignment")
    ap.add_argument("--per-view-finish", action="store_true",
                    help="run the per-Gaussian backward once per view (round-2a protocol) instead of one multi-view pass per range")
    ap.add_argument("--cpu-tiles", type=int, default=96, help="tiles of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = syn.CONFIGS[args.config]

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and args.impl == "reference" and rank != 0:
        return 0   # the reference arm runs on rank 0 alone
    if not torch.cuda.is_available():
        if args.impl == "reference":
            print(json.dumps(run_reference(args, cfg, 0, 1, None)))
            return 0
        raise SystemExit("bench.py needs a CUDA device: the rasterizer has no CPU path")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1 and args.impl == "ours":
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's kernels on a HIGH-priority stream: a finished Gaussian range's all-reduce and the next range's per-Gaussian
        # kernel become runnable at the same moment, and the per-Gaussian kernel fills every SM's register file, so at equal
        # priority the exchange's CTAs (hundreds of threads each) only get in when that grid drains, one range late
        opts = None if args.nccl_normal_priority else dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=device, pg_options=opts)
    try:
        line = run_ours(args, cfg, rank, world, device) if args.impl == "ours" else run_reference(args, cfg, rank, world, device)
        if line is not None:
            print(json.dumps(line))
            sys.stdout.flush()
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

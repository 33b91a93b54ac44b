#!/usr/bin/env python
"""bench.py — BASELINE.json configs[1]: cam_bp 256x256 depth -> 128^3 voxel TDF, batch 32 per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path (one rank per GPU)
    python bench.py --impl reference [...]                         the reference's CPU path (rank 0 only)

One JSON line on stdout (rank 0).  A "step" is one cam_bp forward over a [32,1,256,256] depth batch.
  value     whole-job shapes/s with the inputs already resident in HBM (device-timed, max over ranks)
  e2e       the same through the public module with HOST buffers: pinned depth -> H2D -> kernels -> D2H of the
            [32,1,128,128,128] result, every step, inside the timed region
  roofline  the dominant kernel (vox_splat_kernel) timed alone with CUDA events on its stream:
            algorithmic bytes B*(4*H*W + 4*R^3) per launch / its duration, against the measured HBM peak
  cpu_baseline  the CPU oracle (a port of the reference's CUDA-only op) on rank 0's host, 1 thread
The oracle is only executed by the cpu_baseline leg and by --impl reference.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# the contract is ONE JSON line on stdout: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) off it
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

H = W = 256
RES = 128
FL, CAM_DIST = 418.3, 2.2
METRIC = "cam_bp shapes/sec @128^3 voxel (256x256 depth -> 128^3 TDF)"
UNIT = "shapes/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="depth maps per GPU per step (BASELINE: 32)")
    ap.add_argument("--no-graph", action="store_true", help="launch steps from Python instead of replaying a CUDA graph")
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="budget of the cpu_baseline leg")
    return ap.parse_args()


def config(args, n_gpus):
    return {"workload": "cam_bp 256x256 depth -> 128^3 voxel back-projection, batch=%d per GPU (BASELINE configs[1])" % args.batch,
            "batch_per_gpu": args.batch, "global_batch": args.batch * n_gpus, "depth_hw": [H, W], "voxel_res": RES,
            "fl": FL, "cam_dist": CAM_DIST, "shift_tdf": True, "parallelism": "replicas x%d (batch-sharded, no collective)" % n_gpus,
            "l2": "no explicit flush: each step streams 256 MiB of output + ~13 MB of scratch (> 126 MB L2); inputs rotate over 4 buffers"}


# ----------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, mxc = float(f[1]), float(f[2])
            except ValueError:
                continue
            mx = mxc
            if t0 - 0.05 <= t <= t1 + 0.05:
                sm.append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # region shorter than the sampling period: fall back to every sample taken
            sm = [float(l.split(",")[1]) for _, l in self.rows if len(l.split(",")) > 2 and l.split(",")[1].strip().replace(".", "").isdigit()]
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# reference arm: the reference op is CUDA-only, so its CPU path is the oracle port, all host threads
# ----------------------------------------------------------------------------------------------------
def cpu_forward_batch(oracle, depth, threads):
    """cam_bp forward (+shift) of every map of `depth` on the CPU oracle; maps are independent."""
    if threads <= 1:
        for i in range(depth.shape[0]):
            oracle.cam_bp_forward(depth[i:i + 1], FL, CAM_DIST, RES, shift=True)
        return
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(threads) as ex:  # ctypes releases the GIL inside liboracle.so
        list(ex.map(lambda i: oracle.cam_bp_forward(depth[i:i + 1], FL, CAM_DIST, RES, shift=True), range(depth.shape[0])))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from genre_shapehd_b200.synth import bench_depth_batch
    from oracle import oracle
    oracle.lib()
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 256))            # every host thread: maps are independent, one map per thread at a time
    import numpy as np
    batch = bench_depth_batch(args.batch)
    # each step = a bounded sample of the workload: whole 32-map batches, as many as it takes to occupy every thread once
    n_batches = max(1, -(-threads // args.batch))
    depth = np.ascontiguousarray(np.tile(batch, (n_batches, 1, 1, 1)))
    sample = depth.shape[0]
    cpu_forward_batch(oracle, depth[:2], 1)  # touch pages
    t = time.time()
    cpu_forward_batch(oracle, depth[:sample], threads)
    per_batch = time.time() - t
    budget = 90.0
    steps, warmup = args.steps, args.warmup
    if per_batch * (steps + warmup) > budget:  # keep the whole run within a few minutes
        sample = max(min(threads, sample), int(sample * budget / (per_batch * (steps + warmup))))
    for _ in range(warmup):
        cpu_forward_batch(oracle, depth[:sample], threads)
    t0 = time.time()
    for _ in range(steps):
        cpu_forward_batch(oracle, depth[:sample], threads)
    dt = time.time() - t0
    value = sample * steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config(args, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": min(threads, sample), "kind": "port",
                             "sample": "%d maps per step (the %d-map batch repeated), %d threads over maps, host has %d logical CPUs "
                                       "(oracle/genre_oracle.c; the reference op has no CPU implementation)"
                                       % (sample, args.batch, min(threads, sample), cores)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch

    import genre_shapehd_b200
    genre_shapehd_b200.install()
    from genre_shapehd_b200 import _lib, dist_util
    from genre_shapehd_b200.synth import bench_depth_batch
    from toolbox.cam_bp.cam_bp.modules.camera_backprojection_module import Camera_back_projection_layer

    world, rank, local = dist_util.env_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist_util.init("nccl", dev)
    _lib.load()

    B, K, Wm = args.batch, args.steps, max(args.warmup, 3)
    layer = Camera_back_projection_layer()
    host = bench_depth_batch(B)
    n_in = 4
    inputs = [torch.from_numpy(host).to(dev).clone() for _ in range(n_in)]

    def barrier():
        dist_util.barrier(dev)

    def max_over_ranks(ms):
        return dist_util.max_over_ranks(ms, dev)

    # ---- leg 1: inputs resident in HBM --------------------------------------------------------------
    use_graph = not args.no_graph
    graphs = []
    with torch.no_grad():
        if use_graph:
            # one CUDA graph per rotating input buffer (counter memset + project + splat: 2 kernels)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for x in inputs:
                    layer(x)
            torch.cuda.current_stream().wait_stream(side)
            for x in inputs:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    out = layer(x)
                graphs.append((g, out))

        def step(i):
            if use_graph:
                graphs[i % n_in][0].replay()
            else:
                layer(inputs[i % n_in])

        for i in range(Wm):
            step(i)
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
            time.sleep(0.25)
        launches0 = _lib.launch_count
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.time()
        e0.record()
        for i in range(K):
            step(i)
        e1.record()
        barrier()
        t_wall1 = time.time()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        launches = (K * 2) if use_graph else (_lib.launch_count - launches0)
        # nvidia-smi samples every 100 ms and the timed region lasts a few ms: keep the same step running (untimed) for
        # ~0.5 s so that the clock / throttle samples describe this kernel mix under sustained load
        t_load1 = t_wall1
        try:
            n_load = min(20000, max(K, int(0.5 / max(ms_total / K * 1e-3, 1e-6))))
            for i in range(n_load):
                step(i)
            torch.cuda.synchronize()
            t_load1 = time.time()
        except Exception:  # the clock window is informational: never let it fail the measurement
            pass
        clocks = sampler.stop(t_wall0, t_load1) if rank == 0 else None
        if clocks is not None:
            clocks["window"] = "timed region + %.2f s untimed replay of the same step" % (t_load1 - t_wall1)

    value = world * B * K / (ms_total * 1e-3)

    # ---- leg 2: end to end with host buffers ---------------------------------------------------------
    pin_in = torch.from_numpy(host).pin_memory()
    pin_out = torch.empty((B, 1, RES, RES, RES), dtype=torch.float32).pin_memory()
    d_in = torch.empty_like(inputs[0])
    Ke = max(3, min(K, 20))

    def e2e_step():
        d_in.copy_(pin_in, non_blocking=True)
        with torch.no_grad():
            o = layer(d_in)
        pin_out.copy_(o, non_blocking=True)

    for _ in range(3):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(Ke):
        e2e_step()
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * B * Ke / (ms_e2e * 1e-3)
    checksum = float(pin_out.double().sum())

    # ---- leg 3: the dominant kernel on its own (rank 0) ----------------------------------------------
    roofline = None
    if rank == 0:
        x = inputs[0]
        ws, nbytes = _lib.workspace_for(B, H * W, RES, dev)
        tdf = torch.empty((B, 1, RES, RES, RES), device=dev)
        st = _lib.stream_ptr(x)
        fl = torch.full((1, 1), FL, device=dev).expand(B, 1)
        cd = torch.full((1, 1), CAM_DIST, device=dev).expand(B, 1)
        _lib.call("genre_b200_cam_bp_stage_project", x.data_ptr(), B, 1, H, W, *x.stride(), fl.data_ptr(), *fl.stride(),
                  cd.data_ptr(), *cd.stride(), RES, ws.data_ptr(), nbytes, st)

        def splat():
            _lib.call("genre_b200_voxelize_stage_splat", B, H * W, RES, tdf.data_ptr(), None, 1.0, -1.0 / 16777216.0, 0.0,
                      ws.data_ptr(), nbytes, st)
        for _ in range(5):
            splat()
        torch.cuda.synchronize()
        reps = 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            splat()
        e1.record()
        torch.cuda.synchronize()
        splat_ms = e0.elapsed_time(e1) / reps
        alg_bytes = B * (4 * H * W + 4 * RES ** 3)
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        pk = os.path.join(REPO, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            try:
                peak, peak_src = float(json.load(open(pk))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
            except Exception:
                pass
        traffic = None
        tf = os.path.join(REPO, "profiles", "splat_traffic.json")
        if os.path.exists(tf):
            try:
                traffic = json.load(open(tf)).get("dram_bytes_per_launch")
            except Exception:
                pass
        achieved = alg_bytes / (splat_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": "vox_splat_kernel<VEC=true,WRITE_CNT=false>", "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "kernel_us": splat_ms * 1e3,
                    "whole_op_GBps": alg_bytes / (ms_total / K * 1e-3) / 1e9,
                    "whole_op_frac": alg_bytes / (ms_total / K * 1e-3) / 1e9 / peak}

    # ---- leg 4: CPU baseline (rank 0, N=1 only): the oracle port, one thread, bounded sample ----------
    cpu = None
    if rank == 0 and world == 1:
        from oracle import oracle
        oracle.lib()
        n_maps, t0 = 0, time.time()
        cpu_forward_batch(oracle, host[:1], 1)
        t0 = time.time()
        while time.time() - t0 < args.cpu_seconds:
            cpu_forward_batch(oracle, host[n_maps % B:n_maps % B + 1], 1)
            n_maps += 1
        dt = time.time() - t0
        cpu = {"value": n_maps / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "%d maps of the batch, one at a time, %.1f s of oracle/genre_oracle.c (single thread; "
                         "the reference op itself is CUDA-only)" % (n_maps, dt),
               "host_cpus": os.cpu_count()}

    # ---- secondary (rank 0, N=1): the GenRe 3D hot path at batch 16, informational ------------------------------------
    secondary = None   # GENRE_B200_BENCH_SECONDARY=0 skips it (used for the ncu launch list of the headline step alone)
    if rank == 0 and world == 1 and os.environ.get("GENRE_B200_BENCH_SECONDARY", "1") != "0":
        try:
            secondary = genre3d_path(torch, dev)
        except Exception as e:  # informational only: never fail the headline measurement
            secondary = {"error": repr(e)[:200]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config(args, world),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * H * W * 4,
                        "d2h_bytes_per_step": B * RES ** 3 * 4, "steps": Ke, "ms_per_step": ms_e2e / Ke,
                        "result_checksum": checksum},
                "gpu_launches": launches, "launch_mode": "cuda_graph" if use_graph else "python",
                "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "secondary": secondary}
        print(json.dumps(line), flush=True)
    dist_util.finalize()


def genre3d_path(torch, dev, batch=16, reps=5):
    """BASELINE configs[2] without the two 2D U-ResNets (out of scope): cam_bp -> render_spherical -> sph_pad ->
    backproject_spherical glue -> clamp/cat -> Unet_3D (eval).  Timed twice: with the glue lines of the frozen callers
    (depth_pred_with_sph_inpaint.py:120-126, genre_full_model.py:120-143) on the drop-in ops, and with the opt-in
    fused glue (genre_shapehd_b200/fused.py)."""
    from genre_shapehd_b200.fused import GenRe3DGlue
    from genre_shapehd_b200.synth import bench_depth_batch
    from toolbox.cam_bp.cam_bp.functions import SphericalBackProjection
    from toolbox.cam_bp.cam_bp.modules.camera_backprojection_module import Camera_back_projection_layer
    from toolbox.spherical_proj import gen_sph_grid, render_spherical, sph_pad
    import networks.networks as nets
    depth = torch.from_numpy(bench_depth_batch(batch)).to(dev)
    proj, rend = Camera_back_projection_layer(), render_spherical().to(dev)
    grid = gen_sph_grid().to(dev).expand(batch, -1, -1, -1, -1)
    unet = nets.Unet_3D().to(dev).eval()
    glue = GenRe3DGlue().to(dev)

    def step_callers():
        pd = proj(depth)
        sph = sph_pad(rend(torch.clamp(pd * 50, 1e-5, 1 - 1e-5)), 16)
        df, cnt = SphericalBackProjection.apply(1 - sph[:, :, 16:144, 16:144], grid, 128)
        ps = (-df + 1 / 128) * 128 * torch.clamp(cnt, 0, 1)
        return unet(torch.cat((ps, torch.clamp((pd * 50) / 50, 1e-5, 1 - 1e-5)), dim=1))

    def step_fused():
        pd, sph = glue.project_and_render(depth)
        return unet(glue.refine_input(pd, sph))

    def timed(step):
        with torch.no_grad():
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                step()
            e1.record()
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    ms_callers, ms_fused = timed(step_callers), timed(step_fused)
    return {"what": "GenRe 3D hot path (cam_bp + render_spherical + sph_bp + Unet_3D eval), batch %d, 2D nets excluded" % batch,
            "ms_per_batch": ms_fused, "shapes_per_s": batch / ms_fused * 1e3, "glue": "fused (genre_shapehd_b200/fused.py)",
            "callers_glue_ms_per_batch": ms_callers, "callers_glue_shapes_per_s": batch / ms_callers * 1e3}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

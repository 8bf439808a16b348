#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: train points/s of the DiGS hot path.

A "step" = forward jets of one 15 000 + 15 000 point batch (value + gradient on both clouds, Laplacian on the
off-surface cloud) + DiGSLoss + the reverse pass to all parameter gradients (+ the gradient all-reduce when N > 1),
on the SRB recipe network (4x256 sine MLP, MFGI init, BASELINE.json configs[1]).

    python bench.py --gpus 1 --steps 20 --warmup 5              # our arm
    python bench.py --impl reference --steps 5 --warmup 2       # the reference algorithm on the host cores

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for how every field is obtained.
"""
import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NET_KW = dict(latent_size=0, in_dim=3, decoder_hidden_dim=256, nl="sine", encoder_type="none",
              decoder_n_hidden_layers=4, init_type="mfgi", sphere_init_params=[1.6, 0.1])
LOSS_KW = dict(weights=[3e3, 1e2, 1e2, 5e1, 1e2], loss_type="siren_wo_n_w_div", div_decay="linear", div_type="l1",
               div_clamp=50)
MODEL_SEED = 3627473            # surface_recon_args.py:13
N_POINTS = 15000                # run_surf_recon_exp.sh:53
H, LH, D = 256, 4, 3
# algorithmic GEMM FLOPs: forward C*2*Lh*H^2 per point; training = 3x
FLOP_FWD_NONMNFLD = (2 + D) * 2 * LH * H * H
FLOP_FWD_MNFLD = (1 + D) * 2 * LH * H * H
FLOP_STEP = 3 * N_POINTS * (FLOP_FWD_NONMNFLD + FLOP_FWD_MNFLD)
# dram__bytes_read.sum + dram__bytes_write.sum of the forward-jet kernel (C=5, 15000 points, saving for the reverse pass),
# per launch, from the `ncu --set full` capture summarised in profiles/r01_full_tc_kernels.md.  The kernel's algorithmic
# bytes are the saved planes + wgrad operand: 2 * 15000 * 4 layers * 5 channels * 256 units * 4 B = 614 MB.
FWD_DRAM_BYTES = {"bf16x2": 559.65e6}
FWD_SAVED_BYTES = N_POINTS * (LH * 5 * H * 8 + 5 * 4)      # C = 5 channels: fp32 planes + (hi|lo) words per layer, head per point


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return dict(burst=float(p["bf16_tflops"]), sustained=float(p["bf16_tflops_sustained"]),
                    hbm=float(p["hbm_gbs"]), source="measured")
    except Exception:
        return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML, 20 ms period; B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.sm, self.reasons, self.smax, self.stop_flag, self.ok = index, [], set(), None, False, False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.ok = False

    def _poll(self):
        nv = self.nv
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self.stop_flag = True
        self.thread.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.smax,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def host_batches(n_batches, seed):
    from digs_b200 import synth
    cloud, normals = synth.surface_cloud(100000, "torus", seed=0)
    out = []
    for i in range(n_batches):
        m, nrm, nm = synth.sample_batch(cloud, normals, N_POINTS, seed=seed * 100003 + i, box=1.1)
        out.append(tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in (m, nrm, nm)))
    return out


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (autograd with create_graph + per-axis second pass) on the host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_reference_steps(steps, warmup, n_points=N_POINTS, budget_s=150.0):
    """Times oracle/ref_autograd.step (the "port" of models/DiGS.py + models/losses.py + utils.gradient)."""
    import digs_b200
    from oracle import ref_autograd
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(MODEL_SEED)
    net = digs_b200.DiGSNetwork(**NET_KW)
    params = [p.detach().clone().requires_grad_(True) for p in net.decoder.fc_block.flat_params()]
    batches = host_batches(steps + warmup + 1, seed=7)

    def one(i, npts):
        m, nrm, nm = batches[i % len(batches)]
        t0 = time.perf_counter()
        ref_autograd.step(params, m[0, :npts], nm[0, :npts], nrm[0, :npts], (1.6, 0.1), LOSS_KW["weights"],
                          LOSS_KW["loss_type"], LOSS_KW["div_type"], LOSS_KW["div_clamp"])
        return time.perf_counter() - t0

    t_cal = one(0, n_points)                       # calibration (also first warm-up)
    npts = n_points
    if (steps + warmup) * t_cal > budget_s:        # bounded sample: fewer points per step, same network
        npts = max(1000, int(n_points * budget_s / ((steps + warmup) * t_cal)))
    for i in range(warmup):
        one(i + 1, npts)
    ts = [one(warmup + 1 + i, npts) for i in range(steps)]
    sec = float(np.sum(ts))
    value = 2 * npts * steps / sec
    sample = "%d timed + %d warm-up steps of %d+%d points, 4x256 MFGI, torch %s CPU, %d threads" % (
        steps, warmup + 1, npts, npts, torch.__version__, torch.get_num_threads())
    return value, 1e3 * sec / steps, cores, sample


def run_reference_arm(args, rank):
    if rank != 0:
        return
    value, ms, cores, sample = cpu_reference_steps(args.steps, args.warmup)
    line = {"impl": "reference", "metric": "train points/sec (value+grad+div+backward)", "value": value,
            "unit": "points/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": "SRB-style recon step (BASELINE configs[1]): 15000 manifold + 15000 off-surface points per GPU, "
                        "4x256 sine MLP MFGI init, DiGSLoss siren_wo_n_w_div/l1",
            "points_per_step_per_gpu": 2 * N_POINTS, "hidden": H, "hidden_layers": LH, "in_dim": D,
            "precision_mode": args.mode, "parallelism": "dp%d (points sharded, 1 allreduce of 264449 fp32 grads)" % world,
            "step_api": "eager autograd" if args.no_graph else "digs_b200.GraphedStep (CUDA graph replay)",
            "l2": "256 MiB scratch written between timed steps (L2 flush); per-step events exclude the flush"}


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    import digs_b200
    from digs_b200 import _lib, dist as ddist
    from digs_b200 import jet as djet

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.manual_seed(MODEL_SEED)
    net = digs_b200.DiGSNetwork(**NET_KW, precision=args.mode).to(dev)
    crit = digs_b200.DiGSLoss(**copy.deepcopy(LOSS_KW))
    crit.global_counts = (N_POINTS * world, N_POINTS * world)
    flat = ddist.FlatGrads(list(net.parameters()), extra=8)
    net.decoder.fc_block.accumulate_grads_in_place = True   # reverse kernels add straight into the flat buffer
    L = _lib.lib()
    n_batches = args.steps + args.warmup
    hb = [tuple(t.pin_memory() for t in b) for b in host_batches(n_batches, seed=rank + 1)]
    db = [tuple(t.to(dev) for t in b) for b in hb]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    h2d_bytes = sum(t.numel() * 4 for t in hb[0])

    graphed = None
    if not args.no_graph:
        # public API digs_b200.GraphedStep: zero-fill + both forward jets + DiGSLoss + reverse pass as ONE CUDA graph
        graphed = digs_b200.GraphedStep(net, crit, N_POINTS, N_POINTS, batch=1, flat=flat)

    def step_resident(i):
        m, nrm, nm = db[i]
        if graphed is not None:
            ld = graphed.step(m, nm, nrm)
        else:
            m = m.detach().requires_grad_()
            nm = nm.detach().requires_grad_()
            flat.zero_()
            out = net(nm, m)
            ld, _ = crit(out, m, nm, nrm)
            ld["loss"].backward()
        flat.extra[0] = ld["loss"].detach()
        flat.allreduce()
        return ld

    def step_e2e(i):
        if graphed is not None:
            m, nrm, nm = hb[i]                      # pinned host tensors: GraphedStep.step copies them in asynchronously
            ld = graphed.step(m, nm, nrm)
        else:
            m, nrm, nm = (t.to(dev, non_blocking=True) for t in hb[i])
            m.requires_grad_()
            nm.requires_grad_()
            flat.zero_()
            out = net(nm, m)
            ld, _ = crit(out, m, nm, nrm)
            ld["loss"].backward()
        flat.extra[0] = ld["loss"].detach()
        flat.allreduce()
        return float(flat.extra[0].item())        # D2H read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------------
    for i in range(args.warmup):
        step_resident(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.digs_launch_count()
    evs = []
    barrier()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_resident(args.warmup + i)
        e1.record()
        evs.append((e0, e1))
    barrier()
    launches = (L.digs_launch_count() - launches0) / max(args.steps, 1)
    if graphed is not None:
        launches = graphed.launches_per_step      # library kernels recorded in (and replayed from) the graph each step
    ms_dev = sum(a.elapsed_time(b) for a, b in evs) / args.steps

    # ---- end-to-end timing (host buffers in, loss scalar out, every step) --------------------------
    for i in range(min(args.warmup, 3)):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(args.warmup + i)
    barrier()
    ms_e2e = 1e3 * (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop() if rank == 0 else None

    # ---- dominant kernel family alone: forward jet of the off-surface cloud ------------------------
    fc = net.decoder.fc_block
    params = [p.detach() for p in fc.flat_params()]
    xyz = db[0][2][0].contiguous()
    for _ in range(3):
        djet.jet_forward_raw(fc.spec(), params, xyz, None, 0, True, True, fc.mode())
    reps = 10
    kev = []
    for r in range(reps):
        flush.fill_(r)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        djet.jet_forward_raw(fc.spec(), params, xyz, None, 0, True, True, fc.mode())
        e1.record()
        kev.append((e0, e1))
    torch.cuda.synchronize()
    ms_fwd = sum(a.elapsed_time(b) for a, b in kev) / reps
    # pure-write bandwidth of this part, measured on the 256 MiB flush fill: the saving kernels write far more than they read
    fev = []
    for r in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        flush.fill_(r)
        e1.record()
        fev.append((e0, e1))
    torch.cuda.synchronize()
    fill_gbs = flush.numel() / (min(a.elapsed_time(b) for a, b in fev) * 1e-3) / 1e9

    # ---- second half of the metric: MC grid queries/s, grid_res^3 points sharded by iy slab over the ranks -----
    ms_grid, grid_pts = 0.0, 0
    if args.grid_res > 0:
        axes, _, _ = digs_b200.grid_axes(args.grid_res, np.array([[-1, 1], [-1, 1], [-1, 1]], dtype=float))
        ny = len(axes[1])
        slab = digs_b200.slab_range(ny, rank, world)
        out_vol = torch.empty((slab[1] - slab[0], len(axes[0]), len(axes[2])), dtype=torch.float32, device=dev)
        small = digs_b200.slab_range(ny, 0, 16 * world)
        digs_b200.grid_query(net, axes, iy_range=small)                      # warm-up
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        digs_b200.grid_query(net, axes, iy_range=slab, out=out_vol)
        e1.record()
        barrier()
        ms_grid = e0.elapsed_time(e1)
        grid_pts = len(axes[0]) * ny * len(axes[2])
        # mesh extraction from this rank's slab (utils/utils.py:354: marching cubes on the host in the reference), one pass
        mesh_info = None
        if rank == 0:
            from digs_b200 import mesh as dmesh
            dmesh.marching_tets(out_vol[:8].contiguous(), 0.0)                # warm-up
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record()
            mv, mf, _ = dmesh.marching_tets(out_vol, 0.0, axis_of=(1, 0, 2))
            m1.record()
            torch.cuda.synchronize(dev)
            mesh_info = {"ms": m0.elapsed_time(m1), "vertices": int(mv.shape[0]), "faces": int(mf.shape[0]),
                         "what": "marching tetrahedra on rank 0's slab of the volume (3 kernels + 2 prefix sums)"}
            del mv, mf

    t = torch.tensor([ms_dev, ms_e2e, ms_grid], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, ms_grid = float(t[0]), float(t[1]), float(t[2])
    if rank == 0:
        peaks = measured_peaks()
        pts = 2 * N_POINTS * world
        step_tflops = FLOP_STEP / (ms_dev * 1e-3) / 1e12
        fwd_tflops = N_POINTS * FLOP_FWD_NONMNFLD / (ms_fwd * 1e-3) / 1e12
        line = {"metric": "train points/sec (value+grad+div+backward)", "value": pts / (ms_dev * 1e-3),
                "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32", "bf16x2": "bf16x2-split (f32 accumulate)", "bf16": "bf16"}[args.mode],
                "data": "synthetic", "config": workload_config(args, world),
                "e2e": {"value": pts / (ms_e2e * 1e-3), "unit": "points/s", "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e},
                "gpu_launches": launches,
                "clocks": clocks,
                "roofline": {"bound": "tensor", "achieved": fwd_tflops, "peak": peaks["burst"], "unit": "TFLOP/s",
                             "frac": fwd_tflops / peaks["burst"], "traffic": FWD_DRAM_BYTES.get(args.mode),
                             "kernel": "forward jet of the 15000-point off-surface cloud (C=5 channels), timed alone",
                             "ms": ms_fwd, "algorithmic_gflop": N_POINTS * FLOP_FWD_NONMNFLD / 1e9,
                             "peak_source": peaks["source"] + " bf16 dense burst",
                             "hbm_view": {"what": "the same kernel as an HBM problem: it writes the state the reverse pass needs "
                                                  "(planes + layer-input words + head: Lh*C*H*8 + C*4 bytes per point)",
                                          "algorithmic_bytes": FWD_SAVED_BYTES, "achieved": FWD_SAVED_BYTES / (ms_fwd * 1e-3) / 1e9,
                                          "peak": peaks["hbm"], "unit": "GB/s",
                                          "frac": FWD_SAVED_BYTES / (ms_fwd * 1e-3) / 1e9 / peaks["hbm"],
                                          "fill_GBps_measured": fill_gbs,
                                          "frac_of_fill": FWD_SAVED_BYTES / (ms_fwd * 1e-3) / 1e9 / fill_gbs},
                             "whole_step": {"achieved": step_tflops, "peak": peaks["sustained"],
                                            "frac": step_tflops / peaks["sustained"],
                                            "algorithmic_gflop": FLOP_STEP / 1e9,
                                            "peak_source": peaks["source"] + " bf16 dense sustained"}}}
        if args.grid_res > 0:
            flop_pt = 2 * LH * H * H + 2 * D * H + 2 * H
            gt = grid_pts * flop_pt / (ms_grid * 1e-3) / 1e12
            line["grid_query"] = {"metric": "MC grid queries/sec", "value": grid_pts / (ms_grid * 1e-3), "unit": "queries/s",
                                  "grid": "%d^3, bbox [-1,1]^3 + 0.1, fp16-rounded axes, iy slabs over %d rank(s)" % (args.grid_res, world),
                                  "ms": ms_grid, "output": "fp32 volume resident in HBM",
                                  "roofline": {"bound": "tensor", "achieved": gt, "peak": peaks["sustained"], "unit": "TFLOP/s",
                                               "frac": gt / peaks["sustained"], "algorithmic_mflop_per_query": flop_pt / 1e6}}
            if mesh_info is not None:
                line["grid_query"]["mesh_extraction"] = mesh_info
        if world == 1 and not args.no_cpu:
            v, ms, cores, sample = cpu_reference_steps(3, 1)
            line["cpu_baseline"] = {"value": v, "unit": "points/s", "cores": cores, "kind": "port", "sample": sample,
                                    "ms_per_step": ms}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=os.environ.get("DIGS_BENCH_MODE", "bf16x2"), choices=["fp32", "bf16x2", "bf16"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--grid-res", type=int, default=512, help="resolution of the grid-query leg (0 disables it)")
    ap.add_argument("--no-graph", action="store_true", help="eager autograd step instead of digs_b200.GraphedStep")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

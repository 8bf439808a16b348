#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: train points/s of the DiGS hot path.

A "step" = forward jets of one 15 000 + 15 000 point batch (value + gradient on both clouds, Laplacian on the
off-surface cloud) + DiGSLoss + the reverse pass to all parameter gradients (+ the gradient all-reduce when N > 1),
on the SRB recipe network (4x256 sine MLP, MFGI init, BASELINE.json configs[1]).

    python bench.py --gpus 1 --steps 20 --warmup 5              # our arm (one CUDA-graph replay per step)
    python bench.py --impl reference --steps 5 --warmup 2       # the unmodified reference modules on the host cores
    torchrun ... bench.py --gpus 8 --scaling strong             # one batch's points split over the ranks (default: weak)
    python bench.py --config c4 | c5                            # scene (8x512) / shapespace (latent 256) workloads

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
# Reference arm: the UNMODIFIED reference modules (models/DiGS.py + models/losses.py + utils.gradient, installed by
# __graft_entry__.build() into the git-ignored baseline/_ref/) on the box's host cores -- or, for context, on the
# B200 through stock PyTorch CUDA.  Falls back to the oracle port (oracle/ref_autograd.py) when the install is absent.
# ---------------------------------------------------------------------------------------------------------
def reference_steps(steps, warmup, device="cpu", n_points=N_POINTS, budget_s=150.0):
    """forward + DiGSLoss + backward of train_surface_reconstruction.py:118-128 on `device`.
    Returns dict(value points/s from the MEDIAN step, ms_per_step, cores, kind, sample)."""
    from oracle import reference_loader, ref_autograd
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batches = host_batches(steps + warmup + 1, seed=7)
    kind = "reference" if reference_loader.available() else "port"
    dev = torch.device(device)
    if kind == "reference":
        ref_digs, _, _ = reference_loader.load()
        torch.manual_seed(MODEL_SEED)
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):          # the reference prints a banner: stdout carries the JSON line only
            net = ref_digs.DiGSNetwork(**NET_KW).to(dev)
        crit = ref_digs.DiGSLoss(**copy.deepcopy(LOSS_KW))

        def one(i, npts):
            m, nrm, nm = (t[:, :npts].to(dev) for t in batches[i % len(batches)])
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            m.requires_grad_()                                   # train_surface_reconstruction.py:118-119
            nm.requires_grad_()
            net.zero_grad()
            out = net(nm, m)                                     # :121
            ld, _ = crit(out, m, nm, nrm)                        # :123
            ld["loss"].backward()                                # :128
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)
            return time.perf_counter() - t0
    else:
        import digs_b200
        torch.manual_seed(MODEL_SEED)
        net = digs_b200.DiGSNetwork(**NET_KW)
        params = [p.detach().clone().to(dev).requires_grad_(True) for p in net.decoder.fc_block.flat_params()]

        def one(i, npts):
            m, nrm, nm = (t[0, :npts].to(dev) for t in batches[i % len(batches)])
            t0 = time.perf_counter()
            ref_autograd.step(params, m, nm, nrm, (1.6, 0.1), LOSS_KW["weights"], LOSS_KW["loss_type"],
                              LOSS_KW["div_type"], LOSS_KW["div_clamp"])
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)
            return time.perf_counter() - t0

    t_cal = one(0, n_points)                       # calibration (also first warm-up)
    npts = n_points
    if (steps + warmup) * t_cal > budget_s:        # bounded sample: fewer points per step, same network
        npts = max(1000, int(n_points * budget_s / ((steps + warmup) * t_cal)))
    for i in range(warmup):
        one(i + 1, npts)
    ts = [one(warmup + 1 + i, npts) for i in range(steps)]
    med = float(np.median(ts))
    what = ("unmodified reference modules (baseline/_ref: models/DiGS.py, models/losses.py, utils.gradient)"
            if kind == "reference" else "oracle port of the reference mechanism (oracle/ref_autograd.py)")
    where = ("torch %s CPU, %d threads" % (torch.__version__, torch.get_num_threads()) if dev.type == "cpu"
             else "stock PyTorch CUDA kernels on %s (fp32, allow_tf32=%s)" % (torch.cuda.get_device_name(dev),
                                                                               torch.backends.cuda.matmul.allow_tf32))
    sample = "%s; median of %d timed steps after %d warm-up steps of %d+%d points, 4x256 MFGI; %s" % (
        what, steps, warmup + 1, npts, npts, where)
    return dict(value=2 * npts / med, ms_per_step=1e3 * med, cores=cores, kind=kind, sample=sample, points=npts)


def reference_config(args, r):
    return {"workload": "SRB-style recon step (BASELINE configs[1]): 15000 manifold + 15000 off-surface points, "
                        "4x256 sine MLP MFGI init, DiGSLoss siren_wo_n_w_div/l1",
            "points_per_step_per_gpu": 2 * r["points"], "hidden": H, "hidden_layers": LH, "in_dim": D,
            "precision_mode": "fp32 (torch CPU autograd with create_graph, as the reference runs it)",
            "parallelism": "host cores of rank 0 only (%d threads); the reference has no working multi-GPU path" % r["cores"],
            "step_api": "net(nonmnfld, mnfld) -> DiGSLoss(...) -> loss.backward() (train_surface_reconstruction.py:121-128)",
            "l2": "n/a (CPU)"}


def run_reference_arm(args, rank):
    if rank != 0:
        return
    r = reference_steps(max(args.steps, 5), args.warmup)
    line = {"impl": "reference", "metric": "train points/sec (value+grad+div+backward)", "value": r["value"],
            "unit": "points/s", "n_gpus": args.gpus, "steps": max(args.steps, 5), "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": reference_config(args, r),
            "cpu_baseline": {"value": r["value"], "unit": "points/s", "cores": r["cores"], "kind": r["kind"],
                             "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


CONFIGS = {
    # BASELINE.json configs[1]: SRB recipe (run_surf_recon_exp.sh:19-66)
    "c2": dict(net=NET_KW, loss=LOSS_KW, n_points=N_POINTS, batch=1, box=1.1,
               what="SRB-style recon step (BASELINE configs[1]): 15000 manifold + 15000 off-surface points, "
                    "4x256 sine MLP MFGI init, DiGSLoss siren_wo_n_w_div/l1"),
    # configs[3]: scene recipe (run_scene_recon_exp.sh:20-46): 8x512 SIREN init, identity head
    "c4": dict(net=dict(latent_size=0, in_dim=3, decoder_hidden_dim=512, nl="sine", encoder_type="none",
                        decoder_n_hidden_layers=8, init_type="siren", sphere_init_params=[1.6, 0.1]),
               loss=dict(weights=[3e3, 1e2, 1e2, 5e1, 1e1], loss_type="siren_wo_n_w_div", div_decay="linear", div_type="l1",
                         div_clamp=50), n_points=N_POINTS, batch=1, box=1.1,
               what="scene recon step (BASELINE configs[3]): 15000 + 15000 points, 8x512 sine MLP SIREN init"),
    # configs[4]: DFAUST shapespace (run_shapespace.sh:22-47): latent 256, 8x256, shapes sharded over ranks
    "c5": dict(net=dict(latent_size=256, in_dim=3, decoder_hidden_dim=256, nl="sine", encoder_type="autodecoder",
                        decoder_n_hidden_layers=8, init_type="mfgi", sphere_init_params=[1.6, 0.1]),
               loss=dict(weights=[3e3, 1e2, 1e2, 5e1, 1e2, 1e0], loss_type="siren_wo_n_w_div", div_decay="linear",
                         div_type="l1", div_clamp=50), n_points=N_POINTS, batch=2, box=1.2,
               what="DFAUST shapespace step (BASELINE configs[4]): 2 shapes per GPU x (15000 + 15000) points, "
                    "latent 256 auto-decoder, 8x256 sine MLP MFGI init"),
}


def flops_per_step(cfg, n_m, n_n):
    """Algorithmic GEMM FLOPs: forward C*2*Lh*H^2 per point, training = 3x; C = 2+d off-surface, 1+d surface."""
    nk = cfg["net"]
    Hh, Lh, d = nk["decoder_hidden_dim"], nk["decoder_n_hidden_layers"], nk["in_dim"]
    per = 2 * Lh * Hh * Hh
    return 3 * per * ((2 + d) * n_n + (1 + d) * n_m)


def workload_config(args, world, cfg, n_local):
    par = ("dp%d: points of ONE %d+%d batch split over the ranks (strong scaling), " % (world, cfg["n_points"], cfg["n_points"])
           if args.scaling == "strong" else "dp%d: %d+%d points per GPU (weak scaling), " % (world, n_local, n_local))
    return {"workload": cfg["what"], "config": args.config,
            "points_per_step_per_gpu": 2 * n_local * cfg["batch"], "hidden": cfg["net"]["decoder_hidden_dim"],
            "hidden_layers": cfg["net"]["decoder_n_hidden_layers"], "in_dim": cfg["net"]["in_dim"],
            "precision_mode": args.mode,
            "parallelism": par + "1 in-graph NCCL allreduce of the flat gradient buffer (+8 loss scalars)",
            "step_api": "eager two-call API (DiGSNetwork.forward / DiGSLoss / backward)" if args.no_graph else
                        "digs_b200.GraphedStep: CUDA graph of zero-fill + operand prep + digs_step_fused (counter reset, tile "
                        "kernel, weight-gradient GEMM as a programmatic dependent launch that starts on the SMs the tile "
                        "kernel frees in its last round) + allreduce",
            "l2": "256 MiB scratch written between timed steps (L2 flush); per-step events exclude the flush"}


def traffic_record(mode):
    """dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as fh:
            return json.load(fh).get(mode)
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def build_workload(args, cfg, dev, rank, world, n_batches):
    import digs_b200
    from digs_b200 import dist as ddist, synth
    torch.manual_seed(MODEL_SEED)
    net = digs_b200.DiGSNetwork(**cfg["net"], precision=args.mode).to(dev)
    crit = digs_b200.DiGSLoss(**copy.deepcopy(cfg["loss"]))
    n_full = cfg["n_points"]
    if args.scaling == "strong":
        b, e = ddist.shard_range(n_full, rank, world)
        n_local = e - b
        crit.global_counts = (n_full * cfg["batch"], n_full * cfg["batch"])
    else:
        b, e, n_local = 0, n_full, n_full
        crit.global_counts = (n_full * cfg["batch"] * world, n_full * cfg["batch"] * world)
    cloud, normals = synth.surface_cloud(100000, "torus", seed=0)
    hb = []
    for i in range(n_batches):
        ms, ns, nms = [], [], []
        for s in range(cfg["batch"]):
            seed = ((rank + 1) if args.scaling != "strong" else 1) * 100003 + i * 17 + s
            m, nrm, nm = synth.sample_batch(cloud, normals, n_full, seed=seed, box=cfg["box"])
            ms.append(m[:, b:e]); ns.append(nrm[:, b:e]); nms.append(nm[:, b:e])
        hb.append(tuple(torch.from_numpy(np.ascontiguousarray(np.concatenate(a, 0))).pin_memory() for a in (ms, ns, nms)))
    table = None
    if cfg["net"]["encoder_type"] == "autodecoder":
        g = torch.Generator().manual_seed(1)
        table = (0.01 * torch.randn(cfg["batch"] * world, cfg["net"]["latent_size"], generator=g)).to(dev)
    return net, crit, hb, n_local, table


def time_config(args, cfg, dev, rank, world, steps, warmup, with_e2e=True):
    """Device-timed and end-to-end-timed steps of one workload.  Returns a dict of measurements (max over ranks)."""
    import torch.distributed as dist
    import digs_b200
    from digs_b200 import _lib, dist as ddist
    n_batches = steps + warmup
    net, crit, hb, n_local, table = build_workload(args, cfg, dev, rank, world, n_batches)
    db = [tuple(t.to(dev) for t in b) for b in hb]
    flat = ddist.FlatGrads(list(net.parameters()), extra=8)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    h2d_bytes = sum(t.numel() * 4 for t in hb[0])
    L = _lib.lib()
    bsz = cfg["batch"]
    idx = torch.arange(rank * bsz, (rank + 1) * bsz, device=dev) if table is not None else None
    graphed = None
    if not args.no_graph:
        graphed = digs_b200.GraphedStep(net, crit, n_local, n_local, batch=bsz, flat=flat, latent_table=table)
        # host batches packed like the graph's static inputs (what a data loader would hand over): one H2D copy per step
        packs = [graphed.pack_host_batch(m, nm, nrm) for (m, nrm, nm) in hb]
        hb = [(pk[1], pk[3], pk[2]) for pk in packs]
        packed = [pk[0] for pk in packs]

    def run(batch):
        m, nrm, nm = batch
        if graphed is not None:
            return graphed.step(m, nm, nrm, indices=idx)["loss"]
        m = m.to(dev, non_blocking=True).requires_grad_()
        nm = nm.to(dev, non_blocking=True).requires_grad_()
        nrm = nrm.to(dev, non_blocking=True)
        net.decoder.fc_block.accumulate_grads_in_place = True
        flat.zero_()
        ld, _ = crit(net(nm, m), m, nm, nrm)
        ld["loss"].backward()
        flat.extra[0] = ld["loss"].detach()
        flat.allreduce()
        return flat.extra[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        run(db[i])
    barrier()
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    launches0 = L.digs_launch_count()
    evs = []
    barrier()
    for i in range(steps):
        flush.fill_(i & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(db[warmup + i])
        e1.record()
        evs.append((e0, e1))
    barrier()
    launches = (L.digs_launch_count() - launches0) / max(steps, 1)
    if graphed is not None:
        launches = graphed.launches_per_step      # library kernels recorded in (and replayed from) the graph each step
    ms_dev = sum(a.elapsed_time(b) for a, b in evs) / steps
    ms_e2e, ms_e2e_sync, d2h_bytes = 0.0, 0.0, 4
    E2E_BLOCKS = 3          # the K-step end-to-end loop is a ~20 ms wall-clock interval: repeat it and report the median block
    if with_e2e:
        for i in range(min(warmup, 3)):
            float(run(hb[i]).item())
        blocks = []
        for _ in range(E2E_BLOCKS):
            barrier()
            t0 = time.perf_counter()
            for i in range(steps):
                float(run(hb[warmup + i]).item())      # pinned host batch in, loss scalar out, host blocks on every step
            barrier()
            blocks.append(1e3 * (time.perf_counter() - t0) / steps)
        ms_e2e = ms_e2e_sync = float(np.median(blocks))
        if graphed is not None:
            # the same K steps with the host feed pipelined the way the reference's DataLoader workers feed its loop: the copy
            # of batch i+1 overlaps step i, the loss of step i is read back asynchronously and consumed one step later
            for i in range(min(warmup, 3)):
                graphed.result(graphed.submit(packed[i], indices=idx))
            blocks = []
            for _ in range(E2E_BLOCKS):
                barrier()
                losses, prev = [], None
                t0 = time.perf_counter()
                for i in range(steps):
                    tk = graphed.submit(packed[warmup + i], indices=idx)
                    if prev is not None:
                        losses.append(graphed.result(prev)[0])
                    prev = tk
                losses.append(graphed.result(prev)[0])
                barrier()
                blocks.append(1e3 * (time.perf_counter() - t0) / steps)
                assert len(losses) == steps and all(np.isfinite(v) for v in losses)
            ms_e2e = float(np.median(blocks))
            d2h_bytes = 32
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_dev, ms_e2e, ms_e2e_sync], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return dict(ms_dev=float(t[0]), ms_e2e=float(t[1]), ms_e2e_sync=float(t[2]), d2h_bytes=d2h_bytes, launches=launches,
                clocks=clocks, h2d_bytes=h2d_bytes,
                n_local=n_local, net=net, crit=crit, db=db, flush=flush, graphed=graphed)


def tile_kernel_alone(res, cfg, dev):
    """The dominant kernel (k_tc_step: forward + head + loss + reverse of every tile of both clouds) timed alone with CUDA
    events on its launch stream: DIGS_B200_STEP_NO_WGRAD=1 makes digs_step_fused skip its second launch."""
    import digs_b200
    net, crit, db, flush = res["net"], res["crit"], res["db"], res["flush"]
    if not digs_b200.fused_step_supported(net, crit):
        return None
    m, nrm, nm = db[0]
    os.environ["DIGS_B200_STEP_NO_WGRAD"] = "1"
    try:
        for _ in range(3):
            digs_b200.fused_step(net, crit, m, nm, nrm)
        kev = []
        for r in range(10):
            flush.fill_(r)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            digs_b200.fused_step(net, crit, m, nm, nrm)
            e1.record()
            kev.append((e0, e1))
        torch.cuda.synchronize()
    finally:
        os.environ["DIGS_B200_STEP_NO_WGRAD"] = "0"
    # the events bracket the python wrapper too (output allocation, two tiny fills): take the minimum over repetitions
    return min(a.elapsed_time(b) for a, b in kev)


def accuracy_of_mode(args, dev):
    """max-norm-relative error of the benched precision mode vs the fp64 closed-form oracle on a 2048+2048 sample of the
    benchmark workload (the checker, not the thing measured)."""
    import digs_b200
    from oracle import jet_spec
    cfg = CONFIGS["c2"]
    torch.manual_seed(MODEL_SEED)
    net = digs_b200.DiGSNetwork(**cfg["net"], precision=args.mode).to(dev)
    crit = digs_b200.DiGSLoss(**copy.deepcopy(cfg["loss"]))
    m, nrm, nm = (t[:, :2048].contiguous() for t in host_batches(1, seed=3)[0])
    if digs_b200.fused_step_supported(net, crit):
        ld, out, _ = digs_b200.fused_step(net, crit, m.to(dev), nm.to(dev), nrm.to(dev))
    else:
        mm, nn_ = m.to(dev).requires_grad_(), nm.to(dev).requires_grad_()
        out = net(nn_, mm)
        ld, _ = crit(out, mm, nn_, nrm.to(dev))
        ld["loss"].backward()
    lins = net.decoder.fc_block.linears()
    params = [(l.weight.detach().cpu().double().numpy(), l.bias.detach().cpu().double().numpy()) for l in lins]
    ref = jet_spec.step(params, m[0].double().numpy(), nm[0].double().numpy(), nrm[0].double().numpy(), (1.6, 0.1),
                        cfg["loss"]["weights"])

    def rel(a, b):
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        return float(np.abs(a - b).max() / np.abs(b).max())

    errs = {"f": rel(out["nonmanifold_pnts_pred"].detach().cpu()[0], ref["fn"]["f"]),
            "grad_f": rel(out["nonmanifold_pnts_grad"].detach().cpu()[0], ref["fn"]["grad"]),
            "lap_f": rel(out["nonmanifold_pnts_div"].detach().cpu()[0], ref["fn"]["lap"]),
            "param_grads": {("dW%d" % i): rel(lins[i].weight.grad.cpu(), ref["grads"][i][0]) for i in range(len(lins))}}
    errs["param_grads"].update({("db%d" % i): rel(lins[i].bias.grad.cpu(), ref["grads"][i][1]) for i in range(len(lins))})
    errs["param_grads_worst"] = max(errs["param_grads"].values())
    # the same step through the reference's mechanism in float32 (torch CPU autograd): its distance from fp64 on these very
    # points is the yardstick -- the loss has jump discontinuities (sign f, the divergence clamp), so a point that sits on one
    # flips its O(1/N) seed in ANY finite-precision evaluation, the reference's own included
    from oracle import ref_autograd
    tparams = [p.detach().cpu().clone().requires_grad_(True) for p in net.decoder.fc_block.flat_params()]
    _, g32, j32 = ref_autograd.step(tparams, m[0], nm[0], nrm[0], (1.6, 0.1), cfg["loss"]["weights"])
    errs["reference_fp32_same_points"] = {
        "f": rel(j32["f_n"].numpy(), ref["fn"]["f"]), "grad_f": rel(j32["g_n"].numpy(), ref["fn"]["grad"]),
        "lap_f": rel(j32["lap_n"].numpy(), ref["fn"]["lap"]),
        "param_grads_worst": max(max(rel(g32[2 * i].numpy(), ref["grads"][i][0]), rel(g32[2 * i + 1].numpy(), ref["grads"][i][1]))
                                 for i in range(len(lins)))}
    errs["param_grads_worst_vs_reference_fp32"] = max(
        max(rel(lins[i].weight.grad.cpu(), g32[2 * i].numpy()), rel(lins[i].bias.grad.cpu(), g32[2 * i + 1].numpy()))
        for i in range(len(lins)))
    errs["what"] = "max-norm-relative error vs the fp64 oracle, 2048+2048 points of the benchmark workload, mode %s" % args.mode
    return errs


def whole_iteration_leg(args, dev, iters=200):
    """train_surface_reconstruction.py:114-158 as one CUDA-graph replay per iteration with zero host synchronisations:
    DeviceSampler draws the batch from the HBM-resident cloud, digs_step_fused, global-norm clip + Adam (step count on the
    device); the host only advances the div-weight schedule.  Timed over `iters` back-to-back iterations, one sync at the end."""
    import digs_b200
    from digs_b200 import synth
    cfg = CONFIGS["c2"]
    torch.manual_seed(MODEL_SEED)
    net = digs_b200.DiGSNetwork(**cfg["net"], precision=args.mode).to(dev)
    crit = digs_b200.DiGSLoss(**copy.deepcopy(cfg["loss"]))
    cloud, normals = synth.surface_cloud(100000, "torus", seed=0)
    sampler = digs_b200.DeviceSampler(torch.tensor(cloud, device=dev), torch.tensor(normals, device=dev), cfg["n_points"],
                                      grid_range=cfg["box"], seed=0, device=dev)
    opt = digs_b200.FusedClipAdam(net.parameters(), lr=5e-5, max_norm=10.0, device_step=True)
    it = digs_b200.GraphedStep(net, crit, cfg["n_points"], cfg["n_points"], sampler=sampler, optimizer=opt)
    for i in range(10):
        it.step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(iters):
        ld = it.step()
        crit.update_div_weight(i, iters, (1e2, 0.2, 1e2, 0.4, 0.0, 0.0))
    e1.record()
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1) / iters
    out = {"ms_per_iteration": ms, "value": 2 * cfg["n_points"] / (ms * 1e-3), "unit": "points/s",
           "wall_ms_per_iteration": 1e3 * wall / iters, "iterations": iters, "gpu_launches": it.launches_per_step,
           "final_loss": float(ld["loss"]),
           "what": "sampling + forward + DiGSLoss + backward + clip_grad_norm_(10) + Adam as one CUDA graph replay per "
                   "iteration; no host synchronisation inside the loop, div-weight schedule advanced on the host"}
    it.release()
    return out


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    import digs_b200

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    cfg = CONFIGS[args.config]
    res = time_config(args, cfg, dev, rank, world, args.steps, args.warmup)
    ms_dev, ms_e2e, n_local = res["ms_dev"], res["ms_e2e"], res["n_local"]
    net = res["net"]

    ms_tile = tile_kernel_alone(res, cfg, dev) if rank == 0 else None

    # ---- second half of the metric: MC grid queries/s, grid_res^3 points sharded by iy slab over the ranks -----
    ms_grid, ms_grid_bf16, grid_pts, mesh_info = 0.0, 0.0, 0, None
    if args.grid_res > 0 and args.config == "c2":
        axes, _, _ = digs_b200.grid_axes(args.grid_res, np.array([[-1, 1], [-1, 1], [-1, 1]], dtype=float))
        ny = len(axes[1])
        slab = digs_b200.slab_range(ny, rank, world)
        out_vol = torch.empty((slab[1] - slab[0], len(axes[0]), len(axes[2])), dtype=torch.float32, device=dev)
        small = digs_b200.slab_range(ny, 0, 16 * world)
        digs_b200.grid_query(net, axes, iy_range=small)                      # warm-up
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        digs_b200.grid_query(net, axes, iy_range=slab, out=out_vol)
        e1.record()
        torch.cuda.synchronize()
        # the same query in the single-pass bf16 mode (admissible for mesh extraction: Chamfer kept to 3 s.f.)
        digs_b200.grid_query(net, axes, iy_range=small, precision="bf16")
        torch.cuda.synchronize()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        digs_b200.grid_query(net, axes, iy_range=slab, out=out_vol, precision="bf16")
        b1.record()
        torch.cuda.synchronize()
        tg = torch.tensor([e0.elapsed_time(e1), b0.elapsed_time(b1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        ms_grid, ms_grid_bf16 = float(tg[0]), float(tg[1])
        grid_pts = len(axes[0]) * ny * len(axes[2])
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
        del out_vol

    if rank == 0:
        peaks = measured_peaks()
        pts_rank = 2 * n_local * cfg["batch"]
        pts = pts_rank * world if args.scaling == "weak" else 2 * cfg["n_points"] * cfg["batch"]
        nk = cfg["net"]
        Hh, Lh, d = nk["decoder_hidden_dim"], nk["decoder_n_hidden_layers"], nk["in_dim"]
        flop_rank = flops_per_step(cfg, n_local * cfg["batch"], n_local * cfg["batch"])
        step_tflops = flop_rank / (ms_dev * 1e-3) / 1e12
        line = {"metric": "train points/sec (value+grad+div+backward)", "value": pts / (ms_dev * 1e-3),
                "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": {"fp32": "f32", "bf16x2": "bf16x2-split (f32 accumulate)", "bf16": "bf16"}[args.mode],
                "data": "synthetic", "config": workload_config(args, world, cfg, n_local),
                "e2e": {"value": pts / (ms_e2e * 1e-3), "unit": "points/s", "h2d_bytes_per_step": res["h2d_bytes"],
                        "d2h_bytes_per_step": res["d2h_bytes"], "ms_per_step": ms_e2e,
                        "host_feed": "GraphedStep.submit / result: one graph launch per step whose first node copies the pinned "
                                     "host batch to the device and whose last node copies the 8 loss scalars back; two such "
                                     "graphs alternate over two staging buffers, the loss is consumed one step later",
                        "timing": "wall clock over the K steps, median of 3 such blocks",
                        "blocking_ms_per_step": res["ms_e2e_sync"],
                        "blocking_value": pts / (res["ms_e2e_sync"] * 1e-3) if res["ms_e2e_sync"] > 0 else None},
                "gpu_launches": res["launches"], "clocks": res["clocks"]}
        # algorithmic bytes of a step: 12 B in + 20 B out per off-surface point, 12 + 16 per surface point,
        # 12 B of normals, the parameters read once and their gradients written once
        n_par = sum(p.numel() for p in net.parameters())
        alg_bytes = n_local * cfg["batch"] * (32 + 28 + 12) + 2 * 4 * n_par
        tr = traffic_record(args.mode)
        if ms_tile is not None:
            tile_flops = flop_rank * 2.0 / 3.0          # forward + dgrad of both clouds; the wgrad third is the GEMM launch
            tile_tf = tile_flops / (ms_tile * 1e-3) / 1e12
            line["roofline"] = {
                "bound": "tensor", "achieved": tile_tf, "peak": peaks["burst"], "unit": "TFLOP/s", "frac": tile_tf / peaks["burst"],
                "traffic": None if tr is None else tr.get("k_tc_step_dram_bytes"),
                "kernel": "k_tc_step: forward jet + head + loss + reverse jet of every point tile of both clouds, one "
                          "persistent launch, timed alone (digs_step_fused without its weight-gradient GEMM)",
                "ms": ms_tile, "algorithmic_gflop": tile_flops / 1e9, "peak_source": peaks["source"] + " bf16 dense burst",
                "mma_passes_per_product": {"bf16x2": 3, "bf16": 1}.get(args.mode),
                "algorithmic_bytes_per_step": alg_bytes,
                "traffic_over_algorithmic_bytes": None if tr is None or not tr.get("step_dram_bytes") else tr["step_dram_bytes"] / alg_bytes,
                "traffic_source": None if tr is None else tr.get("source"),
                "whole_step": {"achieved": step_tflops, "peak": peaks["sustained"], "frac": step_tflops / peaks["sustained"],
                               "algorithmic_gflop": flop_rank / 1e9, "dram_bytes": None if tr is None else tr.get("step_dram_bytes"),
                               "peak_source": peaks["source"] + " bf16 dense sustained (per GPU)"}}
        else:
            line["roofline"] = {"bound": "tensor", "achieved": step_tflops, "peak": peaks["sustained"], "unit": "TFLOP/s",
                                "frac": step_tflops / peaks["sustained"], "traffic": None,
                                "kernel": "whole step (precision mode without a fused tile kernel)",
                                "algorithmic_gflop": flop_rank / 1e9}
        if grid_pts:
            flop_pt = 2 * Lh * Hh * Hh + 2 * d * Hh + 2 * Hh
            gt = grid_pts * flop_pt / (ms_grid * 1e-3) / 1e12
            line["grid_query"] = {"metric": "MC grid queries/sec", "value": grid_pts / (ms_grid * 1e-3), "unit": "queries/s",
                                  "grid": "%d^3, bbox [-1,1]^3 + 0.1, fp16-rounded axes, iy slabs over %d rank(s)" % (args.grid_res, world),
                                  "ms": ms_grid, "output": "fp32 volume resident in HBM",
                                  "roofline": {"bound": "tensor", "achieved": gt, "peak": peaks["sustained"] * world,
                                               "unit": "TFLOP/s", "frac": gt / (peaks["sustained"] * world),
                                               "peak_source": "%d x measured bf16 dense sustained" % world,
                                               "algorithmic_mflop_per_query": flop_pt / 1e6}}
            line["grid_query"]["single_pass_bf16"] = {
                "ms": ms_grid_bf16, "value": grid_pts / (ms_grid_bf16 * 1e-3), "unit": "queries/s",
                "frac_of_sustained_peak": grid_pts * flop_pt / (ms_grid_bf16 * 1e-3) / 1e12 / (peaks["sustained"] * world),
                "note": "same kernel, hi*hi MMA pass only; Chamfer of the extracted surface moves by 1.8e-4 relative on the "
                        "trained-weights fixture at 256^3 (bf16x2: 1.2e-5; the bound is 5e-4 = 3 significant figures)"}
            if mesh_info is not None:
                line["grid_query"]["mesh_extraction"] = mesh_info
    del res
    torch.cuda.empty_cache()

    # ---- extra legs (N = 1 only, a few steps each): other precision mode and the other single-GPU configurations ------
    extra = {}
    if world == 1 and not args.no_extra and args.config == "c2":
        for name, cfg_name, mode in (("bf16_single_pass", "c2", "bf16"), ("c4_scene_8x512", "c4", args.mode),
                                     ("c5_shapespace", "c5", args.mode)):
            a2 = copy.copy(args)
            a2.mode, a2.config = mode, cfg_name
            try:
                r2 = time_config(a2, CONFIGS[cfg_name], dev, rank, world, 8, 3, with_e2e=False)
                c2 = CONFIGS[cfg_name]
                p2 = 2 * r2["n_local"] * c2["batch"]
                fl = flops_per_step(c2, r2["n_local"] * c2["batch"], r2["n_local"] * c2["batch"])
                extra[name] = {"value": p2 / (r2["ms_dev"] * 1e-3), "unit": "points/s", "ms_per_step": r2["ms_dev"],
                               "precision_mode": mode, "workload": c2["what"], "gpu_launches": r2["launches"],
                               "tflops_algorithmic": fl / (r2["ms_dev"] * 1e-3) / 1e12,
                               "frac_of_sustained_peak": fl / (r2["ms_dev"] * 1e-3) / 1e12 / measured_peaks()["sustained"]}
                del r2
            except Exception as ex:       # an extra leg must never take the headline line down
                extra[name] = {"error": "%s: %s" % (type(ex).__name__, ex)}
            torch.cuda.empty_cache()
    # ---- the complete training iteration as ONE graph: on-device batch sampling -> fused step -> clip + Adam ---
    if world == 1 and not args.no_extra and args.config == "c2":
        try:
            extra["whole_iteration_one_graph"] = whole_iteration_leg(args, dev)
        except Exception as ex:
            extra["whole_iteration_one_graph"] = {"error": "%s: %s" % (type(ex).__name__, ex)}
        torch.cuda.empty_cache()
    if rank == 0:
        if extra:
            line["extra_legs"] = extra
        if world == 1 and not args.no_cpu:
            r = reference_steps(5, 1, budget_s=30.0)
            line["cpu_baseline"] = {"value": r["value"], "unit": "points/s", "cores": r["cores"], "kind": r["kind"],
                                    "sample": r["sample"], "ms_per_step": r["ms_per_step"]}
            try:
                line["accuracy"] = accuracy_of_mode(args, dev)
            except Exception as ex:
                line["accuracy"] = {"error": "%s: %s" % (type(ex).__name__, ex)}
            if not args.no_ref_cuda:
                try:
                    rc = reference_steps(10, 3, device="cuda:%d" % local_rank)
                    line["reference_on_b200"] = {"value": rc["value"], "unit": "points/s", "ms_per_step": rc["ms_per_step"],
                                                 "kind": rc["kind"], "sample": rc["sample"],
                                                 "note": "context only: the reference through stock PyTorch CUDA (cuBLAS "
                                                         "sgemm + elementwise kernels + autograd engine) on the same GPU"}
                except Exception as ex:
                    line["reference_on_b200"] = {"error": "%s: %s" % (type(ex).__name__, ex)}
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
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the reference-on-B200 (stock PyTorch CUDA) context leg")
    ap.add_argument("--grid-res", type=int, default=512, help="resolution of the grid-query leg (0 disables it)")
    ap.add_argument("--no-graph", action="store_true", help="eager autograd step instead of digs_b200.GraphedStep")
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="workload (BASELINE.json configs; the metric is quoted on c2)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: a full batch per GPU; strong: ONE batch's points split over the ranks")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra legs (bf16 single pass, c4, c5) of the N=1 run")
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

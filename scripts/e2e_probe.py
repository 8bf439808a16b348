"""Where the end-to-end step time goes beyond the device-timed graph replay (C2, bf16x2): python scripts/e2e_probe.py"""
import copy
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import digs_b200  # noqa: E402
import bench  # noqa: E402

cfg = bench.CONFIGS["c2"]
dev = torch.device("cuda:0")
torch.manual_seed(bench.MODEL_SEED)
net = digs_b200.DiGSNetwork(**cfg["net"], precision="bf16x2").to(dev)
crit = digs_b200.DiGSLoss(**copy.deepcopy(cfg["loss"]))
n = cfg["n_points"]
gs = digs_b200.GraphedStep(net, crit, n, n)
hb = bench.host_batches(24, seed=7)
packs = [gs.pack_host_batch(m, nm, nrm) for (m, nrm, nm) in hb]
K = 20


def timed(name, body, sync_each=False):
    for i in range(3):
        body(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        body(i)
    torch.cuda.synchronize()
    print("%-58s %.4f ms/step" % (name, 1e3 * (time.perf_counter() - t0) / K), flush=True)


timed("A graph.replay() back to back", lambda i: gs.graph.replay())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for i in range(K):
    gs.graph.replay()
e1.record()
torch.cuda.synchronize()
print("%-58s %.4f ms/step" % ("A' same, CUDA events around the K replays", e0.elapsed_time(e1) / K))
stage = torch.empty_like(gs._inputs)


def body_b(i):
    gs._inputs.copy_(stage, non_blocking=True)
    gs.graph.replay()


timed("B D2D input copy + replay", body_b)
lh = torch.zeros(8).pin_memory()


def body_c(i):
    gs._inputs.copy_(stage, non_blocking=True)
    gs.graph.replay()
    lh.copy_(gs.flat.extra[:8], non_blocking=True)


timed("C B + async 32-byte loss read-back", body_c)


def body_d(i):
    gs._inputs.copy_(packs[i % 20][0].view(-1), non_blocking=True)
    gs.graph.replay()
    lh.copy_(gs.flat.extra[:8], non_blocking=True)


timed("D H2D on the compute stream + replay + read-back, no sync", body_d)
cs = torch.cuda.Stream()


def body_e(i):
    with torch.cuda.stream(cs):
        stage.copy_(packs[i % 20][0].view(-1), non_blocking=True)
    gs.graph.replay()


timed("E H2D on a copy stream, unsynchronised with the replays", body_e)
prev = [None]


def body_f(i):
    tk = gs.submit(packs[i % 20][0])
    if prev[0] is not None:
        gs.result(prev[0])
    prev[0] = tk


timed("F submit / result, one step of lag", body_f)
t0 = time.perf_counter()
for i in range(200):
    gs.submit(packs[i % 20][0])
    if i % 3 == 2:
        torch.cuda.synchronize()
print("host cost of submit(): about %.1f us" % (1e6 * (time.perf_counter() - t0) / 200 - 1e3 * 0.9 / 1))
t0 = time.perf_counter()
for i in range(50):
    float(gs.step(packs[i % 20][1], packs[i % 20][2], packs[i % 20][3])["loss"].item())
print("%-58s %.4f ms/step" % ("G blocking step() + .item()", 1e3 * (time.perf_counter() - t0) / 50))

# ---- GPU-side anatomy of the pipelined feed: events around every I/O-graph launch
torch.cuda.synchronize()
evs = []
prev = None
host_submit = []
for i in range(K):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    t0 = time.perf_counter()
    tk = gs.submit(packs[i % 20][0])
    host_submit.append(time.perf_counter() - t0)
    b.record()
    evs.append((a, b))
    if prev is not None:
        gs.result(prev)
    prev = tk
gs.result(prev)
torch.cuda.synchronize()
busy = [a.elapsed_time(b) for a, b in evs]
gaps = [evs[i][1].elapsed_time(evs[i + 1][0]) for i in range(K - 1)]
print("I/O graph on the device: mean %.4f ms (min %.4f); gap to the next launch: mean %.4f ms (max %.4f); host time in submit(): mean %.1f us"
      % (sum(busy) / K, min(busy), sum(gaps) / (K - 1), max(gaps), 1e6 * sum(host_submit) / K))
# the same with the plain graph (no copies inside)
evs = []
for i in range(K):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    gs.graph.replay()
    b.record()
    evs.append((a, b))
torch.cuda.synchronize()
busy = [a.elapsed_time(b) for a, b in evs]
gaps = [evs[i][1].elapsed_time(evs[i + 1][0]) for i in range(K - 1)]
print("plain graph on the device: mean %.4f ms (min %.4f); gap: mean %.4f ms" % (sum(busy) / K, min(busy), sum(gaps) / (K - 1)))

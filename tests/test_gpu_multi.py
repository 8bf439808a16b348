"""Multi-GPU (NCCL) parity: point-sharded data-parallel step == the single-GPU step.  Needs >= 2 GPUs (skipped otherwise)."""
import copy
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import copy, faulthandler, os, sys
faulthandler.dump_traceback_later(int(os.environ.get("DIGS_TEST_HANG_S", "150")), exit=True)   # a hung rank prints where it is stuck
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
def stage(msg):
    print("[rank %s] %s" % (os.environ["RANK"], msg), flush=True)
import digs_b200
from digs_b200 import dist as dd
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", rank)
torch.cuda.set_device(dev)
stage("init_process_group")
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
stage("initialised")
kw = dict(latent_size=0, in_dim=3, decoder_hidden_dim=256, nl="sine", encoder_type="none", decoder_n_hidden_layers=4,
          init_type="mfgi", sphere_init_params=[1.6, 0.1])
lk = dict(weights=[3e3, 1e2, 1e2, 5e1, 1e2], loss_type="siren_wo_n_w_div", div_decay="linear", div_type="l1", div_clamp=50)
rng = np.random.RandomState(0)
N = 4096
v = rng.normal(size=(1, N, 3)); v /= np.linalg.norm(v, axis=-1, keepdims=True)
m = torch.tensor(0.5 * v, dtype=torch.float32, device=dev)
nrm = torch.tensor(v, dtype=torch.float32, device=dev)
nm = torch.tensor(rng.uniform(-1.1, 1.1, size=(1, N, 3)), dtype=torch.float32, device=dev)
for precision in ("fp32", "bf16x2"):
    def run(sharded):
        torch.manual_seed(5)
        net = digs_b200.DiGSNetwork(**kw, precision=precision).to(dev)
        crit = digs_b200.DiGSLoss(**copy.deepcopy(lk))
        flat = dd.FlatGrads(list(net.parameters()), extra=8)
        net.decoder.fc_block.accumulate_grads_in_place = True
        if sharded:
            crit.global_counts = (N, N)
            mm, nn_, rr = (dd.shard_points(t, rank, world) for t in (m, nm, nrm))
        else:
            mm, nn_, rr = m, nm, nrm
        mm = mm.contiguous().requires_grad_(); nn_ = nn_.contiguous().requires_grad_()
        ld, _ = crit(net(nn_, mm), mm, nn_, rr.contiguous())
        ld["loss"].backward()
        flat.extra[0] = ld["loss"].detach()
        if sharded:
            flat.allreduce()
        return flat.flat.clone()
    stage("two-call step, " + precision)
    full, shard = run(False), run(True)
    err = float((full - shard).abs().max() / full.abs().max())
    assert err < 2e-5, (precision, err)
    # slab-sharded grid query: rank slabs tile the volume bit-exactly
    torch.manual_seed(5)
    net = digs_b200.DiGSNetwork(**kw, precision=precision).to(dev)
    axes, _, _ = digs_b200.grid_axes(32, np.array([[-1, 1], [-1, 1], [-1, 1]], dtype=float))
    vol = digs_b200.grid_query(net, axes)
    b, e = digs_b200.slab_range(32, rank, world)
    mine = digs_b200.grid_query(net, axes, iy_range=(b, e))
    assert torch.equal(mine, vol[b:e])

stage("graphed fused iteration")
# ---- whole iteration in ONE CUDA graph with the NCCL all-reduce captured inside: points of one batch split over the ranks
#      (strong scaling) == the same iteration on one GPU
from digs_b200.step import fused_step
def flat_of(net):
    return torch.cat([p.grad.reshape(-1) for p in net.parameters()])
torch.manual_seed(5)
ref_net = digs_b200.DiGSNetwork(**kw, precision="bf16x2").to(dev)
ref_crit = digs_b200.DiGSLoss(**copy.deepcopy(lk))
ld_ref, _, _ = fused_step(ref_net, ref_crit, m, nm, nrm)
ref_flat = flat_of(ref_net).clone()
torch.manual_seed(5)
net = digs_b200.DiGSNetwork(**kw, precision="bf16x2").to(dev)
crit = digs_b200.DiGSLoss(**copy.deepcopy(lk))
crit.global_counts = (N, N)
b, e = dd.shard_range(N, rank, world)
gs = digs_b200.GraphedStep(net, crit, e - b, e - b)
out = gs.step(m[:, b:e].contiguous(), nm[:, b:e].contiguous(), nrm[:, b:e].contiguous())
torch.cuda.synchronize()
err = float((gs.flat.flat[:ref_flat.numel()] - ref_flat).abs().max() / ref_flat.abs().max())
assert err < 2e-5, ("graphed sharded step", err)
assert abs(float(out["loss"]) - float(ld_ref["loss"])) < 1e-5 * abs(float(ld_ref["loss"])), (float(out["loss"]), float(ld_ref["loss"]))
# a second replay must give the same answer (buffers are re-zeroed inside the graph, the collective replays)
out = gs.step(m[:, b:e].contiguous(), nm[:, b:e].contiguous(), nrm[:, b:e].contiguous())
torch.cuda.synchronize()
err = float((gs.flat.flat[:ref_flat.numel()] - ref_flat).abs().max() / ref_flat.abs().max())
assert err < 2e-5, ("graphed sharded step, replay 2", err)

stage("shape-sharded auto-decoder")
# ---- DFAUST-style auto-decoder (SURVEY 8e row 3): shapes of the batch sharded over the ranks == all shapes on one GPU
kw5 = dict(latent_size=32, in_dim=3, decoder_hidden_dim=128, nl="sine", encoder_type="autodecoder", decoder_n_hidden_layers=3,
           init_type="mfgi", sphere_init_params=[1.6, 0.1])
lk5 = dict(weights=[3e3, 1e2, 1e2, 5e1, 1e2, 1e0], loss_type="siren_wo_n_w_div", div_decay="linear", div_type="l1", div_clamp=50)
S, per, Np = 12, 2, 777                      # table of 12 shapes; every rank trains `per` shapes of the batch
g = torch.Generator().manual_seed(3)
table = (0.05 * torch.randn(S, 32, generator=g)).to(dev)
batch_idx = torch.tensor([(3 * i + 1) % S for i in range(per * world)], device=dev)
rs = np.random.RandomState(9)
vm = rs.normal(size=(per * world, Np, 3)); vm /= np.linalg.norm(vm, axis=-1, keepdims=True)
m5 = torch.tensor(0.5 * vm, dtype=torch.float32, device=dev)
n5 = torch.tensor(vm, dtype=torch.float32, device=dev)
nm5 = torch.tensor(rs.uniform(-1.2, 1.2, size=(per * world, Np, 3)), dtype=torch.float32, device=dev)
torch.manual_seed(7)
ref_net = digs_b200.DiGSNetwork(**kw5, precision="bf16x2").to(dev)
ref_crit = digs_b200.DiGSLoss(**copy.deepcopy(lk5))
ld_ref, _, lat_grad_ref = fused_step(ref_net, ref_crit, m5, nm5, n5, latent=table[batch_idx])
ref_flat = flat_of(ref_net).clone()
ref_table_grad = torch.zeros_like(table).index_add_(0, batch_idx, lat_grad_ref)
torch.manual_seed(7)
net = digs_b200.DiGSNetwork(**kw5, precision="bf16x2").to(dev)
crit = digs_b200.DiGSLoss(**copy.deepcopy(lk5))
crit.global_counts = (per * world * Np, per * world * Np)
mine = dd.shard_shapes(batch_idx, rank, world)
lo = rank * per
gs = digs_b200.GraphedStep(net, crit, Np, Np, batch=per, latent_table=table)
out = gs.step(m5[lo:lo + per].contiguous(), nm5[lo:lo + per].contiguous(), n5[lo:lo + per].contiguous(), indices=mine)
torch.cuda.synchronize()
err = float((gs.flat.flat[:ref_flat.numel()] - ref_flat).abs().max() / ref_flat.abs().max())
assert err < 2e-5, ("shape-sharded decoder gradients", err)
err = float((gs.latent_grad - ref_table_grad).abs().max() / ref_table_grad.abs().max())
assert err < 2e-5, ("shape-sharded latent-table gradient", err)
assert abs(float(out["loss"]) - float(ld_ref["loss"])) < 1e-5 * abs(float(ld_ref["loss"]))
assert abs(float(out["latent_reg_term"]) - float(ld_ref["latent_reg_term"])) < 1e-5 * abs(float(ld_ref["latent_reg_term"]))
stage("all checks passed")
# CUDA graphs that captured NCCL kernels must be released before the communicator is torn down (destroy hangs otherwise)
del gs, out
import gc
gc.collect()
torch.cuda.synchronize()
dist.destroy_process_group()
print("ok", rank)
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_point_sharded_step_equals_single_gpu(tmp_path):
    """Point-sharded two-call step, the graphed fused iteration with the all-reduce inside the graph, the slab-sharded grid
    query and the shape-sharded auto-decoder step -- each equal to its single-GPU run.  World size: DIGS_TEST_WORLD (default 2)."""
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    world = min(torch.cuda.device_count(), int(os.environ.get("DIGS_TEST_WORLD", "2")))
    procs = []
    logs = [tmp_path / ("rank%d.log" % r) for r in range(world)]
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT="29741")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=open(logs[r], "w"),
                                      stderr=subprocess.STDOUT))
    # a rank that fails leaves its peers waiting inside a collective: stop everybody as soon as one exits non-zero
    import time
    deadline = time.time() + 200
    failed = None
    while time.time() < deadline:
        codes = [p.poll() for p in procs]
        if any(c not in (None, 0) for c in codes):
            failed = [i for i, c in enumerate(codes) if c not in (None, 0)][0]
            break
        if all(c == 0 for c in codes):
            break
        time.sleep(0.5)
    else:
        failed = -1
    for p in procs:
        if p.poll() is None:
            p.kill()
    if failed is not None:
        text = "\n".join("---- rank %d ----\n%s" % (r, logs[r].read_text()[-3000:]) for r in range(world))
        raise AssertionError(("timeout" if failed < 0 else "rank %d failed" % failed) + "\n" + text)

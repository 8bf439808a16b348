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
import copy, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import digs_b200
from digs_b200 import dist as dd
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", rank)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
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
dist.destroy_process_group()
print("ok", rank)
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_point_sharded_step_equals_single_gpu(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    world = 2
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT="29741")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=600)
        assert p.returncode == 0, out.decode()[-3000:]

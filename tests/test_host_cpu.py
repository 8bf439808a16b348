"""CPU: host-side logic of the drop-in layer and the C-ABI surface (no compute calls -- there is no GPU here)."""
import copy
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import digs_b200
from digs_b200 import _lib, dist as ddist
from digs_b200.grid import grid_axes, slab_range
from oracle.make_golden import INIT_CASES, INIT_SEED, LOSSVAR_CASES
from _util import golden, sha, build_net, head_of, maxrel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    L = _lib.lib()
    assert L.digs_abi_version() == 2
    header = open(os.path.join(ROOT, "include", "digs_b200.h")).read()
    declared = set(re.findall(r"\b(digs_[a-z0-9_]+)\s*\(", header))
    assert declared, "no entry points found in the header"
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name


def test_library_rejects_bad_arguments_without_a_gpu():
    L = _lib.lib()
    net = _lib.DigsNet()
    net.d, net.H, net.Lh, net.head = 5, 256, 4, 0
    assert L.digs_saved_bytes(ctypes.byref(net), 10, 1, 0) == 0
    assert b"net.d" in L.digs_last_error()
    rc = L.digs_jet_forward(ctypes.byref(net), None, 10, 3, None, 0, 1, None, None, None, None, 0, None, 0, 0, None)
    assert rc == -1
    net.d = 3
    for l in range(6):
        net.W[l] = 256
        net.b[l] = 256
    net.w0_stride = 3
    # tensor-core modes are explicit: never a silent fallback
    assert L.digs_saved_bytes(ctypes.byref(net), 10, 1, 7) == 0
    assert L.digs_saved_bytes(ctypes.byref(net), 1000, 1, 0) >= 4 * 5 * 1000 * 256 * 4
    # mesh / nearest-neighbour entry points validate before touching the device
    assert L.digs_mt_classify(None, 8, 8, 8, 0.0, None, None, None) == -1
    assert b"NULL" in L.digs_last_error()
    assert L.digs_mt_classify(256, 1, 8, 8, 0.0, 256, 256, None) == -1            # a volume needs 2 points per axis
    ax = (ctypes.c_int * 3)(0, 0, 2)
    f3 = (ctypes.c_float * 3)(1.0, 1.0, 1.0)
    assert L.digs_mt_emit(256, 8, 8, 8, 0.0, ax, f3, f3, 256, 256, 256, 256, 256, None, 256, None) == -1
    assert b"permutation" in L.digs_last_error()
    assert L.digs_nn_query(256, 10, 256, 10, 4, 256, None, None) == -1             # d must be 2 or 3
    assert L.digs_nn_query(256, 0, 256, 10, 3, 256, None, None) == -1              # empty reference cloud
    assert L.digs_nn_query(None, 10, None, 0, 3, None, None, None) == 0            # nothing to query


@pytest.mark.parametrize("name", sorted(INIT_CASES))
def test_init_bitwise_equal_to_reference(name):
    g = golden("init_fingerprints.npz")
    net = build_net(INIT_CASES[name], INIT_SEED)
    keys = [k for k in g.files if k.startswith(name + "/") and not k.endswith("/head")]
    assert len(keys) == len(list(net.named_parameters()))
    for pname, p in net.named_parameters():
        assert sha(p) == str(g["%s/%s" % (name, pname)]), pname
        np.testing.assert_array_equal(p.detach().reshape(-1)[:8].numpy(), g["%s/%s/head" % (name, pname)])


def test_state_dict_keys_and_interface():
    net = build_net(INIT_CASES["mfgi_c2"], 0)
    keys = list(net.state_dict().keys())
    assert keys == ["decoder.fc_block.net.%d.0.%s" % (i, w) for i in range(6) for w in ("weight", "bias")]
    assert net.decoder.fc_block.net[2][0].weight.shape == (256, 256)
    with pytest.raises(ValueError):
        digs_b200.DiGSNetwork(0, encoder_type="vae")
    with pytest.raises(NotImplementedError):
        digs_b200.DiGSNetwork(0, encoder_type="none", nl="relu")


def test_cpu_tensors_fail_loudly():
    net = build_net(INIT_CASES["geo"], 0)
    pts = torch.zeros(1, 8, 3, requires_grad=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(pts, pts)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net.decoder(torch.zeros(8, 3))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "digs_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no oracle", ""), fn
            assert "/root/reference" not in src, fn


@pytest.mark.parametrize("decay", ["linear", "quintic", "step", "none"])
@pytest.mark.parametrize("pname,params", [("recipe", (1e2, 0.2, 1e2, 0.4, 0.0, 0.0)),
                                          ("scene", (1e1, 0.1, 1e1, 0.3, 0.0, 0.0)), ("two", (50.0, 5.0))])
def test_div_weight_schedule(decay, pname, params):
    g = golden("div_weight.npz")
    crit = digs_b200.DiGSLoss(weights=[3e3, 1e2, 1e2, 5e1, 1e2], loss_type="siren_wo_n_w_div", div_decay=decay,
                              div_type="l1")
    vals = []
    for it in range(0, 1001, 25):
        crit.update_div_weight(it, 1000, params)
        vals.append(crit.weights[4])
    np.testing.assert_allclose(np.array(vals), g["%s/%s" % (decay, pname)], rtol=0, atol=0)


def test_loss_rejects_unknown_modes_and_missing_inputs():
    z = torch.zeros(1, 4, 3)
    pred = {"nonmanifold_pnts_pred": z, "manifold_pnts_pred": z, "latent_reg": None, "latent": None}
    with pytest.raises(Warning):
        digs_b200.DiGSLoss(loss_type="bogus")(pred, z, z, z)
    with pytest.raises(Warning):
        digs_b200.DiGSLoss(loss_type="siren_w_div", div_type="huber")(pred, z, z, z)
    with pytest.raises(RuntimeError):                     # no jet in output_pred: it did not come from the CUDA forward
        digs_b200.DiGSLoss(loss_type="siren_w_curv")(pred, z, z, z, curvatures=z[..., :2])
    jets = dict(pred, nonmanifold_pnts_grad=z, manifold_pnts_grad=z)
    with pytest.raises(Warning):                          # models/losses.py:80-81
        digs_b200.DiGSLoss(loss_type="siren_w_curv", div_type="l1")(jets, z, z, z)
    with pytest.raises(RuntimeError):                     # surface Laplacian was not requested from the network
        digs_b200.DiGSLoss(loss_type="siren_w_curv", div_type="l1")(jets, z, z, z, curvatures=z[..., :2])


@pytest.mark.parametrize("name", sorted(LOSSVAR_CASES))
def test_composite_loss_variants_match_reference_fixture_cpu(name):
    """The loss algebra of the curvature / IGR / gt-divergence variants (digs_b200.DiGSLoss._composite_terms) against
    fixtures from the real reference, with the jets supplied by the autograd oracle in float64 (so only the loss
    algebra and its gradient flow are under test here; the CUDA jets are covered by the GPU twin of this test)."""
    from oracle import ref_autograd
    case, g = LOSSVAR_CASES[name], golden("lossvar_%s.npz" % name)
    net = build_net(case["net"], case["seed"]).double()
    params = list(net.decoder.fc_block.flat_params())
    bs, d = case["bs"], case["net"]["in_dim"]
    m = torch.tensor(g["in_mnfld"], dtype=torch.float64).reshape(-1, d).requires_grad_()
    nm = torch.tensor(g["in_nonmnfld"], dtype=torch.float64).reshape(-1, d).requires_grad_()
    f_m, g_m, l_m = ref_autograd.jet(params, m, head_of(net), want_lap=True)
    f_n, g_n, l_n = ref_autograd.jet(params, nm, head_of(net), want_lap=True)
    pred = {"manifold_pnts_pred": f_m.reshape(bs, -1), "nonmanifold_pnts_pred": f_n.reshape(bs, -1), "latent_reg": None,
            "latent": None, "manifold_pnts_grad": g_m.reshape(bs, -1, d), "nonmanifold_pnts_grad": g_n.reshape(bs, -1, d),
            "manifold_pnts_div": l_m.reshape(bs, -1), "nonmanifold_pnts_div": l_n.reshape(bs, -1)}
    crit = digs_b200.DiGSLoss(**copy.deepcopy(case["loss"]))
    t64 = lambda k: torch.tensor(g[k], dtype=torch.float64)
    loss_dict, mnfld_grad = crit(pred, m.reshape(bs, -1, d), nm.reshape(bs, -1, d), t64("in_normals"),
                                 curvatures=t64("in_curvatures"), nonmnfld_div_gt=t64("in_div_gt_n"),
                                 mnfld_div_gt=t64("in_div_gt_m"))
    loss_dict["loss"].backward()
    for key in ("loss", "sdf_term", "inter_term", "eikonal_term", "normals_loss", "div_loss", "curv_loss", "latent_reg_term"):
        ref = float(g["ref64_term_" + key][0])
        assert abs(float(loss_dict[key].reshape(-1)[0]) - ref) <= 1e-9 * max(abs(ref), 1.0), key
    np.testing.assert_array_equal(np.array(crit.weights, dtype=np.float64), g["ref64_weights_after"])
    for pname, p in net.named_parameters():
        assert maxrel(p.grad, g["ref64_grad_" + pname]) < 1e-5, pname       # fixture gradients are stored in float32


@pytest.mark.parametrize("tag", ["cube16", "box20", "zshort12"])
def test_grid_axes_match_reference(tag):
    g = golden("grid.npz")
    axes, length, short = grid_axes(int(g[tag + "/res"]), g[tag + "/bbox"])
    for a, name in zip(axes, "xyz"):
        np.testing.assert_array_equal(a, g["%s/%s" % (tag, name)])
    # flat order [iy][ix][iz] of the reference's meshgrid(indexing='xy') ravel
    pts = g[tag + "/points_f16"]
    nx, ny, nz = (len(a) for a in axes)
    assert pts.shape == (nx * ny * nz, 3)
    iy, ix, iz = np.unravel_index(np.arange(pts.shape[0]), (ny, nx, nz))
    np.testing.assert_array_equal(pts[:, 0], axes[0].astype(np.float16)[ix])
    np.testing.assert_array_equal(pts[:, 1], axes[1].astype(np.float16)[iy])
    np.testing.assert_array_equal(pts[:, 2], axes[2].astype(np.float16)[iz])


def test_slab_and_shard_ranges_partition():
    for n in (0, 1, 7, 512, 15000):
        for world in (1, 2, 3, 8):
            spans = [slab_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
            assert spans == [ddist.shard_range(n, r, world) for r in range(world)]


def test_flat_grads_views():
    lin = torch.nn.Linear(4, 3)
    fg = ddist.FlatGrads(list(lin.parameters()), extra=8)
    assert fg.flat.numel() == 12 + 3 + 8
    lin.weight.grad.fill_(2.0)
    assert float(fg.flat[:12].sum()) == 24.0
    fg.zero_()
    assert float(lin.weight.grad.abs().sum()) == 0.0


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from digs_b200 import dist as dd
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.manual_seed(0)
lin = torch.nn.Linear(6, 5)
pts = torch.arange(2 * 11 * 3, dtype=torch.float32).reshape(2, 11, 3)
mine = dd.shard_points(pts, rank, world)
# a sum over point shards normalised by the GLOBAL count equals the single-process mean
fg = dd.FlatGrads(list(lin.parameters()), extra=2)
fg.flat[:3] = mine.sum(dim=(0, 1)) / (2 * 11)
fg.extra[0] = float(mine.shape[1])
fg.allreduce()
ref = pts.mean(dim=(0, 1))
assert torch.allclose(fg.flat[:3], ref, atol=1e-5), (fg.flat[:3], ref)
assert int(fg.extra[0]) == 11
# shapes sharded over ranks (SURVEY 8e row 3): the batch's shape indices split contiguously, and the (index, gradient row)
# pairs of all ranks scattered into the dense latent-table gradient every rank then holds (duplicates in a batch add up)
batch = torch.tensor([7, 2, 7, 5])
mine = dd.shard_shapes(batch, rank, world)
assert mine.tolist() == batch[2 * rank:2 * rank + 2].tolist()
rows = torch.stack([torch.full((3,), 10.0 * rank + i) for i in range(2)])
dense = torch.ones(9, 3)
dd.gather_latent_grads(dense, mine, rows)
ref = torch.zeros(9, 3)
ref[7] = 0.0 + 10.0          # rank 0 row 0 (value 0) + rank 1 row 0 (value 10)
ref[2] = 1.0
ref[5] = 11.0
assert torch.equal(dense, ref), dense
assert [dd.shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
dist.destroy_process_group()
print("ok", rank)
'''


def test_data_parallel_reduction_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29731")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0, out.decode()


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): ONE JSON line on stdout with the same metric /
    unit / direction as our arm, `impl: reference`, a truthful CPU config, zero copy bytes and no GPU launches."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "5", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train points/sec (value+grad+div+backward)" and d["unit"] == "points/s"
    assert d["higher_is_better"] is True and d["steps"] == 5 and d["warmup"] == 3 and d["gpu_launches"] == 0
    assert d["value"] > 0 and abs(d["value"] - d["e2e"]["value"]) < 1e-6 * d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert "GraphedStep" not in json.dumps(d["config"]) and "bf16" not in d["config"]["precision_mode"]

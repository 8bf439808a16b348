"""GPU parity of the fused step (digs_step_fused, csrc/step_tc.cu): forward jets + DiGSLoss + reverse pass in two launches,
against (a) the fixtures produced by the real reference, (b) the live fp64 oracle at the benchmark size, and (c) the
two-call API of the same library (which must agree to accumulation-order noise)."""
import copy

import numpy as np
import pytest
import torch

import digs_b200
from digs_b200.step import fused_step, fused_step_supported
from oracle import jet_spec, ref_autograd
from oracle.make_golden import STEP_CASES
from _util import golden, maxrel, build_net, np_params, head_of

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
from test_gpu_parity import TOL  # noqa: E402  (same stated bounds as the two-call path)

TC_CASES = [n for n in sorted(STEP_CASES) if STEP_CASES[n]["net"]["decoder_hidden_dim"] % 128 == 0]


def _fused(name, precision):
    case = STEP_CASES[name]
    g = golden("step_%s.npz" % name)
    net = build_net(case["net"], case["seed"], precision=precision).to(DEV)
    crit = digs_b200.DiGSLoss(**copy.deepcopy(case["loss"]))
    m = torch.tensor(g["in_mnfld"], device=DEV)
    nm = torch.tensor(g["in_nonmnfld"], device=DEV)
    nrm = torch.tensor(g["in_normals"], device=DEV)
    lat = torch.tensor(g["in_latent"], device=DEV) if "in_latent" in g.files else None
    loss_dict, out, lat_grad = fused_step(net, crit, m, nm, nrm, latent=lat)
    torch.cuda.synchronize()
    return case, g, net, out, loss_dict, lat_grad


@pytest.mark.parametrize("precision", ["bf16x2", "bf16"])
@pytest.mark.parametrize("name", TC_CASES)
def test_fused_step_matches_reference_fixture(name, precision):
    case, g, net, out, loss_dict, lat_grad = _fused(name, precision)
    bs = case["bs"]
    JET_TOL, GRAD_TOL = TOL[precision]
    assert out["manifold_pnts_pred"].shape == (bs, case["Nm"])
    assert out["nonmanifold_pnts_pred"].shape == (bs, case["Nn"])
    assert maxrel(out["manifold_pnts_pred"].cpu(), g["ref64_f_m"]) < JET_TOL
    assert maxrel(out["nonmanifold_pnts_pred"].cpu(), g["ref64_f_n"]) < JET_TOL
    assert maxrel(out["manifold_pnts_grad"].cpu(), g["ref64_g_m"]) < JET_TOL
    assert maxrel(out["nonmanifold_pnts_grad"].cpu(), g["ref64_g_n"]) < JET_TOL
    assert maxrel(out["nonmanifold_pnts_div"].cpu(), g["ref64_lap_n"]) < JET_TOL
    if GRAD_TOL is None:
        return
    for key in ("loss", "sdf_term", "inter_term", "eikonal_term", "normals_loss", "div_loss", "latent_reg_term"):
        ref = float(g["ref64_term_" + key][0])
        got = float(loss_dict[key].reshape(-1)[0])
        assert abs(got - ref) <= JET_TOL * max(abs(ref), 1e-3), (key, got, ref)
    for pname, p in net.named_parameters():
        assert p.grad is not None, pname
        assert maxrel(p.grad.cpu(), g["ref64_grad_" + pname]) < GRAD_TOL, pname
    if lat_grad is not None:
        assert maxrel(lat_grad.cpu(), g["ref64_grad_latent"]) < GRAD_TOL


@pytest.mark.parametrize("loss_type,div_type", [("siren", "l1"), ("siren_wo_n", "l1"), ("siren_w_div", "l2"),
                                                ("siren_wo_n_w_div", "l2")])
@pytest.mark.parametrize("n_m,n_n", [(1, 1), (23, 25), (700, 333), (4099, 4001)])
def test_fused_step_equals_two_call_api(loss_type, div_type, n_m, n_n):
    """Ragged / tiny clouds, every in-scope loss type: fused step == DiGSNetwork.forward -> DiGSLoss -> backward."""
    case = STEP_CASES["c2_small"]
    lk = dict(copy.deepcopy(case["loss"]), loss_type=loss_type, div_type=div_type)
    rng = np.random.RandomState(n_m + n_n)
    v = rng.normal(size=(1, n_m, 3))
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    m = torch.tensor(0.5 * v, dtype=torch.float32, device=DEV)
    nrm = torch.tensor(v, dtype=torch.float32, device=DEV)
    nm = torch.tensor(rng.uniform(-1.1, 1.1, size=(1, n_n, 3)), dtype=torch.float32, device=DEV)
    res = []
    for fused in (False, True):
        net = build_net(case["net"], case["seed"], precision="bf16x2").to(DEV)
        crit = digs_b200.DiGSLoss(**copy.deepcopy(lk))
        crit.global_counts = (2 * n_m, 3 * n_n)          # as if sharded: means over GLOBAL counts
        if fused:
            ld, out, _ = fused_step(net, crit, m, nm, nrm)
        else:
            mm, nn_ = m.clone().requires_grad_(), nm.clone().requires_grad_()
            out = net(nn_, mm)
            ld, _ = crit(out, mm, nn_, nrm)
            ld["loss"].backward()
        torch.cuda.synchronize()
        flat = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
        res.append((ld, out, flat, list(crit.weights)))
    (ld0, out0, g0, w0), (ld1, out1, g1, w1) = res
    assert w0 == w1                                        # siren_wo_n zeroes weights[2] in both (models/losses.py:153)
    for k in ("manifold_pnts_pred", "nonmanifold_pnts_pred", "manifold_pnts_grad", "nonmanifold_pnts_grad"):
        assert maxrel(out1[k].cpu(), out0[k].detach().cpu()) < 1e-6, k
    if "div" in loss_type:
        assert maxrel(out1["nonmanifold_pnts_div"].cpu(), out0["nonmanifold_pnts_div"].detach().cpu()) < 1e-6
    for k in ("loss", "sdf_term", "inter_term", "eikonal_term", "normals_loss", "div_loss"):
        a, b = float(ld1[k].reshape(-1)[0]), float(ld0[k].detach().reshape(-1)[0])
        assert abs(a - b) <= 2e-6 * max(abs(b), 1e-3), (k, a, b)
    assert maxrel(g1.cpu(), g0.cpu()) < 2e-5


def test_fused_step_full_size_c2_against_live_oracle():
    """15 000 + 15 000 points, 4x256 MFGI (the configuration the metric is quoted on) vs the fp64 closed form."""
    case = STEP_CASES["c2_small"]
    JET_TOL, GRAD_TOL = TOL["bf16x2"]
    net = build_net(case["net"], case["seed"], precision="bf16x2").to(DEV)
    rng = np.random.RandomState(0)
    v = rng.normal(size=(15000, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    m_np = (0.5 * v).astype(np.float32)
    nm_np = rng.uniform(-1.1, 1.1, size=(15000, 3)).astype(np.float32)
    lk = case["loss"]
    ref = jet_spec.step(np_params(net, np.float64), m_np.astype(np.float64), nm_np.astype(np.float64), v, head_of(net),
                        lk["weights"], lk["loss_type"], lk["div_type"], lk["div_clamp"])
    tparams = [p.detach().cpu().clone().requires_grad_(True) for p in net.decoder.fc_block.flat_params()]
    _, g32, j32 = ref_autograd.step(tparams, torch.tensor(m_np), torch.tensor(nm_np), torch.tensor(v.astype(np.float32)),
                                    head_of(net), lk["weights"], lk["loss_type"], lk["div_type"], lk["div_clamp"])
    crit = digs_b200.DiGSLoss(**copy.deepcopy(lk))
    ld, out, _ = fused_step(net, crit, torch.tensor(m_np[None], device=DEV), torch.tensor(nm_np[None], device=DEV),
                            torch.tensor(v[None].astype(np.float32), device=DEV))
    torch.cuda.synchronize()

    def check(name, got, ref64, ref32, tol):
        e_ours, e_ref = maxrel(got, ref64), maxrel(ref32, ref64)
        print("%-8s fused-vs-fp64 %.2e   reference-fp32-vs-fp64 %.2e" % (name, e_ours, e_ref))
        assert e_ours < max(tol, 2 * e_ref), (name, e_ours, e_ref)

    check("f_n", out["nonmanifold_pnts_pred"].cpu()[0], ref["fn"]["f"], j32["f_n"].numpy(), JET_TOL)
    check("grad_n", out["nonmanifold_pnts_grad"].cpu()[0], ref["fn"]["grad"], j32["g_n"].numpy(), JET_TOL)
    check("lap_n", out["nonmanifold_pnts_div"].cpu()[0], ref["fn"]["lap"], j32["lap_n"].numpy(), JET_TOL)
    check("grad_m", out["manifold_pnts_grad"].cpu()[0], ref["fm"]["grad"], j32["g_m"].numpy(), JET_TOL)
    assert abs(float(ld["loss"]) - ref["terms"]["loss"]) < JET_TOL * abs(ref["terms"]["loss"])
    lins = net.decoder.fc_block.linears()
    for li, (dW, db) in enumerate(ref["grads"]):
        check("dW%d" % li, lins[li].weight.grad.cpu(), dW, g32[2 * li].numpy(), GRAD_TOL)
        check("db%d" % li, lins[li].bias.grad.cpu(), db, g32[2 * li + 1].numpy(), GRAD_TOL)


def test_fused_step_refuses_what_it_does_not_cover():
    case = STEP_CASES["c2_small"]
    net = build_net(case["net"], 1, precision="fp32").to(DEV)
    crit = digs_b200.DiGSLoss(**copy.deepcopy(case["loss"]))
    x = torch.zeros(1, 8, 3, device=DEV)
    assert not fused_step_supported(net, crit)
    with pytest.raises(RuntimeError):
        fused_step(net, crit, x, x, x)
    net.precision = "bf16x2"
    crit.loss_type, crit.use_curvs = "siren_w_curv", True
    with pytest.raises(RuntimeError):
        fused_step(net, crit, x, x, x)
    with pytest.raises(RuntimeError):
        fused_step(net, digs_b200.DiGSLoss(**copy.deepcopy(case["loss"])), x.cpu(), x, x)


def test_graphed_iteration_with_optimizer_matches_eager_loop():
    """sampling-free whole iteration in ONE graph (zero-fill, operand prep, fused step, clip + Adam with the step count
    on the device) == the same iteration issued eagerly through the two-call API + FusedClipAdam."""
    case = STEP_CASES["c2_small"]
    lk = case["loss"]
    rng = np.random.RandomState(5)
    N = 2048
    batches = []
    for _ in range(6):
        v = rng.normal(size=(1, N, 3))
        v /= np.linalg.norm(v, axis=-1, keepdims=True)
        batches.append((torch.tensor(0.5 * v, dtype=torch.float32, device=DEV), torch.tensor(v, dtype=torch.float32, device=DEV),
                        torch.tensor(rng.uniform(-1.1, 1.1, size=(1, N, 3)), dtype=torch.float32, device=DEV)))
    outs = []
    for graphed in (False, True):
        net = build_net(case["net"], case["seed"], precision="bf16x2").to(DEV)
        crit = digs_b200.DiGSLoss(**copy.deepcopy(lk))
        opt = digs_b200.FusedClipAdam(net.parameters(), lr=5e-5, max_norm=10.0, device_step=graphed)
        losses = []
        if graphed:
            gs = digs_b200.GraphedStep(net, crit, N, N, optimizer=opt)
            # the warm-up / capture iterations advanced the optimizer: restart from the same state as the eager run
            ref_net = build_net(case["net"], case["seed"], precision="bf16x2").to(DEV)
            with torch.no_grad():
                for p, q in zip(net.parameters(), ref_net.parameters()):
                    p.copy_(q)
            opt.exp_avg.zero_(); opt.exp_avg_sq.zero_(); opt._step_dev.zero_()
            for i, (m, nrm, nm) in enumerate(batches):
                crit.update_div_weight(i, 6, [1e2, 0.2, 1e2, 0.4, 0, 0])
                losses.append(float(gs.step(m, nm, nrm)["loss"]))
        else:
            for i, (m, nrm, nm) in enumerate(batches):
                crit.update_div_weight(i, 6, [1e2, 0.2, 1e2, 0.4, 0, 0])
                opt.zero_grad()
                mm, nn_ = m.clone().requires_grad_(), nm.clone().requires_grad_()
                ld, _ = crit(net(nn_, mm), mm, nn_, nrm)
                ld["loss"].backward()
                opt.step()
                losses.append(float(ld["loss"]))
        outs.append((losses, torch.cat([p.detach().reshape(-1) for p in net.parameters()]).cpu()))
    init = torch.cat([p.detach().reshape(-1) for p in build_net(case["net"], case["seed"]).parameters()])
    (l0, p0), (l1, p1) = outs
    assert np.allclose(l0, l1, rtol=2e-3), (l0, l1)
    upd0, upd1 = p0 - init, p1 - init
    assert float(upd0.abs().mean()) > 0
    assert float((upd1 - upd0).abs().mean()) < 0.02 * float(upd0.abs().mean())


def test_pipelined_host_feed_equals_blocking_steps():
    """GraphedStep.submit / result (host-to-device copy of batch i+1 under step i, loss read back one call later) returns
    the same per-iteration losses as blocking step() calls on the same pinned host batches, and rejects stale tickets."""
    case = STEP_CASES["c2_small"]
    rng = np.random.RandomState(11)
    N = 1536
    net = build_net(case["net"], case["seed"], precision="bf16x2").to(DEV)
    crit = digs_b200.DiGSLoss(**copy.deepcopy(case["loss"]))
    gs = digs_b200.GraphedStep(net, crit, N, N)
    packs = []
    for _ in range(7):
        v = rng.normal(size=(1, N, 3))
        v /= np.linalg.norm(v, axis=-1, keepdims=True)
        packs.append(gs.pack_host_batch(torch.tensor(0.5 * v, dtype=torch.float32), torch.tensor(rng.uniform(-1.1, 1.1, size=(1, N, 3)), dtype=torch.float32),
                                        torch.tensor(v, dtype=torch.float32)))
    blocking = [float(gs.step(pk[1], pk[2], pk[3])["loss"]) for pk in packs]
    grads_blocking = gs.flat.flat.clone()
    piped, prev = [], None
    for i, pk in enumerate(packs):
        if i % 2:                                       # zero-copy path: the loader writes into the staging views itself
            m_s, nm_s, nrm_s = gs.staging()
            m_s.copy_(pk[1]); nm_s.copy_(pk[2]); nrm_s.copy_(pk[3])
            tk = gs.submit()
        else:
            tk = gs.submit(pk[0])
        if prev is not None:
            piped.append(gs.result(prev)[0])
        prev = tk
    piped.append(gs.result(prev)[0])
    torch.cuda.synchronize()
    assert len(set(blocking)) > 1                      # the batches differ
    assert np.allclose(piped, blocking, rtol=1e-5), (piped, blocking)
    assert float((gs.flat.flat - grads_blocking).abs().max()) <= 1e-4 * float(grads_blocking.abs().max())
    with pytest.raises(ValueError):
        gs.result(prev - 5)
    with pytest.raises(ValueError):
        gs.submit(torch.zeros(3))


@pytest.mark.parametrize("n_m,n_n", [(5, 3), (700, 333), (9001, 8000), (15000, 15000)])
def test_overlapped_weight_gradient_launch_equals_stream_ordered_launch(n_m, n_n, monkeypatch):
    """The weight-gradient GEMM as a programmatic dependent launch on the tile kernel's round counters (default) == the
    same GEMM launched in plain stream order (DIGS_B200_WGRAD_OVERLAP=0), for K ranges shorter and longer than its ring."""
    case = STEP_CASES["c2_small"]
    rng = np.random.RandomState(3)
    v = rng.normal(size=(1, n_m, 3))
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    m = torch.tensor(0.5 * v, dtype=torch.float32, device=DEV)
    nrm = torch.tensor(v, dtype=torch.float32, device=DEV)
    nm = torch.tensor(rng.uniform(-1.1, 1.1, size=(1, n_n, 3)), dtype=torch.float32, device=DEV)
    grads = []
    for overlap in ("1", "0", "1"):
        monkeypatch.setenv("DIGS_B200_WGRAD_OVERLAP", overlap)
        net = build_net(case["net"], case["seed"], precision="bf16x2").to(DEV)
        crit = digs_b200.DiGSLoss(**copy.deepcopy(case["loss"]))
        for _ in range(3):                              # back to back: the counters are reset by every call
            for p in net.parameters():
                p.grad = None
            fused_step(net, crit, m, nm, nrm)
        torch.cuda.synchronize()
        grads.append(torch.cat([p.grad.reshape(-1) for p in net.parameters()]).cpu())
    scale = float(grads[1].abs().max())
    assert scale > 0
    assert float((grads[0] - grads[1]).abs().max()) <= 2e-5 * scale
    assert float((grads[2] - grads[1]).abs().max()) <= 2e-5 * scale

"""TEST INFRASTRUCTURE ONLY.  Generate tests/golden/*.npz by RUNNING THE REAL REFERENCE (/root/reference).

Run in the build container (the reference tree does not exist on the GPU box):

    python -m oracle.make_golden

Everything written here is an output of the unmodified reference modules models/DiGS.py,
models/losses.py and utils/utils.py (imported with five stub modules, oracle/reference_loader.py):
  * init_*      sha256 of every parameter tensor after seeded construction (all three init schemes)
  * step_*      jets (f, grad f, lap f), the 8 loss terms and all parameter gradients of one
                forward + DiGSLoss + backward, in fp64 (ground truth) and fp32 (the reference's own noise floor)
  * grid_*      get_3d_grid axes and decoder values in the reference's flat point order
  * divw_*      update_div_weight schedules
  * trained_c2  weights after 200 iterations of the reference's own training loop on a synthetic torus + one step on them
  * c1_training_curves  loss curves / fit metrics of the reference's 2-D sanity loop over three seeds (statistical parity)
"""
import copy
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import reference_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# name -> dict(net kwargs, loss kwargs, Nm, Nn, bs, seed)
STEP_CASES = {
    "c2_small": dict(net=dict(latent_size=0, in_dim=3, decoder_hidden_dim=256, nl="sine", encoder_type="none",
                              decoder_n_hidden_layers=4, init_type="mfgi", sphere_init_params=[1.6, 0.1]),
                     loss=dict(weights=[3e3, 1e2, 1e2, 5e1, 1e2], loss_type="siren_wo_n_w_div", div_decay="linear",
                               div_type="l1", div_clamp=50), Nm=512, Nn=512, bs=1, seed=3627473, box=1.1),
    "c1_small": dict(net=dict(latent_size=0, in_dim=2, decoder_hidden_dim=128, nl="sine", encoder_type="none",
                              decoder_n_hidden_layers=4, init_type="mfgi"),
                     loss=dict(weights=[3e3, 1e2, 1e2, 5e1, 1e2], loss_type="siren_wo_n_w_div", div_decay="linear",
                               div_type="l1", div_clamp=50), Nm=384, Nn=384, bs=1, seed=11, box=1.2),
    "siren_l2_ragged": dict(net=dict(latent_size=0, in_dim=3, decoder_hidden_dim=128, nl="sine", encoder_type="none",
                                     decoder_n_hidden_layers=2, init_type="siren"),
                            loss=dict(weights=[3e3, 1e2, 1e2, 5e1, 1e1], loss_type="siren_w_div", div_decay="none",
                                      div_type="l2", div_clamp=50), Nm=301, Nn=203, bs=1, seed=5, box=1.1),
    "geo_siren_loss": dict(net=dict(latent_size=0, in_dim=3, decoder_hidden_dim=64, nl="sine", encoder_type="none",
                                    decoder_n_hidden_layers=3, init_type="geometric_sine",
                                    sphere_init_params=[1.6, 1.0]),
                           loss=dict(weights=[3e3, 1e2, 1e2, 5e1, 1e2], loss_type="siren", div_decay="none",
                                     div_type="l1", div_clamp=50), Nm=130, Nn=130, bs=2, seed=7, box=1.1),
    "wo_n_2d": dict(net=dict(latent_size=0, in_dim=2, decoder_hidden_dim=64, nl="sine", encoder_type="none",
                             decoder_n_hidden_layers=1, init_type="siren"),
                    loss=dict(weights=[3e3, 1e2, 1e2, 5e1, 1e2], loss_type="siren_wo_n", div_decay="none",
                              div_type="l1", div_clamp=50), Nm=64, Nn=96, bs=1, seed=9, box=1.2),
    "latent_small": dict(net=dict(latent_size=16, in_dim=3, decoder_hidden_dim=128, nl="sine",
                                  encoder_type="autodecoder", decoder_n_hidden_layers=2, init_type="mfgi",
                                  sphere_init_params=[1.6, 0.1]),
                         loss=dict(weights=[3e3, 1e2, 1e2, 5e1, 1e2, 1e0], loss_type="siren_wo_n_w_div",
                                   div_decay="linear", div_type="l1", div_clamp=50), Nm=96, Nn=96, bs=2, seed=13,
                         box=1.2),
}

# Loss variants outside the fused kernel (curvature / IGR / ground-truth divergence, models/losses.py:79-181): small nets,
# extra inputs (two principal curvatures per surface point; divergence targets) drawn from the case seed.
_SMALL3 = dict(latent_size=0, in_dim=3, decoder_hidden_dim=128, nl="sine", encoder_type="none", decoder_n_hidden_layers=2,
               init_type="mfgi", sphere_init_params=[1.6, 0.1])
_SMALL2 = dict(latent_size=0, in_dim=2, decoder_hidden_dim=64, nl="sine", encoder_type="none", decoder_n_hidden_layers=2,
               init_type="siren")
LOSSVAR_CASES = {
    "curv_l1": dict(net=_SMALL3, loss=dict(weights=[3e3, 1e2, 1e2, 5e1, 1e2], loss_type="siren_w_curv", div_type="l1",
                                           div_clamp=50), Nm=160, Nn=160, bs=1, seed=21, box=1.1),
    "div_curv_l2": dict(net=_SMALL3, loss=dict(weights=[3e3, 1e2, 1e2, 5e1, 1e1], loss_type="siren_wo_n_w_div_w_curv",
                                               div_type="l2", div_clamp=50), Nm=128, Nn=192, bs=1, seed=22, box=1.1),
    "n_div_curv_2d": dict(net=_SMALL2, loss=dict(weights=[3e3, 1e2, 1e2, 5e1, 1e2], loss_type="siren_w_div_w_curv",
                                                 div_type="l1", div_clamp=50), Nm=96, Nn=96, bs=2, seed=23, box=1.2),
    "igr": dict(net=_SMALL3, loss=dict(weights=[3e3, 1e2, 1e2, 5e1, 1e2], loss_type="igr"), Nm=100, Nn=140, bs=2, seed=24,
                box=1.1),
    "igr_wo_n": dict(net=_SMALL3, loss=dict(weights=[3e3, 1e2, 1e2, 5e1, 1e2], loss_type="igr_wo_n"), Nm=100, Nn=100, bs=1,
                     seed=25, box=1.1),
    "gt_l2": dict(net=_SMALL3, loss=dict(weights=[3e3, 1e2, 1e2, 5e1, 1e0], loss_type="siren_w_div_w_curv", div_type="gt_l2",
                                         div_clamp=50), Nm=120, Nn=120, bs=1, seed=26, box=1.1),
    "gt_l1": dict(net=_SMALL3, loss=dict(weights=[3e3, 1e2, 1e2, 5e1, 1e0], loss_type="siren_wo_n_w_div_w_curv",
                                         div_type="gt_l1", div_clamp=50), Nm=120, Nn=120, bs=1, seed=27, box=1.1),
}


def lossvar_extras(case):
    """(curvatures (bs,Nm,2), nonmnfld_div_gt (bs,Nn), mnfld_div_gt (bs,Nm)) float32, deterministic per case."""
    rng = np.random.RandomState(1000 + case["seed"])
    bs, Nm, Nn = case["bs"], case["Nm"], case["Nn"]
    curv = rng.uniform(-3.0, 3.0, size=(bs, Nm, 2)).astype(np.float32)
    return curv, rng.uniform(-4, 4, size=(bs, Nn)).astype(np.float32), rng.uniform(-4, 4, size=(bs, Nm)).astype(np.float32)


def run_reference_lossvar(ref_digs, case, dtype):
    torch.manual_seed(case["seed"])
    net = ref_digs.DiGSNetwork(**case["net"]).to(dtype)
    crit = ref_digs.DiGSLoss(**copy.deepcopy(case["loss"]))
    m, nrm, nm, _ = make_inputs(case, dtype)
    curv, dgt_n, dgt_m = lossvar_extras(case)
    m_t, n_t, nm_t = (torch.tensor(a, dtype=dtype) for a in (m, nrm, nm))
    m_t.requires_grad_()
    nm_t.requires_grad_()
    out = net(nm_t, m_t)
    loss_dict, _ = crit(out, m_t, nm_t, n_t, curvatures=torch.tensor(curv, dtype=dtype),
                        nonmnfld_div_gt=torch.tensor(dgt_n, dtype=dtype), mnfld_div_gt=torch.tensor(dgt_m, dtype=dtype))
    loss_dict["loss"].backward()
    res = {"term_" + k: np.asarray(v.detach().numpy()).reshape(-1)[:1] for k, v in loss_dict.items()}
    for name, p in net.named_parameters():
        res["grad_" + name] = p.grad.detach().numpy().astype(np.float32)
    res["weights_after"] = np.array(crit.weights, dtype=np.float64)
    return res, dict(in_mnfld=m, in_normals=nrm, in_nonmnfld=nm, in_curvatures=curv, in_div_gt_n=dgt_n, in_div_gt_m=dgt_m)


def make_lossvar(ref_digs):
    for name, case in LOSSVAR_CASES.items():
        r64, inputs = run_reference_lossvar(ref_digs, case, torch.float64)
        blob = dict(inputs)
        blob.update({"ref64_" + k: v for k, v in r64.items()})
        np.savez_compressed(os.path.join(OUT, "lossvar_%s.npz" % name), **blob)
        print("lossvar", name, {k[5:]: float(v[0]) for k, v in r64.items() if k.startswith("term_")})


INIT_CASES = {
    "mfgi_c2": dict(latent_size=0, in_dim=3, decoder_hidden_dim=256, nl="sine", encoder_type="none",
                    decoder_n_hidden_layers=4, init_type="mfgi", sphere_init_params=[1.6, 0.1]),
    "mfgi_c1": dict(latent_size=0, in_dim=2, decoder_hidden_dim=128, nl="sine", encoder_type="none",
                    decoder_n_hidden_layers=4, init_type="mfgi"),
    "siren_c4": dict(latent_size=0, in_dim=3, decoder_hidden_dim=512, nl="sine", encoder_type="none",
                     decoder_n_hidden_layers=8, init_type="siren"),
    "geo": dict(latent_size=0, in_dim=3, decoder_hidden_dim=64, nl="sine", encoder_type="none",
                decoder_n_hidden_layers=3, init_type="geometric_sine"),
    "mfgi_c5": dict(latent_size=256, in_dim=3, decoder_hidden_dim=256, nl="sine", encoder_type="autodecoder",
                    decoder_n_hidden_layers=8, init_type="mfgi", sphere_init_params=[1.6, 0.1]),
}
INIT_SEED = 3627473


def sha(t):
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


def make_inputs(case, dtype):
    """Deterministic synthetic clouds: manifold on a sphere/circle r=0.5 (+ analytic normals), off-surface uniform."""
    rng = np.random.RandomState(case["seed"])
    d = case["net"]["in_dim"]
    bs, Nm, Nn = case["bs"], case["Nm"], case["Nn"]
    v = rng.normal(size=(bs, Nm, d))
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    m = (0.5 * v).astype(np.float32)
    nrm = v.astype(np.float32)
    nm = rng.uniform(-case["box"], case["box"], size=(bs, Nn, d)).astype(np.float32)
    lat = None
    if case["net"]["latent_size"] > 0:
        lat = (0.01 * rng.normal(size=(bs, case["net"]["latent_size"]))).astype(np.float32)
    return m, nrm, nm, lat


def run_reference_step(ref_digs, case, dtype, state_dict=None, inputs=None):
    torch.manual_seed(case["seed"])
    net = ref_digs.DiGSNetwork(**case["net"])
    if state_dict is not None:
        net.load_state_dict(state_dict)
    crit = ref_digs.DiGSLoss(**copy.deepcopy(case["loss"]))
    m, nrm, nm, lat = make_inputs(case, dtype) if inputs is None else inputs
    net = net.to(dtype)
    m_t, n_t, nm_t = (torch.tensor(a, dtype=dtype) for a in (m, nrm, nm))
    m_t.requires_grad_()
    nm_t.requires_grad_()
    lat_t = None
    if lat is not None:
        lat_t = torch.tensor(lat, dtype=dtype, requires_grad=True)
        m_in = torch.cat([m_t, lat_t[:, None, :].expand(-1, m_t.shape[1], -1)], dim=-1)
        nm_in = torch.cat([nm_t, lat_t[:, None, :].expand(-1, nm_t.shape[1], -1)], dim=-1)
    else:
        m_in, nm_in = m_t, nm_t
    out = net(nm_in, m_in)
    loss_dict, mnfld_grad = crit(out, m_t, nm_t, n_t)
    # jets as the loss sees them (utils/utils.py:108-117, models/losses.py:72-106)
    import utils.utils as U
    g_n = U.gradient(nm_t, out["nonmanifold_pnts_pred"])
    d = m_t.shape[-1]
    lap = sum(U.gradient(nm_t, g_n[:, :, k])[:, :, k] for k in range(d))
    loss_dict["loss"].backward()
    res = dict(f_m=out["manifold_pnts_pred"], g_m=mnfld_grad, f_n=out["nonmanifold_pnts_pred"], g_n=g_n, lap_n=lap)
    res = {k: v.detach().numpy() for k, v in res.items()}
    for k, v in loss_dict.items():
        res["term_" + k] = np.asarray(v.detach().numpy()).reshape(-1)[:1]
    for name, p in net.named_parameters():
        res["grad_" + name] = p.grad.detach().numpy()
    if lat_t is not None:
        res["grad_latent"] = lat_t.grad.detach().numpy()
    shas = {name: sha(p.float()) for name, p in net.named_parameters()}
    return res, shas, (m, nrm, nm, lat)


# ---- trained weights (SURVEY 7.3-1 / 8d "second data point"): where f ~ 0 the max-norm-relative error of f is worst ------
TRAINED = dict(steps=200, n_batch=4096, n_eval=600, lr=5e-5, clip=10.0, decay=(1e2, 0.2, 1e2, 0.4, 0.0, 0.0))


def reference_training_loop(ref_digs, net, crit, batches, lr, clip, decay, dtype=torch.float32):
    """The reference's own iteration (train_surface_reconstruction.py:114-158 / train_basic_shape.py:94-135):
    forward, DiGSLoss, backward, [clip_grad_norm_ 10], Adam.step, update_div_weight.  Returns the list of losses."""
    opt = torch.optim.Adam(net.parameters(), lr=lr)
    n_iter = len(batches)
    losses = []
    for it, (m, nrm, nm) in enumerate(batches):
        m = torch.as_tensor(m, dtype=dtype).requires_grad_()
        nm = torch.as_tensor(nm, dtype=dtype).requires_grad_()
        net.zero_grad()
        out = net(nm, m)
        ld, _ = crit(out, m, nm, torch.as_tensor(nrm, dtype=dtype))
        ld["loss"].backward()
        if clip:
            torch.nn.utils.clip_grad_norm_(net.parameters(), clip)
        opt.step()
        if decay is not None:
            crit.update_div_weight(it, n_iter, decay)
        losses.append(float(ld["loss"].detach()))
    return losses


def make_trained(ref_digs):
    from digs_b200 import synth                     # analytic clouds only (numpy)
    case = STEP_CASES["c2_small"]
    cloud, normals = synth.surface_cloud(100000, "torus", seed=0)
    batches = [synth.sample_batch(cloud, normals, TRAINED["n_batch"], seed=1000 + i, box=1.1) for i in range(TRAINED["steps"])]
    torch.manual_seed(case["seed"])
    net = ref_digs.DiGSNetwork(**case["net"])
    crit = ref_digs.DiGSLoss(**copy.deepcopy(case["loss"]))
    losses = reference_training_loop(ref_digs, net, crit, batches, TRAINED["lr"], TRAINED["clip"], TRAINED["decay"])
    state = {k: v.detach().clone() for k, v in net.state_dict().items()}
    m, nrm, nm = synth.sample_batch(cloud, normals, TRAINED["n_eval"], seed=77, box=1.1)
    inputs = (m, nrm, nm, None)
    r64, _, _ = run_reference_step(ref_digs, case, torch.float64, state_dict=state, inputs=inputs)
    r32, _, _ = run_reference_step(ref_digs, case, torch.float32, state_dict=state, inputs=inputs)
    blob = dict(in_mnfld=m, in_normals=nrm, in_nonmnfld=nm, train_losses=np.array(losses))
    for k, v in state.items():
        blob["w_" + k] = v.numpy()
    for k, v in r64.items():
        blob["ref64_" + k] = v.astype(np.float32) if k.startswith("grad_") else v
    for k, v in r32.items():
        if not k.startswith("grad_"):
            blob["ref32_" + k] = v
    np.savez_compressed(os.path.join(OUT, "trained_c2.npz"), **blob)
    print("trained_c2: loss %.3f -> %.3f after %d reference steps; |f_m| mean %.2e" % (
        losses[0], losses[-1], len(losses), float(np.abs(r64["f_m"]).mean())))


# ---- statistical N-step run on config 1: loss curves of the reference's own 2-D sanity loop, several seeds ------
CURVES = dict(steps=300, n_batch=2048, seeds=(11, 12, 13), lr=1e-4)      # sc_args.py:29 lr; train_basic_shape.py does not clip


def c1_batches(seed, steps, n):
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(steps):
        t = rng.uniform(0, 2 * np.pi, n)
        nrm = np.stack([np.cos(t), np.sin(t)], 1).astype(np.float32)
        out.append(((0.5 * nrm)[None], nrm[None], rng.uniform(-1.2, 1.2, size=(1, n, 2)).astype(np.float32)))
    return out


def c1_eval_points():
    t = np.linspace(0, 2 * np.pi, 512, endpoint=False)
    on = 0.5 * np.stack([np.cos(t), np.sin(t)], 1)
    g = np.linspace(-1.0, 1.0, 33)
    off = np.stack(np.meshgrid(g, g), -1).reshape(-1, 2)
    return on.astype(np.float32), off.astype(np.float32)


def make_curves(ref_digs):
    case = STEP_CASES["c1_small"]
    blob = dict(seeds=np.array(CURVES["seeds"]))
    on, off = c1_eval_points()
    for seed in CURVES["seeds"]:
        torch.manual_seed(seed)
        net = ref_digs.DiGSNetwork(**case["net"])
        crit = ref_digs.DiGSLoss(**copy.deepcopy(case["loss"]))
        losses = reference_training_loop(ref_digs, net, crit, c1_batches(seed, CURVES["steps"], CURVES["n_batch"]),
                                         CURVES["lr"], None, (1e2, 0.2, 1e2, 0.4, 0.0, 0.0))
        with torch.no_grad():
            f_on = net.decoder(torch.tensor(on)).numpy()[:, 0]
            f_off = net.decoder(torch.tensor(off)).numpy()[:, 0]
        blob["loss_%d" % seed] = np.array(losses)
        blob["f_on_abs_mean_%d" % seed] = np.array(np.abs(f_on).mean())
        # agreement of the learnt field with the true signed distance of the r = 0.5 circle on a coarse grid
        sdf = np.linalg.norm(off, axis=1) - 0.5
        blob["sdf_err_%d" % seed] = np.array(np.abs(f_off - sdf).mean())
        print("curves seed %d: loss %.2f -> %.2f, mean|f| on the circle %.3e, mean |f - sdf| %.3e" % (
            seed, losses[0], np.mean(losses[-20:]), np.abs(f_on).mean(), np.abs(f_off - sdf).mean()))
    np.savez_compressed(os.path.join(OUT, "c1_training_curves.npz"), **blob)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_digs, ref_losses, ref_utils = reference_loader.load()
    if "--only-lossvar" in sys.argv:          # regenerate just the loss-variant fixtures
        make_lossvar(ref_digs)
        return
    if "--only-trained" in sys.argv:
        make_trained(ref_digs)
        return
    if "--only-curves" in sys.argv:
        make_curves(ref_digs)
        return
    make_trained(ref_digs)
    make_curves(ref_digs)
    make_lossvar(ref_digs)

    # ---- init fingerprints -----------------------------------------------------------------------
    init = {}
    for name, kw in INIT_CASES.items():
        torch.manual_seed(INIT_SEED)
        net = ref_digs.DiGSNetwork(**kw)
        for pname, p in net.named_parameters():
            init["%s/%s" % (name, pname)] = np.array(sha(p))
            init["%s/%s/head" % (name, pname)] = p.detach().reshape(-1)[:8].numpy().copy()
    np.savez_compressed(os.path.join(OUT, "init_fingerprints.npz"), **init)

    # ---- step fixtures -----------------------------------------------------------------------------
    for name, case in STEP_CASES.items():
        r64, shas, (m, nrm, nm, lat) = run_reference_step(ref_digs, case, torch.float64)
        r32, shas32, _ = run_reference_step(ref_digs, case, torch.float32)
        blob = dict(in_mnfld=m, in_normals=nrm, in_nonmnfld=nm)
        if lat is not None:
            blob["in_latent"] = lat
        for k, v in r64.items():
            # jets/terms in fp64; the (large) parameter gradients rounded to fp32 -- tolerance in tests is 1e-4
            blob["ref64_" + k] = v.astype(np.float32) if k.startswith("grad_") else v
        for k, v in r32.items():
            if not k.startswith("grad_"):
                blob["ref32_" + k] = v
        for k, v in shas32.items():
            blob["sha_" + k] = np.array(v)
        np.savez_compressed(os.path.join(OUT, "step_%s.npz" % name), **blob)
        print("step", name, {k: float(v[0]) for k, v in r64.items() if k.startswith("term_")})

    # ---- grid fixtures -----------------------------------------------------------------------------
    grid = {}
    for tag, (res, bbox) in {"cube16": (16, np.array([[-1, 1], [-1, 1], [-1, 1]], dtype=float)),
                             "box20": (20, np.array([[-0.9, 0.8], [-0.5, 0.45], [-0.7, 0.75]])),
                             "zshort12": (12, np.array([[-1.0, 1.0], [-0.8, 0.9], [-0.3, 0.3]]))}.items():
        gd = ref_utils.get_3d_grid(resolution=res, bbox=bbox)
        for a, ax in zip("xyz", gd["xyz"]):
            grid["%s/%s" % (tag, a)] = ax
        grid["%s/bbox" % tag] = bbox
        grid["%s/res" % tag] = np.array(res)
        grid["%s/points_f16" % tag] = gd["grid_points"].numpy()
        torch.manual_seed(INIT_SEED)
        net = ref_digs.DiGSNetwork(latent_size=0, in_dim=3, decoder_hidden_dim=64, nl="sine", encoder_type="none",
                                   decoder_n_hidden_layers=2, init_type="mfgi", sphere_init_params=[1.6, 0.1])
        with torch.no_grad():
            grid["%s/values32" % tag] = net.decoder(gd["grid_points"].float()).numpy()[:, 0]
            grid["%s/values64" % tag] = net.double().decoder(gd["grid_points"].double()).numpy()[:, 0]
    np.savez_compressed(os.path.join(OUT, "grid.npz"), **grid)

    # ---- 2-D derivative-field images (utils.compute_deriv_props, utils/utils.py:144-173) -------------------------------
    deriv = {}
    torch.manual_seed(INIT_SEED)
    net2 = ref_digs.DiGSNetwork(latent_size=0, in_dim=2, decoder_hidden_dim=128, nl="sine", encoder_type="none",
                                decoder_n_hidden_layers=4, init_type="mfgi")
    z_gt = np.zeros((48, 48))
    # (the reference builds the grid in float32 -- utils/utils.py:181 -- so only its fp32 run exists; the fp64 truth for
    #  the test comes from oracle/jet_spec.py, itself pinned by the step fixtures)
    x, y, z_np, z_diff, eik, div, curl = ref_utils.compute_deriv_props(net2.decoder, None, z_gt, None)
    for k, v in (("x", x), ("z", z_np), ("z_diff", z_diff), ("eikonal", eik), ("div", div), ("curl", curl)):
        deriv["f32/%s" % k] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, "deriv_props.npz"), **deriv)

    # ---- div-weight schedules ----------------------------------------------------------------------
    sched = {}
    for decay in ("linear", "quintic", "step", "none"):
        for pname, params in {"recipe": (1e2, 0.2, 1e2, 0.4, 0.0, 0.0), "scene": (1e1, 0.1, 1e1, 0.3, 0.0, 0.0),
                              "two": (50.0, 5.0)}.items():
            crit = ref_digs.DiGSLoss(weights=[3e3, 1e2, 1e2, 5e1, 1e2], loss_type="siren_wo_n_w_div",
                                     div_decay=decay, div_type="l1")
            vals = []
            for it in range(0, 1001, 25):
                crit.update_div_weight(it, 1000, params)
                vals.append(crit.weights[4])
            sched["%s/%s" % (decay, pname)] = np.array(vals, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "div_weight.npz"), **sched)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()

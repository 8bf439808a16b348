"""CPU: both oracle forms (closed-form jet, autograd port) against fixtures produced by the REAL reference
(tests/golden/step_*.npz, generator oracle/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from oracle import jet_spec, ref_autograd
from oracle.make_golden import STEP_CASES
from _util import golden, maxrel, build_net, np_params, head_of, param_names, sha

CASES = sorted(STEP_CASES)


def _latent_bias(net, lat, n_per_shape, dtype):
    """(bs*N, H) per-point extra bias of layer 0 from the latent columns."""
    W0 = net.decoder.fc_block.linears()[0].weight.detach().numpy().astype(dtype)
    d = net.in_dim
    eb = lat.astype(dtype) @ W0[:, d:].T
    return np.repeat(eb, n_per_shape, axis=0)


@pytest.mark.parametrize("name", CASES)
def test_weights_match_reference_bitwise(name):
    case = STEP_CASES[name]
    g = golden("step_%s.npz" % name)
    net = build_net(case["net"], case["seed"])
    for pname, p in net.named_parameters():
        assert sha(p) == str(g["sha_" + pname]), pname


@pytest.mark.parametrize("name", CASES)
def test_jet_spec_fp64_matches_reference(name):
    case = STEP_CASES[name]
    g = golden("step_%s.npz" % name)
    net = build_net(case["net"], case["seed"])
    params = np_params(net, np.float64)
    head = head_of(net)
    d = case["net"]["in_dim"]
    m = g["in_mnfld"].astype(np.float64).reshape(-1, d)
    nm = g["in_nonmnfld"].astype(np.float64).reshape(-1, d)
    nrm = g["in_normals"].astype(np.float64).reshape(-1, d)
    ebm = ebn = None
    lat = None
    if "in_latent" in g.files:
        lat = g["in_latent"]
        ebm = _latent_bias(net, lat, case["Nm"], np.float64)
        ebn = _latent_bias(net, lat, case["Nn"], np.float64)
    fm = jet_spec.forward(params, m, head, ebm)
    fn = jet_spec.forward(params, nm, head, ebn)
    tol = 1e-11
    assert maxrel(fm["f"], g["ref64_f_m"].reshape(-1)) < tol
    assert maxrel(fm["grad"], g["ref64_g_m"].reshape(-1, d)) < tol
    assert maxrel(fn["f"], g["ref64_f_n"].reshape(-1)) < tol
    assert maxrel(fn["grad"], g["ref64_g_n"].reshape(-1, d)) < tol
    assert maxrel(fn["lap"], g["ref64_lap_n"].reshape(-1)) < 1e-10
    lk = case["loss"]
    latent_reg = None
    if lat is not None:
        latent_reg = np.linalg.norm(lat.astype(np.float64), axis=-1).mean()   # models/DiGS.py:49
    terms, seeds = jet_spec.loss_and_seeds(fm["f"], fm["grad"], fn["f"], fn["grad"], fn["lap"], nrm, lk["weights"],
                                           lk["loss_type"], lk["div_type"], lk["div_clamp"], latent_reg)
    for key in ("loss", "sdf_term", "inter_term", "eikonal_term", "normals_loss", "div_loss"):
        ref = float(g["ref64_term_" + key][0])
        assert abs(float(terms[key]) - ref) <= 1e-10 * max(1.0, abs(ref)), key
    gm, ebm_bar = jet_spec.backward(params, fm, seeds["f_m"], seeds["g_m"], np.zeros_like(fm["f"]), head)
    gn, ebn_bar = jet_spec.backward(params, fn, seeds["f_n"], seeds["g_n"], seeds["lap_n"], head)
    names = param_names(net)
    for li, ((dWm, dbm), (dWn, dbn)) in enumerate(zip(gm, gn)):
        dW, db = dWm + dWn, dbm + dbn
        if lat is not None and li == 0:
            # latent columns of W0 and the latent itself receive gradient through the folded bias
            W0 = params[0][0]
            latm = np.repeat(lat.astype(np.float64), case["Nm"], axis=0)
            latn = np.repeat(lat.astype(np.float64), case["Nn"], axis=0)
            dW[:, d:] = ebm_bar.T @ latm + ebn_bar.T @ latn
        refW = g["ref64_grad_" + names[2 * li]]
        refb = g["ref64_grad_" + names[2 * li + 1]]
        assert maxrel(dW, refW) < 2e-6, names[2 * li]      # fixture gradients are stored rounded to fp32
        assert maxrel(db, refb) < 2e-6, names[2 * li + 1]
    if lat is not None:
        W0 = params[0][0]
        bs = lat.shape[0]
        gl = (ebm_bar.reshape(bs, case["Nm"], -1).sum(1) + ebn_bar.reshape(bs, case["Nn"], -1).sum(1)) @ W0[:, d:]
        nrm_l = np.linalg.norm(lat.astype(np.float64), axis=-1, keepdims=True)
        gl = gl + lk["weights"][5] * lat.astype(np.float64) / nrm_l / bs
        assert maxrel(gl, g["ref64_grad_latent"]) < 2e-6


@pytest.mark.parametrize("name", [c for c in CASES if "latent" not in c])
def test_autograd_port_matches_reference(name):
    case = STEP_CASES[name]
    g = golden("step_%s.npz" % name)
    net = build_net(case["net"], case["seed"]).double()
    params = [p.detach().clone().requires_grad_(True) for p in net.decoder.fc_block.flat_params()]
    d = case["net"]["in_dim"]
    m = torch.tensor(g["in_mnfld"].reshape(-1, d), dtype=torch.float64)
    nm = torch.tensor(g["in_nonmnfld"].reshape(-1, d), dtype=torch.float64)
    nrm = torch.tensor(g["in_normals"].reshape(-1, d), dtype=torch.float64)
    lk = case["loss"]
    terms, grads, jets = ref_autograd.step(params, m, nm, nrm, head_of(net), lk["weights"], lk["loss_type"],
                                           lk["div_type"], lk["div_clamp"])
    assert maxrel(jets["f_n"].numpy(), g["ref64_f_n"].reshape(-1)) < 1e-12
    assert maxrel(jets["g_n"].numpy(), g["ref64_g_n"].reshape(-1, d)) < 1e-12
    if jets["lap_n"] is not None:
        assert maxrel(jets["lap_n"].numpy(), g["ref64_lap_n"].reshape(-1)) < 1e-11
    assert abs(float(terms["loss"]) - float(g["ref64_term_loss"][0])) < 1e-9 * abs(float(g["ref64_term_loss"][0]))
    for name_p, gr in zip(param_names(net), grads):
        assert maxrel(gr.numpy(), g["ref64_grad_" + name_p]) < 2e-6, name_p


def test_reference_fp32_noise_floor_recorded():
    """The fixture also carries the reference's own fp32 result; its distance to fp64 bounds what 1e-4 can mean."""
    g = golden("step_c2_small.npz")
    assert maxrel(g["ref32_f_n"], g["ref64_f_n"]) < 1e-4
    assert maxrel(g["ref32_g_n"], g["ref64_g_n"]) < 1e-4
    assert maxrel(g["ref32_lap_n"], g["ref64_lap_n"]) < 1e-4


def test_grid_values_port():
    g = golden("grid.npz")
    net = build_net(dict(latent_size=0, in_dim=3, decoder_hidden_dim=64, nl="sine", encoder_type="none",
                         decoder_n_hidden_layers=2, init_type="mfgi", sphere_init_params=[1.6, 0.1]), 3627473)
    params = [p.detach() for p in net.decoder.fc_block.flat_params()]
    pts = torch.tensor(g["cube16/points_f16"].astype(np.float32))
    v = ref_autograd.grid_values(params, pts, head_of(net)).numpy()
    assert maxrel(v, g["cube16/values32"]) < 1e-5


def test_closed_form_oracle_on_trained_weights():
    """The oracle is pinned on weights after 200 reference training iterations too (tests/golden/trained_c2.npz)."""
    g = golden("trained_c2.npz")
    n_layers = len([k for k in g.files if k.startswith("w_") and k.endswith("weight")])
    params = [(g["w_decoder.fc_block.net.%d.0.weight" % i].astype(np.float64),
               g["w_decoder.fc_block.net.%d.0.bias" % i].astype(np.float64)) for i in range(n_layers)]
    case = STEP_CASES["c2_small"]
    res = jet_spec.step(params, g["in_mnfld"][0].astype(np.float64), g["in_nonmnfld"][0].astype(np.float64),
                        g["in_normals"][0].astype(np.float64), (1.6, 0.1), case["loss"]["weights"],
                        case["loss"]["loss_type"], case["loss"]["div_type"], case["loss"]["div_clamp"])
    assert maxrel(res["fn"]["f"], g["ref64_f_n"][0]) < 1e-9
    assert maxrel(res["fn"]["grad"], g["ref64_g_n"][0]) < 1e-9
    assert maxrel(res["fn"]["lap"], g["ref64_lap_n"][0]) < 1e-9
    assert abs(res["terms"]["loss"] - float(g["ref64_term_loss"][0])) < 1e-9 * abs(float(g["ref64_term_loss"][0]))
    for i, (dW, db) in enumerate(res["grads"]):
        assert maxrel(dW, g["ref64_grad_decoder.fc_block.net.%d.0.weight" % i]) < 1e-6     # fixture gradients are stored in fp32
        assert maxrel(db, g["ref64_grad_decoder.fc_block.net.%d.0.bias" % i]) < 1e-6

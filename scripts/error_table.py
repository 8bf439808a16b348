"""Measured parity errors per configuration and precision mode (GPU): max-norm-relative vs the fp64 closed-form oracle,
next to the fp32 noise floor (the same closed form evaluated in float32 on the same points).  Writes a markdown table."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import digs_b200
from oracle import jet_spec

def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

CONFIGS = [
    ("C1 2-D sanity 4x128 mfgi (1.6,1.0)", dict(latent_size=0, in_dim=2, decoder_hidden_dim=128, nl="sine", encoder_type="none", decoder_n_hidden_layers=4, init_type="mfgi"), 15000, 1.2),
    ("C2/C3 SRB 4x256 mfgi (1.6,0.1)", dict(latent_size=0, in_dim=3, decoder_hidden_dim=256, nl="sine", encoder_type="none", decoder_n_hidden_layers=4, init_type="mfgi", sphere_init_params=[1.6, 0.1]), 15000, 1.1),
    ("C4 scene 8x512 siren", dict(latent_size=0, in_dim=3, decoder_hidden_dim=512, nl="sine", encoder_type="none", decoder_n_hidden_layers=8, init_type="siren"), 4000, 1.1),
    ("C5 decoder 8x256 mfgi (no latent)", dict(latent_size=0, in_dim=3, decoder_hidden_dim=256, nl="sine", encoder_type="none", decoder_n_hidden_layers=8, init_type="mfgi", sphere_init_params=[1.6, 0.1]), 6000, 1.2),
]
W = [3e3, 1e2, 1e2, 5e1, 1e2]
TRAINED = None
_g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "trained_c2.npz")
if os.path.isfile(_g):
    _z = np.load(_g)
    TRAINED = {k[2:]: torch.tensor(_z[k]) for k in _z.files if k.startswith("w_")}
    CONFIGS.insert(2, ("C2 trained weights (200 reference iterations, torus)", CONFIGS[1][1], 15000, 1.1))
print("| configuration | points | mode | f | grad f | lap f | loss | max dW | max db |")
print("|---|---:|---|---:|---:|---:|---:|---:|---:|")
for name, kw, n, box in CONFIGS:
    d = kw["in_dim"]
    rng = np.random.RandomState(1)
    v = rng.normal(size=(n, d)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    m_np, nm_np = (0.5 * v).astype(np.float32), rng.uniform(-box, box, size=(n, d)).astype(np.float32)
    torch.manual_seed(3627473)
    base = digs_b200.DiGSNetwork(**kw)
    if TRAINED is not None and name.startswith("C2 trained"):
        base.load_state_dict(TRAINED)
        from digs_b200 import synth
        cloud, cn = synth.surface_cloud(100000, "torus", seed=0)
        mb, nb, nmb = synth.sample_batch(cloud, cn, n, seed=5, box=box)
        m_np, v, nm_np = mb[0], nb[0].astype(np.float64), nmb[0]
    head = tuple(base.decoder.fc_block.sphere_init_params) if kw["init_type"] != "siren" else None
    P64 = [(l.weight.detach().double().numpy(), l.bias.detach().double().numpy()) for l in base.decoder.fc_block.linears()]
    ref = jet_spec.step(P64, m_np.astype(np.float64), nm_np.astype(np.float64), v, head, W)
    P32 = [(a.astype(np.float32), b.astype(np.float32)) for a, b in P64]
    r32 = jet_spec.step(P32, m_np, nm_np, v.astype(np.float32), head, W)
    def row(mode, fn, terms, grads):
        dW = max(rel(g[0], r[0]) for g, r in zip(grads, ref["grads"]))
        db = max(rel(g[1], r[1]) for g, r in zip(grads, ref["grads"]))
        print("| %s | %d+%d | %s | %.1e | %.1e | %.1e | %.1e | %.1e | %.1e |" % (name, n, n, mode, rel(fn["f"], ref["fn"]["f"]),
              rel(fn["grad"], ref["fn"]["grad"]), rel(fn["lap"], ref["fn"]["lap"]), abs(terms["loss"] / ref["terms"]["loss"] - 1), dW, db), flush=True)
    row("float32 closed form (noise floor)", r32["fn"], r32["terms"], r32["grads"])
    for mode in ("fp32", "bf16x2", "bf16"):
        torch.manual_seed(3627473)
        net = digs_b200.DiGSNetwork(**kw, precision=mode).cuda()
        if TRAINED is not None and name.startswith("C2 trained"):
            net.load_state_dict(TRAINED)
        crit = digs_b200.DiGSLoss(copy.deepcopy(W), "siren_wo_n_w_div", "linear", "l1", 50)
        m = torch.tensor(m_np[None], device="cuda").requires_grad_(); nm = torch.tensor(nm_np[None], device="cuda").requires_grad_()
        out = net(nm, m)
        ld, _ = crit(out, m, nm, torch.tensor(v[None].astype(np.float32), device="cuda"))
        ld["loss"].backward()
        fn = dict(f=out["nonmanifold_pnts_pred"][0].detach().cpu().numpy(), grad=out["nonmanifold_pnts_grad"][0].detach().cpu().numpy(),
                  lap=out["nonmanifold_pnts_div"][0].detach().cpu().numpy())
        grads = [(l.weight.grad.cpu().numpy(), l.bias.grad.cpu().numpy()) for l in net.decoder.fc_block.linears()]
        row(mode, fn, dict(loss=float(ld["loss"].detach())), grads)
        if mode != "fp32":       # the same step through digs_step_fused (tile kernel + weight-gradient GEMM)
            torch.manual_seed(3627473)
            net = digs_b200.DiGSNetwork(**kw, precision=mode).cuda()
            if TRAINED is not None and name.startswith("C2 trained"):
                net.load_state_dict(TRAINED)
            crit = digs_b200.DiGSLoss(copy.deepcopy(W), "siren_wo_n_w_div", "linear", "l1", 50)
            ld, out, _ = digs_b200.fused_step(net, crit, torch.tensor(m_np[None], device="cuda"), torch.tensor(nm_np[None], device="cuda"),
                                              torch.tensor(v[None].astype(np.float32), device="cuda"))
            fn = dict(f=out["nonmanifold_pnts_pred"][0].cpu().numpy(), grad=out["nonmanifold_pnts_grad"][0].cpu().numpy(),
                      lap=out["nonmanifold_pnts_div"][0].cpu().numpy())
            grads = [(l.weight.grad.cpu().numpy(), l.bias.grad.cpu().numpy()) for l in net.decoder.fc_block.linears()]
            row(mode + " (fused step)", fn, dict(loss=float(ld["loss"])), grads)

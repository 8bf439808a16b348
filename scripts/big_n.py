"""Scale check (GPU): large clouds through the tensor-core path (index arithmetic, memory), plus tiny / empty clouds."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import digs_b200
from digs_b200 import jet as djet, _lib
from oracle import jet_spec

torch.manual_seed(0)
for (H, Lh, init, n) in ((256, 4, "mfgi", 1_000_000), (512, 8, "siren", 262_144)):
    net = digs_b200.DiGSNetwork(0, 3, H, "sine", "none", Lh, init, [1.6, 0.1], precision="bf16x2").cuda()
    crit = digs_b200.DiGSLoss([3e3, 1e2, 1e2, 5e1, 1e2], "siren_wo_n_w_div", "linear", "l1", 50)
    m = (torch.rand(1, n, 3, device="cuda") - 0.5).requires_grad_()
    nm = (torch.rand(1, n, 3, device="cuda") * 2.2 - 1.1).requires_grad_()
    nrm = torch.nn.functional.normalize(torch.randn(1, n, 3, device="cuda"), dim=-1)
    for it in range(2):
        net.zero_grad()
        torch.cuda.synchronize(); t0 = time.time()
        out = net(nm, m)
        ld, _ = crit(out, m, nm, nrm)
        ld["loss"].backward()
        torch.cuda.synchronize(); dt = time.time() - t0
    ok = all(torch.isfinite(p.grad).all().item() for p in net.parameters())
    flops = 3 * n * 9 * 2 * Lh * H * H
    # spot-check 2000 random points of the big cloud against the fp64 oracle
    sel = torch.randint(0, n, (2000,), device="cuda")
    params = [(l.weight.detach().cpu().double().numpy(), l.bias.detach().cpu().double().numpy()) for l in net.decoder.fc_block.linears()]
    ref = jet_spec.forward(params, nm[0, sel].detach().cpu().double().numpy(), (1.6, 0.1) if init != "siren" else None)
    e = float(np.abs(out["nonmanifold_pnts_div"][0, sel].detach().cpu().numpy() - ref["lap"]).max() / np.abs(ref["lap"]).max())
    print("H=%d Lh=%d n=%d+%d: step %.1f ms -> %.1f M pts/s, %.0f TFLOP/s algorithmic, grads finite %s, lap err %.1e, peak mem %.1f GB" % (
        H, Lh, n, n, dt * 1e3, 2 * n / dt / 1e6, flops / dt / 1e12, ok, e, torch.cuda.max_memory_allocated() / 2**30), flush=True)
    del net, m, nm, nrm, out, ld
    torch.cuda.empty_cache()
net = digs_b200.DiGSNetwork(0, 3, 256, "sine", "none", 4, "mfgi", [1.6, 0.1], precision="bf16x2").cuda()
fc = net.decoder.fc_block
params = [p.detach() for p in fc.flat_params()]
for n in (0, 1, 3, 23, 25):
    xyz = torch.rand(n, 3, device="cuda") - 0.5
    f, g, lap, saved = djet.jet_forward_raw(fc.spec(), params, xyz, None, 0, True, True, 1)
    gr, _ = (djet.jet_backward_raw(fc.spec(), params, xyz, None, 0, True, torch.ones_like(f), torch.ones_like(g), torch.ones_like(f), saved, 1)
             if n else ([torch.zeros(1)], None))
    torch.cuda.synchronize()
    print("n=%d ok, finite %s" % (n, all(torch.isfinite(t).all().item() for t in gr)))

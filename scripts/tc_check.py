"""Debug helper (GPU): tcgen05 forward vs the fp32 CUDA-core forward and the fp64 oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import digs_b200
from digs_b200 import _lib, jet as djet
from oracle import jet_spec

def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))

def run(d, H, Lh, init, n, want_lap, seed=0):
    torch.manual_seed(seed)
    net = digs_b200.DiGSNetwork(0, d, H, "sine", "none", Lh, init, [1.6, 0.1]).cuda()
    fc = net.decoder.fc_block
    params = [p.detach() for p in fc.flat_params()]
    xyz = (torch.rand(n, d, device="cuda") * 2.2 - 1.1)
    pn = [(l.weight.detach().cpu().double().numpy(), l.bias.detach().cpu().double().numpy()) for l in fc.linears()]
    head = (1.6, 0.1) if init != "siren" else None
    ref = jet_spec.forward(pn, xyz.cpu().double().numpy(), head)
    out = {}
    for name, mode in (("fp32", 0), ("bf16x2", 1), ("bf16", 2)):
        f, g, lap, saved = djet.jet_forward_raw(fc.spec(), params, xyz, None, 0, want_lap, True, mode)
        torch.cuda.synchronize()
        e = [rel(f.cpu(), ref["f"]), rel(g.cpu(), ref["grad"]), rel(lap.cpu(), ref["lap"]) if want_lap else -1]
        fv = djet.value_forward_raw(fc.spec(), params, xyz, None, 0, mode)
        e.append(rel(fv.cpu(), ref["f"]))
        out[name] = (e, saved)
        print("d=%d H=%d Lh=%d %-14s n=%6d lap=%d %-7s f %.2e grad %.2e lap %.2e value-only %.2e" % (d, H, Lh, init, n, want_lap, name, *e), flush=True)
    # saved planes of tc vs simt
    s0 = out["fp32"][1].view(torch.float32); s1 = out["bf16x2"][1].view(torch.float32)
    m = min(s0.numel(), s1.numel())
    print("   saved-plane rel diff bf16x2 vs fp32: %.2e" % rel(s1[:m].cpu(), s0[:m].cpu()), flush=True)

if __name__ == "__main__":
    run(3, 256, 4, "mfgi", 1000, True)
    run(3, 256, 4, "mfgi", 15000, True)
    run(3, 256, 4, "mfgi", 15000, False)
    run(3, 256, 4, "siren", 5003, True)
    run(2, 128, 4, "mfgi", 3001, True)
    run(2, 128, 4, "mfgi", 3001, False)
    run(3, 512, 8, "siren", 2000, True)
    run(3, 128, 2, "geometric_sine", 777, True)
    # timing
    torch.manual_seed(0)
    net = digs_b200.DiGSNetwork(0, 3, 256, "sine", "none", 4, "mfgi", [1.6, 0.1]).cuda()
    fc = net.decoder.fc_block
    params = [p.detach() for p in fc.flat_params()]
    xyz = (torch.rand(15000, 3, device="cuda") * 2.2 - 1.1)
    for name, mode in (("fp32", 0), ("bf16x2", 1), ("bf16", 2)):
        for save in (True, False):
            for _ in range(3):
                djet.jet_forward_raw(fc.spec(), params, xyz, None, 0, True, save, mode)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                djet.jet_forward_raw(fc.spec(), params, xyz, None, 0, True, save, mode)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print("forward jet 15000 pts C=5 %-7s save=%d: %.3f ms  -> %.1f TFLOP/s algorithmic" % (name, save, ms, 15000 * 5 * 2 * 4 * 256 * 256 / ms / 1e9), flush=True)
    # grid 256^3
    axes, _, _ = digs_b200.grid_axes(256, np.array([[-1, 1], [-1, 1], [-1, 1]], dtype=float))
    for prec in ("fp32", "bf16x2", "bf16"):
        net.precision = prec
        v = digs_b200.grid_query(net, axes); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); v = digs_b200.grid_query(net, axes); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if prec == "fp32": v0 = v
        print("grid 256^3 %-7s %.2f ms -> %.1f M queries/s, %.1f TFLOP/s; rel vs fp32 %.2e" % (prec, ms, 256**3 / ms / 1e3, 256**3 * 0.526e6 / ms / 1e9, rel(v.cpu(), v0.cpu())), flush=True)

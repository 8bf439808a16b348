"""Timeline helper (GPU): kernel start / duration per stream for one eager training step (torch.profiler / CUPTI)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import digs_b200
from digs_b200 import dist as ddist

torch.manual_seed(3627473)
net = digs_b200.DiGSNetwork(0, 3, 256, "sine", "none", 4, "mfgi", [1.6, 0.1], precision="bf16x2").cuda()
crit = digs_b200.DiGSLoss([3e3, 1e2, 1e2, 5e1, 1e2], "siren_wo_n_w_div", "linear", "l1", 50)
flat = ddist.FlatGrads(list(net.parameters()), extra=8)
net.decoder.fc_block.accumulate_grads_in_place = True
m = torch.randn(1, 15000, 3, device="cuda"); m = 0.5 * m / m.norm(dim=-1, keepdim=True)
nrm = m * 2
nm = torch.rand(1, 15000, 3, device="cuda") * 2.2 - 1.1
m.requires_grad_(); nm.requires_grad_()

def step():
    flat.zero_()
    out = net(nm, m)
    ld, _ = crit(out, m, nm, nrm)
    ld["loss"].backward()

for _ in range(5): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# last step only
if evs:
    t_end = evs[-1].time_range.end
    starts = [e for e in evs if "FillFunctor" in e.name]
    t0 = evs[0].time_range.start
    # find beginning of the last step: the last zero-fill of the flat buffer that precedes a k_tc_prep
    cut = [i for i, e in enumerate(evs) if "k_tc_prep" in e.name]
    first = cut[-4] - 1 if len(cut) >= 4 else 0
    base = evs[first].time_range.start
    for e in evs[first:]:
        print("%8.1f us  +%7.1f us  stream %-3s %s" % (e.time_range.start - base, e.time_range.end - e.time_range.start,
                                                     getattr(e, "device_index", "?"), e.name[:70]))
    print("step span %.1f us" % (t_end - base))

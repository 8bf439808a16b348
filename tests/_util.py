"""Shared helpers for the test-suite (test infrastructure)."""
import hashlib
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def maxrel(a, b):
    """max-norm relative error ||a-b||_inf / ||b||_inf (SURVEY 7.3-1 definition of the tolerance)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    if den == 0:
        return float(np.abs(a - b).max())
    return float(np.abs(a - b).max() / den)


def sha(t):
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


def build_net(net_kwargs, seed, **extra):
    import digs_b200
    torch.manual_seed(seed)
    return digs_b200.DiGSNetwork(**net_kwargs, **extra)


def np_params(net, dtype=np.float64):
    lins = net.decoder.fc_block.linears()
    return [(l.weight.detach().cpu().numpy().astype(dtype), l.bias.detach().cpu().numpy().astype(dtype)) for l in lins]


def head_of(net):
    fc = net.decoder.fc_block
    if fc.init_type in ("mfgi", "geometric_sine"):
        return tuple(float(x) for x in fc.sphere_init_params)
    return None


def param_names(net):
    return [n for n, _ in net.named_parameters()]

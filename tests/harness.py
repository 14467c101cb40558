"""Test helper: drive the device Engine (MPDOSimulator/_engine/steps.py) from the layer list of an
OracleCircuit, so the same program runs through the oracle and through the engine under test."""
import math

import torch

from MPDOSimulator._engine.steps import Engine


def run_engine(oc, prims, dtype, device='cpu', npass=None, env_sweep=False):
    E = Engine(prims, dtype, npass)
    n = oc.qn
    Ts = [torch.tensor([1, 0], dtype=dtype, device=device).reshape(1, 1, 2, 1, 1) for _ in range(n)]
    bond = [False] * (n - 1)
    inner = [False] * n
    for layer in oc.layers:
        if layer[0] == 'truncate':
            if n <= 1 or all(bond):
                if not (oc.chi is None and oc.max_truncation_err is None):
                    if env_sweep and oc.max_truncation_err is None:
                        E.bond_truncate_env(Ts, oc.chi)
                    else:
                        E.qr_left2right(Ts)
                        E.svd_right2left(Ts, oc.chi, oc.max_truncation_err)
                if not oc.ideal and not (oc.kappa is None and oc.max_truncation_err is None):
                    E.svd_kappa(Ts, oc.kappa, oc.max_truncation_err, inner)
        elif layer[0] == 'barrier':
            pass
        else:
            _, name, G, single, var, ideal, oqs = layer
            if single:
                noisy = (oc.idealNoise or oc.unified) and not ideal
                if noisy:
                    G = torch.einsum('nlm, ljk, ji -> nimk', oc.noise['decay'], oc.noise['dephasing'], G).reshape(2, 2, -1)
                else:
                    G = G.reshape(2, 2, 1)
                for q in oqs:
                    Ts[q] = E.absorb_1q(Ts[q], G.reshape(1, 2, 2, -1).contiguous().to(device))
                    if noisy:
                        inner[q] = True
            else:
                lo, hi = min(oqs), max(oqs)
                g_noise = (oc.idealNoise and not ideal) or oc.realNoise
                if g_noise and not oc.realNoise:
                    G = torch.einsum('ijklp, klmn -> ijmnp', oc.noise['dpc2'], G)
                if G.dim() == 4:
                    G = G.unsqueeze(-1)
                if oqs[0] != lo:
                    G = G.permute(1, 0, 3, 2, 4)
                Ts[lo], Ts[hi] = E.split_2q(Ts[lo], Ts[hi], G.unsqueeze(0).contiguous().to(device))
                bond[lo] = True
                if g_noise:
                    inner[hi] = True
    if not oc.ideal and oc.layers[-1][0] != 'truncate' and not (oc.kappa is None and oc.max_truncation_err is None):
        E.svd_kappa(Ts, oc.kappa, oc.max_truncation_err, inner)
    return E, Ts


def brickwork(oc, n, depth, seed=0, ghz_prefix=True, entangler='cz'):
    g = torch.Generator().manual_seed(seed)
    if ghz_prefix:
        oc.h(0)
        for i in range(n - 1):
            oc.cnot(i, i + 1)
        oc.truncate()
    for d in range(depth):
        for q in range(n):
            th, ph, la = (torch.rand(3, generator=g) * 2 * math.pi).tolist()
            oc.u3(th, ph, la, [q])
        oc.truncate()
        for q in range(d % 2, n - 1, 2):
            getattr(oc, entangler)(q, q + 1)
        oc.truncate()


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()

"""Helpers shared by the -m gpu parity tests: build Vars / Params on the device, run an op under the
tape with a random upstream gradient, and compare against torch-CPU float64 autograd of the oracle."""
import numpy as np
import torch

import tgan
from tgan import core, ops


def setup(math='fp32', injected=None):
    tgan.init('cuda:0', math=math)
    core.ctx.store = core.VariableStore()
    if injected is not None:
        core.ctx.rng = core.InjectedSource(injected)
    return core.ctx


def param(a, requires_grad=True, name='p'):
    a = np.asarray(a, np.float32)
    p = core.Param(name, a.shape, True, a)
    p.data = torch.from_numpy(a.copy()).cuda()
    p.grad = torch.zeros_like(p.data)
    p.requires_grad = requires_grad
    return p


def var(a, requires_grad=False, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, np.float32))).cuda()
    if dtype is not None:
        t = t.to(dtype)
    return ops.Var(t, tuple(t.shape), requires_grad=requires_grad)


def run_bwd(out, gy):
    """seed out.grad = gy (numpy) and run the current tape."""
    out.grad = torch.from_numpy(np.ascontiguousarray(gy.astype(np.float32))).cuda().to(out.data.dtype)
    core.ctx.tape.backward()


def relerr(a, b):
    """max |a-b| / max|b| (relative-to-max-abs error, SURVEY.md §8c parity contract)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def tnp(t):
    return t.detach().float().cpu().numpy()

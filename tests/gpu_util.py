import numpy as np
import pytest
import torch

gpu = pytest.mark.gpu
DEV = "cuda"
TDT = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}
# the reference's own fp tolerances (tests/ops_test.cc:1434-1445)
TOL = {"float32": 1e-5, "float16": 1e-2, "bfloat16": 4e-2}


def to_np(t):
    return t.detach().float().cpu().numpy() if t.is_floating_point() else t.detach().cpu().numpy()


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t.to(dtype) if dtype is not None else t


def round_through(a, name):
    """numpy float32 array rounded through the torch dtype `name` (what the device holds)."""
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(TDT[name]).float().numpy()

"""CPU: the scalar math the loss kernels run on the device (csrc/loss_math.h), compiled for the host with g++,
against torch autograd of the oracle's CIoU.  Catches derivation bugs of the hand-written backward without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import port

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hm(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hm") / "libhostmath.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                           os.path.join(HERE, "hostmath", "hostmath.cpp")])
    lib = C.CDLL(so)
    lib.hm_bce.restype = C.c_float
    lib.hm_bce.argtypes = [C.c_float, C.c_float]
    return lib


def test_row_ciou_forward_backward(hm):
    r = np.random.RandomState(3)
    n = 4096
    logits = (r.standard_normal((n, 4)) * 2).astype(np.float32)
    anch = r.uniform(1, 12, (n, 2)).astype(np.float32)
    tbox = np.concatenate([r.uniform(-0.5, 1.5, (n, 2)), r.uniform(0.2, 40, (n, 2))], 1).astype(np.float32)
    ciou = np.zeros(n, np.float32)
    grad = np.zeros((n, 4), np.float32)
    fp = C.POINTER(C.c_float)
    hm.hm_row_ciou(logits.ctypes.data_as(fp), anch.ctypes.data_as(fp), tbox.ctypes.data_as(fp), n,
                   ciou.ctypes.data_as(fp), grad.ctypes.data_as(fp))
    lt = torch.from_numpy(logits).requires_grad_(True)
    pxy = lt[:, :2].sigmoid() * 2.0 - 0.5
    pwh = (lt[:, 2:4].sigmoid() * 2) ** 2 * torch.from_numpy(anch)
    c = port.ciou(torch.cat([pxy, pwh], 1), torch.from_numpy(tbox))
    c.sum().backward()
    np.testing.assert_allclose(ciou, c.detach().numpy(), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(grad, lt.grad.numpy(), rtol=2e-4, atol=2e-6)


def test_bce(hm):
    for x, z in ((-8.0, 0.0), (0.3, 1.0), (12.0, 0.37), (0.0, 0.5)):
        ref = torch.nn.functional.binary_cross_entropy_with_logits(torch.tensor([x]), torch.tensor([z])).item()
        assert abs(hm.hm_bce(x, z) - ref) < 1e-6

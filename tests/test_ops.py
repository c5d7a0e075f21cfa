"""CPU tests of ``adaptdl_b200.ops`` and the mixed-precision helpers: on a
machine without a GPU every op is the plain PyTorch composition, with the
same signature, state-dict layout and numerics as the modules it replaces
(the sm_100a kernels are compared against fp32 references in
``tests/test_gpu_kernels.py``)."""

import io


import torch
import torch.nn.functional as F

import adaptdl_b200.torch as adl
from adaptdl_b200 import ops
from adaptdl_b200.models.bert import aligned_linear


def test_linear_gelu_fallback_matches_composition():
    torch.manual_seed(0)
    layer = ops.LinearGELU(64, 128)
    plain = torch.nn.Linear(64, 128)
    plain.load_state_dict(layer.state_dict())       # same parameter names
    x = torch.randn(3, 5, 64, requires_grad=True)
    y = layer(x)
    assert torch.allclose(y, F.gelu(plain(x)), atol=1e-6)
    y.sum().backward()
    assert layer.weight.grad is not None and x.grad is not None
    assert torch.allclose(ops.linear_act(x, plain.weight, plain.bias, None),
                          plain(x), atol=1e-6)


def test_batch_norm_act_is_a_drop_in_batchnorm():
    torch.manual_seed(1)
    fused = ops.BatchNormAct2d(8)
    plain = torch.nn.BatchNorm2d(8)
    plain.load_state_dict(fused.state_dict())
    assert list(fused.state_dict()) == list(plain.state_dict())
    x = torch.randn(4, 8, 5, 5)
    r = torch.randn(4, 8, 5, 5)
    for residual, relu in ((None, True), (r, True), (r, False),
                           (None, False)):
        want = plain(x)
        if residual is not None:
            want = want + residual
        if relu:
            want = torch.relu(want)
        got = fused(x, residual, relu)
        assert torch.allclose(got, want, atol=1e-6)
    assert int(fused.num_batches_tracked) == int(plain.num_batches_tracked)
    assert torch.allclose(fused.running_var, plain.running_var)
    fused.eval(), plain.eval()
    assert torch.allclose(fused(x, None, False), plain(x), atol=1e-6)
    assert int(fused.num_batches_tracked) == 4       # eval does not count
    # cumulative moving average (momentum=None) keeps PyTorch's behaviour
    a, b = ops.BatchNormAct2d(8, momentum=None), \
        torch.nn.BatchNorm2d(8, momentum=None)
    a(x), b(x)
    assert torch.allclose(a.running_mean, b.running_mean)
    assert int(a.num_batches_tracked) == 1


def test_dropout_add_layer_norm_fallback():
    torch.manual_seed(2)
    x, h = torch.randn(6, 16), torch.randn(6, 16)
    w, b = torch.rand(16) + 0.5, torch.randn(16)
    want = F.layer_norm(x + h, (16,), w, b, 1e-5)
    assert torch.allclose(
        ops.dropout_add_layer_norm(x, h, w, b, 0.3, False), want, atol=1e-6)
    mask = (torch.rand(6, 16) > 0.3).to(torch.uint8)
    got = ops.dropout_add_layer_norm(x, h, w, b, 0.3, True, mask=mask)
    want = F.layer_norm(x + h * mask.float() / 0.7, (16,), w, b, 1e-5)
    assert torch.allclose(got, want, atol=1e-5)


def test_aligned_linear_is_linear():
    torch.manual_seed(3)
    x, w, b = torch.randn(4, 7, 16), torch.randn(13, 16), torch.randn(13)
    assert torch.equal(aligned_linear(x, w, b), F.linear(x, w, b))


def test_mixed_precision_params_casts_matrices_only():
    net = torch.nn.Sequential(
        torch.nn.Conv2d(3, 8, 3, bias=True), torch.nn.BatchNorm2d(8),
        torch.nn.Flatten(), torch.nn.Linear(8 * 30 * 30, 4))
    net = net.to(memory_format=torch.channels_last)
    assert adl.mixed_precision_params(net) is net
    conv, bn, _, fc = net
    assert conv.weight.dtype == torch.bfloat16
    assert conv.weight.is_contiguous(memory_format=torch.channels_last)
    assert fc.weight.dtype == torch.bfloat16
    assert conv.bias.dtype == bn.weight.dtype == fc.bias.dtype == \
        torch.float32
    assert bn.running_mean.dtype == torch.float32
    # an optimizer built afterwards sees the 16-bit parameters
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    dtypes = {p.dtype for g in opt.param_groups for p in g["params"]}
    assert dtypes == {torch.bfloat16, torch.float32}


def test_checkpoint_state_keeps_reference_layout_without_engine():
    """Host path (no device engine): the saved object is the reference's
    3-tuple with a 4-entry state list; a fifth entry only appears with the
    engine's fp32 masters, and the loader accepts both."""
    from adaptdl_b200.torch.parallel import _AdaptiveDataParallelState
    model = torch.nn.Linear(4, 2)
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9)
    opt.state["gns"] = {"progress": 0.0}
    state = _AdaptiveDataParallelState(model, opt, None, None, "ckpt-layout")
    buf = io.BytesIO()
    state.save(buf)
    dicts, gain, lr_factor = torch.load(io.BytesIO(buf.getvalue()),
                                        weights_only=False)
    assert len(dicts) == 4 and gain == 1.0 and lr_factor == 1.0
    state.load(io.BytesIO(buf.getvalue()))
    assert state.wide_state is None
    # a checkpoint written by an engine run carries the extra entry
    dicts.append({(0, 0, "master"): torch.zeros(2, 4)})
    buf2 = io.BytesIO()
    torch.save((dicts, gain, lr_factor), buf2)
    state.load(io.BytesIO(buf2.getvalue()))
    assert (0, 0, "master") in state.wide_state


def test_gemm_epilogue_gelu_polynomial_meets_its_error_bound():
    """The tcgen05 GEMM's epilogue evaluates GELU as x * sat(0.5 + 0.5 * x *
    Q(x^2)) with the coefficients in csrc/adl_gemm.cu; check the documented
    bounds (|Phi error| < 1.4e-5, |GELU error| < 6e-5) in fp32 arithmetic,
    including far outside the fitted range where the saturation supplies the
    tails."""
    import os
    import re
    import numpy as np
    from math import erf, sqrt
    src = open(os.path.join(os.path.dirname(__file__), "..", "csrc",
                            "adl_gemm.cu")).read()
    body = src[src.index("uint64_t gelu2("):]
    body = body[:body.index("float t0, t1;")]
    coefs = [np.float32(c) for c in re.findall(
        r"pk\((-?[0-9.]+e[-+][0-9]+)f,", body)]
    assert len(coefs) == 9                     # degree 8 in s = x^2
    x = np.concatenate([np.linspace(-8, 8, 400001),
                        np.logspace(1, 18, 2000), -np.logspace(1, 18, 2000)
                        ]).astype(np.float32)
    with np.errstate(over="ignore", invalid="ignore"):
        s = x * x
        q = np.full_like(s, coefs[0])
        for c in coefs[1:]:
            q = (q * s + c).astype(np.float32)
        t = (x * q).astype(np.float32)
        phi = np.clip(t * np.float32(0.5) + np.float32(0.5), 0.0, 1.0)
        phi = np.where(np.isnan(phi), 0.0, phi).astype(np.float32)
        y = (x * phi).astype(np.float64)
    xd = x.astype(np.float64)
    ref_phi = np.array([0.5 * (1.0 + erf(v / sqrt(2.0))) for v in xd])
    assert np.abs(phi - ref_phi).max() < 1.4e-5
    finite = np.abs(xd) < 1e6
    assert np.abs(y[finite] - (xd * ref_phi)[finite]).max() < 6e-5
    # exact tails: identity for large x, zero for very negative x
    assert np.all(y[xd > 6] == xd[xd > 6])
    assert np.all(y[xd < -6] == 0.0)


def test_bn_launch_geometry(monkeypatch):
    """Host-side launch arithmetic of the fused BN kernels for the layer
    shapes of the model zoo (no GPU needed: the SM count is injected)."""
    import importlib
    bn_act = importlib.import_module("adaptdl_b200.ops.bn_act")
    dev = torch.device("cuda", 0)
    monkeypatch.setitem(bn_act._SM, 0, 148)
    for m, c, vec in ((131072, 64, 8), (32768, 128, 8), (8192, 256, 8),
                      (2048, 512, 8), (2048, 512, 4), (7, 16, 8),
                      (100000, 32, 4), (64, 2048, 8)):
        cb, gy = bn_act._reduce_grid(dev, m, c, vec, 4)
        assert cb == min(c, 64) and c % cb == 0
        assert 256 % (cb // vec) == 0           # threads per row divide the CTA
        assert 1 <= gy and gy * (c // cb) <= 2 * 148 + (c // cb)
        rows_per_cta_iter = (256 // (cb // vec)) * 4
        assert gy <= max(1, -(-m // rows_per_cta_iter))    # no idle CTAs
        ga = bn_act._grid(dev, m, c, vec, 4)
        assert 1 <= ga <= 8 * 148


def test_ln_and_gemm_support_predicates():
    import importlib
    layer_norm = importlib.import_module("adaptdl_b200.ops.layer_norm")
    linear_act = importlib.import_module("adaptdl_b200.ops.linear_act")
    cpu = torch.randn(4, 768)
    assert not layer_norm.supported(cpu)        # CPU -> PyTorch composition
    assert not linear_act.supported(cpu, torch.randn(3072, 768))
    # shape rules of the GEMM kernel (checked before any device work)
    class Fake(object):
        is_cuda = True
        shape = (16, 768)

        def numel(self):
            return 16 * 768
    assert linear_act.supported(Fake(), torch.empty(3072, 768))
    assert not linear_act.supported(Fake(), torch.empty(3000, 768))   # N % 128
    assert not linear_act.supported(Fake(), torch.empty(3072, 700))   # K mismatch







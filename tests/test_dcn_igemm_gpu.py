"""The tcgen05 implicit-GEMM forward of the wide mask-head layers (devis_dcn_igemm_forward, dcn_igemm.cuh) against the
torchvision fixtures (float64, tests/golden/dcn_*.npz) and against this library's im2col + cuBLAS form at a real
mask-head layer shape.  3xTF32 (default): forward <= 1e-5 on the fixtures (<= 2e-5 at K = 2376, see below), gradients <= 1e-4;
single-pass TF32 (torch.backends.cuda.matmul.allow_tf32): <= 2e-3, the TF32 rounding of two operands."""
import numpy as np
import pytest
import torch

from conftest import load_golden, nmax

pytestmark = pytest.mark.gpu

IGEMM_CASES = ["dcn_fused_c72_o32", "dcn_fused_c40_o64", "dcn_fused_c8_o4_s2", "dcn_fused_c136_o8_k1",
               "dcn_fused_c16_o16_s2", "dcn_fused_c32_o16"]
GRAD_KEYS = (("x", "gx"), ("offset", "goffset"), ("weight", "gweight"), ("bias", "gbias"), ("mask", "gmask"))


@pytest.fixture(autouse=True)
def _no_tf32():
    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def _run(g, fn):
    t = lambda k: torch.from_numpy(g[k]).to("cuda", torch.float32)
    st, pd, dl, use_mask = [int(v) for v in g["cfg"]]
    names = ["x", "offset", "weight", "bias"] + (["mask"] if use_mask else [])
    leaves = [t(k).clone().requires_grad_(True) for k in names]
    out = fn.apply(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4] if use_mask else None, (st, st), (pd, pd), (dl, dl))
    out.backward(t("gout"))
    return out.detach(), dict(zip(names, [x.grad for x in leaves]))


@pytest.mark.parametrize("name", IGEMM_CASES)
def test_igemm_matches_torchvision_fixture(name):
    from devis_b200 import _lib
    from devis_b200.deform_conv import IGemmDeformConv2dFunction
    g = load_golden(name)
    cout, cin, kh, kw = g["weight"].shape
    assert _lib.load().devis_dcn_igemm_supported(cin, cout, kh, kw, _lib.F32)
    before = _lib.kernel_launches(_lib.KERNEL_DCN_IGEMM)
    out, grads = _run(g, IGemmDeformConv2dFunction)
    assert _lib.kernel_launches(_lib.KERNEL_DCN_IGEMM) == before + 1
    assert nmax(out.cpu().numpy(), g["out"]) < 1e-5
    for k, key in GRAD_KEYS:
        if k in grads:
            assert nmax(grads[k].cpu().numpy(), g[key]) < 1e-4, k


@pytest.mark.parametrize("c,cout,hw,n", [(264, 264, (12, 20), 6), (264, 128, (12, 20), 6), (136, 64, (23, 40), 3),
                                         (72, 32, (45, 80), 2)])
def test_igemm_at_mask_head_layer_shapes(c, cout, hw, n):
    """the three wide layers of MaskHeadConv (deformable_segmentation.py:323-380) and the 72 -> 32 layer: dispatch picks
    the tensor-core forward, and it agrees with the CUDA-core forms (3xTF32) / within TF32 rounding (allow_tf32)"""
    from devis_b200 import _lib, deform_conv
    from devis_b200.deform_conv import deform_conv2d
    gen = torch.Generator(device="cuda").manual_seed(c + cout)
    h, w = hw
    x = torch.randn(n, c, h, w, device="cuda", generator=gen)
    off = 1.5 * torch.randn(n, 18, h, w, device="cuda", generator=gen)
    msk = torch.rand(n, 9, h, w, device="cuda", generator=gen) * 2
    wt = torch.randn(cout, c, 3, 3, device="cuda", generator=gen) / (9 * c) ** 0.5
    b = torch.randn(cout, device="cuda", generator=gen)
    old = deform_conv.set_tensor_core(False)
    try:
        with torch.no_grad():
            want = deform_conv2d(x, off, wt, b, padding=1, mask=msk)
    finally:
        deform_conv.set_tensor_core(old)
    before = _lib.kernel_launches(_lib.KERNEL_DCN_IGEMM)
    with torch.no_grad():
        got = deform_conv2d(x, off, wt, b, padding=1, mask=msk)
    assert _lib.kernel_launches(_lib.KERNEL_DCN_IGEMM) == before + 1          # the dispatch took the tcgen05 kernel
    # Tolerance at the REAL contraction length (K = 9 x 264 = 2376): 2e-5 of max|out| against float64.  Measured
    # (benchmarks/igemm_debug.py, profiles/r2d_igemm_precision.txt): cuBLAS fp32 1.1-1.3e-6, 3xTF32 5.7e-6 (K = 1224) to
    # 1.3e-5 (K = 2376), of which 5.8e-6 is the tensor core's own fp32 accumulation (operands exact in TF32, plain 3 x 3
    # convolution: every product exact) -- tcgen05 adds each 8-term block to the accumulator with less than IEEE
    # round-to-nearest care, which no operand split repairs.  The small fixtures (K <= 1224) hold 1e-5.
    with torch.no_grad():
        want64 = deform_conv2d(x.double(), off.double(), wt.double(), b.double(), padding=1, mask=msk.double())
    assert nmax(want.cpu().numpy(), want64.cpu().numpy()) < 5e-6              # the cuBLAS fp32 form
    assert nmax(got.cpu().numpy(), want64.cpu().numpy()) < 2e-5               # 3xTF32 on the tensor cores
    torch.backends.cuda.matmul.allow_tf32 = True
    with torch.no_grad():
        fast = deform_conv2d(x, off, wt, b, padding=1, mask=msk)
    err = nmax(fast.cpu().numpy(), want.cpu().numpy())
    assert 1e-6 < err < 2e-3, err                                             # a TF32 pass really ran, within its rounding

    # training path: gradients of the tensor-core form == gradients of the im2col form (same backward kernels)
    torch.backends.cuda.matmul.allow_tf32 = False
    leaves = [t.clone().requires_grad_(True) for t in (x, off, wt, b, msk)]
    gout = torch.randn_like(want)
    launches = _lib.kernel_launches(_lib.KERNEL_DCN_IGEMM)
    deform_conv2d(leaves[0], leaves[1], leaves[2], leaves[3], padding=1, mask=leaves[4]).backward(gout)
    # with gradients the wide layers stay on the tensor cores (columns recomputed); 72 -> 32 keeps its columns (im2col form)
    assert _lib.kernel_launches(_lib.KERNEL_DCN_IGEMM) == launches + (1 if c >= 128 else 0)
    old = deform_conv.set_tensor_core(False)
    try:
        ref = [t.clone().requires_grad_(True) for t in (x, off, wt, b, msk)]
        deform_conv2d(ref[0], ref[1], ref[2], ref[3], padding=1, mask=ref[4]).backward(gout)
    finally:
        deform_conv.set_tensor_core(old)
    for a, r in zip(leaves, ref):
        assert nmax(a.grad.cpu().numpy(), r.grad.cpu().numpy()) < 1e-4

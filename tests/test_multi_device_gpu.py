"""Single-process use of a second GPU (nn.DataParallel-style, per-thread devices): the library keeps no device-global state
for the attention op and PER-DEVICE state for the constant-bank deformable-conv kernels.  Skipped on a
one-GPU box (the driver's GPU tier); run with `gpurun --gpus 2`."""
import pytest
import torch

from conftest import load_golden, nmax

pytestmark = pytest.mark.gpu


def _need_two():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")


def test_attention_op_on_the_second_device_matches_the_first():
    _need_two()
    from devis_b200 import MSDeformAttnFunction
    g = load_golden("op_d32")
    res = []
    for dev in ("cuda:0", "cuda:1"):
        t = lambda k, dt=torch.float32: torch.from_numpy(g[k]).to(dev, dt)
        v, loc, aw = (t(k).requires_grad_(True) for k in ("value", "loc", "aw"))
        shapes, lsi = torch.from_numpy(g["shapes"]).to(dev), torch.from_numpy(g["lsi"]).to(dev)
        out = MSDeformAttnFunction.apply(v, shapes, lsi, loc, aw, 64)
        out.backward(t("gout"))
        res.append([x.detach().cpu() for x in (out, v.grad, loc.grad, aw.grad)])
    assert nmax(res[1][0].numpy(), g["out"]) < 1e-5
    for a, b in zip(*res):
        assert nmax(a.numpy(), b.numpy()) < 1e-5          # same kernels, float-atomic order aside


def test_deform_conv_forms_on_both_devices_interleaved():
    _need_two()
    from devis_b200.deform_conv import deform_conv2d
    outs = {}
    for name in ("dcn_fused_c32_o16", "dcn_fused_c72_o32", "dcn_fused_c16_o1"):
        g = load_golden(name)
        st, pd, dl, use_mask = [int(v) for v in g["cfg"]]
        for rep in range(2):                                # alternate devices: constant bank + events are per device
            for dev in ("cuda:1", "cuda:0"):
                t = lambda k: torch.from_numpy(g[k]).to(dev, torch.float32)
                with torch.no_grad(), torch.cuda.device(dev):
                    out = deform_conv2d(t("x"), t("offset"), t("weight"), t("bias"), stride=st, padding=pd, dilation=dl,
                                        mask=t("mask") if use_mask else None)
                outs[(name, dev, rep)] = out.cpu()
        for key, out in outs.items():
            if key[0] == name:
                assert nmax(out.numpy(), g["out"]) < 1e-5, key


def test_tensor_core_deform_conv_on_the_second_device():
    _need_two()
    from devis_b200 import _lib
    from devis_b200.deform_conv import deform_conv2d
    g = load_golden("dcn_fused_c40_o64")
    st, pd, dl, use_mask = [int(v) for v in g["cfg"]]
    t = lambda k: torch.from_numpy(g[k]).to("cuda:1", torch.float32)
    x = t("x").repeat(1, 4, 1, 1)[:, :136].contiguous()      # 136 channels: not served by the CUDA-core fused forms
    w = torch.randn(64, 136, 3, 3, device="cuda:1") * 0.03
    before = _lib.kernel_launches(_lib.KERNEL_DCN_IGEMM)
    with torch.no_grad(), torch.cuda.device("cuda:1"):
        got = deform_conv2d(x, t("offset"), w, None, stride=st, padding=pd, dilation=dl, mask=t("mask") if use_mask else None)
        want = deform_conv2d(x.double(), t("offset").double(), w.double(), None, stride=st, padding=pd, dilation=dl,
                             mask=t("mask").double() if use_mask else None)
    assert _lib.kernel_launches(_lib.KERNEL_DCN_IGEMM) == before + 1
    assert nmax(got.cpu().numpy(), want.cpu().numpy()) < 1e-5

"""Quick look at the tcgen05 implicit-GEMM forward: errors per fixture and timing against the im2col + cuBLAS form."""
import os, sys, statistics
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import load_golden, nmax
from devis_b200 import _lib, deform_conv
from devis_b200.deform_conv import IGemmDeformConv2dFunction, deform_conv2d

torch.backends.cuda.matmul.allow_tf32 = False
for name in ["dcn_fused_c8_o4_s2", "dcn_fused_c32_o16", "dcn_fused_c72_o32", "dcn_fused_c40_o64", "dcn_fused_c136_o8_k1", "dcn_fused_c16_o16_s2"]:
    g = load_golden(name)
    t = lambda k: torch.from_numpy(g[k]).to("cuda", torch.float32)
    st, pd, dl, use_mask = [int(v) for v in g["cfg"]]
    with torch.no_grad():
        out = IGemmDeformConv2dFunction.apply(t("x"), t("offset"), t("weight"), t("bias"), t("mask") if use_mask else None, (st, st), (pd, pd), (dl, dl))
    torch.cuda.synchronize()
    print(name, "fwd err", nmax(out.cpu().numpy(), g["out"]), flush=True)

def med(fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev) * 1e3

for c, cout, (h, w), n in [(264, 264, (12, 20), 60), (264, 128, (12, 20), 60), (136, 64, (23, 40), 60)]:
    gen = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(n, c, h, w, device="cuda", generator=gen).contiguous(memory_format=torch.channels_last)
    off = 1.5 * torch.randn(n, 18, h, w, device="cuda", generator=gen)
    msk = torch.rand(n, 9, h, w, device="cuda", generator=gen)
    wt = torch.randn(cout, c, 3, 3, device="cuda", generator=gen) / (9 * c) ** 0.5
    b = torch.randn(cout, device="cuda", generator=gen)
    row = {}
    with torch.no_grad():
        deform_conv.set_tensor_core(False)
        want = deform_conv2d(x, off, wt, b, padding=1, mask=msk)
        want64 = deform_conv2d(x.double(), off.double(), wt.double(), b.double(), padding=1, mask=msk.double())
        row["err_cublas_fp32_vs_fp64"] = nmax(want.cpu().numpy(), want64.cpu().numpy())
        row["im2col_cublas_fp32_us"] = med(lambda: deform_conv2d(x, off, wt, b, padding=1, mask=msk))
        deform_conv.set_tensor_core(True)
        got = deform_conv2d(x, off, wt, b, padding=1, mask=msk)
        row["err_3xtf32"] = nmax(got.cpu().numpy(), want.cpu().numpy())
        row["err_3xtf32_vs_fp64"] = nmax(got.cpu().numpy(), want64.cpu().numpy())
        d = (got.double() - want64)
        row["bias_3xtf32"] = float((d * want64.sign()).mean() / want64.abs().max())   # < 0: accumulation truncates toward zero
        row["igemm_3xtf32_us"] = med(lambda: deform_conv2d(x, off, wt, b, padding=1, mask=msk))
        torch.backends.cuda.matmul.allow_tf32 = True
        fast = deform_conv2d(x, off, wt, b, padding=1, mask=msk)
        row["err_tf32"] = nmax(fast.cpu().numpy(), want.cpu().numpy())
        row["igemm_tf32_us"] = med(lambda: deform_conv2d(x, off, wt, b, padding=1, mask=msk))
        deform_conv.set_tensor_core(False)
        row["im2col_cublas_tf32_us"] = med(lambda: deform_conv2d(x, off, wt, b, padding=1, mask=msk))
        deform_conv.set_tensor_core(True)
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            import torchvision
            row["torchvision_fp32_us"] = med(lambda: torchvision.ops.deform_conv2d(x, off, wt, b, padding=1, mask=msk))
        except Exception as exc:
            row["torchvision"] = str(exc)[:80]
    print(f"{c}->{cout} @{h}x{w} x{n}", row, flush=True)

# where does the 1e-5 of the 3xTF32 form come from?  Operands that are EXACT in TF32 and a plain 3x3 convolution (zero
# offsets, unit mask: the columns are input values) leave only the tensor core's accumulation: every product is exact.
def tf32_exact(t):
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)

for c, cout, (h, w), n in [(264, 264, (12, 20), 60)]:
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = tf32_exact(torch.randn(n, c, h, w, device="cuda", generator=gen))
    off = torch.zeros(n, 18, h, w, device="cuda")
    msk = torch.ones(n, 9, h, w, device="cuda")
    wt = tf32_exact(torch.randn(cout, c, 3, 3, device="cuda", generator=gen) / (9 * c) ** 0.5)
    b = torch.zeros(cout, device="cuda")
    with torch.no_grad():
        want64 = deform_conv2d(x.double(), off.double(), wt.double(), b.double(), padding=1, mask=msk.double())
        deform_conv.set_tensor_core(False)
        e_cublas = nmax(deform_conv2d(x, off, wt, b, padding=1, mask=msk).cpu().numpy(), want64.cpu().numpy())
        deform_conv.set_tensor_core(True)
        e_3x = nmax(deform_conv2d(x, off, wt, b, padding=1, mask=msk).cpu().numpy(), want64.cpu().numpy())
        torch.backends.cuda.matmul.allow_tf32 = True
        e_1x = nmax(deform_conv2d(x, off, wt, b, padding=1, mask=msk).cpu().numpy(), want64.cpu().numpy())
        torch.backends.cuda.matmul.allow_tf32 = False
    print("tf32-exact operands, accumulation only:", {"cublas_fp32": e_cublas, "igemm_3xtf32": e_3x, "igemm_tf32": e_1x}, flush=True)

"""Mask-head modules of DeVIS with the reference's names, constructor arguments and parameters (state dicts load
unchanged), evaluated with this library's deformable convolution.

Mirrors src/models/deformable_segmentation.py of the reference:
  ModulatedDeformableConv2d  :244-267   offset / modulator convolutions (cuDNN) + torchvision.ops.deform_conv2d
  Conv2d                     :270-274   plain convolution with the reference's initialisation
  MultiScaleMHAttentionMap   :276-320   per-level query x key attention maps (einsum + softmax; stays ATen/cuBLAS)
  MaskHeadConv               :323-380   FPN-style convolutional head with GroupNorm
Only the deformable convolution is native code here (devis_b200.deform_conv); everything dense stays cuDNN / cuBLAS,
like the projection GEMMs of the attention modules.
"""
import torch
import torch.nn.functional as F
from torch import nn

from .deform_conv import deform_conv2d


class ModulatedDeformableConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=False):
        super().__init__()
        self.padding = padding
        k2 = kernel_size * kernel_size
        self.offset_conv = nn.Conv2d(in_channels, 2 * k2, kernel_size=kernel_size, stride=stride, padding=padding, bias=True)
        self.modulator_conv = nn.Conv2d(in_channels, k2, kernel_size=kernel_size, stride=stride, padding=padding, bias=True)
        for conv in (self.offset_conv, self.modulator_conv):       # zero init: starts as a plain convolution (:249-256)
            nn.init.constant_(conv.weight, 0.)
            nn.init.constant_(conv.bias, 0.)
        self.regular_conv = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                                      padding=padding, bias=bias)

    def forward(self, x):
        offset = self.offset_conv(x)
        modulator = 2. * torch.sigmoid(self.modulator_conv(x))
        # NB the reference passes padding but not the stride to deform_conv2d (:265): kept, a stride != 1 layer would
        # fail there with a shape error as well
        return deform_conv2d(x, offset, self.regular_conv.weight, self.regular_conv.bias, padding=self.padding,
                             mask=modulator)


class Conv2d(nn.Conv2d):
    def __init__(self, in_channels, out_channels, kernel_size, padding):
        super().__init__(in_channels, out_channels, kernel_size=kernel_size, padding=padding)
        nn.init.kaiming_uniform_(self.weight, a=1)
        nn.init.constant_(self.bias, 0)


class MultiScaleMHAttentionMap(nn.Module):
    def __init__(self, query_dim, hidden_dim, num_heads, num_levels, dropout=0, bias=True):
        super().__init__()
        self.num_heads, self.num_levels, self.hidden_dim = num_heads, num_levels, hidden_dim
        self.dropout = nn.Dropout(dropout)
        for i in range(num_levels):
            suffix = "" if i == 0 else f"_{i}"
            for name in ("q_linear", "k_linear"):
                lin = nn.Linear(query_dim, hidden_dim, bias=bias)
                nn.init.zeros_(lin.bias)
                nn.init.xavier_uniform_(lin.weight)
                setattr(self, name + suffix, lin)
        self.normalize_fact = float(hidden_dim / self.num_heads) ** -0.5

    def forward(self, q, k, mask=None):
        assert len(k) == self.num_levels and (mask is None or len(mask) == self.num_levels)
        maps = []
        for i, k_lvl in enumerate(k):
            suffix = "" if i == 0 else f"_{i}"
            q_lin, k_lin = getattr(self, "q_linear" + suffix), getattr(self, "k_linear" + suffix)
            q_lvl = q_lin(q)
            k_lvl = F.conv2d(k_lvl, k_lin.weight[:, :, None, None], k_lin.bias)
            d = self.hidden_dim // self.num_heads
            qh = q_lvl.view(q_lvl.shape[0], q_lvl.shape[1], self.num_heads, d)
            kh = k_lvl.view(k_lvl.shape[0], self.num_heads, d, k_lvl.shape[-2], k_lvl.shape[-1])
            weights = torch.einsum("bqnc,bnchw->bqnhw", qh * self.normalize_fact, kh)
            if mask is not None:
                weights.masked_fill_(mask[i][:, None, None], float("-inf"))
            maps.append(F.softmax(weights.flatten(2), dim=-1).view_as(weights))
        return maps


class MaskHeadConv(nn.Module):
    """FPN-style mask head: lay1/gn1, lay2/gn2 on the coarsest feature + attention maps, then one (adapter, lay, gn)
    per finer feature, optional 1-channel output layer (:323-380)."""

    def __init__(self, dim, fpn_dims, nheads, use_deformable_conv, multi_scale_att_maps, num_levels, out_layer=True):
        super().__init__()
        widths = [dim // (2 ** e) for e in range(num_levels + 2)]
        in_dims = list(widths)
        for i in range(len(multi_scale_att_maps)):
            in_dims[i] += nheads
        self.multi_scale_att_maps = len(multi_scale_att_maps) > 1
        conv = ModulatedDeformableConv2d if use_deformable_conv else Conv2d
        self.lay1, self.gn1 = conv(in_dims[0], in_dims[0], 3, padding=1), nn.GroupNorm(8, in_dims[0])
        self.lay2, self.gn2 = conv(in_dims[0], widths[1], 3, padding=1), nn.GroupNorm(8, widths[1])
        last = widths[1]
        for i in range(1, len(fpn_dims) + 1):
            setattr(self, f"lay{i + 2}", conv(in_dims[i], widths[i + 1], 3, padding=1))
            setattr(self, f"gn{i + 2}", nn.GroupNorm(8, widths[i + 1]))
            setattr(self, f"adapter{i}", Conv2d(fpn_dims[i - 1], widths[i], 1, padding=0))
            last = widths[i + 1]
        self.out_lay = conv(last, 1, 3, padding=1) if out_layer else None

    def forward(self, features, bbox_mask, instances_per_batch, expand_func):
        x = torch.cat([expand_func(features[0], instances_per_batch), bbox_mask[0]], 1)
        x = F.relu(self.gn1(self.lay1(x)))
        x = F.relu(self.gn2(self.lay2(x)))
        for lvl, feature in enumerate(features[1:]):
            fpn = expand_func(getattr(self, f"adapter{lvl + 1}")(feature), instances_per_batch)
            x = fpn + F.interpolate(x, size=fpn.shape[-2:], mode="nearest")
            if self.multi_scale_att_maps and lvl + 1 < len(bbox_mask):
                x = torch.cat([x, bbox_mask[lvl + 1]], 1)
            x = F.relu(getattr(self, f"gn{lvl + 3}")(getattr(self, f"lay{lvl + 3}")(x)))
        return self.out_lay(x) if self.out_lay is not None else x

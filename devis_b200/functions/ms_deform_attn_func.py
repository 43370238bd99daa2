"""``MSDeformAttnFunction`` with the reference's signature and autograd contract
(functions/ms_deform_attn_func.py:21-38): ``apply(value, spatial_shapes, level_start_index,
sampling_locations, attention_weights, im2col_step)``; saves the five input tensors and nothing
else, backward is once-differentiable and returns ``(grad_value, None, None, grad_sampling_loc,
grad_attn_weight, None)``.  The arithmetic runs in the sm_100a kernels behind the C ABI.

(The reference file also holds ``ms_deform_attn_core_pytorch``; that is the oracle and lives in
oracle/msda_torch.py -- the product package has no PyTorch implementation of the op.)
"""
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import MultiScaleDeformableAttention as MSDA


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        output = MSDA.ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index,
                                             sampling_locations, attention_weights, im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, loc, aw = ctx.saved_tensors
        grad_value, grad_loc, grad_aw = MSDA.ms_deform_attn_backward(
            value, shapes, lsi, loc, aw, grad_output.contiguous(), ctx.im2col_step,
            need_grad_value=ctx.needs_input_grad[0])
        return grad_value, None, None, grad_loc, grad_aw, None

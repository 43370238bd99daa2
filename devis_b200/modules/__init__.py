from .ms_deform_attn import (MSDeformAttn, TemporalMSDeformAttnBase, TemporalMSDeformAttnDecoder,
                             TemporalMSDeformAttnEncoder)

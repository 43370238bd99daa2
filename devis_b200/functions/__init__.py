from .ms_deform_attn_func import MSDeformAttnFunction
from .temporal_func import (TemporalMSDeformAttnFunction, TemporalMSDeformAttnFusedFunction,
                            temporal_ms_deform_attn)

"""The direct callers of the temporal attention modules: DeVIS's deformable transformer trunk.

Mirror of the reference's devis_transformer.py (/root/reference/src/models/devis_transformer.py: DeVISTransformer,
its encoder / decoder and their layers) together with the pieces of deformable_transformer.py those classes inherit (prepare_data :69-97,
get_reference_points :184-198, the layer bodies :143-173 and :215-281, iterative box refinement :284-310).
Same class names, constructor arguments, parameter names (reference checkpoints load with load_state_dict),
forward signatures and return tuples.  SURVEY.md section 8 row A8 / section 8(f): the callers either side of the hot
path, needed to run the path the way DeVIS runs it (integration parity, trunk-level benchmarks).  The dense parts
(LayerNorm, FFN, nn.MultiheadAttention) are ATen / cuBLAS, as in the reference.

What differs from the reference is bookkeeping only:
  * the temporal tables (frame offsets, repeated level shapes / start indices) are built ONCE per forward on the
    host and handed to all layers -- the reference rebuilds device tensors in the encoder and again in the decoder
    (devis_transformer.py:94-118,147-154), and our attention modules would have to read them back;
  * ``spatial_shapes`` / ``level_start_index`` carry a host copy (clip_geometry.attach_host_copy), so no layer
    synchronises the stream to learn the pyramid shape.
"""
import copy

import torch
import torch.nn.functional as F
from torch import nn

from . import clip_geometry
from .modules import MSDeformAttn, TemporalMSDeformAttnDecoder, TemporalMSDeformAttnEncoder


def inverse_sigmoid(x, eps=1e-5):
    """util/misc.py:430-434"""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def _activation(name):
    try:
        return {"relu": F.relu, "gelu": F.gelu, "glu": F.glu}[name]
    except KeyError:
        raise RuntimeError(f"activation should be relu/gelu, not {name}.") from None


def _clones(module, n):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(n)])


def _temporal_pyramid(spatial_shapes, n_slots):
    """(shapes repeated n_slots times, their cumulative start index): what the reference hands the modules as the
    'temporal' half of the shape / start-index pairs (devis_transformer.py:97,118)."""
    shapes = spatial_shapes.repeat(n_slots, 1)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    # the numbers are known on the host (pyramid_tensors attached them): hand them on, so that the modules' argument
    # validation (clip_geometry.from_reference_args) never reads these tensors back from the device
    host = getattr(spatial_shapes, clip_geometry._ATTR, None)
    if host is not None and host[0] == spatial_shapes._version:
        rows = [list(r) for r in host[1]] * n_slots
        starts, acc = [], 0
        for h, w in rows:
            starts.append(acc)
            acc += h * w
        clip_geometry.attach_host_copy(shapes, rows)
        clip_geometry.attach_host_copy(lsi, starts)
    return shapes, lsi


class DeVISTransformerEncoderLayer(nn.Module):
    """temporal deformable self-attention -> add & norm -> FFN -> add & norm (deformable_transformer.py:143-173)"""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_frames=6, t_window=2, n_levels=4,
                 n_heads=8, n_curr_points=4, n_temporal_points=2):
        super().__init__()
        self.self_attn = TemporalMSDeformAttnEncoder(n_frames, d_model, n_levels, t_window, n_heads, n_curr_points,
                                                     n_temporal_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = _activation(activation)
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def forward_ffn(self, src):
        return self.norm2(src + self.dropout3(self.linear2(self.dropout2(self.activation(self.linear1(src))))))

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, **kwargs):
        attended = self.self_attn(self.with_pos_embed(src, pos), reference_points, src, spatial_shapes,
                                  level_start_index, **kwargs)[0]
        return self.forward_ffn(self.norm1(src + self.dropout1(attended)))


class DeVISTransformerEncoder(nn.Module):
    def __init__(self, encoder_layer, num_layers, t_window, enc_connect_all_embeddings):
        super().__init__()
        self.layers = _clones(encoder_layer, num_layers)
        self.num_layers = num_layers
        self.t_window = t_window
        self.enc_connect_all_embeddings = enc_connect_all_embeddings

    @staticmethod
    def get_reference_points(spatial_shapes, valid_ratios, device):
        """pixel centres of every level, normalised by the valid part of the map, then expressed in every level's
        valid ratio: (T, S, L, 2) -- deformable_transformer.py:184-198"""
        shapes = clip_geometry.host_list(spatial_shapes)
        per_level = []
        for lvl, (h, w) in enumerate(shapes):
            ys = torch.linspace(0.5, h - 0.5, h, dtype=torch.float32, device=device)
            xs = torch.linspace(0.5, w - 0.5, w, dtype=torch.float32, device=device)
            gy, gx = torch.meshgrid(ys, xs, indexing="ij")
            gy = gy.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * h)
            gx = gx.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * w)
            per_level.append(torch.stack((gx, gy), -1))
        return torch.cat(per_level, 1)[:, :, None] * valid_ratios[:, None]

    def frame_offsets(self, n_frames):
        """per query frame, the offsets of the frames it also samples (devis_transformer.py:94-112)"""
        if self.enc_connect_all_embeddings:
            table = clip_geometry.all_frames_table(n_frames)
        else:
            table = clip_geometry.window_table(n_frames, self.t_window)
        return [[f - t for f in row] for t, row in enumerate(table)]

    def forward(self, src, spatial_shapes, level_start_index, valid_ratios, pos=None, padding_mask=None):
        # padding_mask is accepted and ignored, like the reference's (devis_transformer.py:90,120)
        n_frames = src.shape[0]
        reference_points = self.get_reference_points(spatial_shapes, valid_ratios, device=src.device)
        offsets = self.frame_offsets(n_frames)
        t_shapes, t_lsi = _temporal_pyramid(spatial_shapes, len(offsets[0]))
        output = src
        for layer in self.layers:
            output = layer(output, pos, reference_points, (spatial_shapes, t_shapes), (level_start_index, t_lsi),
                           temporal_offsets=offsets)
        return output


class DeVISTransformerDecoderLayer(nn.Module):
    """query self-attention -> temporal deformable cross-attention -> FFN, post-norm (deformable_transformer.py:215-281)"""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_frames=36, t_window=2,
                 dec_instance_aware_att=True, n_levels=4, n_heads=8, n_curr_points=4, n_temporal_points=2):
        super().__init__()
        self.cross_attn = TemporalMSDeformAttnDecoder(n_frames, d_model, n_levels, t_window, n_heads, n_curr_points,
                                                      n_temporal_points, dec_instance_aware_att)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        # the reference builds the query self-attention with 8 heads whatever n_heads is (devis_transformer.py:131-132)
        self.self_attn = nn.MultiheadAttention(d_model, 8, dropout=dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = _activation(activation)
        self.dropout3 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout4 = nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_model)

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def forward_ffn(self, tgt):
        return self.norm3(tgt + self.dropout4(self.linear2(self.dropout3(self.activation(self.linear1(tgt))))))

    def forward(self, tgt, query_pos, reference_points, src, src_spatial_shapes, level_start_index, **kwargs):
        qk = self.with_pos_embed(tgt, query_pos).transpose(0, 1)
        attended = self.self_attn(qk, qk, tgt.transpose(0, 1))[0].transpose(0, 1)
        tgt = self.norm2(tgt + self.dropout2(attended))
        sampled = self.cross_attn(self.with_pos_embed(tgt, query_pos), reference_points, src, src_spatial_shapes,
                                  level_start_index, **kwargs)[0]
        return self.forward_ffn(self.norm1(tgt + self.dropout1(sampled)))


class DeVISTransformerDecoder(nn.Module):
    def __init__(self, decoder_layer, num_layers, with_gradient=False, instance_aware_att=True):
        super().__init__()
        self.layers = _clones(decoder_layer, num_layers)
        self.num_layers = num_layers
        self.with_gradient = with_gradient
        self.instance_aware_att = instance_aware_att
        # set from outside by the detection head for iterative box refinement (deformable_detr.py), as in the reference
        self.bbox_embed = None
        self.class_embed = None
        self.ref_point_embed = None

    def refine_reference_point(self, lid, output, reference_points, intermediate, intermediate_reference_points):
        """deformable_transformer.py:284-310"""
        if self.bbox_embed is not None:
            delta = self.bbox_embed[lid](output)
            if reference_points.shape[-1] == 4:
                refined = (delta + inverse_sigmoid(reference_points)).sigmoid()
            else:
                assert reference_points.shape[-1] == 2
                refined = torch.cat([delta[..., :2] + inverse_sigmoid(reference_points), delta[..., 2:]], -1).sigmoid()
            reference_points = refined if self.with_gradient else refined.detach()
        if self.ref_point_embed is not None:
            reference_points = (self.ref_point_embed[lid](output) + inverse_sigmoid(reference_points)).sigmoid()
        intermediate.append(output)
        intermediate_reference_points.append(reference_points)
        return reference_points, intermediate, intermediate_reference_points

    def forward(self, tgt, reference_points, src, src_spatial_shapes, src_level_start_index, src_valid_ratios,
                query_pos=None, src_padding_mask=None):
        n_frames = src.shape[0]
        offsets = [[f - t for f in row] for t, row in enumerate(clip_geometry.all_frames_table(n_frames))]
        t_shapes, t_lsi = _temporal_pyramid(src_spatial_shapes, n_frames - 1)
        # every frame's points / boxes are scaled by FRAME 0's valid ratios (devis_transformer.py:161-166)
        ratio0 = src_valid_ratios[0, None]
        output, intermediate, intermediate_refs = tgt, [], []
        for lid, layer in enumerate(self.layers):
            if reference_points.shape[-1] == 4:
                ref_in = reference_points[:, :, None] * torch.cat([ratio0, ratio0], -1)[:, None]
            else:
                assert reference_points.shape[-1] == 2
                ref_in = reference_points[:, :, None] * ratio0
            output = layer(output, query_pos, ref_in, src, (src_spatial_shapes, t_shapes),
                           (src_level_start_index, t_lsi), temporal_offsets=offsets)
            reference_points, intermediate, intermediate_refs = self.refine_reference_point(
                lid, output, reference_points, intermediate, intermediate_refs)
        return torch.stack(intermediate), torch.stack(intermediate_refs)


class DeVISTransformer(nn.Module):
    def __init__(self, d_model=256, num_frames=6, nhead=8, num_encoder_layers=6, num_decoder_layers=6,
                 dim_feedforward=1024, dropout=0.1, activation="relu", num_feature_levels=4,
                 enc_connect_all_embeddings=True, enc_temporal_window=2, enc_n_curr_points=4, enc_n_temporal_points=2,
                 dec_n_curr_points=4, dec_n_temporal_points=2, dec_instance_aware_att=True, with_gradient=False):
        super().__init__()
        self.d_model = d_model
        self.nhead = nhead
        self.level_embed = nn.Parameter(torch.Tensor(num_feature_levels, d_model))
        self.reference_points = nn.Linear(d_model, 2)
        if enc_connect_all_embeddings:
            enc_temporal_window = num_frames - 1
        self.encoder = DeVISTransformerEncoder(
            DeVISTransformerEncoderLayer(d_model, dim_feedforward, dropout, activation, num_frames, enc_temporal_window,
                                         num_feature_levels, nhead, enc_n_curr_points, enc_n_temporal_points),
            num_encoder_layers, enc_temporal_window, enc_connect_all_embeddings)
        self.decoder = DeVISTransformerDecoder(
            DeVISTransformerDecoderLayer(d_model, dim_feedforward, dropout, activation, num_frames, num_frames - 1,
                                         dec_instance_aware_att, num_feature_levels, nhead, dec_n_curr_points,
                                         dec_n_temporal_points),
            num_decoder_layers, with_gradient)
        self._reset_parameters()

    def _reset_parameters(self):
        """deformable_transformer.py:48-58"""
        for prm in self.parameters():
            if prm.dim() > 1:
                nn.init.xavier_uniform_(prm)
        for mod in self.modules():
            if isinstance(mod, (MSDeformAttn, TemporalMSDeformAttnEncoder, TemporalMSDeformAttnDecoder)):
                mod._reset_parameters()
        nn.init.xavier_uniform_(self.reference_points.weight.data, gain=1.0)
        nn.init.constant_(self.reference_points.bias.data, 0.)
        nn.init.normal_(self.level_embed)

    @staticmethod
    def get_valid_ratio(mask):
        """fraction of each (T, H, W) padding mask that is image, as (w, h) -- deformable_transformer.py:60-67"""
        _, h, w = mask.shape
        valid_h = (~mask[:, :, 0]).sum(1).float() / h
        valid_w = (~mask[:, 0, :]).sum(1).float() / w
        return torch.stack([valid_w, valid_h], -1)

    def prepare_data(self, srcs, masks, pos_embeds):
        """per-level (T, C, H, W) maps -> one (T, S, C) sequence (+ masks, positional + level embeddings, shapes)"""
        shapes = [tuple(int(x) for x in src.shape[-2:]) for src in srcs]
        src_flatten = torch.cat([src.flatten(2).transpose(1, 2) for src in srcs], 1)
        mask_flatten = torch.cat([mask.flatten(1) for mask in masks], 1)
        pos_flatten = torch.cat([pos.flatten(2).transpose(1, 2) + self.level_embed[lvl].view(1, 1, -1)
                                 for lvl, pos in enumerate(pos_embeds)], 1)
        spatial_shapes, level_start_index = clip_geometry.pyramid_tensors(shapes, src_flatten.device)
        valid_ratios = torch.stack([self.get_valid_ratio(m) for m in masks], 1)
        return src_flatten, mask_flatten, pos_flatten, spatial_shapes, level_start_index, valid_ratios

    def forward(self, srcs, masks, pos_embeds, query_embed=None):
        src, mask, pos, spatial_shapes, level_start_index, valid_ratios = self.prepare_data(srcs, masks, pos_embeds)
        memory = self.encoder(src, spatial_shapes, level_start_index, valid_ratios, pos, mask)

        n_frames, _, channels = memory.shape
        query_embed, tgt = torch.split(query_embed, channels, dim=1)
        query_embed, tgt = query_embed.unsqueeze(0), tgt.unsqueeze(0)
        reference_points = self.reference_points(query_embed).sigmoid()
        hs, inter_references = self.decoder(tgt, reference_points, memory, spatial_shapes, level_start_index,
                                            valid_ratios, query_embed)

        memories, start = [], 0
        for lvl_src in srcs:
            h, w = lvl_src.shape[-2:]
            memories.append(memory[:, start:start + h * w].permute(2, 0, 1).view(1, channels, n_frames, h, w))
            start += h * w
        # the pyramid tensors are process-cached (clip_geometry.pyramid_tensors): hand the caller its own copies, so an
        # in-place edit there cannot corrupt later forwards with the same image size
        return (hs, query_embed, memories, reference_points, inter_references, level_start_index.clone(), valid_ratios,
                spatial_shapes.clone())

"""TEST INFRASTRUCTURE - CPU restatement of the SAG decoder, the step that produces the
`init_image` of LivelySpeaker sampling (scripts/test_LivelySpeaker_ted.py:80-113).

Follows scripts/model/motionclip_module.py:
  PositionalEncoding            :12-29
  Decoder_TRANSFORMER.__init__  :98-134  (nn.TransformerDecoder: 3 post-norm layers, 4 heads, GELU FFN 512-1024-512)
  Decoder_TRANSFORMER.forward   :137-183
with torch's nn.TransformerDecoderLayer / F.multi_head_attention_forward written out in plain ops (fp32, CPU).
The memory `z` is a single token, so the cross-attention softmax is over one key and its output is
out_proj(v_proj(z)) for every query.  Pinned against the reference module itself by
tests/golden/make_golden_sag.py (tests/golden/sag.npz).  Only tests/, smoke() and bench.py may import this.
"""
import math

import torch
import torch.nn.functional as F

N_HEADS = 4


def positional_encoding(n, d):
    """motionclip_module.py:17-23 (pe[:n, 0, :])."""
    pe = torch.zeros(n, d)
    position = torch.arange(0, n, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d, 2).float() * (-math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def _self_attention(sd, p, x):
    """x [T, B, d] -> [T, B, d]; torch.nn.functional.multi_head_attention_forward without masks or dropout."""
    T, B, d = x.shape
    hd = d // N_HEADS
    qkv = x @ sd[p + "in_proj_weight"].t() + sd[p + "in_proj_bias"]
    q, k, v = qkv.split(d, dim=-1)
    q = q.reshape(T, B * N_HEADS, hd).transpose(0, 1) * (1.0 / math.sqrt(hd))
    k = k.reshape(T, B * N_HEADS, hd).transpose(0, 1)
    v = v.reshape(T, B * N_HEADS, hd).transpose(0, 1)
    w = torch.softmax(q @ k.transpose(1, 2), dim=-1)
    o = (w @ v).transpose(0, 1).reshape(T, B, d)
    return o @ sd[p + "out_proj.weight"].t() + sd[p + "out_proj.bias"]


def _cross_attention_one_token(sd, p, z, d):
    """memory of length 1: softmax over one key is 1, the result is out_proj(v_proj(z)) [B, d]."""
    wv, bv = sd[p + "in_proj_weight"][2 * d:], sd[p + "in_proj_bias"][2 * d:]
    v = z @ wv.t() + bv
    return v @ sd[p + "out_proj.weight"].t() + sd[p + "out_proj.bias"]


def decode(sd, x, z, mask, n_pre_poses=4):
    """Decoder_TRANSFORMER.forward: x [B,J,D,F] (only the first n_pre_poses frames are used), z [B,512],
    mask [B,F] bool -> output [B,J,D,F].  sd uses the reference module's state_dict keys."""
    B, J, D, Fn = x.shape
    d = z.shape[1]
    motion = x.permute(3, 0, 1, 2).reshape(Fn, B, J * D).clone()
    pre = torch.zeros(Fn, B, J * D + 1)
    pre[:n_pre_poses, :, :-1] = motion[:n_pre_poses]
    pre[:n_pre_poses, :, -1] = 1
    h = pre @ sd["mapping.weight"].t() + sd["mapping.bias"] + positional_encoding(Fn, d)[:, None, :]
    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("seqTransDecoder.layers."))
    for l in range(n_layers):
        p = "seqTransDecoder.layers.%d." % l
        h = F.layer_norm(h + _self_attention(sd, p + "self_attn.", h), (d,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
        h = F.layer_norm(h + _cross_attention_one_token(sd, p + "multihead_attn.", z, d)[None], (d,), sd[p + "norm2.weight"],
                         sd[p + "norm2.bias"], 1e-5)
        ff = F.gelu(h @ sd[p + "linear1.weight"].t() + sd[p + "linear1.bias"]) @ sd[p + "linear2.weight"].t() + sd[p + "linear2.bias"]
        h = F.layer_norm(h + ff, (d,), sd[p + "norm3.weight"], sd[p + "norm3.bias"], 1e-5)
    out = (h @ sd["finallayer.weight"].t() + sd["finallayer.bias"]).reshape(Fn, B, J, D)
    out[~mask.t()] = 0
    return out.permute(1, 2, 3, 0)

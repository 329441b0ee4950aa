"""Test double with the module tree, parameter names and call conventions of the reference's LAVIS BlipITM
(VIT:205-300, MED:56-124, MED:126-311, MED:312-530, BITM:43-57, BITM:217-249), small enough for CPU tests.

LAVIS itself cannot be imported offline; this restates only what the bridge in pnp_ovss_b200/lavis_compat.py relies
on: the state_dict key layout, `model(samples, match_head="itm")`, the path
`text_encoder.base_model.base_model.encoder.layer[i].crossattention.self` and that module's capture protocol and
forward signature / return tuple."""
import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F


class BertSelfAttention(nn.Module):
    def __init__(self, hidden, heads, kv_width):
        super().__init__()
        self.num_attention_heads = heads
        self.attention_head_size = hidden // heads
        self.all_head_size = hidden
        self.query = nn.Linear(hidden, hidden)
        self.key = nn.Linear(kv_width, hidden)
        self.value = nn.Linear(kv_width, hidden)
        self.dropout = nn.Dropout(0.0)
        self.position_embedding_type = "absolute"
        self.save_attention = False
        self.attention_map = None
        self.attn_gradients = None

    def save_attn_gradients(self, g):
        self.attn_gradients = g

    def get_attn_gradients(self):
        return self.attn_gradients

    def save_attention_map(self, a):
        self.attention_map = a

    def get_attention_map(self):
        return self.attention_map

    def transpose_for_scores(self, x):
        return x.view(x.shape[0], x.shape[1], self.num_attention_heads, self.attention_head_size).permute(0, 2, 1, 3)

    def forward(self, hidden_states, attention_mask=None, head_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, past_key_value=None, output_attentions=False):
        cross = encoder_hidden_states is not None
        src = encoder_hidden_states if cross else hidden_states
        if cross:
            attention_mask = encoder_attention_mask
        q = self.transpose_for_scores(self.query(hidden_states))
        k = self.transpose_for_scores(self.key(src))
        v = self.transpose_for_scores(self.value(src))
        scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(self.attention_head_size)
        if attention_mask is not None:
            scores = scores + attention_mask
        probs = torch.softmax(scores, -1)
        if cross and self.save_attention:
            self.save_attention_map(probs)
            probs.register_hook(self.save_attn_gradients)
        ctx = torch.matmul(self.dropout(probs), v).permute(0, 2, 1, 3).contiguous()
        ctx = ctx.view(ctx.shape[0], ctx.shape[1], self.all_head_size)
        return ((ctx, probs) if output_attentions else (ctx,)) + ((k, v),)


class _DenseLN(nn.Module):
    def __init__(self, d_in, d_out, eps):
        super().__init__()
        self.dense = nn.Linear(d_in, d_out)
        self.LayerNorm = nn.LayerNorm(d_out, eps=eps)

    def forward(self, x, residual):
        return self.LayerNorm(self.dense(x) + residual)


class BertAttention(nn.Module):
    def __init__(self, hidden, heads, kv_width, eps):
        super().__init__()
        self.self = BertSelfAttention(hidden, heads, kv_width)
        self.output = _DenseLN(hidden, hidden, eps)

    def forward(self, x, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None):
        out = self.self(x, attention_mask, None, encoder_hidden_states, encoder_attention_mask)
        return self.output(out[0], x)


class _Dense(nn.Module):
    def __init__(self, d_in, d_out):
        super().__init__()
        self.dense = nn.Linear(d_in, d_out)


class BertLayer(nn.Module):
    def __init__(self, hidden, heads, inter, enc_width, eps=1e-12):
        super().__init__()
        self.attention = BertAttention(hidden, heads, hidden, eps)
        self.crossattention = BertAttention(hidden, heads, enc_width, eps)
        self.intermediate = _Dense(hidden, inter)
        self.output = _DenseLN(inter, hidden, eps)

    def forward(self, x, attention_mask, enc, enc_mask):
        x = self.attention(x, attention_mask)
        x = self.crossattention(x, None, enc, enc_mask)
        return self.output(F.gelu(self.intermediate.dense(x)), x)


class BertEmbeddings(nn.Module):
    def __init__(self, vocab, hidden, max_pos, eps=1e-12):
        super().__init__()
        self.word_embeddings = nn.Embedding(vocab, hidden, padding_idx=0)
        self.position_embeddings = nn.Embedding(max_pos, hidden)
        self.LayerNorm = nn.LayerNorm(hidden, eps=eps)
        self.register_buffer("position_ids", torch.arange(max_pos).expand((1, -1)))

    def forward(self, ids):
        pos = self.position_ids[:, :ids.shape[1]]
        return self.LayerNorm(self.word_embeddings(ids) + self.position_embeddings(pos))


class XBertEncoder(nn.Module):
    def __init__(self, vocab, hidden, layers, heads, inter, enc_width, max_pos):
        super().__init__()
        self.embeddings = BertEmbeddings(vocab, hidden, max_pos)
        self.encoder = nn.Module()
        self.encoder.layer = nn.ModuleList([BertLayer(hidden, heads, inter, enc_width) for _ in range(layers)])
        self.config = SimpleNamespace(hidden_size=hidden)

    @property
    def base_model(self):
        return self

    def forward(self, input_ids, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                return_dict=True):
        ext = ((1.0 - attention_mask[:, None, None, :].float()) * -10000.0)
        enc_ext = ((1.0 - encoder_attention_mask[:, None, None, :].float()) * -10000.0)
        x = self.embeddings(input_ids)
        for lyr in self.encoder.layer:
            x = lyr(x, ext, encoder_hidden_states, enc_ext)
        return SimpleNamespace(last_hidden_state=x)


class _VitAttention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.qkv = nn.Linear(dim, dim * 3)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, L, D = x.shape
        qkv = self.qkv(x).reshape(B, L, 3, self.num_heads, D // self.num_heads).permute(2, 0, 3, 1, 4)
        attn = ((qkv[0] @ qkv[1].transpose(-2, -1)) * (D // self.num_heads) ** -0.5).softmax(-1)
        return self.proj((attn @ qkv[2]).transpose(1, 2).reshape(B, L, D))


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(F.gelu(self.fc1(x)))


class _Block(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _VitAttention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, dim * 4)

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class VisionTransformer(nn.Module):
    def __init__(self, img_size, dim, depth, heads, patch=16):
        super().__init__()
        self.patch_embed = nn.Module()
        self.patch_embed.proj = nn.Conv2d(3, dim, patch, patch)
        self.patch_embed.num_patches = (img_size // patch) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, dim))
        self.blocks = nn.ModuleList([_Block(dim, heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.vision_width = dim

    def forward_features(self, x):
        x = self.patch_embed.proj(x).flatten(2).transpose(1, 2)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], 1)
        x = x + self.pos_embed[:, :x.shape[1]]
        for blk in self.blocks:
            x = blk(x)
        return self.norm(x)


class LavisShapedBlipITM(nn.Module):
    def __init__(self, tokenizer, img_size=32, vit_dim=32, vit_depth=2, vit_heads=2, hidden=24, layers=3, heads=2, inter=48,
                 vocab=30524, max_pos=512, embed_dim=8):
        super().__init__()
        self.tokenizer = tokenizer
        self.max_txt_len = 500
        self.visual_encoder = VisionTransformer(img_size, vit_dim, vit_depth, vit_heads)
        self.text_encoder = XBertEncoder(vocab, hidden, layers, heads, inter, vit_dim, max_pos)
        self.vision_proj = nn.Linear(vit_dim, embed_dim)
        self.text_proj = nn.Linear(hidden, embed_dim)
        self.itm_head = nn.Linear(hidden, 2)
        for p in self.parameters():
            nn.init.normal_(p, std=0.05)
        for m in self.modules():
            if isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)

    def forward(self, samples, match_head="itm"):
        image, caption = samples["image"], samples["text_input"]
        image_embeds = self.visual_encoder.forward_features(image)
        image_atts = torch.ones(image_embeds.shape[:-1], dtype=torch.long, device=image.device)
        text = self.tokenizer(caption, padding="longest", truncation=True, max_length=self.max_txt_len, return_tensors="pt")
        ids = text.input_ids.to(image.device).clone()
        ids[:, 0] = self.tokenizer.enc_token_id
        out = self.text_encoder(ids, attention_mask=text.attention_mask.to(image.device), encoder_hidden_states=image_embeds,
                                encoder_attention_mask=image_atts, return_dict=True)
        return self.itm_head(out.last_hidden_state[:, 0, :])

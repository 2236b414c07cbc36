"""TEST INFRASTRUCTURE — CPU oracle for the Qwen3-TTS talker / code-predictor decode path.

This file is NOT part of the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker / the timed CPU baseline.  The product path (``qwen_megakernel``) never falls back to it.

What it restates (upstream = jayanth-kumar-morem/qwen-megakernel-tts, paths relative to its root):
  * the PyTorch talker decode step      validate_kernel.py:118-201  (``PyTorchTalkerReference._forward``)
  * the PyTorch code-predictor decode   qwen_megakernel/model_tts.py:576-619 (``CodePredictor._layer_decode``)
  * the code-predictor frame loop       qwen_megakernel/model_tts.py:439-504 and :729-773
  * greedy / top-k sampling             qwen_megakernel/model_tts.py:756-764
  * the text side of the prefill        qwen_megakernel/model_tts.py:361-374 (``TextProjection.embed_text_ids``) and
                                        :776-864 (``build_prefill_embeddings``); pinned by tests/golden/text_projection.npz
                                        (tests/golden/make_golden_text.py, the upstream functions themselves)
The arithmetic lives in third-party PyTorch (``torch>=2.7`` unpinned upstream, requirements.txt:1;
2.11.0+cu128 in this image): bf16 ``mv`` with fp32 accumulation, fp32 RMSNorm / softmax.

Parity pin: upstream ships NO golden vectors (its tests compare two live implementations).  This
restatement is pinned against fixtures produced by importing the upstream classes themselves
(tests/golden/make_golden.py -> tests/golden/*.npz); tests/test_oracle_golden.py checks it.
Rounding convention ("r" = round fp32 -> bf16, RNE), one decode step at position p:

    res = float(x)                                   x = embed[token] or the caller's bf16 vector
    per layer:
        n      = r(rmsnorm(r(res)) * w_in)           fp32 math: x / sqrt(mean(x^2) + 1e-6) * w
        q,k,v  = r(Wq n), r(Wk n), r(Wv n)           bf16 GEMV, fp32 accumulate
        q,k    = r(rmsnorm_128(q) * w_qn), r(rmsnorm_128(k) * w_kn)
        q,k    = rope(q), rope(k)                    bf16 elementwise: r(r(t1*c) - r(t2*s)), r(r(t2*c) + r(t1*s))
        cache[l,:,p] = k, v                          bf16, layout [L][8][S][128]
        a      = r(softmax(q K^T / sqrt(128)) V)     fp32 scores/softmax/PV, GQA kv = h // 2
        res    = res + float(r(Wo a))                fp32 residual (talker)  |  r(res + o) (code predictor)
        n2     = r(rmsnorm(r(res)) * w_post)
        m      = r(r(silu(r(Wg n2))) * r(Wu n2))
        res    = res + float(r(Wd m))                (same residual convention)
    hn     = r(rmsnorm(r(res)) * w_final)            returned hidden (bf16 values in an fp32 container)
    logits = r(W_head hn);  token = argmax, lowest index on ties
"""

from __future__ import annotations

import math
from typing import Optional

import torch

NUM_Q_HEADS, NUM_KV_HEADS, HEAD_DIM = 16, 8, 128
HIDDEN, INTER = 1024, 3072
EPS = 1e-6
ATTN_SCALE = 1.0 / math.sqrt(HEAD_DIM)
BF16 = torch.bfloat16


def _gemv(w_bf16: torch.Tensor, x_bf16: torch.Tensor) -> torch.Tensor:
    """bf16 [out,in] x bf16 [in] -> bf16 [out]; fp32 accumulation inside, one rounding (F.linear, as upstream)."""
    return torch.nn.functional.linear(x_bf16, w_bf16)


def _rmsnorm(x_bf16: torch.Tensor, w_bf16: torch.Tensor) -> torch.Tensor:
    """validate_kernel.py:87-90 / model_tts.py:506-509 — fp32 math, bf16 result."""
    xf = x_bf16.float()
    rms = torch.sqrt(xf.pow(2).mean(-1, keepdim=True) + EPS)
    return (xf / rms * w_bf16.float()).to(BF16)


def _rope(t_bf16: torch.Tensor, cos_row: torch.Tensor, sin_row: torch.Tensor) -> torch.Tensor:
    """Rotate-half RoPE carried out in bf16 (validate_kernel.py:92-99 / model_tts.py:511-520)."""
    half = HEAD_DIM // 2
    t1, t2 = t_bf16[..., :half], t_bf16[..., half:]
    c, s = cos_row[:half], sin_row[:half]
    return torch.cat([t1 * c - t2 * s, t2 * c + t1 * s], dim=-1)


def mrope_axis_map(section, interleaved: bool = False) -> list:
    """Axis (0 = temporal, 1 = height, 2 = width) of each of the 64 rotary frequencies under multimodal RoPE.

    Third-party algorithm (the upstream repo implements standard RoPE only and documents M-RoPE as its gap,
    README.md:208): HF transformers 5.5.0, ``models/qwen2_vl/modeling_qwen2_vl.py::apply_multimodal_rotary_pos_emb``
    (chunked: ``cos.split(mrope_section * 2)``, chunk i takes axis i % 3) and
    ``models/qwen3_vl/modeling_qwen3_vl.py::Qwen3VLTextRotaryEmbedding.apply_interleaved_mrope`` (interleaved: index
    ``i`` with ``i % 3 == a`` and ``i < 3 * section[a]`` takes axis a for a = 1, 2; everything else axis 0).
    Pinned by tests/golden/mrope.npz, generated from those two functions (tests/golden/make_golden_mrope.py)."""
    assert sum(section) == HEAD_DIM // 2
    if interleaved:
        return [1 if (i % 3 == 1 and i < 3 * section[1]) else 2 if (i % 3 == 2 and i < 3 * section[2]) else 0
                for i in range(HEAD_DIM // 2)]
    return [0 if i < section[0] else 1 if i < section[0] + section[1] else 2 for i in range(HEAD_DIM // 2)]


def mrope_rows(cos_table: torch.Tensor, sin_table: torch.Tensor, rope_pos, axis_map):
    """cos / sin row [128] of one token whose three axes sit at positions ``rope_pos`` (halves duplicated like the tables)."""
    idx = torch.tensor([rope_pos[a] for a in axis_map] * 2, dtype=torch.long)
    col = torch.arange(HEAD_DIM)
    return cos_table[idx, col], sin_table[idx, col]


class LayerStackOracle:
    """N identical Qwen3 decoder layers + final RMSNorm with a bf16 KV cache (one token per call)."""

    def __init__(self, layer_tensors, final_norm, cos_table, sin_table, max_seq: int,
                 residual_fp32: bool):
        assert len(layer_tensors) % 11 == 0
        self.num_layers = len(layer_tensors) // 11
        self.layers = [layer_tensors[i * 11:(i + 1) * 11] for i in range(self.num_layers)]
        self.final_norm = final_norm
        self.cos, self.sin = cos_table, sin_table
        self.max_seq = max_seq
        self.residual_fp32 = residual_fp32
        self.mrope_axis = None          # list of 64 axes -> multimodal RoPE (see mrope_axis_map); None = standard RoPE
        dev = final_norm.device
        self.k_cache = torch.zeros(self.num_layers, NUM_KV_HEADS, max_seq, HEAD_DIM, dtype=BF16, device=dev)
        self.v_cache = torch.zeros_like(self.k_cache)

    def reset(self):
        self.k_cache.zero_()
        self.v_cache.zero_()

    @torch.no_grad()
    def forward(self, x_bf16: torch.Tensor, pos: int, rope_pos=None) -> torch.Tensor:
        """One decode step at position ``pos`` (KV row); returns the post-final-norm hidden (bf16[1024]).
        ``rope_pos``: (t, h, w) positions of the three M-RoPE axes (default: ``pos`` on all of them)."""
        assert 0 <= pos < self.max_seq
        if self.mrope_axis is not None:
            cos_row, sin_row = mrope_rows(self.cos, self.sin, rope_pos if rope_pos is not None else (pos, pos, pos),
                                          self.mrope_axis)
        else:
            assert rope_pos is None or tuple(rope_pos) == (pos, pos, pos)
            cos_row, sin_row = self.cos[pos], self.sin[pos]
        res = x_bf16.float() if self.residual_fp32 else x_bf16.to(BF16)
        for li, (w_in, wq, wk, wv, w_qn, w_kn, wo, w_post, wg, wu, wd) in enumerate(self.layers):
            n = _rmsnorm(res.to(BF16), w_in)
            q = _gemv(wq, n).view(NUM_Q_HEADS, HEAD_DIM)
            k = _gemv(wk, n).view(NUM_KV_HEADS, HEAD_DIM)
            v = _gemv(wv, n).view(NUM_KV_HEADS, HEAD_DIM)
            q = _rope(_rmsnorm(q, w_qn), cos_row, sin_row)
            k = _rope(_rmsnorm(k, w_kn), cos_row, sin_row)
            self.k_cache[li, :, pos] = k
            self.v_cache[li, :, pos] = v
            kf = self.k_cache[li, :, :pos + 1].float()                       # [8, p+1, 128]
            vf = self.v_cache[li, :, :pos + 1].float()
            qf = q.float().view(NUM_KV_HEADS, NUM_Q_HEADS // NUM_KV_HEADS, HEAD_DIM)   # GQA: head h -> kv h//2
            scores = torch.einsum("ghd,gsd->ghs", qf, kf) * ATTN_SCALE
            probs = torch.softmax(scores, dim=-1)
            a = torch.einsum("ghs,gsd->ghd", probs, vf).to(BF16).reshape(-1)  # [2048]
            o = _gemv(wo, a)
            res = res + o.float() if self.residual_fp32 else res + o          # bf16 + bf16 rounds once
            n2 = _rmsnorm(res.to(BF16), w_post)
            m = torch.nn.functional.silu(_gemv(wg, n2)) * _gemv(wu, n2)
            d = _gemv(wd, m)
            res = res + d.float() if self.residual_fp32 else res + d
        self.last_residual = res
        return _rmsnorm(res.to(BF16), self.final_norm)


def top2_margin(logits_f32: torch.Tensor) -> float:
    top = torch.topk(logits_f32, 2).values
    return float(top[0] - top[1])


class TalkerOracle:
    """Restates ``PyTorchTalkerReference`` (validate_kernel.py:25-201): fp32 residual stream."""

    def __init__(self, weights: dict, max_seq: Optional[int] = None):
        self.embed_weight = weights["embed_weight"]
        self.lm_head_weight = weights["lm_head_weight"]
        max_seq = max_seq or weights["cos_table"].shape[0]
        self.stack = LayerStackOracle(weights["layer_weights"], weights["final_norm_weight"],
                                      weights["cos_table"], weights["sin_table"], max_seq,
                                      residual_fp32=True)
        self.position = 0
        self.last_logits: Optional[torch.Tensor] = None

    def reset(self):
        self.stack.reset()
        self.position = 0

    def set_mrope(self, section=(24, 20, 20), interleaved: bool = False):
        self.stack.mrope_axis = None if section is None else mrope_axis_map(section, interleaved)

    @torch.no_grad()
    def step_with_embed(self, embed_bf16: torch.Tensor, rope_pos=None):
        hn = self.stack.forward(embed_bf16.to(BF16), self.position, rope_pos)
        logits = _gemv(self.lm_head_weight, hn).float()        # bf16 logits widened (validate_kernel.py:197)
        self.last_logits = logits
        self.position += 1
        return int(logits.argmax()), hn.float()

    def step(self, token_id: int, rope_pos=None):
        return self.step_with_embed(self.embed_weight[token_id], rope_pos)


def select_token(logits_bf16: torch.Tensor, do_sample: bool, temperature: float, top_k: int,
                 generator: Optional[torch.Generator] = None) -> int:
    """model_tts.py:756-764 — temperature, top-k threshold (ties kept), softmax, multinomial | argmax."""
    if do_sample and temperature > 0:
        z = logits_bf16.float() / temperature
        if top_k > 0:
            kth = torch.topk(z, min(top_k, z.numel())).values[-1]
            z = torch.where(z < kth, torch.full_like(z, float("-inf")), z)
        probs = torch.softmax(z, dim=-1)
        return int(torch.multinomial(probs, 1, generator=generator))
    return int(logits_bf16.float().argmax())


class CodePredictorOracle:
    """Restates the code-predictor frame (model_tts.py:439-504 PyTorch, :729-773 kernel path).

    bf16 residual stream; 16 single-token decode steps at positions 0..15 (the upstream PyTorch
    class runs positions 0-1 as one causal 2-token pass, which is the same computation).
    """

    def __init__(self, weights: dict, max_seq: int = 64):
        from_cp = weights["code_predictor"]
        keys = ["input_layernorm.weight", "self_attn.q_proj.weight", "self_attn.k_proj.weight",
                "self_attn.v_proj.weight", "self_attn.q_norm.weight", "self_attn.k_norm.weight",
                "self_attn.o_proj.weight", "post_attention_layernorm.weight", "mlp.gate_proj.weight",
                "mlp.up_proj.weight", "mlp.down_proj.weight"]
        flat = [from_cp[f"layers.{i}.{k}"] for i in range(5) for k in keys]
        cos, sin = weights["cos_table"][:max_seq], weights["sin_table"][:max_seq]
        self.stack = LayerStackOracle(flat, from_cp["norm.weight"], cos, sin, max_seq, residual_fp32=False)
        self.lm_heads = [from_cp[f"lm_head.{g}.weight"] for g in range(15)]
        self.codec_embeddings = [from_cp[f"codec_embedding.{g}.weight"] for g in range(15)]

    @torch.no_grad()
    def predict(self, talker_hidden: torch.Tensor, first_codebook_token: int,
                talker_embed_weight: torch.Tensor, do_sample: bool = True, temperature: float = 0.9,
                top_k: int = 50, generator: Optional[torch.Generator] = None,
                forced_tokens=None, record: Optional[list] = None) -> torch.Tensor:
        """Returns int64[16] = [first_token, g0..g14].

        ``forced_tokens`` (15 ints) teacher-forces the fed-back tokens; ``record`` collects
        per-group dicts {logits (f32[2048]), hidden (f32[1024]), token, margin}.
        """
        self.stack.reset()
        self.stack.forward(talker_hidden.to(BF16), 0)
        hn = self.stack.forward(talker_embed_weight[first_codebook_token], 1)
        out = [int(first_codebook_token)]
        for g in range(15):
            logits = _gemv(self.lm_heads[g], hn)                       # bf16[2048]
            tok = select_token(logits, do_sample, temperature, top_k, generator)
            if record is not None:
                lf = logits.float()
                record.append(dict(logits=lf, hidden=hn.float(), token=tok, margin=top2_margin(lf)))
            out.append(tok)
            fed = tok if forced_tokens is None else int(forced_tokens[g])
            if g < 14:
                hn = self.stack.forward(self.codec_embeddings[g][fed], 2 + g)
        return torch.tensor(out, dtype=torch.int64)


def frame_embed_sum(codes, talker_embed: torch.Tensor, cp_embeds, extra_bf16: torch.Tensor) -> torch.Tensor:
    """Next talker input (tts_engine.py:319-333): bf16 adds in this exact order."""
    e = talker_embed[int(codes[0])]
    for g in range(15):
        e = e + cp_embeds[g][int(codes[g + 1])]
    return e + extra_bf16.to(BF16)


class TextProjectionOracle:
    """Upstream ``TextProjection.embed_text_ids`` (model_tts.py:361-374) written out in fp32 with its rounding points:

        x = table[ids];  y = r(x W1^T + b1);  y = r(y / (1 + exp(-y)));  out = r(y W2^T + b2)

    (``F.linear`` on bf16 accumulates in fp32, adds the bias in fp32 and rounds once; ``F.silu`` on bf16 computes in fp32
    and rounds once.)  The matrix products are fp32 matmuls of the bf16 values, i.e. a different summation order than
    torch's bf16 GEMM: equal up to one bf16 ulp, the tolerance of tests/test_text_projection.py."""

    def __init__(self, weights: dict):
        self.table = weights["text_embedding"]
        self.w1, self.b1 = weights["text_proj_fc1_w"].float(), weights["text_proj_fc1_b"].float()
        self.w2, self.b2 = weights["text_proj_fc2_w"].float(), weights["text_proj_fc2_b"].float()

    @torch.no_grad()
    def embed_text_ids(self, token_ids: torch.Tensor) -> torch.Tensor:
        x = self.table[token_ids.long()].float()
        y = (x @ self.w1.T + self.b1).to(BF16).float()
        y = (y / (1.0 + torch.exp(-y))).to(BF16).float()
        return (y @ self.w2.T + self.b2).to(BF16)


def build_prefill_oracle(text_token_ids: torch.Tensor, text_projection, codec_embed_weight: torch.Tensor, pad, bos, eos):
    """Upstream ``build_prefill_embeddings`` with cached pad / bos / eos embeddings (model_tts.py:803-864; the engine's
    configuration, tts_engine.py:107-118): prefill = [3 role tokens | (pad, pad, pad, bos) + codec tags (nothink, think_bos,
    think_eos, codec_pad) | first content token + codec_bos]; trailing = content[1:-5] + [eos].  bf16 adds."""
    emb = text_projection.embed_text_ids(text_token_ids)
    role, content = emb[:3], emb[3:]
    tags = codec_embed_weight[torch.tensor([2155, 2156, 2157, 2148, 2149])]     # model_tts.py:45-53 codec control ids
    fused = torch.cat([pad.expand(3, -1), bos], dim=0) + tags[:4]
    prefill = torch.cat([role, fused, content[:1] + tags[4:5]], dim=0)
    trailing = torch.cat([content[1:-5], eos], dim=0)
    return prefill, trailing

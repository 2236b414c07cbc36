#!/usr/bin/env python3
"""Golden vectors for multimodal RoPE, generated from HF transformers' own functions (the upstream repo implements standard
RoPE only and names M-RoPE with ``mrope_section = [24, 20, 20]`` as its gap, README.md:208):

  * transformers.models.qwen2_vl.modeling_qwen2_vl.apply_multimodal_rotary_pos_emb     (chunked sections)
  * transformers.models.qwen3_vl.modeling_qwen3_vl.Qwen3VLTextRotaryEmbedding.apply_interleaved_mrope (interleaved)

    python tests/golden/make_golden_mrope.py        ->  tests/golden/mrope.npz

cos / sin tables are the upstream bf16 tables (model_tts.py:90-96: fp32 math, halves duplicated, cast to bf16); q / k are
seeded bf16 vectors; all arithmetic in bf16 exactly as HF evaluates it on a bf16 model.
"""

import os
import sys

import numpy as np
import torch
import transformers
from transformers.models.qwen2_vl.modeling_qwen2_vl import apply_multimodal_rotary_pos_emb
from transformers.models.qwen3_vl.modeling_qwen3_vl import Qwen3VLTextRotaryEmbedding

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "qwen-megakernel-tts_b200"))
from qwen_megakernel.synthetic import rope_tables, synthetic_inputs  # noqa: E402

SECTION = [24, 20, 20]
POSITIONS = [(0, 0, 0), (5, 5, 5), (7, 3, 11), (40, 0, 63), (100, 17, 2), (9, 200, 31)]


def bits(t):
    return t.to(torch.bfloat16).contiguous().view(torch.int16).numpy().astype(np.uint16)


def main():
    cos_t, sin_t = rope_tables(256)                                   # bf16 [256, 128]
    pos = torch.tensor(POSITIONS, dtype=torch.long).t().contiguous()    # [3, n]
    n = pos.shape[1]
    # HF layout: cos / sin [3, batch, seq, head_dim], q / k [batch, heads, seq, head_dim]
    cos3 = cos_t[pos].unsqueeze(1)                                    # [3, 1, n, 128]
    sin3 = sin_t[pos].unsqueeze(1)
    q = synthetic_inputs(31, n * 2)[:, :128].reshape(1, 2, n, 128).contiguous()
    k = synthetic_inputs(32, n)[:, :128].reshape(1, 1, n, 128).contiguous()
    q_c, k_c = apply_multimodal_rotary_pos_emb(q, k, cos3, sin3, SECTION)
    # interleaved: HF mixes the FREQUENCIES (64 per axis) before cos / sin; mixing the cos / sin half-rows is the same selection
    cos_i = Qwen3VLTextRotaryEmbedding.apply_interleaved_mrope(None, cos3[..., :64].clone(), SECTION)
    sin_i = Qwen3VLTextRotaryEmbedding.apply_interleaved_mrope(None, sin3[..., :64].clone(), SECTION)
    cos_i, sin_i = torch.cat([cos_i, cos_i], -1), torch.cat([sin_i, sin_i], -1)     # [1, n, 128]
    np.savez_compressed(os.path.join(HERE, "mrope.npz"), section=np.array(SECTION, np.int32),
                        positions=np.array(POSITIONS, np.int32), q_bits=bits(q), k_bits=bits(k),
                        q_chunked_bits=bits(q_c), k_chunked_bits=bits(k_c),
                        cos_interleaved_bits=bits(cos_i[0]), sin_interleaved_bits=bits(sin_i[0]))
    print("wrote mrope.npz; transformers", transformers.__version__, "torch", torch.__version__)


if __name__ == "__main__":
    main()

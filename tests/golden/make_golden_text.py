#!/usr/bin/env python3
"""Golden fixture for the text side of the prefill, from the UPSTREAM classes.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_text.py

Imports, unmodified, upstream qwen_megakernel/model_tts.py (loaded by path under a private alias; its top level needs
only math / struct / typing / torch) and records on CPU (bf16, torch as installed), for the seeded synthetic weights of
qwen_megakernel/synthetic.py with a 512-row text table:
  * TextProjection.embed_text_ids (model_tts.py:361-374) of 150 seeded ids (three 64-token blocks of the kernel: 64 + 64 + 22);
  * build_prefill_embeddings (model_tts.py:776-864) of a 3 + 17 token utterance with cached pad / bos / eos embeddings
    (the engine's configuration, tts_engine.py:107-118): prefill [8, 1024] and trailing text [12, 1024].
Output (committed): tests/golden/text_projection.npz; bf16 values are stored as bit patterns (uint16).
"""

import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("QMK_REFERENCE_ROOT", "/root/reference")
TEXT_VOCAB, N_IDS, SEED_IDS = 512, 150, 2468


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def bf16_bits(t: torch.Tensor) -> np.ndarray:
    return t.to(torch.bfloat16).contiguous().view(torch.int16).numpy().astype(np.uint16)


def text_case():
    """The ids of the fixture (shared with the tests through the npz)."""
    gen = torch.Generator().manual_seed(SEED_IDS)
    ids = torch.randint(0, TEXT_VOCAB, (N_IDS,), generator=gen)
    utter = torch.randint(0, TEXT_VOCAB - 3, (20,), generator=gen)     # 3 role ids + 17 content ids (the last 5 are format tokens)
    special = torch.tensor([TEXT_VOCAB - 3, TEXT_VOCAB - 2, TEXT_VOCAB - 1])   # stand-ins for TTS_PAD / TTS_BOS / TTS_EOS
    return ids, utter, special


def main():
    synth = _load("_qmk_synthetic", os.path.join(REPO, "qwen-megakernel-tts_b200", "qwen_megakernel", "synthetic.py"))
    ref = _load("_ref_model_tts", os.path.join(REF, "qwen_megakernel", "model_tts.py"))
    torch.set_num_threads(os.cpu_count() or 1)
    w = synth.synthetic_tts_weights(seed=1234, num_layers=1, max_seq_len=64, text_vocab=TEXT_VOCAB)
    tp = ref.TextProjection(w, device="cpu")
    ids, utter, special = text_case()
    out = tp.embed_text_ids(ids)
    sp = tp.embed_text_ids(special)
    cached = {"pad": sp[0:1], "bos": sp[1:2], "eos": sp[2:3]}
    prefill, trailing = ref.build_prefill_embeddings(utter, tp, w["embed_weight"], device="cpu", cached_tts_embeds=cached)
    assert out.shape == (N_IDS, 1024) and prefill.shape == (8, 1024) and trailing.shape == (12, 1024)
    np.savez_compressed(os.path.join(HERE, "text_projection.npz"), ids=ids.numpy(), out=bf16_bits(out), utter=utter.numpy(),
                        special=special.numpy(), prefill=bf16_bits(prefill), trailing=bf16_bits(trailing),
                        text_vocab=np.int64(TEXT_VOCAB), torch_version=np.array(torch.__version__))
    print("wrote text_projection.npz:", out.shape, prefill.shape, trailing.shape, "torch", torch.__version__)


if __name__ == "__main__":
    sys.exit(main())

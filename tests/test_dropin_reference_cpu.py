"""CPU: the drop-in boundary checked against the REFERENCE tree itself (skipped where /root/reference is absent, i.e. on the
GPU box -- the reference cannot travel and is never copied into this repo).

  1. upstream ``qwen_megakernel/tts_engine.py`` is loaded FROM ITS ORIGINAL PATH as a submodule of THIS repo's
     ``qwen_megakernel`` package: its ``from .model_tts import ...`` must resolve to the new model layer.
  2. every callable the engine uses binds with the exact call patterns of tts_engine.py / pipecat_tts.py.
  3. the engine's own ``initialize()`` and ``_generate_codec_frames()`` run unmodified over this package with a tokenizer stub
     and no vocoder; the two kernel-backed classes (which need a GPU) are replaced by CPU doubles that restate the same
     contract on the oracle, so what is exercised is upstream's control flow on top of this package's loader contract,
     ``TextProjection`` and constants.  Its frames must equal the Appendix-B loop evaluated directly on the oracle.
  4. the re-typed ``TextProjection`` / ``build_prefill_embeddings`` / ``CodePredictor`` equal the reference's, bit for bit.
"""

import importlib.util
import inspect
import os
import sys
import types

import pytest
import torch

REF = os.environ.get("QMK_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "qwen_megakernel", "tts_engine.py")),
                                reason="reference tree not present (GPU box)")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def ref_model_tts():
    """The reference model layer under a private alias (its top level imports only math / struct / typing / torch)."""
    return _load("_ref_model_tts", os.path.join(REF, "qwen_megakernel", "model_tts.py"))


@pytest.fixture(scope="module")
def ref_engine_module():
    import qwen_megakernel  # this repo's package
    assert os.path.realpath(os.path.dirname(qwen_megakernel.__file__)).startswith(os.path.realpath(os.path.dirname(os.path.dirname(__file__))))
    mod = _load("qwen_megakernel.tts_engine", os.path.join(REF, "qwen_megakernel", "tts_engine.py"))
    yield mod
    sys.modules.pop("qwen_megakernel.tts_engine", None)


def test_reference_engine_binds_to_this_model_layer(ref_engine_module, ref_model_tts):
    from qwen_megakernel import model_tts as ours
    eng = ref_engine_module
    for name in ("CodePredictorKernel", "TextProjection", "TTSDecoder", "load_tts_weights"):
        assert getattr(eng, name) is getattr(ours, name), name
    for name in ("CODEC_BOS", "CODEC_EOS", "CODEC_NOTHINK", "CODEC_PAD", "CODEC_THINK_BOS", "CODEC_THINK_EOS", "NUM_CODE_GROUPS",
                 "TTS_BOS", "TTS_EOS", "TTS_PAD"):
        assert getattr(eng, name) == getattr(ref_model_tts, name), name
    # every public constant of the reference model layer exists here with the same value
    for name, val in vars(ref_model_tts).items():
        if name.isupper() and isinstance(val, (int, float)):
            assert getattr(ours, name) == val, name


def test_call_patterns_of_the_reference_engine_bind(ref_model_tts):
    """tts_engine.py:82-96, 141-148, 281-335: the exact argument patterns (positional / keyword) must bind, and must bind the
    same way on the reference's own signatures."""
    from qwen_megakernel import model_tts as ours
    w, t = object(), object()
    for mod in (ours, ref_model_tts):
        inspect.signature(mod.load_tts_weights).bind("Qwen/Qwen3-TTS-12Hz-0.6B-Base", device="cuda", verbose=True)
        inspect.signature(mod.TTSDecoder.__init__).bind(None, weights=w)
        inspect.signature(mod.TextProjection.__init__).bind(None, w, device="cuda")
        inspect.signature(mod.CodePredictorKernel.__init__).bind(None, w, device="cuda")
        inspect.signature(mod.TTSDecoder.step).bind(None, 2149)
        inspect.signature(mod.TTSDecoder.step_with_embed).bind(None, t)
        inspect.signature(mod.TTSDecoder.reset).bind(None)
        inspect.signature(mod.CodePredictorKernel.predict).bind(None, talker_hidden=t, first_codebook_token=3, talker_embed_weight=t,
                                                               do_sample=True, temperature=0.9, top_k=50)
        inspect.signature(mod.CodePredictorKernel.predict).bind(None, t, 0, t, do_sample=False, temperature=0.9, top_k=50)
        inspect.signature(mod.build_prefill_embeddings).bind(t, t, t, language="Auto", device="cuda", cached_tts_embeds=None)
        assert isinstance(mod.TTSDecoder.position, property) and isinstance(mod.TTSDecoder.embed_weight, property)
    # same parameter names in the same order for the upstream part of every signature (ours may append keyword-only extras)
    for cls, meth in (("TTSDecoder", "__init__"), ("TTSDecoder", "step"), ("TTSDecoder", "step_with_embed"),
                      ("CodePredictorKernel", "__init__"), ("CodePredictorKernel", "predict"), ("TextProjection", "embed_text_ids")):
        ref_p = list(inspect.signature(getattr(getattr(ref_model_tts, cls), meth)).parameters.values())
        our_p = list(inspect.signature(getattr(getattr(ours, cls), meth)).parameters.values())
        assert [p.name for p in our_p[:len(ref_p)]] == [p.name for p in ref_p], (cls, meth)
        assert [p.default for p in our_p[:len(ref_p)]] == [p.default for p in ref_p], (cls, meth)
        assert all(p.kind is inspect.Parameter.KEYWORD_ONLY for p in our_p[len(ref_p):]), (cls, meth)


class _Tokenizer:
    """Stand-in for AutoTokenizer: deterministic ids < 1000, the 3 role tokens in front, 5 format tokens behind."""

    def encode(self, text, return_tensors=None):
        words = text.replace("<|im_start|>", " \x01 ").replace("<|im_end|>", " \x02 ").replace("\n", " \x03 ").split()
        ids = [sum(ord(c) for c in w) % 997 + 1 for w in words]
        return torch.tensor([ids], dtype=torch.long)


def test_reference_engine_runs_unmodified_over_this_package(ref_engine_module, cpu_weights, monkeypatch):
    from oracle.tts_oracle import CodePredictorOracle, TalkerOracle, frame_embed_sum
    from qwen_megakernel import model_tts as ours
    eng_mod = ref_engine_module
    w = dict(cpu_weights)
    gen = torch.Generator().manual_seed(11)
    w["text_embedding"] = torch.empty(151936, 2048, dtype=torch.bfloat16).normal_(generator=gen)   # TTS_PAD/BOS/EOS ids are looked up
    calls = {"talker_init": 0, "cp_init": 0, "step": 0, "embed": 0, "predict": 0, "reset": 0}

    class TalkerDouble:            # the TTSDecoder contract (model_tts.py:196-345) on the CPU oracle
        def __init__(self, weights=None, model_path="x", verbose=True):
            assert weights is w
            calls["talker_init"] += 1
            self._o = TalkerOracle(weights, max_seq=256)

        def reset(self):
            calls["reset"] += 1
            self._o.reset()

        def step(self, token_id):
            calls["step"] += 1
            assert isinstance(token_id, int)
            return self._o.step(token_id)

        def step_with_embed(self, e):
            calls["embed"] += 1
            assert e.shape == (1024,) and e.dtype == torch.bfloat16
            return self._o.step_with_embed(e)

    class PredictorDouble:         # the CodePredictorKernel contract (model_tts.py:622-773)
        def __init__(self, weights, device="cuda"):
            calls["cp_init"] += 1
            self._o = CodePredictorOracle(weights)

        def predict(self, talker_hidden, first_codebook_token, talker_embed_weight, do_sample=True, temperature=0.9, top_k=50):
            calls["predict"] += 1
            assert talker_hidden.dtype == torch.float32 and talker_hidden.shape == (1024,)
            return self._o.predict(talker_hidden, int(first_codebook_token), talker_embed_weight, do_sample=False)

    monkeypatch.setattr(eng_mod, "load_tts_weights", lambda path, device="cuda", verbose=True: w)
    monkeypatch.setattr(eng_mod, "TTSDecoder", TalkerDouble)
    monkeypatch.setattr(eng_mod, "CodePredictorKernel", PredictorDouble)
    import transformers
    monkeypatch.setattr(transformers.AutoTokenizer, "from_pretrained", classmethod(lambda cls, *a, **k: _Tokenizer()))
    monkeypatch.setattr(eng_mod.MegakernelTTSEngine, "_load_vocoder",
                        lambda self, path: (setattr(self, "speech_tokenizer", None), setattr(self, "sample_rate", 24000)))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)

    engine = eng_mod.MegakernelTTSEngine(eng_mod.TTSConfig(subtalker_do_sample=False), device="cpu")
    engine.initialize()
    assert isinstance(engine.text_projection, ours.TextProjection)          # THIS package's class, constructed by upstream code
    assert calls["talker_init"] == 1 and calls["cp_init"] == 1 and calls["predict"] == 5     # warm-up loop, tts_engine.py:141-148
    text = "hello world"
    frames = [f.clone() for f in engine._generate_codec_frames(text)]
    assert len(frames) == 25 and all(f.dtype == torch.int64 and f.shape == (16,) for f in frames)   # max(int(2 words / 2.5 * 12.5 * 2), 25), tts_engine.py:294-298
    assert calls["embed"] == 8 + 25 and calls["step"] == 5 + 1

    # the same utterance evaluated directly with this package's helpers and the oracle
    tok = _Tokenizer()
    ids = tok.encode(f"<|im_start|>assistant\n{text}<|im_end|>\n<|im_start|>assistant\n")[0]
    tp = ours.TextProjection(w, device="cpu")
    prefill, trailing = ours.build_prefill_embeddings(ids, tp, w["embed_weight"], device="cpu")
    assert prefill.shape == (8, 1024) and trailing.shape[1] == 1024
    talker, cp = TalkerOracle(w, max_seq=256), CodePredictorOracle(w)
    pad = tp.embed_text_ids(torch.tensor([ours.TTS_PAD]))[0].to(torch.bfloat16)
    for i in range(8):
        talker.step_with_embed(prefill[i])
    t, h = talker.step(ours.CODEC_BOS)
    for f in range(25):
        codes = cp.predict(h, t, w["embed_weight"], do_sample=False)
        assert codes.tolist() == frames[f].tolist(), f"frame {f}"
        t, h = talker.step_with_embed(frame_embed_sum(codes, w["embed_weight"], cp.codec_embeddings,
                                                      trailing[f] if f < trailing.shape[0] else pad))


def test_retyped_helpers_equal_the_reference(ref_model_tts, cpu_weights):
    """TextProjection / build_prefill_embeddings / CodePredictor are plain PyTorch glue re-typed in this repo (the reference's
    files are not copied); they must reproduce the reference's outputs bit for bit on the same weights."""
    from qwen_megakernel import model_tts as ours
    w = dict(cpu_weights)
    gen = torch.Generator().manual_seed(5)
    w["text_embedding"] = torch.empty(151936, 2048, dtype=torch.bfloat16).normal_(generator=gen)
    ids = torch.randint(0, 151000, (19,), generator=gen)
    tp_o, tp_r = ours.TextProjection(w, device="cpu"), ref_model_tts.TextProjection(w, device="cpu")
    assert torch.equal(tp_o.embed_text_ids(ids), tp_r.embed_text_ids(ids))
    assert torch.equal(tp_o.embed_text_ids(ids.view(1, -1)), tp_r.embed_text_ids(ids.view(1, -1)))
    for cached in (None, "cached"):
        c = None
        if cached:
            sp = tp_r.embed_text_ids(torch.tensor([ours.TTS_PAD, ours.TTS_BOS, ours.TTS_EOS]))
            c = {"pad": sp[0:1], "bos": sp[1:2], "eos": sp[2:3]}
        a = ours.build_prefill_embeddings(ids, tp_o, w["embed_weight"], device="cpu", cached_tts_embeds=c)
        b = ref_model_tts.build_prefill_embeddings(ids, tp_r, w["embed_weight"], device="cpu", cached_tts_embeds=c)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    # CodePredictor (free-running greedy): equal to the fixtures recorded from the reference class up to the first group whose
    # reference top-2 margin is inside the tie rule (the two evaluate attention in a different summation order)
    from conftest import bf16_from_bits
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cp_config2.npz"))
    cp_o = ours.CodePredictor(w, device="cpu")
    for f in range(3):
        th = bf16_from_bits(g["talker_hidden_bits"][f]).float()
        out = cp_o.predict(th, int(g["first_tokens"][f]), w["embed_weight"], do_sample=False).tolist()
        for grp in range(15):
            if out[grp + 1] != int(g["tokens"][f][grp]):
                assert g["margins"][f][grp] <= 1e-2, f"frame {f} group {grp}: diverged at margin {g['margins'][f][grp]}"
                break

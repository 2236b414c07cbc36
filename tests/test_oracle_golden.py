"""CPU: pin the oracle restatement (oracle/tts_oracle.py) against fixtures recorded from the UPSTREAM
classes (tests/golden/make_golden.py).  Teacher-forced: each step is fed the reference's token."""

import torch

from conftest import bf16_from_bits
from oracle.tts_oracle import CodePredictorOracle, TalkerOracle
from parity import assert_parity, compare

CODEC_BOS = 2149


def test_talker_config1_oracle_matches_reference(cpu_weights, golden):
    g = golden["talker_config1"]
    prefill = bf16_from_bits(g["prefill_bits"])
    orc = TalkerOracle(cpu_weights, max_seq=128)
    toks, hids = [], []
    for i in range(len(g["tokens"])):
        if i < prefill.shape[0]:
            t, h = orc.step_with_embed(prefill[i])
        elif i == prefill.shape[0]:
            t, h = orc.step(CODEC_BOS)
        else:
            t, h = orc.step(int(g["tokens"][i - 1]))
        toks.append(t)
        hids.append(h)
    ref_h = [bf16_from_bits(b).float() for b in g["hidden_bits"]]
    assert_parity(compare("oracle-vs-upstream talker config1", toks, hids, g["tokens"], g["margins"], ref_h))


def test_talker_mixed_oracle_matches_reference(cpu_weights, golden):
    g = golden["talker_mixed"]
    emb = bf16_from_bits(g["embed_bits"])
    orc = TalkerOracle(cpu_weights, max_seq=128)
    toks, hids = [], []
    for i in range(len(g["tokens"])):
        if i == 0:
            t, h = orc.step(CODEC_BOS)
        elif i % 2 == 0:
            t, h = orc.step_with_embed(emb[i])
        else:
            t, h = orc.step(int(g["tokens"][i - 1]))
        toks.append(t)
        hids.append(h)
    ref_h = [bf16_from_bits(b).float() for b in g["hidden_bits"]]
    assert_parity(compare("oracle-vs-upstream talker mixed", toks, hids, g["tokens"], g["margins"], ref_h))


def test_code_predictor_config2_oracle_matches_reference(cpu_weights, golden):
    g = golden["cp_config2"]
    cp = CodePredictorOracle(cpu_weights)
    for f in range(g["tokens"].shape[0]):
        rec = []
        th = bf16_from_bits(g["talker_hidden_bits"][f]).float()
        out = cp.predict(th, int(g["first_tokens"][f]), cpu_weights["embed_weight"], do_sample=False,
                         forced_tokens=g["tokens"][f], record=rec)
        assert out.dtype == torch.int64 and out.shape == (16,) and int(out[0]) == int(g["first_tokens"][f])
        ref_h = [bf16_from_bits(b).float() for b in g["hidden_bits"][f]]
        rep = compare(f"oracle-vs-upstream cp frame {f}", [r["token"] for r in rec], [r["hidden"] for r in rec],
                      g["tokens"][f], g["margins"][f], ref_h)
        assert_parity(rep)


def test_oracle_sampling_is_topk_restricted(cpu_weights):
    """Sampled tokens must lie in the top-k set of the teacher-forced logits (model_tts.py:756-762)."""
    cp = CodePredictorOracle(cpu_weights)
    gen = torch.Generator().manual_seed(5)
    rec = []
    th = torch.zeros(1024)
    th[::3] = 1.0
    out = cp.predict(th, 7, cpu_weights["embed_weight"], do_sample=True, temperature=0.9, top_k=50,
                     generator=gen, record=rec)
    for g, r in enumerate(rec):
        kth = torch.topk(r["logits"], 50).values[-1]
        assert r["logits"][int(out[g + 1])] >= kth

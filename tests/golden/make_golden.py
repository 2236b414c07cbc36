#!/usr/bin/env python3
"""Generate the golden fixtures by running the UPSTREAM reference classes on the synthetic checkpoint.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports, unmodified, from /root/reference:
  * validate_kernel.PyTorchTalkerReference            (validate_kernel.py:25-201)
  * qwen_megakernel.model_tts.CodePredictor           (model_tts.py:377-619)
and records what they compute on CPU (bf16, torch as installed) for the seeded synthetic weights of
qwen-megakernel-tts_b200/qwen_megakernel/synthetic.py.  Outputs (committed):
  tests/golden/talker_config1.npz   8 prefill embeds + 50 greedy steps (BASELINE.json configs[0])
  tests/golden/talker_mixed.npz     BOS + 39 steps alternating caller embeddings / token feedback
  tests/golden/cp_config2.npz       code-predictor frames, greedy        (BASELINE.json configs[1])
  tests/golden/meta.json            torch version, cpu capability, weights fingerprint
  tests/golden/frame_loop.npz       (``--only frames``) the full frame loop of tts_engine.py:246-335, greedy: 8 prefill
                                    steps, step(CODEC_BOS), then N_FRAMES x {CodePredictor.predict, 16-way embedding sum +
                                    trailing text / pad, step_with_embed}: talker positions run to 8 + N_FRAMES, i.e. past
                                    the attention boundaries at 40 / 80 cached positions at full depth (28 layers)
Hidden states are stored as bf16 bit patterns (uint16); logits margins as float32.
"""

import importlib.util
import json
import os
import platform
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("QMK_REFERENCE_ROOT", "/root/reference")

CODEC_BOS = 2149
SEED_WEIGHTS = 1234
SEED_PREFILL = 99
N_PREFILL, N_DECODE = 8, 50
N_CP_FRAMES = 6
N_MIXED = 40
N_FRAMES, N_TRAILING = 112, 40
SEED_TRAILING, SEED_PAD = 4321, 777


def _load_synth():
    path = os.path.join(REPO, "qwen-megakernel-tts_b200", "qwen_megakernel", "synthetic.py")
    spec = importlib.util.spec_from_file_location("_qmk_synthetic", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def bf16_bits(t: torch.Tensor) -> np.ndarray:
    return t.to(torch.bfloat16).contiguous().view(torch.int16).numpy().astype(np.uint16)


def cp_frame_replay(cp, w, talker_hidden, first_token):
    """One greedy frame through the upstream CodePredictor's own layer methods, recording what predict() hides:
    per-group tokens, top-2 margins and post-norm hidden states."""
    cp._reset_cache()
    first_embed = w["embed_weight"][first_token]
    h = torch.stack([talker_hidden.to(torch.bfloat16), first_embed], 0).unsqueeze(0)
    for li, lw in enumerate(cp.layers):
        h = cp._layer_prefill(h, lw, li, seq_len=2)
    last = cp._rms_norm(h, cp.final_norm)[:, -1:, :]
    toks, mar, hid_bits, top1 = [], [], [], []
    pos = 2
    for g in range(cp.num_groups):
        logits = torch.nn.functional.linear(last, cp.lm_heads[g]).reshape(-1).float()
        t2 = torch.topk(logits, 2).values
        tok = int(logits.argmax())
        toks.append(tok)
        mar.append(float(t2[0] - t2[1]))
        top1.append(float(t2[0]))
        hid_bits.append(bf16_bits(last.reshape(-1)))
        if g < cp.num_groups - 1:
            h = cp.codec_embeddings[g][tok].view(1, 1, -1)
            for li, lw in enumerate(cp.layers):
                h = cp._layer_decode(h, lw, li, pos)
            last = cp._rms_norm(h, cp.final_norm)
            pos += 1
    return toks, mar, hid_bits, top1


def make_frame_loop(synth, ref_vk, ref_mt, w):
    """tts_engine.py:281-335 with the upstream reference classes (greedy), synthetic prefill / trailing text."""
    ref = ref_vk.PyTorchTalkerReference(w, device="cpu")
    cp = ref_mt.CodePredictor(w, device="cpu")
    prefill = synth.synthetic_inputs(SEED_PREFILL, N_PREFILL)
    trailing = synth.synthetic_inputs(SEED_TRAILING, N_TRAILING)
    pad = synth.synthetic_inputs(SEED_PAD, 1)[0]
    cp_embeds = [w["code_predictor"][f"codec_embedding.{g}.weight"] for g in range(15)]
    ref.reset()
    for i in range(N_PREFILL):
        ref.step_with_embed(prefill[i])
    tok, hid = ref.step(CODEC_BOS)
    rec = dict(codes=[], cp_margins=[], cp_top1=[], talker_top1=[], talker_tokens=[], talker_margins=[], talker_hidden=[], in_hidden=[bf16_bits(hid)],
               in_token=tok)
    for f in range(N_FRAMES):
        toks, mar, _, top1 = cp_frame_replay(cp, w, hid, tok)
        want = cp.predict(hid, tok, w["embed_weight"], do_sample=False)
        assert want.tolist() == [tok] + toks
        codes = [tok] + toks
        e = torch.nn.functional.embedding(torch.tensor(codes[0:1]), w["embed_weight"]).squeeze(0)
        for g in range(15):
            e = e + torch.nn.functional.embedding(torch.tensor(codes[g + 1:g + 2]), cp_embeds[g]).squeeze(0)
        e = e + (trailing[f].to(torch.bfloat16) if f < N_TRAILING else pad)
        tok, hid = ref.step_with_embed(e)
        logits = torch.nn.functional.linear(hid.to(torch.bfloat16), w["lm_head_weight"]).float()
        t2 = torch.topk(logits, 2).values
        assert int(logits.argmax()) == tok
        rec["codes"].append(codes); rec["cp_margins"].append(mar); rec["talker_tokens"].append(tok)
        rec["cp_top1"].append(top1); rec["talker_top1"].append(float(t2[0]))
        rec["talker_margins"].append(float(t2[0] - t2[1])); rec["talker_hidden"].append(bf16_bits(hid))
        if f % 10 == 0:
            print("frame", f, codes[:4], "->", tok, flush=True)
    np.savez_compressed(
        os.path.join(HERE, "frame_loop.npz"),
        prefill_bits=bf16_bits(prefill), trailing_bits=bf16_bits(trailing), pad_bits=bf16_bits(pad),
        first_token=np.int32(rec["in_token"]), first_hidden_bits=rec["in_hidden"][0],
        codes=np.array(rec["codes"], np.int32), cp_margins=np.array(rec["cp_margins"], np.float32),
        cp_top1=np.array(rec["cp_top1"], np.float32), talker_top1=np.array(rec["talker_top1"], np.float32),
        talker_tokens=np.array(rec["talker_tokens"], np.int32), talker_margins=np.array(rec["talker_margins"], np.float32),
        talker_hidden_bits=np.stack(rec["talker_hidden"]),
    )
    print("wrote frame_loop.npz:", N_FRAMES, "frames; min cp margin", float(np.min(rec["cp_margins"])),
          "min talker margin", min(rec["talker_margins"]))


def main():
    warnings.filterwarnings("ignore", category=UserWarning)
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    synth = _load_synth()
    sys.path.insert(0, REF)                       # upstream package wins the name qwen_megakernel here
    import validate_kernel as ref_vk              # noqa: E402
    from qwen_megakernel import model_tts as ref_mt  # noqa: E402

    w = synth.synthetic_tts_weights(seed=SEED_WEIGHTS)
    fp = synth.weights_fingerprint(w)
    print("weights fingerprint", fp)
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "frames":
        make_frame_loop(synth, ref_vk, ref_mt, w)
        return

    # ── config 1: talker ────────────────────────────────────────────────────────────────────
    ref = ref_vk.PyTorchTalkerReference(w, device="cpu")
    prefill = synth.synthetic_inputs(SEED_PREFILL, N_PREFILL)
    tokens, margins, hiddens, top1 = [], [], [], []

    def record(tok, hid):
        logits = torch.nn.functional.linear(hid.to(torch.bfloat16), w["lm_head_weight"]).float()
        t2 = torch.topk(logits, 2).values
        assert int(logits.argmax()) == tok
        tokens.append(tok)
        margins.append(float(t2[0] - t2[1]))
        top1.append(float(t2[0]))
        hiddens.append(bf16_bits(hid))

    ref.reset()
    for i in range(N_PREFILL):
        record(*ref.step_with_embed(prefill[i]))
    tok, hid = ref.step(CODEC_BOS)
    record(tok, hid)
    for _ in range(N_DECODE - 1):
        tok, hid = ref.step(tok)
        record(tok, hid)
    np.savez_compressed(
        os.path.join(HERE, "talker_config1.npz"),
        prefill_bits=bf16_bits(prefill), tokens=np.array(tokens, np.int32),
        margins=np.array(margins, np.float32), top1=np.array(top1, np.float32),
        hidden_bits=np.stack(hiddens),
    )
    print("talker tokens", tokens[:16], "min margin", min(margins))

    # ── scenario "mixed" (validate_kernel.py:305-337 pattern): BOS, then alternate caller-supplied
    #    embeddings (sentinel path) with token feedback, so both input modes and many distinct
    #    tokens are covered ────────────────────────────────────────────────────────────────────────
    tokens, margins, hiddens, top1 = [], [], [], []
    emb = synth.synthetic_inputs(SEED_PREFILL + 1, N_MIXED)
    ref.reset()
    tok, hid = ref.step(CODEC_BOS)
    record(tok, hid)
    for i in range(1, N_MIXED):
        tok, hid = ref.step_with_embed(emb[i]) if i % 2 == 0 else ref.step(tok)
        record(tok, hid)
    np.savez_compressed(
        os.path.join(HERE, "talker_mixed.npz"),
        embed_bits=bf16_bits(emb), tokens=np.array(tokens, np.int32),
        margins=np.array(margins, np.float32), top1=np.array(top1, np.float32),
        hidden_bits=np.stack(hiddens),
    )
    print("mixed tokens", tokens[:16], "min margin", min(margins))
    del ref

    # ── config 2: code predictor (greedy), driven through the upstream class's own layer methods so
    #    that logits and hidden states can be recorded (predict() itself only returns tokens) ─────────
    cp = ref_mt.CodePredictor(w, device="cpu")
    frames = []
    for f in range(N_CP_FRAMES):
        talker_hidden = synth.synthetic_inputs(1000 + f, 1)[0].float()          # bf16-valued fp32
        first_token = int((1335 + 977 * f) % 3072)
        want = cp.predict(talker_hidden, first_token, w["embed_weight"], do_sample=False)

        cp._reset_cache()
        first_embed = w["embed_weight"][first_token]
        h = torch.stack([talker_hidden.to(torch.bfloat16), first_embed], 0).unsqueeze(0)
        for li, lw in enumerate(cp.layers):
            h = cp._layer_prefill(h, lw, li, seq_len=2)
        last = cp._rms_norm(h, cp.final_norm)[:, -1:, :]
        toks, mar, hid_bits = [], [], []
        pos = 2
        for g in range(cp.num_groups):
            logits = torch.nn.functional.linear(last, cp.lm_heads[g]).reshape(-1).float()
            t2 = torch.topk(logits, 2).values
            tok = int(logits.argmax())
            toks.append(tok)
            mar.append(float(t2[0] - t2[1]))
            hid_bits.append(bf16_bits(last.reshape(-1)))
            if g < cp.num_groups - 1:
                h = cp.codec_embeddings[g][tok].view(1, 1, -1)
                for li, lw in enumerate(cp.layers):
                    h = cp._layer_decode(h, lw, li, pos)
                last = cp._rms_norm(h, cp.final_norm)
                pos += 1
        assert want.tolist() == [first_token] + toks, "replayed loop must equal upstream predict()"
        frames.append(dict(hidden=bf16_bits(talker_hidden), first=first_token, toks=toks, mar=mar, hid=np.stack(hid_bits)))
        print("cp frame", f, [first_token] + toks, "min margin", min(mar))
    np.savez_compressed(
        os.path.join(HERE, "cp_config2.npz"),
        talker_hidden_bits=np.stack([fr["hidden"] for fr in frames]),
        first_tokens=np.array([fr["first"] for fr in frames], np.int32),
        tokens=np.array([fr["toks"] for fr in frames], np.int32),
        margins=np.array([fr["mar"] for fr in frames], np.float32),
        hidden_bits=np.stack([fr["hid"] for fr in frames]),
    )

    meta = dict(
        generator="tests/golden/make_golden.py",
        reference="jayanth-kumar-morem/qwen-megakernel-tts (validate_kernel.PyTorchTalkerReference, model_tts.CodePredictor)",
        torch=torch.__version__, numpy=np.__version__,
        cpu_capability=torch.backends.cpu.get_cpu_capability(), machine=platform.machine(),
        threads=torch.get_num_threads(),
        weights_fingerprint=fp, seed_weights=SEED_WEIGHTS, seed_prefill=SEED_PREFILL,
        n_prefill=N_PREFILL, n_decode=N_DECODE, n_cp_frames=N_CP_FRAMES, n_mixed=N_MIXED,
        recipe=synth.RECIPE_VERSION,
    )
    with open(os.path.join(HERE, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote goldens:", meta)


if __name__ == "__main__":
    main()

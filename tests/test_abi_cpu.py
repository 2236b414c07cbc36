"""CPU: the C-ABI shared library loads and exports every symbol include/qmk_b200.h declares; host-side
logic of the Python surface that needs no GPU."""

import os
import re

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "qmk_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:qmk_|launch_ldg_)[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    from qwen_megakernel import build_tts
    path = build_tts.build()
    lib = build_tts.load_library(path)
    names = _declared_symbols()
    assert "launch_ldg_decode_direct" in names and "qmk_decode_step" in names and "qmk_cp_predict" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in qmk_b200.h but not exported"
        assert n in build_tts.SIGNATURES, f"{n} has no ctypes signature"
    assert lib.qmk_abi_version() == 3


def test_library_is_never_stale():
    """The library carries the hash of the sources it was compiled from; build() rebuilds on any difference (every csrc/*.cu,
    csrc/*.cuh and include/*.h is hashed, so an edited kernel header cannot leave a stale binary behind)."""
    from qwen_megakernel import build_tts
    path = build_tts.build()
    names = {os.path.basename(f) for f in build_tts.source_files()}
    assert {"qmk_engine.cu", "qmk_batched.cu", "qmk_device.cuh", "qmk_device2.cuh", "qmk_bgemm.cuh", "qmk_b200.h"} <= names
    assert build_tts.library_hash(path) == build_tts.source_hash()


def test_generate_args_mirror_matches_the_c_struct():
    """ctypes mirror of qmk_generate_args: same size as the C struct the library was compiled with."""
    import ctypes
    from qwen_megakernel import build_tts
    lib = build_tts.load_library(build_tts.build())
    assert lib.qmk_generate_args_size() == ctypes.sizeof(build_tts.GenerateArgs)
    assert lib.qmk_generate_nosync(None, None) == -1          # null args: QMK_ERR_ARG, no GPU touched


def test_import_does_not_load_native_code():
    import importlib
    import qwen_megakernel
    importlib.reload(qwen_megakernel)
    assert qwen_megakernel.__all__ == []


def test_constants_match_upstream_surface():
    from qwen_megakernel import model_tts as m
    assert (m.NUM_LAYERS, m.NUM_KV_HEADS, m.NUM_Q_HEADS, m.HEAD_DIM, m.HIDDEN_SIZE, m.INTERMEDIATE_SIZE) == (28, 8, 16, 128, 1024, 3072)
    assert (m.Q_SIZE, m.KV_SIZE, m.VOCAB_SIZE, m.MAX_SEQ_LEN, m.ROPE_THETA) == (2048, 1024, 3072, 8192, 1000000.0)
    assert (m.NUM_CODE_GROUPS, m.CODE_PREDICTOR_LAYERS, m.CODE_PREDICTOR_VOCAB) == (16, 5, 2048)
    assert (m.CODEC_BOS, m.CODEC_EOS, m.CODEC_PAD, m.EMBED_FROM_BUFFER) == (2149, 2150, 2148, -1)
    assert (m.TTS_BOS, m.TTS_EOS, m.TTS_PAD) == (151672, 151673, 151671)
    for name in ("load_tts_weights", "TTSDecoder", "CodePredictorKernel", "CodePredictor", "TextProjection",
                 "build_prefill_embeddings", "_pack_layer_weights"):
        assert hasattr(m, name)


def test_product_path_fails_loudly_without_cuda(cpu_weights):
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from qwen_megakernel.model_tts import CodePredictorKernel, TTSDecoder
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        TTSDecoder(weights=cpu_weights, verbose=False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CodePredictorKernel(cpu_weights, device="cuda")


def test_pytorch_code_predictor_matches_oracle(cpu_weights):
    """The kept pure-PyTorch CodePredictor (same role as upstream's) agrees with the oracle on CPU."""
    from oracle.tts_oracle import CodePredictorOracle
    from qwen_megakernel.model_tts import CodePredictor
    from qwen_megakernel.synthetic import synthetic_inputs
    h = synthetic_inputs(4242, 1)[0].float()
    a = CodePredictor(cpu_weights, device="cpu").predict(h, 11, cpu_weights["embed_weight"], do_sample=False)
    b = CodePredictorOracle(cpu_weights).predict(h, 11, cpu_weights["embed_weight"], do_sample=False)
    assert a.dtype == torch.int64 and a.shape == (16,)
    assert (a == b).float().mean() >= 14 / 16   # free-running; near-ties may diverge late


def test_load_tts_weights_key_layout(tmp_path):
    """Loader reads the upstream safetensors key names (model_tts.py:101-146) — tiny fake shapes."""
    from safetensors.torch import save_file
    from qwen_megakernel import model_tts as m
    t = lambda *s: torch.zeros(*s, dtype=torch.bfloat16)
    state = {}
    for i in range(28):
        for f in m._LAYER_FIELDS:
            state[f"talker.model.layers.{i}.{f}"] = t(2)
    for i in range(5):
        for f in m._LAYER_FIELDS:
            state[f"talker.code_predictor.model.layers.{i}.{f}"] = t(2)
    for k in ("talker.model.codec_embedding.weight", "talker.codec_head.weight", "talker.model.norm.weight",
              "talker.model.text_embedding.weight", "talker.text_projection.linear_fc1.weight",
              "talker.text_projection.linear_fc1.bias", "talker.text_projection.linear_fc2.weight",
              "talker.text_projection.linear_fc2.bias", "talker.code_predictor.model.norm.weight",
              "speaker_encoder.x"):
        state[k] = t(2)
    for g in range(15):
        state[f"talker.code_predictor.lm_head.{g}.weight"] = t(2)
        state[f"talker.code_predictor.model.codec_embedding.{g}.weight"] = t(2)
    save_file(state, str(tmp_path / "model.safetensors"))
    w = m.load_tts_weights(str(tmp_path), device="cpu", verbose=False)
    assert len(w["layer_weights"]) == 308 and w["cos_table"].shape == (8192, 128)
    assert set(w) >= {"embed_weight", "lm_head_weight", "final_norm_weight", "layer_weights", "cos_table", "sin_table",
                      "text_embedding", "text_proj_fc1_w", "text_proj_fc1_b", "text_proj_fc2_w", "text_proj_fc2_b",
                      "code_predictor", "speaker_encoder"}
    assert len(w["code_predictor"]) == 5 * 11 + 1 + 30 and list(w["speaker_encoder"]) == ["speaker_encoder.x"]


def test_pack_layer_weights_blob_cpu():
    import struct
    from qwen_megakernel.model_tts import _pack_layer_weights
    ts = [torch.zeros(4, dtype=torch.bfloat16) for _ in range(22)]
    blob = _pack_layer_weights(ts, 2, device="cpu")
    assert blob.dtype == torch.uint8 and blob.numel() == 2 * 88
    ptrs = struct.unpack("22Q", bytes(blob.tolist()))
    assert list(ptrs) == [t.data_ptr() for t in ts]


def test_entry_points_reject_null_arguments_without_a_gpu():
    """Argument validation happens before any CUDA call: null handles / pointers give QMK_ERR_ARG (-1) on a CPU-only host."""
    from qwen_megakernel import build_tts
    lib = build_tts.load_library(build_tts.build())
    assert lib.qmk_model_set_mrope(None, None, 0) == -1
    assert lib.qmk_decode_step(None, 0, 0, None, None, None, None, None, None, None, None, 0, 1, 0.0, 0, None) == -1
    assert lib.qmk_decode_step_mrope(None, 0, 0, None, None, None, None, None, None, None, None, 0, None, 1, 0.0, None) == -1
    assert lib.qmk_cp_predict(None, None, 0, None, None, None, None, None, 64, 0, 1.0, 0, 0, 0, None, None, None, None, None) == -1
    assert lib.qmk_batched_step_ex(None, None, None) == -1
    assert lib.qmk_batched_prefill(None, None, 1, 0, None, None, None, None, None) == -1
    assert lib.qmk_batched_add_head(None, None, 128) == -1
    assert lib.qmk_batched_embed_sum(0, None, None, 0, None, 0, None, 0, None, None) == -1
    assert lib.qmk_batched_counter_add(None, 1, None) == -1
    assert lib.qmk_batched_chain_trace(0, None, None, 0) == 0          # never armed: nothing to copy, no CUDA call
    assert b"null" in lib.qmk_last_error() or b"null" in lib.qmk_batched_last_error()
    assert lib.qmk_legacy_check_blob(None, 5, None) == 0
    lib.qmk_legacy_invalidate(None)                     # no cached models: a no-op


def test_ctypes_mirrors_have_the_header_field_order():
    """GenerateArgs / BatchedStepArgs mirror the C structs field by field (names in header order)."""
    import re
    from qwen_megakernel import build_tts
    text = open(os.path.join(REPO, "include", "qmk_b200.h")).read()
    for cname, mirror in (("qmk_generate_args", build_tts.GenerateArgs), ("qmk_batched_step_args", build_tts.BatchedStepArgs)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), text, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        c_names = re.findall(r"(\w+)\s*(?:,|$)", " , ".join(re.sub(r"^[\w\s\*]*?([\w]+(?:\s*,\s*\w+)*)$", r"\1", d.strip().replace("*", " ")) for d in body.split(";") if d.strip()))
        assert [n for n, _ in mirror._fields_] == c_names, (cname, c_names)

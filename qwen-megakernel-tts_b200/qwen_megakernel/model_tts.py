"""Model layer of the B200 decode engine — same public surface as upstream qwen_megakernel/model_tts.py.

Replaced (new code, B200-native): ``load_tts_weights``, ``_pack_layer_weights``, ``TTSDecoder``,
``CodePredictorKernel``.  Kept for callers that import them from here (``tts_engine.py:19-34`` upstream):
``TextProjection``, ``CodePredictor``, ``build_prefill_embeddings`` — plain PyTorch helpers that sit
outside the accelerated path.

Differences a caller can observe (all additive):
  * constructors accept an optional ``device=`` keyword (upstream hard-wires ``"cuda"``,
    model_tts.py:227-247) so eight replicas can live on eight GPUs;
  * ``reset()`` is O(1): attention reads rows ``0..position`` only, so the 939 MB KV memset upstream
    performs per utterance (model_tts.py:332-336) is not needed;
  * argument validation: bad token ids / positions raise instead of writing out of bounds;
  * numerics follow the upstream *PyTorch* talker / code-predictor path (bf16 rounding points, fp32 vs
    bf16 residual), which is what parity is judged against — not upstream kernel.cu's rounding.
"""

from __future__ import annotations

import ctypes
import math
import os
import struct
from typing import Optional

import torch

# ─── Talker decoder constants (upstream model_tts.py:19-29) ─────────────────────────────────────────
NUM_LAYERS = 28
NUM_KV_HEADS = 8
NUM_Q_HEADS = 16
HEAD_DIM = 128
HIDDEN_SIZE = 1024
INTERMEDIATE_SIZE = 3072
Q_SIZE = NUM_Q_HEADS * HEAD_DIM
KV_SIZE = NUM_KV_HEADS * HEAD_DIM
VOCAB_SIZE = 3072
MAX_SEQ_LEN = 8192
ROPE_THETA = 1000000.0

# ─── Code predictor constants (upstream model_tts.py:31-34) ─────────────────────────────────────────
NUM_CODE_GROUPS = 16
CODE_PREDICTOR_LAYERS = 5
CODE_PREDICTOR_VOCAB = 2048

# ─── Special token ids (upstream model_tts.py:36-53) ────────────────────────────────────────────────
CODEC_BOS = 2149
CODEC_EOS = 2150
CODEC_PAD = 2148
CODEC_NOTHINK = 2155
CODEC_THINK_BOS = 2156
CODEC_THINK_EOS = 2157
TTS_BOS = 151672
TTS_EOS = 151673
TTS_PAD = 151671
EMBED_FROM_BUFFER = -1

_LAYER_FIELDS = (
    "input_layernorm.weight", "self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight",
    "self_attn.q_norm.weight", "self_attn.k_norm.weight", "self_attn.o_proj.weight",
    "post_attention_layernorm.weight", "mlp.gate_proj.weight", "mlp.up_proj.weight", "mlp.down_proj.weight",
)
_LAYER_SHAPES = (
    (HIDDEN_SIZE,), (Q_SIZE, HIDDEN_SIZE), (KV_SIZE, HIDDEN_SIZE), (KV_SIZE, HIDDEN_SIZE), (HEAD_DIM,), (HEAD_DIM,),
    (HIDDEN_SIZE, Q_SIZE), (HIDDEN_SIZE,), (INTERMEDIATE_SIZE, HIDDEN_SIZE), (INTERMEDIATE_SIZE, HIDDEN_SIZE),
    (HIDDEN_SIZE, INTERMEDIATE_SIZE),
)


def _rope_tables(max_seq: int, device) -> tuple[torch.Tensor, torch.Tensor]:
    inv_freq = 1.0 / (ROPE_THETA ** (torch.arange(0, HEAD_DIM, 2, dtype=torch.float32) / HEAD_DIM))
    freqs = torch.outer(torch.arange(max_seq, dtype=torch.float32), inv_freq)
    cos = torch.cos(freqs).repeat(1, 2).to(torch.bfloat16).to(device).contiguous()
    sin = torch.sin(freqs).repeat(1, 2).to(torch.bfloat16).to(device).contiguous()
    return cos, sin


def load_tts_weights(model_path: str = "Qwen/Qwen3-TTS-12Hz-0.6B-Base", device: str = "cuda",
                     verbose: bool = True) -> dict:
    """safetensors checkpoint -> the weights dict every other class consumes.

    Same keys, layouts and dtypes as upstream ``load_tts_weights`` (model_tts.py:56-179): ordinary torch
    tensors in the original ``[out, in]`` layout (kept code such as ``TextProjection`` and the frame loop
    of tts_engine.py index this dict directly); the engine re-packs its own copy at construction time.
    """
    if verbose:
        print(f"Loading TTS weights from {model_path}...")
    if os.path.isdir(model_path):
        path = os.path.join(model_path, "model.safetensors")
    else:
        from huggingface_hub import hf_hub_download
        path = hf_hub_download(model_path, "model.safetensors")
    from safetensors.torch import load_file
    state = load_file(path, device=device)

    def take(key):
        return state[key].contiguous()

    layer_weights = [take(f"talker.model.layers.{i}.{f}") for i in range(NUM_LAYERS) for f in _LAYER_FIELDS]
    cos_table, sin_table = _rope_tables(MAX_SEQ_LEN, device)
    cp = {}
    for i in range(CODE_PREDICTOR_LAYERS):
        for f in _LAYER_FIELDS:
            cp[f"layers.{i}.{f}"] = state[f"talker.code_predictor.model.layers.{i}.{f}"]
    cp["norm.weight"] = state["talker.code_predictor.model.norm.weight"]
    for g in range(NUM_CODE_GROUPS - 1):
        cp[f"lm_head.{g}.weight"] = state[f"talker.code_predictor.lm_head.{g}.weight"]
        cp[f"codec_embedding.{g}.weight"] = state[f"talker.code_predictor.model.codec_embedding.{g}.weight"]
    weights = dict(
        embed_weight=take("talker.model.codec_embedding.weight"),
        lm_head_weight=take("talker.codec_head.weight"),
        final_norm_weight=take("talker.model.norm.weight"),
        layer_weights=layer_weights,
        cos_table=cos_table,
        sin_table=sin_table,
        text_embedding=take("talker.model.text_embedding.weight"),
        text_proj_fc1_w=take("talker.text_projection.linear_fc1.weight"),
        text_proj_fc1_b=take("talker.text_projection.linear_fc1.bias"),
        text_proj_fc2_w=take("talker.text_projection.linear_fc2.weight"),
        text_proj_fc2_b=take("talker.text_projection.linear_fc2.bias"),
        code_predictor=cp,
        speaker_encoder={k: v for k, v in state.items() if k.startswith("speaker_encoder.")},
    )
    if verbose:
        print(f"Loaded {len(state)} tensors ({sum(v.numel() for v in state.values()) / 1e6:.1f}M params)")
    del state
    if torch.cuda.is_available():
        torch.cuda.empty_cache()
    return weights


def _check_layer_tensors(layer_weights, num_layers: int, device) -> None:
    if len(layer_weights) != 11 * num_layers:
        raise ValueError(f"expected {11 * num_layers} layer tensors, got {len(layer_weights)}")
    for i, t in enumerate(layer_weights):
        want = _LAYER_SHAPES[i % 11]
        if tuple(t.shape) != want or t.dtype != torch.bfloat16 or not t.is_contiguous():
            raise ValueError(f"layer tensor {i} ({_LAYER_FIELDS[i % 11]}): need contiguous bf16 {want}, "
                             f"got {t.dtype} {tuple(t.shape)}")
        if device is not None and t.device != device:
            raise ValueError(f"layer tensor {i} is on {t.device}, expected {device}")


def _pack_layer_weights(layer_weights: list[torch.Tensor], num_layers: int = NUM_LAYERS,
                        device=None) -> torch.Tensor:
    """11 data pointers per layer -> uint8[num_layers * 88] blob of ``LDGLayerWeights`` structs
    (include/qmk_b200.h; upstream model_tts.py:182-193).  The caller keeps the tensors alive."""
    buf = bytearray(num_layers * 88)
    for i in range(num_layers * 11):
        struct.pack_into("Q", buf, i * 8, layer_weights[i].data_ptr())
    blob = torch.frombuffer(buf, dtype=torch.uint8)
    dev = device if device is not None else (layer_weights[0].device if layer_weights[0].is_cuda else "cuda")
    return blob.to(dev)


def _require_cuda_bf16(t: torch.Tensor, n: int, what: str) -> None:
    if not (t.is_cuda and t.dtype == torch.bfloat16 and t.numel() == n and t.is_contiguous()):
        raise ValueError(f"{what}: need a contiguous CUDA bf16 tensor with {n} elements, got "
                         f"{t.dtype} {tuple(t.shape)} on {t.device}")


class _Native:
    """Process-wide handle on libqmk_b200.so plus one engine per CUDA device.

    Every ``TTSDecoder`` / ``CodePredictorKernel`` on a device shares that device's engine (exchange words, accumulator
    totals, epoch counter), so their launches execute in submission order: the C layer makes a launch on a different
    CUDA stream wait for the engine's previous stream (include/qmk_b200.h).  Launches from several host threads are
    serialised by the engine's mutex; they never overlap on the GPU."""

    _engines: dict = {}

    @classmethod
    def lib(cls):
        from .build_tts import get_extension
        return get_extension().lib

    @classmethod
    def engine(cls, device: torch.device):
        from .build_tts import check
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if idx not in cls._engines:
            lib = cls.lib()
            h = ctypes.c_void_p()
            check(lib, lib.qmk_engine_create(idx, 0, ctypes.byref(h)), "qmk_engine_create")
            cls._engines[idx] = h
        return cls._engines[idx]

    @classmethod
    def raise_kernel_status(cls, device: torch.device, what: str):
        from .build_tts import NativeError, check
        lib = cls.lib()
        detail = (ctypes.c_int32 * 4)()
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream().cuda_stream
            check(lib, lib.qmk_engine_sync_status(cls.engine(device), stream, detail), what)
        raise NativeError(f"{what}: kernel reported failure but the status word is clear")


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class TTSDecoder:
    """Stateful talker decoder (28 layers) on the fused sm_100a kernel.

    Same contract as upstream ``TTSDecoder`` (model_tts.py:196-345): ``step(token_id)`` /
    ``step_with_embed(embed_bf16)`` return ``(next_token:int, hidden:float32[1024])`` where ``hidden`` is the
    post-final-RMSNorm state the code predictor consumes; ``position`` is host state.
    """

    def __init__(self, weights: Optional[dict] = None, model_path: str = "Qwen/Qwen3-TTS-12Hz-0.6B-Base",
                 verbose: bool = True, *, device=None, max_seq_len: int = MAX_SEQ_LEN, mode: int = 0,
                 num_layers: Optional[int] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("TTSDecoder needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device(device if device is not None else "cuda")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        if weights is None:
            weights = load_tts_weights(model_path, device=str(dev), verbose=verbose)
        self.device = dev
        self._weights = weights
        self._position = 0
        self._mode = mode
        self._mrope_delta = None
        self._max_seq = int(max_seq_len)
        # num_layers is a runtime parameter of the kernel (upstream passes 28 or 5, model_tts.py:278,724)
        self._num_layers = int(num_layers) if num_layers is not None else len(weights["layer_weights"]) // 11
        self._embed_weight = weights["embed_weight"]
        self._final_norm_weight = weights["final_norm_weight"]
        self._lm_head_weight = weights["lm_head_weight"]
        self._cos_table = weights["cos_table"]
        self._sin_table = weights["sin_table"]
        if self._cos_table.shape[0] < self._max_seq:
            raise ValueError("RoPE tables are shorter than max_seq_len")
        _check_layer_tensors(weights["layer_weights"][:11 * self._num_layers], self._num_layers, dev)
        for name, t, shape in (("embed_weight", self._embed_weight, (VOCAB_SIZE, HIDDEN_SIZE)),
                               ("lm_head_weight", self._lm_head_weight, (VOCAB_SIZE, HIDDEN_SIZE)),
                               ("final_norm_weight", self._final_norm_weight, (HIDDEN_SIZE,))):
            if tuple(t.shape) != shape or t.dtype != torch.bfloat16 or t.device != dev:
                raise ValueError(f"{name}: need bf16 {shape} on {dev}")
        self._attn_scale = 1.0 / math.sqrt(HEAD_DIM)

        from .build_tts import check
        self._lib = _Native.lib()
        self._engine = _Native.engine(dev)
        with torch.cuda.device(dev):
            self._layer_weights_packed = _pack_layer_weights(weights["layer_weights"], self._num_layers, dev)
            h = ctypes.c_void_p()
            check(self._lib, self._lib.qmk_model_create(
                self._engine, self._layer_weights_packed.data_ptr(), self._num_layers, self._final_norm_weight.data_ptr(),
                1, _stream_ptr(dev), ctypes.byref(h)), "qmk_model_create")
            self._model = h
            self._head = check(self._lib, self._lib.qmk_model_add_head(
                self._model, self._lm_head_weight.data_ptr(), VOCAB_SIZE, _stream_ptr(dev)), "qmk_model_add_head")
            self._k_cache = torch.zeros(self._num_layers, NUM_KV_HEADS, self._max_seq, HEAD_DIM,
                                        dtype=torch.bfloat16, device=dev)
            self._v_cache = torch.zeros_like(self._k_cache)
            self._hidden = torch.zeros(HIDDEN_SIZE, dtype=torch.bfloat16, device=dev)
            self._norm_out = torch.zeros(HIDDEN_SIZE, dtype=torch.float32, device=dev)
            self._out_token = torch.zeros(1, dtype=torch.int32, device=dev)
        if verbose:
            mb = self._lib.qmk_model_packed_bytes(self._model) / 1e6
            print(f"TTSDecoder: {self._lib.qmk_engine_num_ctas(self._engine)} persistent CTAs, "
                  f"{mb:.0f} MB re-packed weights on {dev}")

    def __del__(self):
        try:
            if getattr(self, "_model", None):
                self._lib.qmk_model_destroy(self._model)
                self._model = None
        except Exception:
            pass

    def _launch(self, token_id: int, input_ptr: int, rope_pos=None) -> None:
        from .build_tts import check
        if self._position >= self._max_seq:
            raise IndexError(f"KV cache is full (position {self._position} == max_seq_len)")
        if rope_pos is None and self._mrope_delta is not None:
            rope_pos = tuple(self._position + d for d in self._mrope_delta)
        if rope_pos is not None:
            rp = (ctypes.c_int32 * 3)(*[int(v) for v in rope_pos])
            if min(rp) < 0 or max(rp) >= self._cos_table.shape[0]:
                raise ValueError(f"rope_pos {tuple(rp)} outside the RoPE tables")
            check(self._lib, self._lib.qmk_decode_step_mrope(
                self._model, self._head, token_id, self._embed_weight.data_ptr(), self._cos_table.data_ptr(),
                self._sin_table.data_ptr(), self._k_cache.data_ptr(), self._v_cache.data_ptr(), input_ptr,
                self._norm_out.data_ptr(), self._out_token.data_ptr(), self._position, rp, self._max_seq,
                self._attn_scale, _stream_ptr(self.device)), "qmk_decode_step_mrope")
        else:
            check(self._lib, self._lib.qmk_decode_step(
                self._model, self._head, token_id, self._embed_weight.data_ptr(), self._cos_table.data_ptr(),
                self._sin_table.data_ptr(), self._k_cache.data_ptr(), self._v_cache.data_ptr(), input_ptr,
                self._norm_out.data_ptr(), self._out_token.data_ptr(), self._position, self._max_seq,
                self._attn_scale, self._mode, _stream_ptr(self.device)), "qmk_decode_step")
        self._position += 1

    def set_mrope(self, section=(24, 20, 20), interleaved: bool = False, delta=(0, 0, 0)) -> None:
        """Switch the talker to multimodal RoPE (upstream's documented gap, README.md:208: the checkpoint is trained with
        ``mrope_section = [24, 20, 20]`` while upstream's kernel applies standard RoPE).  Rotary frequency ``i`` then takes
        the cos/sin row of the position of its axis (temporal / height / width).  ``delta``: per-axis offset added to the
        sequence position when ``step`` / ``step_with_embed`` are called without explicit ``rope_pos`` (text-only TTS
        uses the same position on all three axes, i.e. ``(0, 0, 0)``).  ``section=None`` restores standard RoPE."""
        from .build_tts import check
        if section is None:
            check(self._lib, self._lib.qmk_model_set_mrope(self._model, None, 0), "qmk_model_set_mrope")
            self._mrope_delta = None
            return
        sec = (ctypes.c_int32 * 3)(*[int(v) for v in section])
        check(self._lib, self._lib.qmk_model_set_mrope(self._model, sec, int(bool(interleaved))), "qmk_model_set_mrope")
        self._mrope_delta = tuple(int(v) for v in delta)

    def _finish(self) -> tuple[int, torch.Tensor]:
        hidden = self._norm_out.clone()
        token = int(self._out_token.item())          # the one host sync per step the upstream API implies
        if token < 0:
            _Native.raise_kernel_status(self.device, "TTSDecoder.step")
        return token, hidden

    def step(self, token_id: int, *, rope_pos=None) -> tuple[int, torch.Tensor]:
        """Decode one token via embedding lookup. Returns (next_token, hidden_state_f32).
        ``rope_pos``: optional (t, h, w) M-RoPE positions of this step (after ``set_mrope``)."""
        token_id = int(token_id)
        if not 0 <= token_id < VOCAB_SIZE:
            raise ValueError(f"token id {token_id} outside [0, {VOCAB_SIZE})")
        self._launch(token_id, self._hidden.data_ptr(), rope_pos)
        return self._finish()

    def step_with_embed(self, embed_bf16: torch.Tensor, *, rope_pos=None) -> tuple[int, torch.Tensor]:
        """Decode from a precomputed bf16[1024] embedding (upstream sentinel path, token_id = -1)."""
        if embed_bf16.dtype != torch.bfloat16:
            embed_bf16 = embed_bf16.to(torch.bfloat16)
        embed_bf16 = embed_bf16.reshape(-1)
        if embed_bf16.device != self.device:
            embed_bf16 = embed_bf16.to(self.device)
        _require_cuda_bf16(embed_bf16.contiguous(), HIDDEN_SIZE, "step_with_embed(embed_bf16)")
        self._hidden.copy_(embed_bf16)
        self._launch(EMBED_FROM_BUFFER, self._hidden.data_ptr(), rope_pos)
        return self._finish()

    def step_with_codes(self, codes: torch.Tensor, code_embeddings, extra_embed_bf16: torch.Tensor, *, sync: bool = True):
        """Decode from a frame's 16 codes: the embedding sum of the upstream frame loop (tts_engine.py:319-335,
        ``embed[codes[0]] + sum_g code_embeddings[g][codes[g+1]] + extra``, bf16 adds in that order) is evaluated
        inside the step's launch instead of by 32 torch kernels.  ``codes``: int64[16] on this device (what
        ``CodePredictorKernel.predict`` returns); ``code_embeddings``: the 15 ``codec_embedding.{g}.weight`` tables;
        ``extra_embed_bf16``: trailing-text or tts_pad embedding, bf16[1024].

        ``sync=False`` returns the device tensors ``(token int32[1], hidden f32[1024])`` that the next step overwrites,
        without the host round trip of ``.item()``: feed them straight to ``CodePredictorKernel.predict`` and read
        tokens / codes back asynchronously."""
        from .build_tts import check
        if codes.dtype != torch.int64 or codes.numel() != NUM_CODE_GROUPS or codes.device != self.device:
            raise ValueError("codes: need an int64[16] tensor on the decoder's device")
        if len(code_embeddings) != NUM_CODE_GROUPS - 1:
            raise ValueError("code_embeddings: need the 15 code-predictor embedding tables")
        for t in code_embeddings:
            if t.dtype != torch.bfloat16 or t.device != self.device or not t.is_contiguous() or t.shape[-1] != HIDDEN_SIZE:
                raise ValueError("code_embeddings: need contiguous bf16 [*, 1024] tables on the decoder's device")
        extra = extra_embed_bf16.to(self.device, torch.bfloat16).reshape(-1).contiguous()
        _require_cuda_bf16(extra, HIDDEN_SIZE, "step_with_codes(extra_embed_bf16)")
        if self._position >= self._max_seq:
            raise IndexError(f"KV cache is full (position {self._position} == max_seq_len)")
        tables = (ctypes.c_void_p * 15)(*[t.data_ptr() for t in code_embeddings])
        codes = codes.contiguous()
        check(self._lib, self._lib.qmk_decode_step_codes(
            self._model, self._head, codes.data_ptr(), self._embed_weight.data_ptr(), int(self._embed_weight.shape[0]),
            tables, int(min(t.shape[0] for t in code_embeddings)), extra.data_ptr(),
            self._cos_table.data_ptr(), self._sin_table.data_ptr(), self._k_cache.data_ptr(), self._v_cache.data_ptr(),
            self._hidden.data_ptr(), self._norm_out.data_ptr(), self._out_token.data_ptr(), self._position,
            self._max_seq, self._attn_scale, _stream_ptr(self.device)), "qmk_decode_step_codes")
        self._position += 1
        return self._finish() if sync else (self._out_token, self._norm_out)

    def generate_frames(self, code_predictor: "CodePredictorKernel", n_frames: int, trailing_text=None,
                        pad_embed_bf16: Optional[torch.Tensor] = None, *, do_sample: bool = True,
                        temperature: float = 0.9, top_k: int = 50, eos_token: int = CODEC_EOS, trailing_offset: int = 0,
                        host_visible: bool = False, sync: bool = True):
        """Up to ``n_frames`` codec frames of the upstream loop (tts_engine.py:301-335) WITHOUT the host in it: the
        persistent kernel itself iterates ``predict -> 16-way embedding sum + trailing text -> talker step``, checks EOS on
        the device and stops there (``qmk_generate_nosync``; upstream analogue: ``generate_nosync``).  Must follow a talker
        step (``step(CODEC_BOS)`` in the upstream flow): the loop starts from that step's token and hidden state.

        ``trailing_text``: bf16[T, 1024] embeddings added to frame ``trailing_offset + f`` while it lasts, afterwards
        ``pad_embed_bf16`` (the tts_pad embedding).  ``host_visible=True`` puts codes / tokens / progress word into pinned
        host memory that the kernel writes directly, so another host thread can consume frames while generation runs
        (``state[0]`` = frames finished).  ``sync=True`` waits and returns ``(codes[n_done, 16], tokens[n_done], n_done)``
        with ``position`` advanced by ``n_done``; ``sync=False`` returns ``(codes, tokens, state)`` buffers immediately --
        call ``finish_generate(state)`` before the next talker step."""
        from .build_tts import GenerateArgs, check
        cp = code_predictor
        n_frames = int(n_frames)
        if n_frames < 1:
            raise ValueError("n_frames must be >= 1")
        if cp.device != self.device:
            raise ValueError("talker and code predictor must live on the same device")
        if self._position + n_frames > self._max_seq:
            raise IndexError(f"{n_frames} frames from position {self._position} exceed max_seq_len {self._max_seq}")
        dev = self.device
        with torch.cuda.device(dev):
            if pad_embed_bf16 is None:
                pad_embed_bf16 = torch.zeros(HIDDEN_SIZE, dtype=torch.bfloat16, device=dev)
            pad = pad_embed_bf16.to(dev, torch.bfloat16).reshape(-1).contiguous()
            _require_cuda_bf16(pad, HIDDEN_SIZE, "generate_frames(pad_embed_bf16)")
            trail, n_trail = None, 0
            if trailing_text is not None and trailing_text.numel() > 0:
                trail = trailing_text.to(dev, torch.bfloat16).reshape(-1, HIDDEN_SIZE).contiguous()
                n_trail = trail.shape[0]
            if host_visible:
                codes = torch.zeros(n_frames, NUM_CODE_GROUPS, dtype=torch.int64).pin_memory()
                tokens = torch.zeros(n_frames, dtype=torch.int32).pin_memory()
                state = torch.zeros(4, dtype=torch.int32).pin_memory()
            else:
                codes = torch.zeros(n_frames, NUM_CODE_GROUPS, dtype=torch.int64, device=dev)
                tokens = torch.zeros(n_frames, dtype=torch.int32, device=dev)
                state = torch.zeros(4, dtype=torch.int32, device=dev)
            tables = (ctypes.c_void_p * 15)(*[t.data_ptr() for t in cp.codec_embeddings])
            rope = None
            if self._mrope_delta is not None:
                rope = (ctypes.c_int32 * 3)(*[self._position + d for d in self._mrope_delta])
            cp._frame_counter += 1
            a = GenerateArgs(
                talker=self._model, talker_head=self._head, talker_vocab=int(self._embed_weight.shape[0]),
                talker_embed_weight=self._embed_weight.data_ptr(), talker_cos=self._cos_table.data_ptr(),
                talker_sin=self._sin_table.data_ptr(), talker_k_cache=self._k_cache.data_ptr(),
                talker_v_cache=self._v_cache.data_ptr(), talker_max_seq=self._max_seq, position=self._position,
                rope_pos=rope, hidden_buffer=self._hidden.data_ptr(), talker_hidden=self._norm_out.data_ptr(),
                talker_token=self._out_token.data_ptr(), cp=cp._model, cp_cos=cp._cos_table.data_ptr(),
                cp_sin=cp._sin_table.data_ptr(), cp_k_cache=cp._k_cache.data_ptr(), cp_v_cache=cp._v_cache.data_ptr(),
                cp_max_seq=cp._max_seq, cp_vocab=CODE_PREDICTOR_VOCAB, group_embedding_tables=tables,
                n_frames=n_frames, eos_token=int(eos_token), trailing_text=trail.data_ptr() if trail is not None else None,
                n_trailing=n_trail, trailing_offset=int(trailing_offset), pad_embed=pad.data_ptr(),
                do_sample=int(bool(do_sample) and temperature > 0), top_k=int(top_k), temperature=float(temperature),
                reset_state=1, seed=torch.initial_seed() & 0xFFFFFFFFFFFFFFFF, frame_counter=cp._frame_counter,
                codes_out=codes.data_ptr(), tokens_out=tokens.data_ptr(), gen_state=state.data_ptr())
            check(self._lib, self._lib.qmk_generate_nosync(ctypes.byref(a), _stream_ptr(dev)), "qmk_generate_nosync")
            cp._frame_counter += n_frames
            cp._position = NUM_CODE_GROUPS
            self._gen_keep = (a, tables, rope, pad, trail, codes, tokens, state)   # buffers the queued launches still use
            self._gen_pending = n_frames
            if not sync:
                return codes, tokens, state
            n_done = self.finish_generate(state)
            return codes[:n_done], tokens[:n_done], n_done

    def finish_generate(self, state: torch.Tensor) -> int:
        """Wait for a ``generate_frames(sync=False)`` run; advances ``position`` by the frames it produced and returns
        their number (fewer than requested if the talker emitted EOS)."""
        with torch.cuda.device(self.device):
            torch.cuda.current_stream().synchronize()
        st = state.cpu().tolist()
        n_done = int(st[0])
        if n_done < 0 or n_done > getattr(self, "_gen_pending", 0):
            raise RuntimeError(f"generate_frames: inconsistent progress word {st}")
        if int(self._out_token.item()) < 0:
            _Native.raise_kernel_status(self.device, "TTSDecoder.generate_frames")
        self._position += n_done
        self._gen_pending = 0
        self._gen_keep = None
        return n_done

    def frame(self, code_predictor: "CodePredictorKernel", extra_embed_bf16: torch.Tensor, *, do_sample: bool = True,
              temperature: float = 0.9, top_k: int = 50, eos_token: int = CODEC_EOS):
        """ONE launch for one frame of the upstream loop (predict + embedding sum + talker step), with the device-side
        EOS flag: returns ``(codes int64[16] or None if the preceding token was EOS, next_token, hidden)``."""
        codes, tokens, n = self.generate_frames(code_predictor, 1, None, extra_embed_bf16, do_sample=do_sample,
                                                temperature=temperature, top_k=top_k, eos_token=eos_token)
        if n == 0:
            return None, int(eos_token), self._norm_out.clone()
        return codes[0], int(tokens[0].item()), self._norm_out.clone()

    def prefill(self, embeds_bf16: torch.Tensor) -> tuple[int, torch.Tensor]:
        """Feed n prefill embeddings (bf16[n, 1024], n <= 16) as ONE batched pass instead of n sequential
        ``step_with_embed`` calls (upstream tts_engine.py:281-282 loops; the prefill is 24.9 of its 50.5 ms time to first
        chunk, README.md:23): the projections run once for all positions on the tcgen05 tensor cores (every weight byte is
        read once instead of n times) with causal attention, writing the same KV rows (the batched launch chain with lane =
        position; ``QMK_PREFILL_PERSISTENT=1`` selects the persistent kernel's prefill mode instead).  Returns what the LAST sequential
        step would return: ``(token, hidden)``.  M-RoPE: with the same position on all three axes (``delta == (0, 0, 0)``, what
        text-only TTS uses) the rotation is bit-identical to standard RoPE and the pass applies as it is; different per-axis
        offsets need ``step_with_embed``."""
        if self._mrope_delta is not None and any(self._mrope_delta):
            raise NotImplementedError("prefill() needs the same position on all three M-RoPE axes (delta (0, 0, 0)); "
                                      "use step_with_embed for per-axis offsets")
        e = embeds_bf16.to(self.device, torch.bfloat16).reshape(-1, HIDDEN_SIZE)
        n = e.shape[0]
        if n < 1 or n > 16:
            raise ValueError("prefill(): between 1 and 16 positions")
        if self._position + n > self._max_seq:
            raise IndexError(f"KV cache is full (position {self._position} + {n} > max_seq_len)")
        if getattr(self, "_prefiller", None) is None:
            self._prefiller = BatchedTTSDecoder(self._weights, 16, device=self.device, max_seq_len=self._max_seq,
                                                num_layers=self._num_layers)
        with torch.cuda.device(self.device):
            self._prefiller.prefill_into(e, self._position, self._k_cache, self._v_cache, self._norm_out, self._out_token)
        self._position += n
        return self._finish()

    def reset(self):
        """New utterance.  O(1): rows beyond ``position`` are never read."""
        self._position = 0

    @property
    def position(self) -> int:
        return self._position

    @property
    def embed_weight(self) -> torch.Tensor:
        """Codec embedding table [3072, 1024] bf16."""
        return self._embed_weight


def _capture(device, fn):
    """Record the launches ``fn`` enqueues on the current stream into a CUDA graph (nothing executes during the capture)."""
    g = torch.cuda.CUDAGraph()
    with torch.cuda.device(device):
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                result = fn()
        torch.cuda.current_stream(device).wait_stream(side)
    return g, result


class BatchedTTSDecoder:
    """B concurrent utterance streams (B = 16, 32, 48 or 64) decoded together; no upstream counterpart.

    Every stream is numerically the B = 1 ``TTSDecoder`` step (same rounding points, own position, own KV cache);
    the projections run as tcgen05 / TMEM tensor-core GEMMs so each weight byte is read once per step for all
    streams.  ``step`` / ``step_with_embed`` are asynchronous: they return device tensors and never sync the host.
    """

    def __init__(self, weights: dict, batch: int, *, device=None, max_seq_len: int = 2048,
                 num_layers: Optional[int] = None, residual_fp32: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedTTSDecoder needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device(device if device is not None else "cuda")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device, self.batch, self._max_seq = dev, int(batch), int(max_seq_len)
        self._num_layers = int(num_layers) if num_layers is not None else len(weights["layer_weights"]) // 11
        self._weights = weights
        lw = weights["layer_weights"][:11 * self._num_layers]
        _check_layer_tensors(lw, self._num_layers, dev)
        if weights["cos_table"].shape[0] < self._max_seq:
            raise ValueError("RoPE tables are shorter than max_seq_len")
        from .build_tts import NativeError
        self._lib = _Native.lib()
        host_blob = (ctypes.c_void_p * (11 * self._num_layers))(*[t.data_ptr() for t in lw])
        h = ctypes.c_void_p()
        with torch.cuda.device(dev):
            rc = self._lib.qmk_batched_create(
                dev.index, host_blob, self._num_layers, weights["final_norm_weight"].data_ptr(),
                weights["lm_head_weight"].data_ptr(), weights["lm_head_weight"].shape[0], weights["embed_weight"].data_ptr(),
                weights["cos_table"].data_ptr(), weights["sin_table"].data_ptr(), int(residual_fp32), self.batch,
                self._max_seq, ctypes.byref(h))
            if rc < 0:
                raise NativeError(f"qmk_batched_create: {self._lib.qmk_batched_last_error().decode()} (code {rc})")
            self._handle = h
            B, L = self.batch, self._num_layers
            self._k_cache = torch.zeros(B, L, NUM_KV_HEADS, self._max_seq, HEAD_DIM, dtype=torch.bfloat16, device=dev)
            self._v_cache = torch.zeros_like(self._k_cache)
            self.positions = torch.zeros(B, dtype=torch.int32, device=dev)
            self._tokens = torch.zeros(B, dtype=torch.int32, device=dev)
            self._hidden = torch.zeros(B, HIDDEN_SIZE, dtype=torch.float32, device=dev)
        self._steps = 0

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                self._lib.qmk_batched_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def _run(self, token_ptr, embed_ptr):
        from .build_tts import NativeError
        if self._steps >= self._max_seq:
            raise IndexError("KV cache is full for at least one stream")
        if getattr(self, "_pd", None) is None:
            self._pd = self.persistent_decode         # fixed at create time (QMK_BATCHED_PERSISTENT)
        if self._pd:
            rc = self._lib.qmk_batched_step(self._handle, token_ptr, embed_ptr, self.positions.data_ptr(),
                                            self._k_cache.data_ptr(), self._v_cache.data_ptr(), self._hidden.data_ptr(),
                                            self._tokens.data_ptr(), _stream_ptr(self.device))
        else:   # the launch chain, with the host's bound on the positions (deep contexts are split over several CTAs)
            from .build_tts import BatchedStepArgs
            a = BatchedStepArgs(token_ids=token_ptr, embeds_bf16=embed_ptr, positions=self.positions.data_ptr(),
                                k_cache=self._k_cache.data_ptr(), v_cache=self._v_cache.data_ptr(), hidden_out=self._hidden.data_ptr(),
                                tokens_out=self._tokens.data_ptr(), head=0, group=-1, depth_hint=self._steps)
            rc = self._lib.qmk_batched_step_ex(self._handle, ctypes.byref(a), _stream_ptr(self.device))
        if rc < 0:
            raise NativeError(f"qmk_batched_step: {self._lib.qmk_batched_last_error().decode()} (code {rc})")
        self._steps += 1
        return self._tokens, self._hidden

    def step(self, token_ids: torch.Tensor):
        """token_ids: int[B] on this device.  Returns (next_tokens int32[B], hidden f32[B, 1024]) — views of internal
        buffers that the next step overwrites."""
        t = token_ids.to(self.device, torch.int32).contiguous()
        if t.numel() != self.batch:
            raise ValueError(f"token_ids must have {self.batch} elements")
        self._keep = t
        return self._run(t.data_ptr(), None)

    def step_with_embed(self, embeds_bf16: torch.Tensor):
        """embeds_bf16: bf16[B, 1024] (the upstream sentinel path, one precomputed embedding per stream)."""
        e = embeds_bf16.to(self.device, torch.bfloat16).contiguous()
        if tuple(e.shape) != (self.batch, HIDDEN_SIZE):
            raise ValueError(f"embeds must be [{self.batch}, {HIDDEN_SIZE}]")
        self._keep = e
        return self._run(None, e.data_ptr())

    def step_graph(self, token_ids: Optional[torch.Tensor] = None):
        """``step`` replayed from a CUDA graph: the 227 launches of the chain (programmatic-launch edges included) are
        captured once -- every per-step quantity (positions, tokens) lives in device memory -- and each later call is ONE
        graph launch.  ``token_ids`` None feeds the previous step's tokens back (no copy at all); a tensor is copied into the
        graph's input buffer first.  Same results as ``step``."""
        if self._steps >= self._max_seq:
            raise IndexError("KV cache is full for at least one stream")
        if not hasattr(self, "_g_tok"):
            self._g_tok = torch.zeros(self.batch, dtype=torch.int32, device=self.device)
            self._graphs, self._graph_warm = {}, False
        feedback = token_ids is None
        deep = (self._steps >= 256) + (self._steps >= 1024)   # (the captured chain switches to split attention at depth: qmk_batched_step_args.depth_hint)
        src = self._tokens if feedback else self._g_tok
        if not feedback:
            self._g_tok.copy_(token_ids.to(self.device, torch.int32).reshape(self.batch))
        if not self._graph_warm:                      # first call: plain launches (also the warm-up the capture needs)
            self._graph_warm = True
            return self.step(src)
        if (feedback, deep) not in self._graphs:
            self._graphs[(feedback, deep)], _ = _capture(self.device, lambda: self.step(src))
            self._steps -= 1                          # the capture only recorded the step
        self._graphs[(feedback, deep)].replay()
        self._steps += 1
        return self._tokens, self._hidden

    def reset(self):
        """New utterances on every stream (O(1): rows beyond a stream's position are never read)."""
        self.positions.zero_()
        self._steps = 0

    @property
    def persistent(self) -> bool:
        """True if the persistent tcgen05 step kernel (csrc/qmk_bstep.cuh) is available on this device (one-pass prefill)."""
        return bool(self._lib.qmk_batched_is_persistent(self._handle) & 1)

    @property
    def persistent_decode(self) -> bool:
        """True if decode steps run as ONE persistent launch (QMK_BATCHED_PERSISTENT=1) instead of the launch chain."""
        return bool(self._lib.qmk_batched_is_persistent(self._handle) & 2)

    def sync_status(self) -> None:
        """Synchronise the current stream and raise if a wait inside the persistent kernel timed out."""
        from .build_tts import NativeError
        rc = self._lib.qmk_batched_sync_status(self._handle, _stream_ptr(self.device))
        if rc < 0:
            raise NativeError(f"BatchedTTSDecoder: {self._lib.qmk_batched_last_error().decode()} (code {rc})")

    def prefill_into(self, embeds_bf16: torch.Tensor, position0: int, k_cache: torch.Tensor, v_cache: torch.Tensor,
                     hidden_out: torch.Tensor, token_out: torch.Tensor, graph: bool = True) -> None:
        """One batched pass over n <= batch consecutive positions of ONE utterance (causal), writing the KV rows into a
        B = 1 cache ``[L, 8, S, 128]`` and the last position's hidden state / token (``qmk_batched_prefill``)."""
        from .build_tts import NativeError
        e = embeds_bf16.to(self.device, torch.bfloat16).reshape(-1, HIDDEN_SIZE).contiguous()
        n = e.shape[0]
        if not 1 <= n <= self.batch:
            raise ValueError(f"prefill of {n} positions needs a batched decoder with batch >= {n}")
        if tuple(k_cache.shape) != (self._num_layers, NUM_KV_HEADS, self._max_seq, HEAD_DIM) or k_cache.shape != v_cache.shape:
            raise ValueError("prefill_into: cache must be [L, 8, max_seq_len, 128] with this decoder's max_seq_len")
        # The ~255 launches of the pass are replayed from a CUDA graph from the third call with the same (n, position, buffers) on:
        # an engine prefills every utterance at position 0 into the same cache.
        if not hasattr(self, "_pre_e"):
            self._pre_e = torch.zeros(self.batch, HIDDEN_SIZE, dtype=torch.bfloat16, device=self.device)
            self._pre_graphs, self._pre_seen = {}, {}
        self._pre_e[:n].copy_(e)
        key = (n, int(position0), k_cache.data_ptr(), v_cache.data_ptr(), hidden_out.data_ptr(), token_out.data_ptr())

        def launch():
            rc = self._lib.qmk_batched_prefill(self._handle, self._pre_e.data_ptr(), n, int(position0), k_cache.data_ptr(), v_cache.data_ptr(),
                                               hidden_out.data_ptr(), token_out.data_ptr(), _stream_ptr(self.device))
            if rc < 0:
                raise NativeError(f"qmk_batched_prefill: {self._lib.qmk_batched_last_error().decode()} (code {rc})")

        if len(self._pre_seen) > 64:
            self._pre_seen.clear()
        seen = self._pre_seen.get(key, 0)
        self._pre_seen[key] = seen + 1
        if seen < 2 or not graph or os.environ.get("QMK_PREFILL_PERSISTENT", "0") not in ("", "0"):   # (the persistent form copies from host memory)
            launch()
            return
        if key not in self._pre_graphs:
            if len(self._pre_graphs) >= 8:
                self._pre_graphs.clear()
            self._pre_graphs[key], _ = _capture(self.device, launch)
        self._pre_graphs[key].replay()


class BatchedCodePredictor:
    """The code predictor for B concurrent streams (BASELINE.json configs[3]-[4] in codec frames/s; no upstream counterpart):
    one frame = 16 five-layer batched steps (tcgen05 projections, every weight byte read once for all streams) + the 15 group
    heads with greedy or temperature / top-k / multinomial selection on the device.  Lane b is numerically the B = 1
    ``CodePredictorKernel`` (bf16 residual stream, same rounding points) and, when sampling, draws like a B = 1 engine seeded
    ``seed + b * 0x632BE59BD9B4E019``."""

    def __init__(self, weights: dict, batch: int, *, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedCodePredictor needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device(device if device is not None else "cuda")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        from .build_tts import NativeError
        self.device, self.batch = dev, int(batch)
        cp = weights["code_predictor"]
        self.num_groups = NUM_CODE_GROUPS - 1
        self._layer_weights = [cp[f"layers.{i}.{f}"].contiguous() for i in range(CODE_PREDICTOR_LAYERS) for f in _LAYER_FIELDS]
        _check_layer_tensors(self._layer_weights, CODE_PREDICTOR_LAYERS, dev)
        self._final_norm = cp["norm.weight"].contiguous()
        self.codec_embeddings = [cp[f"codec_embedding.{g}.weight"].contiguous() for g in range(self.num_groups)]
        self.lm_heads = [cp[f"lm_head.{g}.weight"].contiguous() for g in range(self.num_groups)]
        self._max_seq = 64
        self._cos, self._sin = _rope_tables(self._max_seq, dev)
        self._lib = _Native.lib()
        host_blob = (ctypes.c_void_p * (11 * CODE_PREDICTOR_LAYERS))(*[t.data_ptr() for t in self._layer_weights])
        h = ctypes.c_void_p()
        with torch.cuda.device(dev):
            rc = self._lib.qmk_batched_create(
                dev.index, host_blob, CODE_PREDICTOR_LAYERS, self._final_norm.data_ptr(), self.lm_heads[0].data_ptr(),
                CODE_PREDICTOR_VOCAB, self.codec_embeddings[0].data_ptr(), self._cos.data_ptr(), self._sin.data_ptr(), 0,
                self.batch, self._max_seq, ctypes.byref(h))
            if rc < 0:
                raise NativeError(f"qmk_batched_create: {self._lib.qmk_batched_last_error().decode()} (code {rc})")
            self._handle = h
            for g in range(1, self.num_groups):
                rc = self._lib.qmk_batched_add_head(self._handle, self.lm_heads[g].data_ptr(), CODE_PREDICTOR_VOCAB)
                if rc != g:
                    raise NativeError(f"qmk_batched_add_head: {self._lib.qmk_batched_last_error().decode()} (code {rc})")
            B = self.batch
            self._k_cache = torch.zeros(B, CODE_PREDICTOR_LAYERS, NUM_KV_HEADS, self._max_seq, HEAD_DIM, dtype=torch.bfloat16, device=dev)
            self._v_cache = torch.zeros_like(self._k_cache)
            self._positions = torch.zeros(B, dtype=torch.int32, device=dev)
            self._tokens = torch.zeros(B, dtype=torch.int32, device=dev)
            self._frame_counter = torch.zeros(1, dtype=torch.int64, device=dev)   # on the device: a captured frame draws fresh numbers per replay

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                self._lib.qmk_batched_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    def _step(self, **kw):
        from .build_tts import BatchedStepArgs, NativeError
        kw.setdefault("group", -1)
        kw.setdefault("head", -1)
        a = BatchedStepArgs(positions=self._positions.data_ptr(), k_cache=self._k_cache.data_ptr(), v_cache=self._v_cache.data_ptr(), **kw)
        rc = self._lib.qmk_batched_step_ex(self._handle, ctypes.byref(a), _stream_ptr(self.device))
        if rc < 0:
            raise NativeError(f"qmk_batched_step_ex: {self._lib.qmk_batched_last_error().decode()} (code {rc})")

    @torch.no_grad()
    def predict(self, talker_hidden: torch.Tensor, first_tokens: torch.Tensor, talker_embed_weight: torch.Tensor,
                do_sample: bool = True, temperature: float = 0.9, top_k: int = 50) -> torch.Tensor:
        """All 16 codebook groups of one frame for every stream: int64[B, 16] on the device = [first_token, g0..g14].
        ``talker_hidden``: f32[B, 1024]; ``first_tokens``: int32[B] (device; the talker's tokens).  Asynchronous."""
        B = self.batch
        hid = talker_hidden.to(self.device, torch.float32).reshape(B, HIDDEN_SIZE).contiguous()
        tok = first_tokens.to(self.device, torch.int32).reshape(B).contiguous()
        sample = bool(do_sample) and temperature > 0
        with torch.cuda.device(self.device):
            codes = torch.empty(B, NUM_CODE_GROUPS, dtype=torch.int64, device=self.device)
            codes[:, 0] = tok
            self._positions.zero_()
            rc = self._lib.qmk_batched_counter_add(self._frame_counter.data_ptr(), 1, _stream_ptr(self.device))   # frames count from 1
            if rc < 0:
                from .build_tts import NativeError
                raise NativeError(f"qmk_batched_counter_add: {self._lib.qmk_batched_last_error().decode()} (code {rc})")
            sel = dict(do_sample=int(sample), top_k=int(top_k), temperature=float(temperature),
                       seed=torch.initial_seed() & 0xFFFFFFFFFFFFFFFF, counter=0, counter_ptr=self._frame_counter.data_ptr(),
                       tokens_out=self._tokens.data_ptr(), codes_out=codes.data_ptr(), codes_stride=NUM_CODE_GROUPS)
            self._step(embeds_f32=hid.data_ptr())                                     # position 0: the talker's hidden state
            self._keep = (hid, tok, codes)
            for g in range(self.num_groups):                                          # position 1 + g: head g
                if g == 0:
                    src = dict(token_ids=tok.data_ptr(), token_table=talker_embed_weight.data_ptr(),
                               table_rows=int(talker_embed_weight.shape[0]))
                else:
                    src = dict(token_ids=self._tokens.data_ptr(), token_table=self.codec_embeddings[g - 1].data_ptr(),
                               table_rows=CODE_PREDICTOR_VOCAB)
                kw = dict(sel, **src)
                kw.update(head=g, group=g, codes_col=g + 1)
                self._step(**kw)
        return codes


class BatchedFrameLoop:
    """The upstream frame loop (tts_engine.py:301-335) for B concurrent utterances on one GPU: per frame one batched
    code-predictor frame, the 16-way embedding sum for every stream and one batched talker step -- all asynchronous on the
    device (tokens and hidden states never visit the host)."""

    def __init__(self, weights: dict, batch: int, *, device=None, max_seq_len: int = 2048, graph: bool = True):
        """``graph``: replay a frame (~915 launches) from a CUDA graph captured on the second call with the same sampling
        settings (the first call runs the plain launches); ``codes`` is then a buffer the next frame overwrites."""
        self.talker = BatchedTTSDecoder(weights, batch, device=device, max_seq_len=max_seq_len)
        self.cp = BatchedCodePredictor(weights, batch, device=self.talker.device)
        self.device, self.batch = self.talker.device, int(batch)
        self._embed = weights["embed_weight"]
        self._tables = (ctypes.c_void_p * 15)(*[t.data_ptr() for t in self.cp.codec_embeddings])
        self._e = torch.zeros(batch, HIDDEN_SIZE, dtype=torch.bfloat16, device=self.device)
        self._extra = torch.zeros(batch, HIDDEN_SIZE, dtype=torch.bfloat16, device=self.device)
        self._use_graph, self._graphs, self._warm = bool(graph), {}, set()
        self.tokens = self.hidden = None

    def start(self, prefill_bf16: torch.Tensor, bos_token: int = CODEC_BOS):
        """``prefill_bf16``: bf16[n, B, 1024] prefill embeddings (n steps for every stream), then step(bos)."""
        self.talker.reset()
        for i in range(prefill_bf16.shape[0]):
            self.talker.step_with_embed(prefill_bf16[i])
        tok = torch.full((self.batch,), int(bos_token), dtype=torch.int32, device=self.device)
        self.tokens, self.hidden = self.talker.step(tok)

    def frame(self, extra_bf16: torch.Tensor, do_sample: bool = True, temperature: float = 0.9, top_k: int = 50) -> torch.Tensor:
        """One codec frame for every stream; ``extra_bf16``: bf16[B, 1024] (per-stream trailing-text / pad embedding) or
        bf16[1024] (same for all).  Returns codes int64[B, 16] (device)."""
        if self._use_graph:
            if self.talker._steps >= self.talker._max_seq:
                raise IndexError("KV cache is full for at least one stream")
            self._extra.copy_(extra_bf16.to(self.device, torch.bfloat16).expand(self.batch, HIDDEN_SIZE))
            key = (bool(do_sample), float(temperature), int(top_k), (self.talker._steps >= 256) + (self.talker._steps >= 1024))
            if key not in self._warm:                 # first frame with these settings: plain launches
                self._warm.add(key)
                return self._frame(self._extra, do_sample, temperature, top_k)
            if key not in self._graphs:
                self._graphs[key] = _capture(self.device, lambda: self._frame(self._extra, do_sample, temperature, top_k))
                self.talker._steps -= 1               # the capture only recorded the frame
            g, codes = self._graphs[key]
            g.replay()
            self.talker._steps += 1
            self.tokens, self.hidden = self.talker._tokens, self.talker._hidden
            return codes
        return self._frame(extra_bf16, do_sample, temperature, top_k)

    def _frame(self, extra_bf16, do_sample, temperature, top_k):
        from .build_tts import NativeError
        codes = self.cp.predict(self.hidden, self.tokens, self._embed, do_sample, temperature, top_k)
        extra = extra_bf16.to(self.device, torch.bfloat16).contiguous()
        stride = HIDDEN_SIZE if extra.dim() == 2 else 0
        rc = self.talker._lib.qmk_batched_embed_sum(self.batch, codes.data_ptr(), self._embed.data_ptr(), int(self._embed.shape[0]),
                                                    self._tables, CODE_PREDICTOR_VOCAB, extra.data_ptr(), stride, self._e.data_ptr(),
                                                    _stream_ptr(self.device))
        if rc < 0:
            raise NativeError(f"qmk_batched_embed_sum: {self.talker._lib.qmk_batched_last_error().decode()} (code {rc})")
        self._keep = extra
        self.tokens, self.hidden = self.talker.step_with_embed(self._e)
        return codes


class TextProjection:
    """text ids -> talker hidden size: embedding(151936 -> 2048) -> fc1 + SiLU -> fc2 (-> 1024).

    Outside the accelerated path (once per utterance); same interface as upstream model_tts.py:348-374.
    """

    def __init__(self, weights: dict, device: str = "cuda"):
        self.text_embedding = weights["text_embedding"]
        self.fc1_w, self.fc1_b = weights["text_proj_fc1_w"], weights["text_proj_fc1_b"]
        self.fc2_w, self.fc2_b = weights["text_proj_fc2_w"], weights["text_proj_fc2_b"]

    @torch.no_grad()
    def embed_text_ids(self, token_ids: torch.Tensor) -> torch.Tensor:
        F = torch.nn.functional
        x = F.embedding(token_ids, self.text_embedding)
        return F.linear(F.silu(F.linear(x, self.fc1_w, self.fc1_b)), self.fc2_w, self.fc2_b)


class TextProjectionKernel:
    """``TextProjection`` on the native library (the text side of the prefill): same constructor and
    ``embed_text_ids`` contract as upstream ``TextProjection`` (model_tts.py:348-374), evaluated by
    ``qmk_text_proj_embed`` -- row gather, fc1 and fc2 on tcgen05 with the weights read in place, bias / SiLU / bf16
    rounding fused into the split-K epilogues, one chain of five launches per 512 tokens (csrc/qmk_text.cuh).

    Relates to ``TextProjection`` as upstream's ``CodePredictorKernel`` relates to its ``CodePredictor``: the PyTorch
    class stays the reference, this one is the accelerated path and has no fallback (CUDA tensors only).
    ``tts_engine.py:87`` constructs ``TextProjection(weights, device=...)``; the drop-in is that one name.
    Ids outside the table are clamped on the device (upstream's ``F.embedding`` would trap).
    """

    def __init__(self, weights: dict, device: str = "cuda"):
        if not torch.cuda.is_available():
            raise RuntimeError("TextProjectionKernel needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device(device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.text_embedding = weights["text_embedding"]
        self.fc1_w, self.fc1_b = weights["text_proj_fc1_w"], weights["text_proj_fc1_b"]
        self.fc2_w, self.fc2_b = weights["text_proj_fc2_w"], weights["text_proj_fc2_b"]
        for name, t, shape in (("text_embedding", self.text_embedding, (self.text_embedding.shape[0], 2048)),
                               ("text_proj_fc1_w", self.fc1_w, (2048, 2048)), ("text_proj_fc1_b", self.fc1_b, (2048,)),
                               ("text_proj_fc2_w", self.fc2_w, (HIDDEN_SIZE, 2048)), ("text_proj_fc2_b", self.fc2_b, (HIDDEN_SIZE,))):
            if not t.is_cuda or t.device != dev or t.dtype != torch.bfloat16 or not t.is_contiguous() or tuple(t.shape) != shape:
                raise ValueError(f"TextProjectionKernel: {name} must be a contiguous bf16 tensor of shape {shape} on {dev}")
        self._lib = _Native.lib()
        h = ctypes.c_void_p()
        rc = self._lib.qmk_text_proj_create(dev.index, self.text_embedding.data_ptr(), int(self.text_embedding.shape[0]),
                                            self.fc1_w.data_ptr(), self.fc1_b.data_ptr(), self.fc2_w.data_ptr(),
                                            self.fc2_b.data_ptr(), ctypes.byref(h))
        if rc < 0:
            from .build_tts import NativeError
            raise NativeError(f"qmk_text_proj_create: {self._lib.qmk_batched_last_error().decode()} (code {rc})")
        self._handle = h

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                self._lib.qmk_text_proj_destroy(self._handle)
                self._handle = None
        except Exception:
            pass

    @torch.no_grad()
    def embed_text_ids(self, token_ids: torch.Tensor) -> torch.Tensor:
        """``[seq_len]`` or ``[batch, seq_len]`` integer ids -> ``[*, HIDDEN_SIZE]`` bf16 (asynchronous on the current stream)."""
        if token_ids.dtype not in (torch.int64, torch.int32, torch.int16, torch.uint8):
            raise ValueError("TextProjectionKernel.embed_text_ids: token_ids must be an integer tensor")
        ids = token_ids.to(device=self.device, dtype=torch.int64).contiguous()
        out = torch.empty(*ids.shape, HIDDEN_SIZE, dtype=torch.bfloat16, device=self.device)
        rc = self._lib.qmk_text_proj_embed(self._handle, ids.data_ptr(), ids.numel(), out.data_ptr(), _stream_ptr(self.device))
        if rc < 0:
            from .build_tts import NativeError
            raise NativeError(f"qmk_text_proj_embed: {self._lib.qmk_batched_last_error().decode()} (code {rc})")
        return out


class CodePredictorKernel:
    """Code predictor (5 layers, 15 group heads) on the same kernel as the talker (``num_layers=5``).

    Same contract as upstream ``CodePredictorKernel`` (model_tts.py:622-773).  Group LM heads are registered
    with the engine, so greedy prediction never leaves the kernel; the sampling path applies the upstream
    temperature / top-k / multinomial rule on the kernel's hidden state.
    """

    def __init__(self, weights: dict, device: str = "cuda", *, mode: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("CodePredictorKernel needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device(device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self._mode = mode
        cp = weights["code_predictor"]
        self.num_groups = NUM_CODE_GROUPS - 1
        self._layer_weights = [cp[f"layers.{i}.{f}"].contiguous() for i in range(CODE_PREDICTOR_LAYERS)
                               for f in _LAYER_FIELDS]
        _check_layer_tensors(self._layer_weights, CODE_PREDICTOR_LAYERS, dev)
        self._final_norm_weight = cp["norm.weight"].contiguous()
        self.codec_embeddings = [cp[f"codec_embedding.{g}.weight"].contiguous() for g in range(self.num_groups)]
        self.lm_heads = [cp[f"lm_head.{g}.weight"].contiguous() for g in range(self.num_groups)]
        self._max_seq = 64
        self._cos_table, self._sin_table = _rope_tables(self._max_seq, dev)
        self._attn_scale = 1.0 / math.sqrt(HEAD_DIM)
        self._position = 0

        from .build_tts import check
        self._lib = _Native.lib()
        self._engine = _Native.engine(dev)
        with torch.cuda.device(dev):
            self._layer_weights_packed = _pack_layer_weights(self._layer_weights, CODE_PREDICTOR_LAYERS, dev)
            h = ctypes.c_void_p()
            check(self._lib, self._lib.qmk_model_create(
                self._engine, self._layer_weights_packed.data_ptr(), CODE_PREDICTOR_LAYERS,
                self._final_norm_weight.data_ptr(), 0, _stream_ptr(dev), ctypes.byref(h)), "qmk_model_create")
            self._model = h
            self._heads = []
            for g in range(self.num_groups):
                self._heads.append(check(self._lib, self._lib.qmk_model_add_head(
                    self._model, self.lm_heads[g].data_ptr(), CODE_PREDICTOR_VOCAB, _stream_ptr(dev)),
                    "qmk_model_add_head"))
                check(self._lib, self._lib.qmk_model_set_group_embedding(
                    self._model, g, self.codec_embeddings[g].data_ptr()), "qmk_model_set_group_embedding")
            self._k_cache = torch.zeros(CODE_PREDICTOR_LAYERS, NUM_KV_HEADS, self._max_seq, HEAD_DIM,
                                        dtype=torch.bfloat16, device=dev)
            self._v_cache = torch.zeros_like(self._k_cache)
            self._hidden = torch.zeros(HIDDEN_SIZE, dtype=torch.bfloat16, device=dev)
            self._norm_out = torch.zeros(HIDDEN_SIZE, dtype=torch.float32, device=dev)
            self._out_token = torch.zeros(1, dtype=torch.int32, device=dev)
            self._token_buf = torch.zeros(1, dtype=torch.long, device=dev)
        self._frame_counter = 0

    def __del__(self):
        try:
            if getattr(self, "_model", None):
                self._lib.qmk_model_destroy(self._model)
                self._model = None
        except Exception:
            pass

    def reset(self):
        self._position = 0

    def _step_with_embed(self, embed_bf16: torch.Tensor, head: int = -1):
        """One 5-layer decode step; afterwards ``_norm_out`` holds the hidden state and, if ``head`` >= 0,
        ``_out_token`` the argmax of that group head."""
        from .build_tts import check
        if self._position >= self._max_seq:
            raise IndexError("code-predictor KV cache is full")
        self._hidden.copy_(embed_bf16.reshape(-1))
        check(self._lib, self._lib.qmk_decode_step(
            self._model, head, EMBED_FROM_BUFFER, None, self._cos_table.data_ptr(), self._sin_table.data_ptr(),
            self._k_cache.data_ptr(), self._v_cache.data_ptr(), self._hidden.data_ptr(),
            self._norm_out.data_ptr(), self._out_token.data_ptr(), self._position, self._max_seq,
            self._attn_scale, self._mode, _stream_ptr(self.device)), "qmk_decode_step")
        self._position += 1

    @torch.no_grad()
    def predict(self, talker_hidden: torch.Tensor, first_codebook_token: int, talker_embed_weight: torch.Tensor,
                do_sample: bool = True, temperature: float = 0.9, top_k: int = 50, *,
                forced_tokens: Optional[torch.Tensor] = None, return_debug: bool = False):
        """All 16 codebook groups of one frame: int64[16] on device = [first_token, g0..g14].

        ONE kernel launch (``qmk_cp_predict``): 16 five-layer steps, the 15 group heads, greedy or temperature /
        top-k / multinomial selection and the embedding gather of the next step all stay on the device
        (upstream runs 16 launches plus ~10 torch ops per group, model_tts.py:742-773).  Sampling draws come from a
        counter-based generator keyed by ``torch.initial_seed()`` and a per-instance frame counter.

        ``forced_tokens`` (int32[15], device): teacher forcing -- the token fed to step g+1 (the returned codes are
        still the model's own choices).  ``return_debug``: also return (logits f32[15, 2048], hidden f32[15, 1024]).
        """
        from .build_tts import check
        token_dev = None
        if isinstance(first_codebook_token, torch.Tensor):      # device token of the preceding talker step: no host sync
            token_dev = first_codebook_token
            if token_dev.dtype != torch.int32 or token_dev.numel() != 1 or token_dev.device != self.device:
                raise ValueError("first_codebook_token tensor: need int32[1] on the predictor's device")
            if forced_tokens is not None or return_debug:
                raise ValueError("forced_tokens / return_debug need a host token")
        else:
            first_codebook_token = int(first_codebook_token)
            if not 0 <= first_codebook_token < talker_embed_weight.shape[0]:
                raise ValueError(f"first_codebook_token {first_codebook_token} out of range")
        if talker_embed_weight.dtype != torch.bfloat16 or talker_embed_weight.device != self.device or \
                talker_embed_weight.shape[1] != HIDDEN_SIZE or not talker_embed_weight.is_contiguous():
            raise ValueError("talker_embed_weight: need a contiguous bf16 [*, 1024] tensor on the predictor's device")
        sample = bool(do_sample) and temperature > 0
        with torch.cuda.device(self.device):
            hid = talker_hidden.to(self.device, torch.float32).reshape(-1).contiguous()
            if hid.numel() != HIDDEN_SIZE:
                raise ValueError("talker_hidden must have 1024 elements")
            out = torch.empty(NUM_CODE_GROUPS, dtype=torch.int64, device=self.device)
            logits = hidden = None
            if return_debug:
                logits = torch.empty(self.num_groups, CODE_PREDICTOR_VOCAB, dtype=torch.float32, device=self.device)
                hidden = torch.empty(self.num_groups, HIDDEN_SIZE, dtype=torch.float32, device=self.device)
            forced_ptr = None
            if forced_tokens is not None:
                forced_tokens = forced_tokens.to(self.device, torch.int32).contiguous()
                if forced_tokens.numel() != self.num_groups:
                    raise ValueError("forced_tokens must have 15 elements")
                forced_ptr = forced_tokens.data_ptr()
            self._frame_counter += 1
            if token_dev is not None:
                check(self._lib, self._lib.qmk_cp_predict_dev(
                    self._model, hid.data_ptr(), token_dev.data_ptr(), int(talker_embed_weight.shape[0]),
                    talker_embed_weight.data_ptr(), self._cos_table.data_ptr(), self._sin_table.data_ptr(),
                    self._k_cache.data_ptr(), self._v_cache.data_ptr(), self._max_seq, int(sample), float(temperature),
                    int(top_k), torch.initial_seed() & 0xFFFFFFFFFFFFFFFF, self._frame_counter, out.data_ptr(),
                    _stream_ptr(self.device)), "qmk_cp_predict_dev")
                self._position = NUM_CODE_GROUPS
                return out
            check(self._lib, self._lib.qmk_cp_predict(
                self._model, hid.data_ptr(), first_codebook_token, talker_embed_weight.data_ptr(),
                self._cos_table.data_ptr(), self._sin_table.data_ptr(), self._k_cache.data_ptr(),
                self._v_cache.data_ptr(), self._max_seq, int(sample), float(temperature), int(top_k),
                torch.initial_seed() & 0xFFFFFFFFFFFFFFFF, self._frame_counter, forced_ptr, out.data_ptr(),
                logits.data_ptr() if logits is not None else None,
                hidden.data_ptr() if hidden is not None else None, _stream_ptr(self.device)), "qmk_cp_predict")
            self._position = NUM_CODE_GROUPS
        return (out, logits, hidden) if return_debug else out

    @torch.no_grad()
    def predict_stepwise(self, talker_hidden: torch.Tensor, first_codebook_token: int,
                         talker_embed_weight: torch.Tensor, do_sample: bool = True, temperature: float = 0.9,
                         top_k: int = 50) -> torch.Tensor:
        """The upstream control flow (model_tts.py:742-773), one launch per step with torch glue in between."""
        F = torch.nn.functional
        first_codebook_token = int(first_codebook_token)
        sample = bool(do_sample) and temperature > 0
        with torch.cuda.device(self.device):
            self.reset()
            self._step_with_embed(talker_hidden.to(self.device).to(torch.bfloat16))
            self._token_buf[0] = first_codebook_token
            out = [self._token_buf.clone()]
            embed = F.embedding(self._token_buf, talker_embed_weight).squeeze(0)
            for g in range(self.num_groups):
                self._step_with_embed(embed, head=-1 if sample else self._heads[g])
                if sample:
                    logits = F.linear(self._norm_out.to(torch.bfloat16).unsqueeze(0), self.lm_heads[g]).squeeze(0)
                    z = logits.float() / temperature
                    if top_k > 0:
                        kth = torch.topk(z, min(top_k, z.numel())).values[-1]
                        z = z.masked_fill(z < kth, float("-inf"))
                    tok = torch.multinomial(F.softmax(z, dim=-1), 1)
                else:
                    tok = self._out_token.long()
                out.append(tok)
                if g < self.num_groups - 1:
                    embed = F.embedding(tok, self.codec_embeddings[g]).squeeze(0)
            return torch.cat(out)


class CodePredictor:
    """Pure-PyTorch code predictor with the upstream interface (model_tts.py:377-504).

    Upstream keeps this class as the slow reference implementation; it is not on the accelerated path.
    It is provided so that ``from .model_tts import CodePredictor`` keeps working; it evaluates the same
    decode-step arithmetic with ordinary torch ops on whatever device the weights live on.
    """

    def __init__(self, weights: dict, device: str = "cuda"):
        self.device = device
        cp = weights["code_predictor"]
        self.num_groups = NUM_CODE_GROUPS - 1
        self.codec_embeddings = [cp[f"codec_embedding.{g}.weight"] for g in range(self.num_groups)]
        self.lm_heads = [cp[f"lm_head.{g}.weight"] for g in range(self.num_groups)]
        self.layers = [[cp[f"layers.{i}.{f}"] for f in _LAYER_FIELDS] for i in range(CODE_PREDICTOR_LAYERS)]
        self.final_norm = cp["norm.weight"]
        self._max_seq = 20
        self._cos, self._sin = _rope_tables(64, device)
        self._k = torch.zeros(CODE_PREDICTOR_LAYERS, NUM_KV_HEADS, self._max_seq, HEAD_DIM, dtype=torch.bfloat16,
                              device=device)
        self._v = torch.zeros_like(self._k)

    @staticmethod
    def _norm(x, w, eps=1e-6):
        xf = x.float()
        return (xf / torch.sqrt(xf.pow(2).mean(-1, keepdim=True) + eps) * w.float()).to(x.dtype)

    def _rope(self, t, pos):
        half = HEAD_DIM // 2
        c, s = self._cos[pos, :half], self._sin[pos, :half]
        t1, t2 = t[..., :half], t[..., half:]
        return torch.cat([t1 * c - t2 * s, t2 * c + t1 * s], dim=-1)

    def _token_step(self, x, pos):
        F = torch.nn.functional
        h = x.to(torch.bfloat16)
        for li, (w_in, wq, wk, wv, w_qn, w_kn, wo, w_post, wg, wu, wd) in enumerate(self.layers):
            n = self._norm(h, w_in)
            q = self._rope(self._norm(F.linear(n, wq).view(NUM_Q_HEADS, HEAD_DIM), w_qn), pos)
            k = self._rope(self._norm(F.linear(n, wk).view(NUM_KV_HEADS, HEAD_DIM), w_kn), pos)
            self._k[li, :, pos] = k
            self._v[li, :, pos] = F.linear(n, wv).view(NUM_KV_HEADS, HEAD_DIM)
            kf = self._k[li, :, :pos + 1].float().repeat_interleave(2, dim=0)
            vf = self._v[li, :, :pos + 1].float().repeat_interleave(2, dim=0)
            p = torch.softmax(torch.einsum("hd,hsd->hs", q.float(), kf) / math.sqrt(HEAD_DIM), dim=-1)
            a = torch.einsum("hs,hsd->hd", p, vf).to(torch.bfloat16).reshape(-1)
            h = h + F.linear(a, wo)
            n2 = self._norm(h, w_post)
            h = h + F.linear(F.silu(F.linear(n2, wg)) * F.linear(n2, wu), wd)
        return self._norm(h, self.final_norm)

    @torch.no_grad()
    def predict(self, talker_hidden, first_codebook_token, talker_embed_weight, do_sample=True, temperature=0.9,
                top_k=50):
        F = torch.nn.functional
        self._token_step(talker_hidden, 0)
        hn = self._token_step(talker_embed_weight[int(first_codebook_token)], 1)
        out = [int(first_codebook_token)]
        for g in range(self.num_groups):
            z = F.linear(hn, self.lm_heads[g]).float()
            if do_sample and temperature > 0:
                z = z / temperature
                if top_k > 0:
                    z = z.masked_fill(z < torch.topk(z, min(top_k, z.numel())).values[-1], float("-inf"))
                tok = int(torch.multinomial(F.softmax(z, dim=-1), 1))
            else:
                tok = int(z.argmax())
            out.append(tok)
            if g < self.num_groups - 1:
                hn = self._token_step(self.codec_embeddings[g][tok], 2 + g)
        return torch.tensor(out, dtype=torch.int64, device=self.device)


def build_prefill_embeddings(text_token_ids: torch.Tensor, text_projection: TextProjection,
                             codec_embed_weight: torch.Tensor, language: str = "Auto", device: str = "cuda",
                             cached_tts_embeds: Optional[dict] = None) -> tuple[torch.Tensor, torch.Tensor]:
    """Prefill sequence of the talker (upstream model_tts.py:776-864), outside the accelerated path.

    Returns ``(prefill[8, 1024], trailing_text[T, 1024])``: 3 role tokens, 4 codec tags fused with
    tts_pad/tts_bos, first text token fused with codec_bos; the remaining text (minus the 5 closing
    format tokens) plus tts_eos is fed one embedding per decode step.
    """
    F = torch.nn.functional
    ids = text_token_ids.to(device)
    if cached_tts_embeds is None:
        special = torch.tensor([TTS_PAD, TTS_BOS, TTS_EOS], device=device)
        emb = text_projection.embed_text_ids(torch.cat([ids, special]))
        text, (pad, bos, eos) = emb[:-3], (emb[-3:-2], emb[-2:-1], emb[-1:])
    else:
        text = text_projection.embed_text_ids(ids)
        pad, bos, eos = cached_tts_embeds["pad"], cached_tts_embeds["bos"], cached_tts_embeds["eos"]
    role, content = text[:3], text[3:]
    tags = F.embedding(torch.tensor([CODEC_NOTHINK, CODEC_THINK_BOS, CODEC_THINK_EOS, CODEC_PAD, CODEC_BOS],
                                    device=device), codec_embed_weight)
    fused = torch.cat([pad.expand(3, -1), bos], dim=0) + tags[:4]
    prefill = torch.cat([role, fused, content[:1] + tags[4:5]], dim=0)
    trailing = torch.cat([content[1:-5], eos], dim=0)
    return prefill, trailing

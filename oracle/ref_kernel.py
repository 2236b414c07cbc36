"""TEST / MEASUREMENT INFRASTRUCTURE -- never imported by the product.

ctypes driver for the UPSTREAM decode kernel built for sm_100a by oracle/Makefile (oracle/_ref/libref_kernel_sm100a.so, compiled
from /root/reference/csrc/kernel.cu where it lies; nothing of it is copied into this repository).  It calls upstream's own C
entry point ``launch_ldg_decode_direct`` (kernel.cu:1485-1513) with the buffers upstream's ``TTSDecoder`` allocates
(model_tts.py:208-252), so that the unmodified upstream kernel can be timed on a B200 next to this repository's kernel and its
greedy tokens compared with it.  Parity is NOT judged against this kernel (it keeps fp32 where the PyTorch path rounds to bf16,
SURVEY.md section 8c); it is a GPU-side reference point for the speed of the path.
"""

from __future__ import annotations

import ctypes
import math
import os
import struct

import torch

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_kernel_sm100a.so")


def available() -> bool:
    return os.path.exists(LIB)


class UpstreamKernelDecoder:
    """Upstream ``TTSDecoder`` / ``CodePredictorKernel._step_with_embed`` buffer set + launch (model_tts.py:208-330, 622-727)."""

    def __init__(self, layer_tensors, final_norm, lm_head, embed, cos, sin, num_layers: int, max_seq: int, device="cuda"):
        self.lib = ctypes.CDLL(LIB)
        vp, i32 = ctypes.c_void_p, ctypes.c_int
        self.lib.launch_ldg_decode_direct.restype = None
        self.lib.launch_ldg_decode_direct.argtypes = [i32] + [vp] * 20 + [i32, i32, i32, ctypes.c_float, vp]
        self.L, self.S, self.dev = num_layers, max_seq, device
        self._keep = layer_tensors
        buf = bytearray(num_layers * 88)
        for i in range(num_layers * 11):
            struct.pack_into("Q", buf, i * 8, layer_tensors[i].data_ptr())
        self.blob = torch.frombuffer(buf, dtype=torch.uint8).to(device)
        self.final_norm, self.lm_head, self.embed, self.cos, self.sin = final_norm, lm_head, embed, cos, sin
        f32 = dict(dtype=torch.float32, device=device)
        self.k_cache = torch.zeros(num_layers, 8, max_seq, 128, dtype=torch.bfloat16, device=device)
        self.v_cache = torch.zeros_like(self.k_cache)
        self.hidden = torch.zeros(1024, dtype=torch.bfloat16, device=device)
        self.act, self.res, self.norm_out = torch.zeros(1024, **f32), torch.zeros(1024, **f32), torch.zeros(1024, **f32)
        self.q, self.k, self.v = torch.zeros(2048, **f32), torch.zeros(1024, **f32), torch.zeros(1024, **f32)
        self.attn_out, self.mlp = torch.zeros(2048, **f32), torch.zeros(3072, **f32)
        self.bmv, self.bmi = torch.zeros(4096, **f32), torch.zeros(4096, dtype=torch.int32, device=device)
        self.out_token = torch.zeros(1, dtype=torch.int32, device=device)
        self.position = 0
        self.scale = 1.0 / math.sqrt(128)

    def launch(self, token_id: int) -> None:
        """token_id >= 0: embedding row; -1: ``self.hidden`` (the sentinel path).  Asynchronous on the current stream."""
        self.lib.launch_ldg_decode_direct(
            int(token_id), self.out_token.data_ptr(), self.embed.data_ptr(), self.blob.data_ptr(), self.final_norm.data_ptr(),
            self.lm_head.data_ptr(), self.cos.data_ptr(), self.sin.data_ptr(), self.k_cache.data_ptr(), self.v_cache.data_ptr(),
            self.hidden.data_ptr(), self.act.data_ptr(), self.res.data_ptr(), self.q.data_ptr(), self.k.data_ptr(), self.v.data_ptr(),
            self.attn_out.data_ptr(), self.mlp.data_ptr(), self.norm_out.data_ptr(), self.bmv.data_ptr(), self.bmi.data_ptr(),
            self.L, self.position, self.S, ctypes.c_float(self.scale), torch.cuda.current_stream().cuda_stream)
        self.position += 1

    def step_with_embed(self, e):
        self.hidden.copy_(e)
        self.launch(-1)
        return int(self.out_token.item()), self.norm_out.clone()

    def step(self, token_id: int):
        self.launch(token_id)
        return int(self.out_token.item()), self.norm_out.clone()


def time_upstream_kernel(weights_gpu, prefill, n: int = 50, warmup: int = 10):
    """Mean launch time (us) of the upstream kernel in its talker configuration (28 layers, positions 18 .. 18 + n, the same
    window bench.py times its own kernel on) and in its code-predictor configuration (5 layers, 16 steps, dummy LM head)."""
    w = weights_gpu
    dev = w["embed_weight"].device
    dec = UpstreamKernelDecoder(w["layer_weights"], w["final_norm_weight"], w["lm_head_weight"], w["embed_weight"], w["cos_table"],
                                w["sin_table"], 28, 2048, dev)
    for i in range(prefill.shape[0]):
        dec.step_with_embed(prefill[i])
    dec.hidden.copy_(prefill[0])
    for _ in range(warmup):
        dec.launch(-1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        dec.launch(-1)
    b.record()
    torch.cuda.synchronize()
    talker_us = a.elapsed_time(b) / n * 1e3
    from qwen_megakernel.model_tts import _LAYER_FIELDS
    cp = w["code_predictor"]
    lw = [cp[f"layers.{i}.{f}"].contiguous() for i in range(5) for f in _LAYER_FIELDS]
    zeros = torch.zeros(3072, 1024, dtype=torch.bfloat16, device=dev)
    cpd = UpstreamKernelDecoder(lw, cp["norm.weight"], zeros, zeros, w["cos_table"][:64].contiguous(), w["sin_table"][:64].contiguous(), 5, 64, dev)
    cpd.hidden.copy_(prefill[0])
    for _ in range(3):
        cpd.position = 0
        for _ in range(16):
            cpd.launch(-1)
    torch.cuda.synchronize()
    a.record()
    reps = 10
    for _ in range(reps):
        cpd.position = 0
        for _ in range(16):
            cpd.launch(-1)
    b.record()
    torch.cuda.synchronize()
    cp_step_us = a.elapsed_time(b) / (reps * 16) * 1e3
    return talker_us, cp_step_us

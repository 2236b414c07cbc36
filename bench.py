#!/usr/bin/env python3
"""Benchmark of the accelerated path: B=1 codec-frame loop (talker step + code-predictor frame).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is ONE codec frame of the upstream frame loop (tts_engine.py:301-335): code_predictor.predict
(top-k sampling, T=0.9, k=50) -> 16-way embedding sum + trailing text embedding -> talker.step_with_embed,
on the seeded synthetic checkpoint (random-init weights of the Qwen3-TTS talker / code predictor), KV cache
allocated to 2048 positions, after the 8-step synthetic prefill + step(CODEC_BOS)  (BASELINE.json configs[2],
which contains configs[0] and configs[1] as its two halves; they are also timed separately below).

Printed JSON (one line, rank 0):
  value      codec frames/s, inputs resident in HBM, device-timed (CUDA events), max over ranks: the device-autonomous
             frame loop (TTSDecoder.generate_frames -> qmk_generate_nosync)
  e2e        same loop through the public API with the trailing-text embeddings coming from pinned host memory (H2D
             inside the timed region) and every frame's 16 codes + talker token written to pinned host memory (D2H)
  e2e_per_frame_api / e2e_dropin_loop / e2e_upstream_loop   the per-frame API in three control flows (see notes)
  roofline   qmk_decode_kernel in its talker configuration: algorithmic bytes per launch / mean launch time
  cpu_baseline  the CPU oracle port of the upstream PyTorch path, same frame loop, bounded sample
`--impl reference` times that CPU port as the main line (rank 0 only under torchrun).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.join(REPO, "qwen-megakernel-tts_b200")
for _p in (REPO, PKG_ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

CODEC_BOS = 2149
METRIC = "codec_frames_per_s_b1"
UNIT = "frames/s"
MAX_SEQ = 2048
N_PREFILL = 8
SEED = 1234

# Algorithmic bytes (SURVEY.md §8d / DESIGN.md): bf16 weights + norms read once per launch, KV read + append.
LAYER_BYTES = 31_461_888
TALKER_STEP_BYTES = 28 * LAYER_BYTES + 6_291_456 + 2_048 + 2_048 + 512          # 887,228,928
TALKER_KV_BYTES_PER_POS = 114_688
CP_STEP_BYTES = 5 * LAYER_BYTES + 2_048 + 2_048 + 512                              # 157,314,048
CP_KV_BYTES_PER_POS = 20_480
CP_HEAD_BYTES = 4_194_304


def talker_bytes(position: int) -> int:
    return TALKER_STEP_BYTES + TALKER_KV_BYTES_PER_POS * (position + 2)


def cp_frame_bytes() -> int:
    return sum(CP_STEP_BYTES + CP_KV_BYTES_PER_POS * (p + 2) for p in range(16)) + 15 * CP_HEAD_BYTES


def measured_peaks() -> tuple[float, str]:
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


NCU_RAW = "r02_ncu_b1_raw.csv"   # ncu --set full: launch 0 = one talker launch of the default (group) kernel


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of one talker launch, from the committed `ncu --set full` capture
    (profiles/<NCU_RAW>, transposed raw page); None if the file is missing."""
    path = os.path.join(REPO, "profiles", NCU_RAW)
    try:
        import csv
        vals = {}
        for row in csv.reader(open(path)):
            if row and row[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[row[1]]
                vals[row[0]] = float(row[2]) * scale
        return vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.tmp = None

    def start(self):
        try:
            self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.strip().split(",") for r in open(self.tmp.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.tmp.name)
        sm, reasons = [], set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                out["sm_max_mhz"] = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


# ── CPU port of the upstream frame loop (oracle) ───────────────────────────────────────────────────────
def cpu_frame_loop(weights_cpu, n_frames: int, warmup: int, budget_s: float, sample: bool = True):
    from oracle.tts_oracle import CodePredictorOracle, TalkerOracle, frame_embed_sum
    from qwen_megakernel.synthetic import synthetic_inputs
    torch.set_num_threads(os.cpu_count() or 1)
    talker = TalkerOracle(weights_cpu, max_seq=MAX_SEQ)
    cp = CodePredictorOracle(weights_cpu)
    prefill = synthetic_inputs(99, N_PREFILL)
    trail = synthetic_inputs(4321, n_frames + warmup + 1)
    cp_emb = cp.codec_embeddings
    gen = torch.Generator().manual_seed(7)
    for i in range(N_PREFILL):
        talker.step_with_embed(prefill[i])
    tok, hid = talker.step(CODEC_BOS)
    done, t0 = 0, None
    for f in range(n_frames + warmup):
        if f == warmup:
            t0 = time.perf_counter()
        codes = cp.predict(hid, tok, weights_cpu["embed_weight"], do_sample=sample, temperature=0.9, top_k=50,
                           generator=gen)
        e = frame_embed_sum(codes, weights_cpu["embed_weight"], cp_emb, trail[f])
        tok, hid = talker.step_with_embed(e)
        if f >= warmup:
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
    dt = time.perf_counter() - t0
    return done / dt, done, dt


# ── GPU frame loop through the public API ──────────────────────────────────────────────────────────────
class FrameLoop:
    def __init__(self, weights_gpu, device):
        from qwen_megakernel.model_tts import CodePredictorKernel, TTSDecoder
        from qwen_megakernel.synthetic import synthetic_inputs
        self.dev = device
        self.w = weights_gpu
        self.talker = TTSDecoder(weights=weights_gpu, verbose=False, max_seq_len=MAX_SEQ, device=device)
        self.cp = CodePredictorKernel(weights_gpu, device=str(device))
        self.prefill = synthetic_inputs(99, N_PREFILL).to(device)
        self.pad = synthetic_inputs(777, 1)[0].to(device)
        self.embed = weights_gpu["embed_weight"]
        self.cp_embeds = [weights_gpu["code_predictor"][f"codec_embedding.{g}.weight"] for g in range(15)]
        self.launches = 0

    def start_utterance(self):
        """tts_engine.py:281-289: reset, 8 prefill steps, step(CODEC_BOS)."""
        self.talker.reset()
        for i in range(N_PREFILL):
            self.talker.step_with_embed(self.prefill[i])
        self.tok, self.hid = self.talker.step(CODEC_BOS)
        self.launches += N_PREFILL + 1
        # device-resident (token, hidden) of the last talker step for the sync-free pipeline
        self.tok_dev, self.hid_dev = self.talker._out_token, self.talker._norm_out

    def frame_async(self, extra_bf16, sample=True):
        """One frame on the per-frame API without a host round trip: predict() takes the talker's token from device memory
        and the talker step returns device tensors (two launches per frame)."""
        codes = self.cp.predict(self.hid_dev, self.tok_dev, self.embed, do_sample=sample, temperature=0.9, top_k=50)
        self.tok_dev, self.hid_dev = self.talker.step_with_codes(codes, self.cp_embeds, extra_bf16, sync=False)
        self.launches += 2
        return codes

    def frame(self, extra_bf16, sample=True, torch_glue=False):
        """tts_engine.py:306-335 as written: predict -> embed sum -> step_with_embed, a Python int token per frame."""
        F = torch.nn.functional
        codes = self.cp.predict(self.hid, self.tok, self.embed, do_sample=sample, temperature=0.9, top_k=50)
        if torch_glue:           # the upstream caller's 32 torch launches (tts_engine.py:319-333)
            e = F.embedding(codes[0:1], self.embed).squeeze(0)
            for g in range(15):
                e = e + F.embedding(codes[g + 1:g + 2], self.cp_embeds[g]).squeeze(0)
            e = e + extra_bf16
            self.tok, self.hid = self.talker.step_with_embed(e)
        else:                    # same sum evaluated inside the talker step's launch
            self.tok, self.hid = self.talker.step_with_codes(codes, self.cp_embeds, extra_bf16)
        self.launches += 2
        return codes


def _median(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2] if len(xs) % 2 else 0.5 * (xs[len(xs) // 2 - 1] + xs[len(xs) // 2])


def time_autonomous(loop: FrameLoop, K: int, trail_dev, trail_host, e2e: bool, reps: int, barrier=None):
    """K frames per repetition through TTSDecoder.generate_frames (qmk_generate_nosync): the persistent kernel iterates
    predict -> embedding sum -> talker step itself.  e2e: the utterance's trailing-text embeddings come from pinned host
    memory (H2D inside the timed region) and the kernel writes every frame's codes / token / progress word straight into
    pinned host memory (D2H inside the timed region), which the host reads after the final synchronise.
    Returns (list of elapsed ms per repetition, launches per repetition)."""
    times, sink = [], 0
    launches = -(-K // 170)          # chained launches of <= 201 frames (16-bit epochs); see qmk_generate_nosync
    for _ in range(reps):
        loop.start_utterance()
        torch.cuda.synchronize()
        if barrier:
            barrier()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        if e2e:
            trail = trail_host[:K].to(loop.dev, non_blocking=True)
            codes, tokens, state = loop.talker.generate_frames(loop.cp, K, trail, loop.pad, do_sample=True, temperature=0.9,
                                                               top_k=50, eos_token=-1, host_visible=True, sync=False)
        else:
            codes, tokens, state = loop.talker.generate_frames(loop.cp, K, trail_dev[:K], loop.pad, do_sample=True,
                                                               temperature=0.9, top_k=50, eos_token=-1, sync=False)
        end.record()
        n_done = loop.talker.finish_generate(state)
        assert n_done == K, f"generate_frames produced {n_done} of {K} frames"
        if e2e:
            sink += int(codes[:, 15].sum()) + int(tokens.sum())
        if barrier:
            barrier()
        times.append(start.elapsed_time(end))
    return times, launches


def time_frames(loop: FrameLoop, K: int, W: int, trail_dev, trail_host, mode: str, reps: int, barrier=None):
    """K frames per repetition on the per-frame API.  mode: "async" (sync-free pipeline, device token), "dropin" (upstream
    control flow: blocking Python int per frame, embedding sum inside the talker launch), "upstream" (dropin + the upstream
    caller's 32 torch launches for the embedding sum).  With trail_host the frame's input comes from pinned host memory and
    its 16 codes + token go back to pinned host memory inside the timed region."""
    times, sink = [], 0
    for _ in range(reps):
        loop.start_utterance()
        for f in range(W):
            loop.frame_async(trail_dev[f]) if mode == "async" else loop.frame(trail_dev[f], torch_glue=(mode == "upstream"))
        torch.cuda.synchronize()
        if barrier:
            barrier()
        host_codes = torch.empty(K, 16, dtype=torch.int64).pin_memory()
        host_tok = torch.empty(K, 1, dtype=torch.int32).pin_memory()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for f in range(K):
            extra = trail_host[W + f].to(loop.dev, non_blocking=True) if trail_host is not None else trail_dev[W + f]
            if mode == "async":
                codes = loop.frame_async(extra)
                if trail_host is not None:
                    host_codes[f].copy_(codes, non_blocking=True)
                    host_tok[f].copy_(loop.tok_dev, non_blocking=True)
            else:
                codes = loop.frame(extra, torch_glue=(mode == "upstream"))
                sink += int(codes.cpu()[15])                                      # blocking D2H of this frame's result
        end.record()
        torch.cuda.synchronize()
        if mode == "async" and trail_host is not None:
            sink += int(host_codes[:, 15].sum()) + int(host_tok.sum())
        if barrier:
            barrier()
        times.append(start.elapsed_time(end))
    return times


def time_prefill(loop: FrameLoop, reps: int = 5):
    """The 8-step prefill of an utterance (tts_engine.py:281-282): upstream's loop of sequential step_with_embed calls against
    TTSDecoder.prefill (one batched tcgen05 pass, qmk_batched_prefill); then time to the first codec frame (prefill +
    step(CODEC_BOS) + one predict).  Wall-clock ms around a final synchronise (these are latency numbers)."""
    t = loop.talker
    out = {}
    bos_row = loop.embed[CODEC_BOS:CODEC_BOS + 1].to(loop.prefill.dtype)
    with_bos = torch.cat([loop.prefill.reshape(N_PREFILL, -1), bos_row], dim=0)      # step(CODEC_BOS) == one more prefill row
    for name in ("sequential", "one_pass", "one_pass_bos"):
        best = 1e9
        for _ in range(reps + 1):
            t.reset()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if name == "sequential":
                for i in range(N_PREFILL):
                    t.step_with_embed(loop.prefill[i])
            elif name == "one_pass":
                t.prefill(loop.prefill)
            else:
                tok, hid = t.prefill(with_bos)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            if name != "one_pass_bos":
                tok, hid = t.step(CODEC_BOS)
            codes = loop.cp.predict(hid, tok, loop.embed, do_sample=True, temperature=0.9, top_k=50)
            codes.cpu()
            t2 = time.perf_counter()
            best = min(best, (t1 - t0) * 1e3)
            out[f"first_frame_ms_{name}"] = min(out.get(f"first_frame_ms_{name}", 1e9), (t2 - t0) * 1e3)
        out[f"prefill_ms_{name}"] = best
    t.reset()
    return out


def time_text_projection(weights_gpu, dev, n_tokens=(32, 128), iters: int = 50):
    """Text side of the prefill (tts_engine.py:262-263, once per utterance): TextProjectionKernel.embed_text_ids (one chain of
    five kernels per 512 tokens, fc1 / fc2 on tcgen05) beside upstream's four PyTorch operators (TextProjection), CUDA-event
    ms per call with the ids already on the device."""
    from qwen_megakernel.model_tts import TextProjection, TextProjectionKernel
    native, glue = TextProjectionKernel(weights_gpu, device=str(dev)), TextProjection(weights_gpu, device=str(dev))
    out = {}
    for n in n_tokens:
        ids = torch.randint(0, weights_gpu["text_embedding"].shape[0], (n,), device=dev)
        for name, fn in (("native", native.embed_text_ids), ("torch_glue", glue.embed_text_ids)):
            for _ in range(5):
                fn(ids)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                fn(ids)
            b.record()
            torch.cuda.synchronize()
            out[f"T{n}_{name}_ms"] = a.elapsed_time(b) / iters
    out["algorithmic_bytes"] = 2 * (2048 * 2048 + 1024 * 2048 + 2048 + 1024)     # fc1 + fc2 + biases (+ 4 KB per token)
    out["note"] = ("TextProjectionKernel.embed_text_ids (qmk_text_proj_embed: gather -> tcgen05 fc1 -> bias + SiLU -> tcgen05 fc2 -> bias, "
                   "5 launches per 512 tokens) vs upstream's embedding / linear / silu / linear in PyTorch; launch-latency bound "
                   "(12.6 MB of weights = 2 us at the HBM peak)")
    return out


def time_first_frame_from_text(loop: FrameLoop, n_text: int = 24, reps: int = 6):
    """Time to the first codec frame from TEXT IDS on the host (the part of upstream's TTFC this path covers, README.md:17-25:
    embed build 7.2 + prefill 24.9 + first decode 3.1 + first code predictor 13.0 ms on an RTX 5090): ids H2D -> text projection
    -> build_prefill_embeddings (cached pad / bos / eos, tts_engine.py:107-118) -> one-pass prefill with the BOS row -> predict ->
    16 codes on the host.  Wall-clock ms (best of reps), native text projection vs the PyTorch operators."""
    from qwen_megakernel.model_tts import TextProjection, TextProjectionKernel, build_prefill_embeddings
    dev = loop.dev
    rows = loop.w["text_embedding"].shape[0]
    ids_host = torch.randint(0, rows, (3 + n_text + 5,)).pin_memory()      # 3 role ids, the content, 5 closing format ids
    bos_row = loop.embed[CODEC_BOS:CODEC_BOS + 1]
    out = {}
    for name, tp in (("native", TextProjectionKernel(loop.w, device=str(dev))), ("torch_glue", TextProjection(loop.w, device=str(dev)))):
        sp = tp.embed_text_ids(torch.arange(3, device=dev) % rows)
        cached = {"pad": sp[0:1], "bos": sp[1:2], "eos": sp[2:3]}
        best = 1e9
        for _ in range(reps + 2):
            loop.talker.reset()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            prefill, _trailing = build_prefill_embeddings(ids_host.to(dev, non_blocking=True), tp, loop.embed, device=str(dev),
                                                         cached_tts_embeds=cached)
            tok, hid = loop.talker.prefill(torch.cat([prefill, bos_row], dim=0))
            codes = loop.cp.predict(hid, tok, loop.embed, do_sample=True, temperature=0.9, top_k=50)
            codes.cpu()
            best = min(best, (time.perf_counter() - t0) * 1e3)
        out[f"first_frame_ms_from_text_ids_{name}"] = best
    loop.talker.reset()
    return out


def time_talker_at(loop: FrameLoop, position: int, n: int = 40, warmup: int = 5):
    """Mean duration of one talker launch with `position` cached rows (the launch is repeated at the same position: the KV rows
    it reads are whatever the cache holds, which does not change the work)."""
    t = loop.talker
    t._hidden.copy_(loop.prefill[0])
    for _ in range(warmup):
        t._position = position
        t._launch(-1, t._hidden.data_ptr())
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(n):
        t._position = position
        t._launch(-1, t._hidden.data_ptr())
    end.record()
    torch.cuda.synchronize()
    assert int(t._out_token.item()) >= 0, "talker kernel reported failure"
    t.reset()
    return start.elapsed_time(end) / n


def time_talker_kernel(loop: FrameLoop, n: int = 50, warmup: int = 10):
    """Mean duration of one talker launch over positions 18 .. 18 + n (the short-context end of the frame loop)."""
    t = loop.talker
    t.reset()
    for i in range(N_PREFILL):
        t.step_with_embed(loop.prefill[i])
    t._hidden.copy_(loop.prefill[0])
    for _ in range(warmup):
        t._launch(-1, t._hidden.data_ptr())
    torch.cuda.synchronize()
    p0 = t.position
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(n):
        t._launch(-1, t._hidden.data_ptr())
    end.record()
    torch.cuda.synchronize()
    assert int(t._out_token.item()) >= 0, "talker kernel reported failure"
    ms = start.elapsed_time(end) / n
    bytes_mean = sum(talker_bytes(p0 + i) for i in range(n)) / n
    return ms, bytes_mean


def time_cp_frame(loop: FrameLoop, n: int = 30, warmup: int = 5, sample: bool = False):
    hid = loop.hid if hasattr(loop, "hid") else torch.zeros(1024, device=loop.dev)
    for _ in range(warmup):
        loop.cp.predict(hid, 1335, loop.embed, do_sample=sample)
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(n):
        loop.cp.predict(hid, 1335, loop.embed, do_sample=sample)
    end.record()
    torch.cuda.synchronize()
    return start.elapsed_time(end) / n


def time_batched(weights_gpu, dev, batch: int, steps: int = 100, warmup: int = 10, max_seq: int = 512):
    """BASELINE.json configs[3]: B concurrent utterance streams, tcgen05 projections (BatchedTTSDecoder).  One step =
    one token for every stream; tokens are fed back on the device, no host sync inside the timed region."""
    from qwen_megakernel.model_tts import BatchedTTSDecoder
    bd = BatchedTTSDecoder(weights_gpu, batch, device=dev, max_seq_len=max_seq)
    tok = torch.full((batch,), CODEC_BOS, dtype=torch.int32, device=dev)
    def run(step_fn, n):
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(n):
            step_fn()
        end.record()
        torch.cuda.synchronize()
        return start.elapsed_time(end) / n
    def plain():
        t, _ = bd.step(tok)
        tok.copy_(t)
    run(plain, warmup)
    bd.reset()
    ms_launches = run(plain, steps)                   # 227 launches issued by the host per step
    bd.reset()
    bd.step_graph(tok)
    run(bd.step_graph, max(warmup, 3))                # feedback form (tokens stay on the device); the second call captures
    bd.reset()
    ms = run(bd.step_graph, steps)                    # the same launches replayed from a CUDA graph
    bytes_step = TALKER_STEP_BYTES + batch * sum(TALKER_KV_BYTES_PER_POS * (p + 2) for p in range(steps)) / steps
    del bd
    return {"streams": batch, "ms_per_step": ms, "stream_steps_per_s": batch * 1000.0 / ms,
            "algorithmic_gbs": bytes_step / (ms * 1e-3) / 1e9, "launches_per_step": 8 * 28 + 3,
            "ms_per_step_host_launches": ms_launches,
            "note": "talker step for B streams: tcgen05/TMEM split-K GEMMs + fused epilogues, PDL chain replayed from a CUDA "
                    "graph (BatchedTTSDecoder.step_graph; ms_per_step_host_launches = the same kernels launched one by one); "
                    "positions 0..%d" % steps}


def time_batched_frames(weights_gpu, dev, batch: int, frames: int = 20, warmup: int = 3, max_seq: int = 512):
    """BASELINE.json configs[3]-[4] in codec frames/s: B concurrent utterances through the whole frame loop (batched
    code-predictor frame with top-k sampling + 16-way embedding sum + batched talker step per frame), no host sync inside."""
    from qwen_megakernel.model_tts import BatchedFrameLoop
    from qwen_megakernel.synthetic import synthetic_inputs
    loop = BatchedFrameLoop(weights_gpu, batch, device=dev, max_seq_len=max_seq)
    pre = synthetic_inputs(99, N_PREFILL * batch).to(dev).view(N_PREFILL, batch, 1024)
    extra = synthetic_inputs(4321, batch).to(dev)
    loop.start(pre)
    for _ in range(warmup):
        loop.frame(extra)
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(frames):
        loop.frame(extra)
    end.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(end) / frames
    del loop
    return {"ms_per_frame_all_streams": ms, "codec_frames_per_s": batch * 1000.0 / ms,
            "launches_per_frame": 16 * (8 * 5 + 1) + 15 * 2 + 1 + (8 * 28 + 3),
            "note": "BatchedFrameLoop.frame: 16 batched code-predictor steps + 15 heads (T 0.9 / top-k 50) + embedding sum + talker "
                    "step, one CUDA-graph replay per frame"}


def time_upstream_kernel_subprocess(limit_s: float = 60.0) -> dict:
    """Launch times of the unmodified upstream kernel (oracle/ref_kernel.py) measured in a child process."""
    lib = os.path.join(REPO, "oracle", "_ref", "libref_kernel_sm100a.so")
    if not os.path.exists(lib):
        return {"available": False, "why": "oracle/_ref/libref_kernel_sm100a.so not built (needs the reference tree at build time)"}
    code = ("import sys, json; sys.path.insert(0, %r); sys.path.insert(0, %r); import torch\n"
            "from oracle import ref_kernel\n"
            "from qwen_megakernel.synthetic import synthetic_inputs, synthetic_tts_weights, weights_to\n"
            "w = weights_to(synthetic_tts_weights(seed=%d, max_seq_len=%d), 'cuda'); x = synthetic_inputs(99, %d).cuda()\n"
            "t, c = ref_kernel.time_upstream_kernel(w, x); print(json.dumps({'talker_step_us': t, 'cp_step_us': c}))\n"
            % (REPO, PKG_ROOT, SEED, MAX_SEQ, N_PREFILL))
    try:
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=limit_s)
        d = json.loads(out.stdout.strip().splitlines()[-1])
        d.update(available=True, frame_us_kernels_only=d["talker_step_us"] + 16 * d["cp_step_us"],
                 note="upstream csrc/kernel.cu (decode kernel + LM-head kernel + memset per step) recompiled with -gencode "
                      "arch=compute_100a,code=sm_100a and upstream's build_tts.py flags; same weights, positions 18..68")
        return d
    except subprocess.TimeoutExpired:
        return {"available": False, "why": f"upstream kernel hung (> {limit_s:.0f} s): its grid barrier deadlocks intermittently on B200"}
    except Exception as ex:  # noqa: BLE001
        return {"available": False, "why": f"{type(ex).__name__}: {ex}"}


def pin_rank_to_cores(local_rank: int, world: int) -> list:
    """Give every rank its own slice of the host cores (8 Python launchers otherwise migrate and share cores)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(world, 1))
        mine = cores[local_rank * per:(local_rank + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        return mine
    except Exception:
        return []


def gather_floats(x: float, world: int, dev) -> list:
    if world == 1:
        return [float(x)]
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(o[0]) for o in out]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500, help="codec frames per utterance (BASELINE.json configs[2]: 500)")
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--min-seconds", type=float, default=0.6, help="every timed leg is repeated until it covers this much device time")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batched", action="store_true")
    ap.add_argument("--upstream-kernel", action="store_true",
                    help="also time upstream's kernel.cu built for sm_100a (oracle/_ref; child process with a time limit)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    K, W = args.steps, max(args.warmup, 3 if args.impl == "b200" else 1)
    if K + N_PREFILL + 1 + W > MAX_SEQ:
        raise SystemExit(f"--steps {K}: an utterance must fit the {MAX_SEQ}-position KV cache")
    # identical in both arms (the driver compares the dicts); implementation details live under "impl_detail"
    config = {"workload": "qwen3-tts B=1 frame loop: code predictor (5L, 16 steps + 15 heads, top-k sampling T=0.9 k=50) + "
                          "talker (28L, hidden 1024, vocab 3072) step per frame; synthetic 8-step prefill + step(CODEC_BOS); "
                          "KV cache to 2048 positions",
              "frames_per_utterance": K, "global_batch": world, "streams_per_gpu": 1, "kv_max_positions": MAX_SEQ,
              "parallelism": f"replicas x{world} (one engine per GPU, no collective on the path)",
              "l2_policy": "weights per step (887 MB talker / 157 MB code predictor x16) exceed the 126 MB L2; no flush needed"}

    from qwen_megakernel.synthetic import synthetic_inputs, synthetic_tts_weights, weights_to

    if args.impl == "reference":
        if rank != 0:
            return
        w = synthetic_tts_weights(seed=SEED, max_seq_len=MAX_SEQ)
        # >= 6 s of CPU work regardless of --steps (a 20-frame sample is 2 s of a noisy host), bounded at 150 s
        est = 10.0
        n_frames = max(K, int(6.0 * est))
        fps, done, dt = cpu_frame_loop(w, n_frames, W, 150.0)
        line = {"metric": METRIC, "value": fps, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": K,
                "steps_timed": done, "warmup": W, "ms_per_step": 1000.0 / fps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
                "impl_detail": "oracle port of the upstream PyTorch path (validate_kernel.PyTorchTalkerReference + model_tts.CodePredictor + "
                               "tts_engine.py:319-333 glue) on the host cores; the reference's own GPU path needs an sm_120a build",
                "cpu_baseline": {"value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                 "sample": f"{done} frames of the same loop in {dt:.1f} s (oracle/tts_oracle.py)"},
                "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU for --impl b200 (there is no CPU fallback)"
    cores = pin_rank_to_cores(local_rank, world)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    barrier = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout during the first collective; stdout must carry ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        barrier = lambda: (dist.barrier(), torch.cuda.synchronize())  # noqa: E731

    w_cpu = synthetic_tts_weights(seed=SEED, max_seq_len=MAX_SEQ)
    w_gpu = weights_to(w_cpu, str(dev))
    trail_cpu = synthetic_inputs(4321, K + W + 1)
    trail_dev = trail_cpu.to(dev)
    trail_host = trail_cpu.pin_memory()
    loop = FrameLoop(w_gpu, dev)

    # kernel-only legs (explain the headline): talker launch at the short-context end and at fixed depths, code-predictor frame
    talker_ms, talker_b = time_talker_kernel(loop)
    by_pos = {}
    for pos in (50, 500, 2047):
        ms = time_talker_at(loop, pos)
        by_pos[f"p{pos}"] = {"launch_us": ms * 1e3, "algorithmic_bytes": talker_bytes(pos),
                             "achieved_gbs": talker_bytes(pos) / (ms * 1e-3) / 1e9}
    prefill = time_prefill(loop)
    loop.start_utterance()
    cp_ms = time_cp_frame(loop)
    cp_ms_sampled = time_cp_frame(loop, sample=True)

    # repetitions: every leg covers >= --min-seconds of device time whatever --steps is
    est_ms_frame = cp_ms_sampled + talker_ms * 1.25
    reps = max(1, int(args.min_seconds * 1000.0 / (est_ms_frame * K) + 0.999))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = loop.launches
    t_dev, launches_auto = time_autonomous(loop, K, trail_dev, trail_host, False, reps, barrier)
    launches_value = loop.launches - l0 + launches_auto * reps
    t_e2e, _ = time_autonomous(loop, K, trail_dev, trail_host, True, reps, barrier)
    t_async = time_frames(loop, K, W, trail_dev, trail_host, "async", reps, barrier)
    t_dropin = time_frames(loop, K, W, trail_dev, trail_host, "dropin", reps, barrier)
    t_upstream = time_frames(loop, K, W, trail_dev, trail_host, "upstream", max(1, reps // 2), barrier)
    clocks = sampler.stop() if rank == 0 else {}

    from qwen_megakernel.replicas import combine
    def agg(times):
        ms_local = _median(times)
        frames, ms = combine(K, ms_local, device=dev)      # sum of frames over ranks, max over ranks of the (median) device time
        return frames / (ms / 1000.0), ms, gather_floats(ms_local, world, dev)
    value, ms_dev, per_rank = agg(t_dev)
    e2e, ms_e2e, per_rank_e2e = agg(t_e2e)
    v_async, _, _ = agg(t_async)
    v_dropin, _, _ = agg(t_dropin)
    v_upstream, _, _ = agg(t_upstream)

    batched = None
    if not args.no_batched:
        batched = {}
        for b in (16, 64):
            r = time_batched(w_gpu, dev, b)
            tot, ms_b = combine(r["stream_steps_per_s"], r["ms_per_step"], device=dev)   # replicas: sum of rates, max of step times
            r["stream_steps_per_s_all_gpus"] = tot
            r["ms_per_step_max_over_ranks"] = ms_b
            fr = time_batched_frames(w_gpu, dev, b)
            fr["codec_frames_per_s_all_gpus"], fr["ms_per_frame_max_over_ranks"] = combine(fr["codec_frames_per_s"], fr["ms_per_frame_all_streams"], device=dev)
            r["frame_loop"] = fr
            batched[f"B{b}"] = r
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # GPU-side reference point (opt-in, --upstream-kernel): upstream's own kernel.cu compiled for sm_100a (oracle/Makefile ->
    # oracle/_ref/*.so, built from the reference tree in the build container).  On B200 that kernel deadlocks intermittently in
    # its grid barrier (2 of 5 attempts hung on the first launch in round 2), so it runs in a child process with a time limit
    # and never in the default (driver) run.
    upstream_kernel = {"available": False, "why": "not requested (--upstream-kernel)"}
    if args.upstream_kernel:
        upstream_kernel = time_upstream_kernel_subprocess()

    peak, peak_src = measured_peaks()
    n_ctas = loop.talker._lib.qmk_engine_num_ctas(loop.talker._engine)
    kernel_name = "qmk2_decode_kernel" if n_ctas == 128 else "qmk_decode_kernel"
    achieved = talker_b / (talker_ms * 1e-3) / 1e9
    for v in by_pos.values():
        v["frac"] = v["achieved_gbs"] / peak
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": config,
        "impl_detail": {"engine": ("group kernel: 8 kv-head groups x 16 CTAs, K-split O/down reduced in L2" if n_ctas == 128
                                   else "row-split kernel: one CTA per SM"),
                        "value_path": "TTSDecoder.generate_frames -> qmk_generate_nosync: the persistent kernel runs the whole frame loop "
                                      "(predict + 16-way embedding sum + talker step per frame, device-side EOS flag), prefill + step(BOS) outside "
                                      "the timed region",
                        "timing": f"median of {reps} repetitions of the {K}-frame utterance per leg (CUDA events), max over ranks",
                        "host_cores_of_rank0": len(cores)},
        "reps": reps, "per_rank_ms": {"value": per_rank, "e2e": per_rank_e2e},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 2048, "d2h_bytes_per_step": 16 * 8 + 4 + 4,
                "note": "public API (TTSDecoder.generate_frames, host_visible=True): the utterance's trailing-text embeddings are copied from "
                        "pinned host memory inside the timed region (2048 B per frame), every frame's 16 codes + talker token + progress "
                        "word are written by the kernel into pinned host memory and read by the host after the final synchronise"},
        "e2e_per_frame_api": {"value": v_async, "unit": UNIT,
                              "note": "two launches per frame (predict + step_with_codes, device token), input H2D / codes D2H per frame"},
        "e2e_dropin_loop": {"value": v_dropin, "unit": UNIT,
                            "note": "upstream caller's control flow: step() returns a Python int (blocking D2H per frame)"},
        "e2e_upstream_loop": {"value": v_upstream, "unit": UNIT,
                              "note": "tts_engine.py:301-335 as written: blocking token per frame AND the caller's 32 torch launches for the embedding sum"},
        "gpu_launches": launches_value,
        "roofline": {"bound": "hbm", "kernel": f"{kernel_name} (talker: 28 layers + LM head, one launch per step, {n_ctas} CTAs)",
                     "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic_bytes(),
                     "traffic_source": f"static: profiles/{NCU_RAW} (ncu --set full of one talker launch, committed; not measured in this run)",
                     "algorithmic_bytes_per_launch": talker_b, "launch_us": talker_ms * 1e3,
                     "positions": "18..68 (short-context end of the utterance); see roofline_by_position"},
        "roofline_by_position": by_pos,
        "talker_steps_per_s": 1000.0 / talker_ms,
        "cp_frame": {"ms": cp_ms, "ms_sampled": cp_ms_sampled, "frames_per_s": 1000.0 / cp_ms, "algorithmic_bytes": cp_frame_bytes(),
                     "achieved_gbs": cp_frame_bytes() / (cp_ms * 1e-3) / 1e9, "frac": cp_frame_bytes() / (cp_ms * 1e-3) / 1e9 / peak,
                     "note": "predict(): one fused launch (16 steps + 15 heads + selection); ms = greedy, ms_sampled = T 0.9 / top-k 50"},
        "prefill": dict(prefill, note="8 prefill embeddings: sequential step_with_embed calls (upstream's loop, one blocking token read each) vs "
                                      "TTSDecoder.prefill (one batched tcgen05 pass: launch chain with lane = position, CUDA-graph replay from the "
                                      "third call); first_frame = prefill + step(BOS) + predict + codes on the host; one_pass_bos = the same pass "
                                      "over 9 rows with the codec BOS embedding as the last one (replaces step(BOS))"),
        "upstream_kernel_sm100a": upstream_kernel,
        "clocks": clocks,
    }
    if batched is not None:
        line["batched"] = batched
    try:      # last GPU leg: nothing measured above depends on it
        line["text_projection"] = time_text_projection(w_gpu, dev)
        line["text_projection"].update(time_first_frame_from_text(loop))
    except Exception as e:  # noqa: BLE001
        line["text_projection"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if world == 1 and not args.no_cpu_baseline:
        fps, done, dt = cpu_frame_loop(w_cpu, 10_000, 1, args.cpu_budget)
        line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{done} frames of the same loop in {dt:.1f} s (oracle/tts_oracle.py)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

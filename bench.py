#!/usr/bin/env python3
"""Benchmark of the accelerated path: B=1 codec-frame loop (talker step + code-predictor frame).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is ONE codec frame of the upstream frame loop (tts_engine.py:301-335): code_predictor.predict
(top-k sampling, T=0.9, k=50) -> 16-way embedding sum + trailing text embedding -> talker.step_with_embed,
on the seeded synthetic checkpoint (random-init weights of the Qwen3-TTS talker / code predictor), KV cache
allocated to 2048 positions, after the 8-step synthetic prefill + step(CODEC_BOS)  (BASELINE.json configs[2],
which contains configs[0] and configs[1] as its two halves; they are also timed separately below).

Printed JSON (one line, rank 0):
  value      codec frames/s, inputs resident in HBM, device-timed (CUDA events), max over ranks
  e2e        same loop through the public API with the per-frame trailing-text embedding coming from pinned
             host memory (H2D inside the timed region) and the frame's 16 codes + talker token read back (D2H)
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


NCU_RAW = "r01_ncu_group_raw.csv"   # ncu --set full of one talker launch of the default (group) kernel


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
                 "-lms", "100"], stdout=self.tmp, stderr=subprocess.DEVNULL)
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
    def __init__(self, weights_gpu, device, torch_glue=False):
        self.torch_glue = torch_glue
        from qwen_megakernel.model_tts import CodePredictorKernel, TTSDecoder
        from qwen_megakernel.synthetic import synthetic_inputs
        self.dev = device
        self.w = weights_gpu
        self.talker = TTSDecoder(weights=weights_gpu, verbose=False, max_seq_len=MAX_SEQ, device=device)
        self.cp = CodePredictorKernel(weights_gpu, device=str(device))
        self.prefill = synthetic_inputs(99, N_PREFILL).to(device)
        self.embed = weights_gpu["embed_weight"]
        self.cp_embeds = [weights_gpu["code_predictor"][f"codec_embedding.{g}.weight"] for g in range(15)]
        self.launches = 0

    def start_utterance(self):
        self.talker.reset()
        for i in range(N_PREFILL):
            self.talker.step_with_embed(self.prefill[i])
        self.tok, self.hid = self.talker.step(CODEC_BOS)
        # device-resident (token, hidden) of the last talker step for the sync-free pipeline
        self.tok_dev, self.hid_dev = self.talker._out_token, self.talker._norm_out

    def frame_async(self, extra_bf16, sample=True):
        """Same frame without a host round trip: predict() takes the talker's token from device memory and the talker
        step returns device tensors; the host reads tokens / codes asynchronously (EOS would be seen one frame late)."""
        if self.talker.position >= MAX_SEQ - 1:
            self.start_utterance()
        codes = self.cp.predict(self.hid_dev, self.tok_dev, self.embed, do_sample=sample, temperature=0.9, top_k=50)
        self.tok_dev, self.hid_dev = self.talker.step_with_codes(codes, self.cp_embeds, extra_bf16, sync=False)
        self.launches += 2
        return codes

    def frame(self, extra_bf16, sample=True):
        """tts_engine.py:306-335: predict -> embed sum -> step_with_embed."""
        F = torch.nn.functional
        if self.talker.position >= MAX_SEQ - 1:
            self.start_utterance()
        codes = self.cp.predict(self.hid, self.tok, self.embed, do_sample=sample, temperature=0.9, top_k=50)
        if self.torch_glue:      # the upstream caller's 32 torch launches (tts_engine.py:319-333)
            e = F.embedding(codes[0:1], self.embed).squeeze(0)
            for g in range(15):
                e = e + F.embedding(codes[g + 1:g + 2], self.cp_embeds[g]).squeeze(0)
            e = e + extra_bf16
            self.tok, self.hid = self.talker.step_with_embed(e)
        else:                    # same sum evaluated inside the talker step's launch
            self.tok, self.hid = self.talker.step_with_codes(codes, self.cp_embeds, extra_bf16)
        self.launches += 2          # one fused code-predictor frame launch + one talker step launch
        return codes


def time_frames(loop: FrameLoop, n: int, warmup: int, trail_dev, trail_host=None, barrier=None, dropin=False):
    """Returns elapsed ms for n frames (CUDA events on the current stream).

    dropin=False: the sync-free pipeline (frame_async).  With trail_host, every frame's input comes from pinned host
    memory (H2D inside the timed region) and every frame's 16 codes + talker token are copied to pinned host memory
    (D2H inside the timed region, read by the host after the final synchronise).
    dropin=True: the upstream caller's control flow (tts_engine.py:301-335): step() returns a Python int, i.e. one
    blocking device->host read per frame."""
    loop.start_utterance()
    step = loop.frame if dropin else loop.frame_async
    for f in range(warmup):
        step(trail_dev[f])
    torch.cuda.synchronize()
    if barrier:
        barrier()
    host_codes = torch.empty(n, 16, dtype=torch.int64).pin_memory() if trail_host is not None else None
    host_tok = torch.empty(n, 1, dtype=torch.int32).pin_memory() if trail_host is not None else None
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = loop.launches
    start.record()
    sink = 0
    for f in range(n):
        if trail_host is None:
            codes = step(trail_dev[warmup + f])
        else:
            extra = trail_host[warmup + f].to(loop.dev, non_blocking=True)       # H2D of this frame's input
            codes = step(extra)
            if dropin:
                sink += int(codes.cpu()[15])                                      # blocking D2H of this frame's result
            else:
                host_codes[f].copy_(codes, non_blocking=True)                     # D2H of this frame's result
                host_tok[f].copy_(loop.tok_dev, non_blocking=True)
    end.record()
    torch.cuda.synchronize()
    if host_codes is not None and not dropin:
        sink += int(host_codes[:, 15].sum()) + int(host_tok.sum())
    if barrier:
        barrier()
    return start.elapsed_time(end), loop.launches - l0


def time_talker_kernel(loop: FrameLoop, n: int = 50, warmup: int = 10):
    """Mean duration of one talker launch of qmk_decode_kernel: n back-to-back launches, no host sync inside."""
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
    tok = int(t._out_token.item())
    assert tok >= 0, "talker kernel reported failure"
    ms = start.elapsed_time(end) / n
    bytes_mean = sum(talker_bytes(p0 + i) for i in range(n)) / n
    return ms, bytes_mean


def time_cp_frame(loop: FrameLoop, n: int = 30, warmup: int = 5):
    hid = loop.hid if hasattr(loop, "hid") else torch.zeros(1024, device=loop.dev)
    for _ in range(warmup):
        loop.cp.predict(hid, 1335, loop.embed, do_sample=False)
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(n):
        loop.cp.predict(hid, 1335, loop.embed, do_sample=False)
    end.record()
    torch.cuda.synchronize()
    return start.elapsed_time(end) / n


def time_batched(weights_gpu, dev, batch: int, steps: int = 100, warmup: int = 10, max_seq: int = 512):
    """BASELINE.json configs[3]: B concurrent utterance streams, tcgen05 projections (BatchedTTSDecoder).  One step =
    one token for every stream; tokens are fed back on the device, no host sync inside the timed region."""
    from qwen_megakernel.model_tts import BatchedTTSDecoder
    bd = BatchedTTSDecoder(weights_gpu, batch, device=dev, max_seq_len=max_seq)
    tok = torch.full((batch,), CODEC_BOS, dtype=torch.int32, device=dev)
    for _ in range(warmup):
        t, _ = bd.step(tok)
        tok.copy_(t)
    bd.reset()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(steps):
        t, _ = bd.step(tok)
        tok.copy_(t)
    end.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(end) / steps
    bytes_step = TALKER_STEP_BYTES + batch * sum(TALKER_KV_BYTES_PER_POS * (p + 2) for p in range(steps)) / steps
    del bd
    return {"streams": batch, "ms_per_step": ms, "stream_steps_per_s": batch * 1000.0 / ms,
            "algorithmic_gbs": bytes_step / (ms * 1e-3) / 1e9, "launches_per_step": 6 * 28 + 3,
            "note": "talker step for B streams: tcgen05/TMEM split-K GEMMs + fused epilogues, PDL chain; positions 0..%d" % steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--torch-glue", action="store_true",
                    help="do the per-frame embedding sum with torch ops like upstream tts_engine.py instead of step_with_codes")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    K, W = args.steps, max(args.warmup, 3 if args.impl == "b200" else 1)
    config = {"workload": "qwen3-tts B=1 frame loop: code predictor (5L, 16 steps + 15 heads, top-k sampling) + "
                          "talker (28L, hidden 1024, vocab 3072) step per frame; synthetic 8-step prefill; KV to 2048",
              "frame_glue": "torch ops (upstream tts_engine.py:319-333)" if args.torch_glue else "fused into the talker launch (TTSDecoder.step_with_codes)",
              "global_batch": world, "streams_per_gpu": 1, "kv_max_positions": MAX_SEQ,
              "parallelism": f"replicas x{world} (one engine per GPU, no collective on the path)",
              "l2_policy": "weights per step (887 MB talker / 157 MB code predictor x16) exceed the 126 MB L2; no flush needed"}

    from qwen_megakernel.synthetic import synthetic_tts_weights, weights_to

    if args.impl == "reference":
        if rank != 0:
            return
        w = synthetic_tts_weights(seed=SEED, max_seq_len=MAX_SEQ)
        budget = 150.0
        config["frame_glue"] = "torch ops on the CPU (oracle port of tts_engine.py:319-333)"
        fps, done, dt = cpu_frame_loop(w, K, W, budget)
        line = {"metric": METRIC, "value": fps, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": K,
                "steps_timed": done, "warmup": W, "ms_per_step": 1000.0 / fps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                 "sample": f"{done} frames of the same loop in {dt:.1f} s (oracle/tts_oracle.py)"},
                "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), "bench.py needs a GPU for --impl b200 (there is no CPU fallback)"
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
    from qwen_megakernel.synthetic import synthetic_inputs
    trail_cpu = synthetic_inputs(4321, K + W + 1)
    trail_dev = trail_cpu.to(dev)
    trail_host = trail_cpu.pin_memory()
    loop = FrameLoop(w_gpu, dev, torch_glue=args.torch_glue)

    # kernel-only legs (explain the headline): talker launch and code-predictor frame
    talker_ms, talker_b = time_talker_kernel(loop)
    loop.start_utterance()
    cp_ms = time_cp_frame(loop)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, launches = time_frames(loop, K, W, trail_dev, None, barrier)
    ms_e2e, _ = time_frames(loop, K, W, trail_dev, trail_host, barrier)
    ms_dropin, _ = time_frames(loop, K, W, trail_dev, trail_host, barrier, dropin=True)
    clocks = sampler.stop() if rank == 0 else {}

    from qwen_megakernel.replicas import combine
    frames_dev, ms_dev = combine(K, ms_dev, device=dev)      # sum of frames over ranks, max of device time
    frames_e2e, ms_e2e = combine(K, ms_e2e, device=dev)
    frames_dropin, ms_dropin = combine(K, ms_dropin, device=dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    n_ctas = loop.talker._lib.qmk_engine_num_ctas(loop.talker._engine)
    kernel_name = "qmk2_decode_kernel" if n_ctas == 128 else "qmk_decode_kernel"
    config["engine"] = ("group kernel: 8 kv-head groups x 16 CTAs, K-split O/down reduced in L2" if n_ctas == 128
                        else "row-split kernel: one CTA per SM")
    value = frames_dev / (ms_dev / 1000.0)
    e2e = frames_e2e / (ms_e2e / 1000.0)
    achieved = talker_b / (talker_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": config,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 2048, "d2h_bytes_per_step": 16 * 8 + 4,
                "note": "public API, sync-free pipeline: input embedding from pinned host memory, codes + token copied to pinned "
                        "host memory every frame"},
        "e2e_dropin_loop": {"value": frames_dropin / (ms_dropin / 1000.0), "unit": UNIT,
                            "note": "upstream caller's control flow: step() returns a Python int (blocking D2H per frame)"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": f"{kernel_name} (talker: 28 layers + LM head, one launch per step, {n_ctas} CTAs)",
                     "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic_bytes(),
                     "traffic_source": f"profiles/{NCU_RAW} (ncu --set full, one talker launch, bytes)",
                     "algorithmic_bytes_per_launch": talker_b, "launch_us": talker_ms * 1e3},
        "talker_steps_per_s": 1000.0 / talker_ms,
        "cp_frame": {"ms": cp_ms, "frames_per_s": 1000.0 / cp_ms, "algorithmic_bytes": cp_frame_bytes(),
                     "achieved_gbs": cp_frame_bytes() / (cp_ms * 1e-3) / 1e9, "note": "greedy predict(): one fused launch (16 steps + 15 heads + selection)"},
        "clocks": clocks,
    }
    if world == 1:
        line["batched"] = {f"B{b}": time_batched(w_gpu, dev, b) for b in (16, 64)}
    if world == 1 and not args.no_cpu_baseline:
        fps, done, dt = cpu_frame_loop(w_cpu, 10_000, 1, args.cpu_budget)
        line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{done} frames of the same loop in {dt:.1f} s (oracle/tts_oracle.py)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""GPU perf probe: talker / code-predictor step time under engine settings, plus a per-phase trace.

    python scripts/perf_probe.py [--configs "500;800"] [--trace]
"""
import argparse
import ctypes
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "qwen-megakernel-tts_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from qwen_megakernel import model_tts  # noqa: E402
from qwen_megakernel.synthetic import synthetic_inputs, synthetic_tts_weights, weights_to  # noqa: E402


def time_steps(dec, x, n=50, warm=10):
    dec.reset()
    for i in range(8):
        dec.step_with_embed(x[i])
    dec._hidden.copy_(x[0])
    for _ in range(warm):
        dec._launch(-1, dec._hidden.data_ptr())
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        dec._launch(-1, dec._hidden.data_ptr())
    b.record()
    torch.cuda.synchronize()
    assert int(dec._out_token.item()) >= 0
    return a.elapsed_time(b) / n * 1e3


def time_cp_steps(cp, x, n=64, warm=16):
    cp.reset()
    for _ in range(warm // 16):
        cp.reset()
        for i in range(16):
            cp._step_with_embed(x[i])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for r in range(n // 16):
        cp.reset()
        for i in range(16):
            cp._step_with_embed(x[i], head=cp._heads[i % 15])
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def trace(dec, x, lib, engine, L):
    n_idx = L * 5 + 2
    stride = (n_idx + 1) * 24
    assert lib.qmk_engine_trace_enable(engine, stride) == 0
    dec.reset()
    for i in range(12):
        dec.step_with_embed(x[i % 8])
    G = lib.qmk_engine_num_ctas(engine)
    buf = (ctypes.c_longlong * (G * stride))()
    lib.qmk_engine_trace_read(engine, torch.cuda.current_stream().cuda_stream, buf, G * stride)
    lib.qmk_engine_trace_enable(engine, 0)
    raw = np.frombuffer(buf, dtype=np.int64).reshape(G, n_idx + 1, 24)
    t = raw[:, :, :8].astype(np.float64)
    fine = raw[:, :, 8:24].astype(np.int64)
    start = t[:, :, 0]
    d = np.diff(start, axis=1)
    names = ["qkv", "attn", "o", "gu", "down"]
    gl = ["window", "gather", "bar+aux", "norm", "stages", "reduce+bar", None]
    ga = ["window", "gather", "bar+aux", None, "stages", "reduce+bar", None]
    labels = {0: gl, 1: ["wait+gather+norm/rope", "scores", "merge", "combine+publish", None, None, None], 2: ga, 3: gl, 4: ga}
    layers = list(range(2, L))
    for grp_name, sl in (("attention CTAs 0-7", slice(0, 8)), ("other CTAs", slice(8, G))):
        print(f"--- {grp_name}: mean cycles per sub-step over layers 2..{L-1}")
        for ph in range(5):
            idxs = [l * 5 + ph for l in layers]
            tot = d[sl][:, idxs].mean()
            parts = []
            prev = 0
            for sub in range(1, 8):
                lab = labels[ph][sub - 1]
                if lab is None:
                    continue
                cur = t[sl][:, idxs, sub]
                if not np.any(cur):
                    continue
                ref = t[sl][:, idxs, prev]
                parts.append(f"{lab}={np.mean(cur - ref):.0f}")
                prev = sub
            last = t[sl][:, idxs, prev]
            nxt = start[sl][:, [i + 1 for i in idxs]]
            parts.append(f"tail={np.mean(nxt - last):.0f}")
            print(f"  {names[ph]:5s} total {tot:7.0f} : " + "  ".join(parts))
    # fine-grained stamps of thread 0 inside the GEMV phases (32-bit clock, wrap-safe deltas)
    fl = ["start->window", "issue loads", "shadow work", "data arrived", "sumsq+sts", "gather bar", "normalize+sts",
          "bfrag lds", "stage0 mma", "stage1 mma", "stage2 mma", "partial sts", "partial bar", "finalize+publish"]
    for grp_name, sl in (("attention CTAs 0-7", slice(0, 8)), ("other CTAs", slice(8, G))):
        print(f"--- {grp_name}: fine stamps (cycles between consecutive stamps; 0 = stamp not taken)")
        for ph in (0, 2, 3, 4):
            f = fine[sl][:, [l * 5 + ph for l in layers], :]
            parts = []
            prev = f[:, :, 0]
            for i in range(1, 15):
                cur = f[:, :, i]
                taken = cur != 0
                if not taken.any():
                    continue
                dlt = ((cur - prev) & 0xFFFFFFFF)[taken]
                parts.append(f"{fl[i-1]}={dlt.mean():.0f}")
                prev = np.where(taken, cur, prev)
            print(f"  {names[ph]:5s}: " + "  ".join(parts))
    # publish skew across CTAs (global timer, ns): per phase, spread of the publish time and who is last
    for ph in (0, 2, 3, 4):
        pub = t[:, [l * 5 + ph for l in layers], 7]                      # [G, layers] ns
        rel = pub - np.median(pub, axis=0, keepdims=True)
        last = np.argmax(pub, axis=0)
        cnt = np.bincount(last, minlength=G)
        worst = np.argsort(-cnt)[:5]
        print(f"  {names[ph]:5s} publish skew ns: p10={np.percentile(rel,10):6.0f} p50={np.percentile(rel,50):6.0f} "
              f"p90={np.percentile(rel,90):6.0f} max={rel.max():6.0f}; mean per-CTA offset min/max "
              f"{rel.mean(1).min():6.0f}/{rel.mean(1).max():6.0f} (cta {int(np.argmax(rel.mean(1)))}); most often last: "
              + ", ".join(f"cta{int(w)}x{int(cnt[w])}" for w in worst))
    print(f"  per-layer total (cta 0): {d[0, :L*5].sum() / L:9.0f} cycles;  kernel total cta0 {start[0, n_idx] - start[0, 0]:.0f} cycles")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="500", help="';'-separated initial poll delays (QMK_POLL_DELAY)")
    ap.add_argument("--trace", action="store_true")
    ap.add_argument("--layers", type=int, default=28)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    w = weights_to(synthetic_tts_weights(max_seq_len=256, num_layers=args.layers), "cuda")
    x = synthetic_inputs(99, 16).cuda()
    for cfg in args.configs.split(";"):
        dly = cfg.split(",")
        os.environ["QMK_POLL_DELAY"] = dly[0]
        if len(dly) > 1:
            os.environ["QMK_POLL_DELAY_O"] = dly[1]
        os.environ["QMK_COOP"] = dly[3] if len(dly) > 3 else "1"
        os.environ["QMK_O_SENTINEL"] = dly[4] if len(dly) > 4 else "0"
        os.environ["QMK_POLL_DELAY_ATTN"] = dly[5] if len(dly) > 5 else dly[0]
        os.environ["QMK_POLL_DELAY_OA"] = dly[6] if len(dly) > 6 else dly[0]
        model_tts._Native._engines.clear()
        dec = model_tts.TTSDecoder(weights=w, verbose=False, max_seq_len=256)
        cp = model_tts.CodePredictorKernel(w, device="cuda")
        us = time_steps(dec, x)
        us_cp = time_cp_steps(cp, x)
        print(f"poll_delay0={cfg:>5s}: talker {us:8.1f} us/step ({args.layers} layers)   cp step {us_cp:7.1f} us", flush=True)
        G = dec._lib.qmk_engine_num_ctas(dec._engine)
        st = (ctypes.c_int32 * (G * 24))()
        dec._lib.qmk_engine_poll_stats(dec._engine, torch.cuda.current_stream().cuda_stream, st, G * 24)
        a = np.frombuffer(st, dtype=np.int32).reshape(G, 3, 8)
        print("   repeated-poll gathers [qkv attn o gu down head argmax token]: attn CTAs", a[:8, 1].sum(0).tolist(),
              " others", a[8:, 1].sum(0).tolist())
        print("   mean weight-wait cycles per launch per CTA [qkv - o gu down head]:",
              (a[:, 2].mean(0) * 16 / 186).round(0).tolist(), flush=True)
        if args.trace:
            trace(dec, x, dec._lib, dec._engine, args.layers)
        del dec, cp


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""GPU probe: talker / code-predictor step time for engine versions and poll delays.

    python scripts/probe2.py --configs "1;2;2,400,200,200"     # engine[,QMK_POLL_DELAY[,ATTN[,DOWN]]]
"""
import argparse
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "qwen-megakernel-tts_b200"))

import torch  # noqa: E402

from qwen_megakernel import model_tts  # noqa: E402
from qwen_megakernel.synthetic import synthetic_inputs, synthetic_tts_weights, weights_to  # noqa: E402
from perf_probe import time_cp_steps, time_steps  # noqa: E402


def time_frames(cp, x, n=20):
    hid = x[0].float()
    emb = cp._talker_embed if hasattr(cp, "_talker_embed") else None
    return None


def time_steps_fixed(dec, x, pos=10, n=60, warm=10):
    """talker step repeated at ONE position (short context: no split attention)"""
    dec.reset()
    for i in range(8):
        dec.step_with_embed(x[i])
    dec._hidden.copy_(x[0])
    for _ in range(warm):
        dec._position = pos
        dec._launch(-1, dec._hidden.data_ptr())
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        dec._position = pos
        dec._launch(-1, dec._hidden.data_ptr())
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def time_steps_at(dec, x, pos, n=40, warm=5):
    """talker step time with the KV cache filled up to `pos` (the step itself is repeated at that position)"""
    dec.reset()
    for i in range(8):
        dec.step_with_embed(x[i])
    dec._position = pos
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


TRACE_POS = -1


def trace2(dec, x, L):
    """per-phase critical-path stamps of thread 0 of every CTA (group kernel)"""
    import ctypes
    import numpy as np
    lib, engine = dec._lib, dec._engine
    n_idx = L * 5 + 2
    stride = (n_idx + 1) * 24
    assert lib.qmk_engine_trace_enable(engine, stride) == 0
    dec.reset()
    for i in range(12):
        dec.step_with_embed(x[i % 8])
    if TRACE_POS >= 0:                      # the traced launch is the last one: put it at the requested depth
        dec._position = TRACE_POS
        dec.step_with_embed(x[0])
    G = lib.qmk_engine_num_ctas(engine)
    buf = (ctypes.c_longlong * (G * stride))()
    lib.qmk_engine_trace_read(engine, torch.cuda.current_stream().cuda_stream, buf, G * stride)
    lib.qmk_engine_trace_enable(engine, 0)
    raw = np.frombuffer(buf, dtype=np.int64).reshape(G, n_idx + 1, 24).astype(np.float64)
    names = ["qkv", "attn", "o", "gu", "down"]
    subs = [0, 1, 2, 3, 4, 5, 6, 8]
    lab = {0: "start", 1: "window", 2: "data", 3: "normbar", 4: "vec ready", 5: "mma", 6: "partial bar", 8: "publish"}
    laba = {0: "start", 1: "qkv gathered+rope", 2: "scores", 3: "merged"}
    layers = list(range(2, L))
    tot_layer = 0.0
    for ph in range(5):
        idxs = [l * 5 + ph for l in layers]
        t = raw[:, idxs, :]
        nxt = raw[:, [i + 1 for i in idxs], 0]
        parts = []
        prev = t[:, :, 0]
        for sname in subs[1:]:
            cur = t[:, :, sname]
            if not np.any(cur):
                continue
            parts.append(f"{(laba if ph == 1 else lab).get(sname, sname)}={np.mean(cur - prev):.0f}")
            prev = cur
        parts.append(f"tail={np.mean(nxt - prev):.0f}")
        tot = np.mean(nxt - t[:, :, 0])
        tot_layer += tot
        print(f"  {names[ph]:5s} total {tot:7.0f} : " + "  ".join(parts))
    # skew: per exchange, the wait (phase start -> data complete) of every CTA; the CTAs with the SHORTEST wait published last
    smid = raw[:, n_idx, 1].astype(np.int64) if raw.shape[1] > n_idx else None
    for ph, sub_done, nm in ((0, 3, "qkv-in (A)"), (1, 1, "attn-in (q/k/v)"), (3, 3, "gu-in (B)"), (4, 4, "down-in (m)")):
        idxs = [l * 5 + ph for l in layers]
        wait = raw[:, idxs, sub_done] - raw[:, idxs, 0]          # [G, layers]
        per_cta = wait.mean(1)
        order = np.argsort(per_cta)
        print(f"  {nm:16s} wait over CTAs: min {per_cta.min():6.0f} p10 {np.percentile(per_cta,10):6.0f} median {np.median(per_cta):6.0f} "
              f"p90 {np.percentile(per_cta,90):6.0f} max {per_cta.max():6.0f}; shortest (= last to publish): "
              + " ".join(f"{int(cc)}({per_cta[cc]:.0f})" for cc in order[:8]))
        grp = per_cta.reshape(8, 16).mean(1)
        print("      per-group mean wait: " + " ".join(f"{v:.0f}" for v in grp))
        if smid is not None and ph == 4:
            print("      smid of CTA 0..127: " + " ".join(str(int(v)) for v in smid))
    print(f"  per layer {tot_layer:.0f} cycles; kernel (cta 0) {raw[0, n_idx, 0] - raw[0, 0, 0]:.0f} cycles")


def trace_cp(cp, w, x):
    """stamps of the LAST step of a fused code-predictor frame (the trace slots are overwritten step by step)"""
    import ctypes
    import numpy as np
    lib, engine = cp._lib, cp._engine
    L = 5
    n_idx = L * 5 + 2
    stride = (n_idx + 1) * 24
    assert lib.qmk_engine_trace_enable(engine, stride) == 0
    hid = x[0].float()
    for sample in (False, True):
        cp.predict(hid, 1335, w["embed_weight"], do_sample=sample, temperature=0.9, top_k=50)
        G = lib.qmk_engine_num_ctas(engine)
        buf = (ctypes.c_longlong * (G * stride))()
        lib.qmk_engine_trace_read(engine, torch.cuda.current_stream().cuda_stream, buf, G * stride)
        raw = np.frombuffer(buf, dtype=np.int64).reshape(G, n_idx + 1, 24).astype(np.float64)
        st = raw[:, :, 0]
        layer = (st[:, 25] - st[:, 5]).mean() / 4
        print(f"  cp last step (sample={sample}): layer {layer:.0f} cycles; head phase {np.mean(st[:, 26] - st[:, 25]):.0f} "
              f"(cta0 {st[0, 26] - st[0, 25]:.0f}); argmax phase cta0 {st[0, 27] - st[0, 26]:.0f}; "
              f"head stamps cta0: " + " ".join(f"{raw[0, 25, s] - raw[0, 25, 0]:.0f}" for s in (1, 2, 3, 4, 5, 6, 8)))
    lib.qmk_engine_trace_enable(engine, 0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trace", action="store_true")
    ap.add_argument("--configs", default="1;2")
    ap.add_argument("--layers", type=int, default=28)
    ap.add_argument("--positions", default="300")
    ap.add_argument("--trace-pos", type=int, default=-1)
    args = ap.parse_args()
    global TRACE_POS
    TRACE_POS = args.trace_pos
    torch.cuda.set_device(0)
    w = weights_to(synthetic_tts_weights(max_seq_len=2048, num_layers=args.layers), "cuda")
    x = synthetic_inputs(99, 16).cuda()
    for cfg in args.configs.split(";"):
        f = cfg.split(",")
        os.environ["QMK_ENGINE"] = f[0]
        for name, i in (("QMK_POLL_DELAY", 1), ("QMK_POLL_DELAY_ATTN", 2), ("QMK_POLL_DELAY_DOWN", 3), ("QMK_POLL_DELAY_TOKEN", 4)):
            if len(f) > i and f[i] != "":
                os.environ[name] = f[i]
            else:
                os.environ.pop(name, None)
        model_tts._Native._engines.clear()
        dec = model_tts.TTSDecoder(weights=w, verbose=False, max_seq_len=2048)
        cp = model_tts.CodePredictorKernel(w, device="cuda")
        G = dec._lib.qmk_engine_num_ctas(dec._engine)
        us = time_steps(dec, x)
        us_cp = time_cp_steps(cp, x)
        extra = f"  fixed pos10: {time_steps_fixed(dec, x):7.1f}"
        for pos in [int(v) for v in args.positions.split(",") if v]:
            try:
                extra += f"  pos{pos}: {time_steps_at(dec, x, pos):7.1f}"
            except Exception as ex:  # noqa: BLE001
                extra += f"  pos{pos}: {type(ex).__name__}"
        if args.trace and f[0] == "2":
            trace2(dec, x, args.layers)
            trace_cp(cp, w, x)
        print(f"cfg={cfg:>16s} ctas={G}: talker {us:8.1f} us/step   cp step {us_cp:7.1f} us {extra}", flush=True)
        del dec, cp


if __name__ == "__main__":
    main()

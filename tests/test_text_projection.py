"""Text side of the prefill: upstream ``TextProjection.embed_text_ids`` (model_tts.py:361-374) and
``build_prefill_embeddings`` (:776-864).

CPU: the oracle restatement and the package's PyTorch ``TextProjection`` against the fixture recorded from the upstream
classes (tests/golden/make_golden_text.py); argument validation of the C entry points.
GPU: ``TextProjectionKernel`` (qmk_text_proj_embed: gather -> tcgen05 fc1 -> bias + SiLU -> tcgen05 fc2 -> bias) against
the fixture, the oracle and PyTorch on the same device, over empty / ragged / multi-pass inputs.

Tolerance (floating point, stated here): every operator rounds to bf16 once, so two correct evaluations that differ in
fp32 summation order differ by at most one bf16 ulp per element after fc1 and by a few ulps of the largest output after
fc2: max|impl - ref| <= 2e-2 max|ref| (the repo-wide bf16 rule, tests/parity.py), mean|impl - ref| <= 2e-3 max|ref|,
cosine > 0.9999 per row.
"""

import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, bf16_from_bits

TEXT_MAX_REL, TEXT_MEAN_REL, TEXT_COS = 2e-2, 2e-3, 0.9999


def assert_text_parity(name, impl, ref):
    impl, ref = impl.float().cpu(), ref.float().cpu()
    assert impl.shape == ref.shape, f"{name}: shape {tuple(impl.shape)} vs {tuple(ref.shape)}"
    if ref.numel() == 0:
        return
    scale = float(ref.abs().max())
    d = (impl - ref).abs()
    cos = torch.nn.functional.cosine_similarity(impl.reshape(-1, impl.shape[-1]), ref.reshape(-1, ref.shape[-1]), dim=-1)
    exact = float((d == 0).float().mean())
    print(f"[{name}] max_rel={float(d.max()) / scale:.5f} mean_rel={float(d.mean()) / scale:.6f} min_cos={float(cos.min()):.6f} "
          f"bit_exact={exact:.3f}")
    assert float(d.max()) <= TEXT_MAX_REL * scale, f"{name}: max rel err {float(d.max()) / scale}"
    assert float(d.mean()) <= TEXT_MEAN_REL * scale, f"{name}: mean rel err {float(d.mean()) / scale}"
    assert float(cos.min()) > TEXT_COS, f"{name}: cosine {float(cos.min())}"


@pytest.fixture(scope="module")
def text_golden():
    return np.load(os.path.join(GOLDEN, "text_projection.npz"))


@pytest.fixture(scope="module")
def text_weights(text_golden):
    from qwen_megakernel.synthetic import synthetic_tts_weights
    torch.set_num_threads(os.cpu_count() or 1)
    return synthetic_tts_weights(seed=1234, num_layers=1, max_seq_len=64, text_vocab=int(text_golden["text_vocab"]))


def _cached(tp, special):
    sp = tp.embed_text_ids(special)
    return sp[0:1], sp[1:2], sp[2:3]


# ---- CPU ---------------------------------------------------------------------------------------------------------------

def test_text_oracle_matches_upstream_fixture(text_golden, text_weights):
    from oracle.tts_oracle import TextProjectionOracle, build_prefill_oracle
    orc = TextProjectionOracle(text_weights)
    ids = torch.from_numpy(text_golden["ids"])
    assert_text_parity("oracle-vs-upstream embed_text_ids", orc.embed_text_ids(ids), bf16_from_bits(text_golden["out"]))
    pad, bos, eos = _cached(orc, torch.from_numpy(text_golden["special"]))
    prefill, trailing = build_prefill_oracle(torch.from_numpy(text_golden["utter"]), orc, text_weights["embed_weight"], pad, bos, eos)
    assert_text_parity("oracle-vs-upstream prefill", prefill, bf16_from_bits(text_golden["prefill"]))
    assert_text_parity("oracle-vs-upstream trailing", trailing, bf16_from_bits(text_golden["trailing"]))


def test_pytorch_text_projection_matches_upstream_fixture(text_golden, text_weights):
    """The package's PyTorch ``TextProjection`` / ``build_prefill_embeddings`` (the reference glue the engine imports) evaluate
    the upstream operators in the upstream order: same bits as the fixture on the torch build that recorded it, within the
    tolerance anywhere else (another CPU may pick another bf16 GEMM kernel)."""
    from qwen_megakernel import model_tts as m
    tp = m.TextProjection(text_weights, device="cpu")
    out = tp.embed_text_ids(torch.from_numpy(text_golden["ids"]))
    ref = bf16_from_bits(text_golden["out"])
    assert out.dtype == torch.bfloat16 and out.shape == (150, 1024)
    if str(text_golden["torch_version"]) == torch.__version__ and torch.equal(out, ref):
        print("bit-exact against the fixture")
    assert_text_parity("TextProjection-vs-upstream", out, ref)
    pad, bos, eos = _cached(tp, torch.from_numpy(text_golden["special"]))
    prefill, trailing = m.build_prefill_embeddings(torch.from_numpy(text_golden["utter"]), tp, text_weights["embed_weight"],
                                                   device="cpu", cached_tts_embeds={"pad": pad, "bos": bos, "eos": eos})
    assert prefill.shape == (8, 1024) and trailing.shape == (12, 1024)
    assert_text_parity("build_prefill_embeddings prefill", prefill, bf16_from_bits(text_golden["prefill"]))
    assert_text_parity("build_prefill_embeddings trailing", trailing, bf16_from_bits(text_golden["trailing"]))


def test_text_proj_entry_points_validate_arguments():
    """Null / malformed arguments are refused before any CUDA call (no GPU here)."""
    from qwen_megakernel import build_tts
    lib = build_tts.load_library(build_tts.build())
    h = ctypes.c_void_p()
    assert lib.qmk_text_proj_create(0, None, 16, None, None, None, None, ctypes.byref(h)) == -1      # QMK_ERR_ARG
    assert b"null argument" in lib.qmk_batched_last_error()
    assert lib.qmk_text_proj_create(0, 16, 0, 16, 16, 16, 16, ctypes.byref(h)) == -1                 # vocab_rows < 1
    assert lib.qmk_text_proj_create(0, 16, 16, 24, 16, 16, 16, ctypes.byref(h)) == -1                # misaligned tensor
    assert b"aligned" in lib.qmk_batched_last_error()
    assert not h.value
    assert lib.qmk_text_proj_embed(None, None, 4, None, None) == -1
    lib.qmk_text_proj_destroy(None)                                                                 # no-op


def test_text_projection_kernel_has_no_cpu_fallback(text_weights):
    from qwen_megakernel import model_tts as m
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.TextProjectionKernel(text_weights, device="cuda")


# ---- GPU ---------------------------------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def text_kernel(text_weights):
    from qwen_megakernel import model_tts as m
    wg = {k: v.cuda() for k, v in text_weights.items() if k.startswith("text_")}
    return m.TextProjectionKernel(wg, device="cuda"), wg


@pytest.mark.gpu
def test_text_kernel_vs_upstream_fixture_and_oracle(text_golden, text_weights, text_kernel):
    """150 ids = one pass of three 64-token blocks along gridDim.z (the last one holds 22 tokens and 42 zero rows)."""
    from oracle.tts_oracle import TextProjectionOracle
    tp, _ = text_kernel
    ids = torch.from_numpy(text_golden["ids"])
    out = tp.embed_text_ids(ids.cuda())
    assert out.dtype == torch.bfloat16 and out.shape == (150, 1024) and out.is_cuda
    assert_text_parity("kernel-vs-upstream fixture", out, bf16_from_bits(text_golden["out"]))
    assert_text_parity("kernel-vs-oracle", out, TextProjectionOracle(text_weights).embed_text_ids(ids))
    assert torch.equal(out, tp.embed_text_ids(ids.cuda())), "the chain is deterministic (fixed split-K summation order)"


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 5, 16, 17, 48, 64, 65, 128, 200, 512, 513, 1100])
def test_text_kernel_ragged_lengths_vs_pytorch_on_device(text_kernel, n):
    """Empty, single, partial-tile, exact-tile, multi-block and multi-pass (> 512) inputs against upstream's operators on the same GPU; rows must
    not depend on what else is in the call (row i of a long call == the same id alone)."""
    from qwen_megakernel import model_tts as m
    tp, wg = text_kernel
    gen = torch.Generator().manual_seed(100 + n)
    ids = torch.randint(0, wg["text_embedding"].shape[0], (n,), generator=gen).cuda()
    out = tp.embed_text_ids(ids)
    assert out.shape == (n, 1024)
    assert_text_parity(f"kernel-vs-torch n={n}", out, m.TextProjection(wg, device="cuda").embed_text_ids(ids))
    if n >= 5:      # a row does not depend on its lane or on the other rows of the call (same K order per column; other UMMA N)
        alone = torch.cat([tp.embed_text_ids(ids[n - 1:]), tp.embed_text_ids(ids[3:4])])
        print(f"n={n}: rows alone bit-equal to rows in the call: {torch.equal(alone, out[[n - 1, 3]])}")
        assert_text_parity(f"row alone vs in a call of {n}", alone, out[[n - 1, 3]])


@pytest.mark.gpu
def test_text_kernel_shapes_dtypes_and_clamping(text_kernel):
    from qwen_megakernel import model_tts as m
    tp, wg = text_kernel
    rows = wg["text_embedding"].shape[0]
    ids2d = torch.arange(14, dtype=torch.int32).view(2, 7)                 # [batch, seq_len], int32, host tensor
    out = tp.embed_text_ids(ids2d)
    assert out.shape == (2, 7, 1024)
    assert torch.equal(out.view(14, 1024), tp.embed_text_ids(torch.arange(14).cuda()))
    bad = torch.tensor([-5, rows + 100, 2]).cuda()                         # outside the table: clamped, never out of bounds
    assert torch.equal(tp.embed_text_ids(bad), tp.embed_text_ids(torch.tensor([0, rows - 1, 2]).cuda()))
    with pytest.raises(ValueError):
        tp.embed_text_ids(torch.zeros(3))
    with pytest.raises(ValueError):
        m.TextProjectionKernel({**wg, "text_proj_fc1_w": wg["text_proj_fc1_w"].float()}, device="cuda")
    ids14 = torch.arange(14).cuda()
    big = torch.arange(1100).cuda() % rows
    torch.cuda.synchronize()
    side = torch.cuda.Stream()                     # calls alternating between two streams share the handle's staging buffers:
    outs = []                                      # the library orders them (no wait_stream here)
    for i in range(6):
        if i % 2:
            with torch.cuda.stream(side):
                outs.append(tp.embed_text_ids(ids14))
        else:
            tp.embed_text_ids(big)                 # a long call on the default stream right before the side-stream call
    torch.cuda.synchronize()
    for o2 in outs:
        assert torch.equal(o2, out.view(14, 1024))


@pytest.mark.gpu
def test_build_prefill_embeddings_on_the_kernel(text_golden, text_weights, text_kernel):
    """upstream build_prefill_embeddings (model_tts.py:776-864) driven with the native projection, both ways the engine
    can call it (cached pad / bos / eos; special ids looked up in the same call), vs the upstream fixture."""
    from qwen_megakernel import model_tts as m
    tp, wg = text_kernel
    emb = text_weights["embed_weight"].cuda()
    special = torch.from_numpy(text_golden["special"]).cuda()
    pad, bos, eos = _cached(tp, special)
    prefill, trailing = m.build_prefill_embeddings(torch.from_numpy(text_golden["utter"]), tp, emb, device="cuda",
                                                   cached_tts_embeds={"pad": pad, "bos": bos, "eos": eos})
    assert_text_parity("prefill on the kernel", prefill, bf16_from_bits(text_golden["prefill"]))
    assert_text_parity("trailing on the kernel", trailing, bf16_from_bits(text_golden["trailing"]))


@pytest.mark.gpu
def test_text_kernel_full_size_table_and_graph_capture():
    """The real checkpoint's table is [151936, 2048] (622 MB): TTS_PAD / TTS_BOS / TTS_EOS (151671-151673) are looked up in it
    (tts_engine.py:107-108).  Also: the entry only enqueues kernels, so a call can be captured and replayed."""
    from qwen_megakernel import model_tts as m
    from qwen_megakernel.synthetic import synthetic_tts_weights
    w = synthetic_tts_weights(seed=7, num_layers=1, max_seq_len=64, include_talker=False, include_code_predictor=False)
    wg = {k: v.cuda() for k, v in w.items() if k.startswith("text_proj")}
    gen = torch.Generator(device="cuda").manual_seed(3)
    wg["text_embedding"] = torch.empty(151936, 2048, dtype=torch.bfloat16, device="cuda").normal_(generator=gen)
    tp = m.TextProjectionKernel(wg, device="cuda")
    ids = torch.tensor([m.TTS_PAD, m.TTS_BOS, m.TTS_EOS, 0, 151935, 77777], device="cuda")
    out = tp.embed_text_ids(ids)
    assert_text_parity("full-size table", out, m.TextProjection(wg, device="cuda").embed_text_ids(ids))
    static_ids, static_out = ids.clone(), torch.empty_like(out)
    lib, h = tp._lib, tp._handle
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            assert lib.qmk_text_proj_embed(h, static_ids.data_ptr(), 6, static_out.data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
        static_ids.copy_(ids.flip(0))
        g.replay()
    s.synchronize()
    print("graph replay bit-equal to the plain call:", torch.equal(static_out, out.flip(0)))
    assert_text_parity("graph replay vs plain call", static_out, out.flip(0))


# ---- CPU: the chain's index arithmetic restated -----------------------------------------------------------------------

@pytest.mark.parametrize("n_ids", [1, 16, 17, 64, 65, 150, 512, 513, 1100])
def test_text_chain_partial_layout_restated(n_ids):
    """csrc/qmk_text.cuh + qmk_bgemm.cuh in numpy index arithmetic: a pass covers <= 512 tokens as nb blocks of N lanes (one
    block: N = tokens rounded up to 16; several: N = 64); the GEMM CTA (row tile, K slice y, block z) writes
    partial[((z * splits + y) * N + n) * M + row]; the epilogue of token t reads slices s = 0 .. splits - 1 at
    ((t // N) * splits * N + t % N) * M + row + s * N * M.  Every token must read exactly the words its own block wrote, every
    word of a pass is written once, and the passes tile the call."""
    LANES, BLOCKS, M, SPLITS = 64, 8, 2048, 8
    covered = 0
    for off in range(0, n_ids, BLOCKS * LANES):
        nv = min(BLOCKS * LANES, n_ids - off)
        nb = (nv + LANES - 1) // LANES
        N = LANES if nb > 1 else (nv + 15) & ~15
        assert N % 16 == 0 and 16 <= N <= 64 and nb * N >= nv and (nb - 1) * N < nv
        owner = {}                                     # word -> (block, slice, lane) for one row of the tile
        for z in range(nb):
            for y in range(SPLITS):
                for n in range(N):
                    w = ((z * SPLITS + y) * N + n) * M
                    assert w not in owner
                    owner[w] = (z, y, n)
        assert max(owner) + M <= BLOCKS * LANES * SPLITS * M        # inside the 8-block partial buffer
        for t in range(nv):
            base = ((t // N) * SPLITS * N + t % N) * M
            assert [owner[base + s * N * M] for s in range(SPLITS)] == [(t // N, s, t % N) for s in range(SPLITS)]
        covered += nv
    assert covered == n_ids

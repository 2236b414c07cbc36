"""Both kernel generations stay covered: the default engine is the group kernel (qmk_device2.cuh, 128 CTAs); the row-split
kernel (qmk_device.cuh, one CTA per SM, QMK_ENGINE=1) keeps the staged debugging mode.  The parity suite runs in this
process on the default engine and once more in a child process on the other one."""

import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_default_engine_is_the_group_kernel(gpu_weights):
    from qwen_megakernel.model_tts import TTSDecoder
    if os.environ.get("QMK_ENGINE", "2") != "2":
        pytest.skip("QMK_ENGINE overrides the default")
    dec = TTSDecoder(weights=gpu_weights, verbose=False, max_seq_len=64, num_layers=2)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    assert dec._lib.qmk_engine_num_ctas(dec._engine) == (128 if sms >= 128 else sms)


def test_cumulative_accumulators_survive_many_launches(gpu_weights):
    """The group kernel never zeroes its L2 accumulators: totals are cumulative and the 8-bit arrival counts wrap every
    32 uses.  Two identical 70-step runs separated by other launches must agree bit for bit."""
    from qwen_megakernel.model_tts import TTSDecoder
    from qwen_megakernel.synthetic import synthetic_inputs
    x = synthetic_inputs(7, 8).cuda()
    dec = TTSDecoder(weights=gpu_weights, verbose=False, max_seq_len=128, num_layers=3)
    runs = []
    for rep in range(2):
        dec.reset()
        out = []
        for i in range(70 + rep * 0):
            t, h = dec.step_with_embed(x[i % 8]) if i < 8 else dec.step(out[-1][0] % 3072)
            out.append((t, h.cpu()))
        runs.append(out)
        for i in range(5 + rep):      # shift the accumulator phase between the two runs
            dec.step(1)
    assert [t for t, _ in runs[0]] == [t for t, _ in runs[1]]
    assert all(torch.equal(a, b) for (_, a), (_, b) in zip(runs[0], runs[1]))


def test_row_split_engine_parity_suite_in_a_child_process():
    """QMK_ENGINE=1: golden / oracle parity, the staged-vs-fused bit identity and the code-predictor frame."""
    env = dict(os.environ, QMK_ENGINE="1")
    sel = "golden or staged or deterministic or teacher_forced or fused_equals_stepwise or long_context"
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(REPO, "tests", "test_gpu_parity.py"), "-m", "gpu", "-x", "-q",
                          "-k", sel, "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, timeout=900)
    tail = (res.stdout + res.stderr)[-2000:]
    assert res.returncode == 0, tail
    assert " passed" in res.stdout, tail

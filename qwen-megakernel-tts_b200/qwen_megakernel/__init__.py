"""qwen_megakernel — B200 (sm_100a) decode engine behind the upstream qwen_megakernel Python surface.

Drop-in for the model layer of jayanth-kumar-morem/qwen-megakernel-tts:

    from qwen_megakernel.model_tts import TTSDecoder, CodePredictorKernel, TextProjectionKernel, load_tts_weights
    from qwen_megakernel.build_tts import get_extension     # registers torch.ops.qwen_megakernel_C.decode

Importing the package neither builds nor loads native code (as upstream, qwen_megakernel/__init__.py:9);
the CUDA library is loaded on first use and a missing library is an error, never a fallback.
"""

__all__ = []

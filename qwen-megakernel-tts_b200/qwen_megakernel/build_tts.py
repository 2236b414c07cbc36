"""Build / load the native engine and register ``torch.ops.qwen_megakernel_C.decode``.

Replaces upstream qwen_megakernel/build_tts.py (JIT ``cpp_extension.load`` with ``-arch=sm_120a``,
build_tts.py:45-71).  Here the kernel is compiled ahead of time for sm_100a with plain nvcc into an
in-tree shared library (``libqmk_b200.so``, the C ABI of include/qmk_b200.h) so that it travels with the
source tree; ``get_extension()`` compiles it only if the library is missing and nvcc is available.

There is deliberately NO CPU or PyTorch fallback: if the library cannot be loaded, every entry point
raises.
"""

from __future__ import annotations

import ctypes
import glob
import hashlib
import os
import shutil
import subprocess
import threading
import weakref

import torch

_DIR = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_DIR)
_CSRC = os.path.join(_ROOT, "csrc")
_INCLUDE = os.path.join(os.path.dirname(_ROOT), "include")
# QMK_LIB_PATH: load another build of the same ABI (A/B timing of kernel variants inside one GPU job)
LIB_PATH = os.environ.get("QMK_LIB_PATH") or os.path.join(_DIR, "libqmk_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr", "--extended-lambda",
    "-Xcompiler", "-fPIC", "-shared",
]

_lock = threading.Lock()
_lib = None
_op_lib = None


def _nvcc() -> str | None:
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if os.path.exists(cand) else shutil.which("nvcc")


def source_files() -> list[str]:
    """Every file the library is compiled from: csrc/*.cu, csrc/*.cuh, include/*.h."""
    return sorted(glob.glob(os.path.join(_CSRC, "*.cu")) + glob.glob(os.path.join(_CSRC, "*.cuh")) +
                  glob.glob(os.path.join(_INCLUDE, "*.h")))


def source_hash() -> str:
    """Content hash of the sources (and the compiler flags); compiled into the library as qmk_source_hash()."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS + os.environ.get("QMK_NVCC_EXTRA", "").split()).encode())
    for path in source_files():
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def library_hash(path: str = LIB_PATH) -> str | None:
    """The source hash a built library carries, read from the file WITHOUT loading it (a dlopen here would pin the old
    image in the process and a rebuild would then be invisible); None if the file predates the hash."""
    import re
    try:
        with open(path, "rb") as f:
            m = re.search(rb"QMK_SRC_HASH:([0-9a-f]{16})", f.read())
    except OSError:
        return None
    return m.group(1).decode() if m else None


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu -> libqmk_b200.so (cross-compiles without a GPU).

    Staleness is decided by CONTENT: the library carries the hash of the sources it was compiled from and is rebuilt
    whenever that differs from the tree (an edited header can no longer leave a stale binary behind)."""
    srcs = [os.path.join(_CSRC, "qmk_engine.cu"), os.path.join(_CSRC, "qmk_batched.cu")]
    if os.environ.get("QMK_LIB_PATH") and os.path.exists(LIB_PATH) and not force:
        return LIB_PATH      # an explicitly named build (A/B timing of kernel variants) is loaded as it is
    want = source_hash()
    if not force and os.path.exists(LIB_PATH) and library_hash(LIB_PATH) == want:
        return LIB_PATH
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("qwen_megakernel: libqmk_b200.so is missing/stale and nvcc was not found")
    # Several ranks of one job may find the library stale at the same moment: one of them compiles (exclusive lock), into a
    # temporary name that replaces the library atomically, the others wait for the lock and then find it current.
    import fcntl
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and os.path.exists(LIB_PATH) and library_hash(LIB_PATH) == want:
                return LIB_PATH
            tmp = f"{LIB_PATH}.{os.getpid()}.tmp"
            cmd = [nvcc] + NVCC_FLAGS + os.environ.get("QMK_NVCC_EXTRA", "").split() + ["-Xptxas", "-v"] * int(verbose) + [f'-DQMK_SRC_HASH="{want}"', f"-I{_INCLUDE}", f"-I{_CSRC}", "-o", tmp] + srcs
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
            os.replace(tmp, LIB_PATH)
            if verbose:
                print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


_vp, _i32, _f32, _u64, _i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_uint64, ctypes.c_int64

# symbol -> (restype, argtypes); must list every function declared in include/qmk_b200.h
SIGNATURES = {
    "qmk_abi_version": (_i32, []),
    "qmk_last_error": (ctypes.c_char_p, []),
    "qmk_source_hash": (ctypes.c_char_p, []),
    "qmk_engine_create": (_i32, [_i32, _i32, ctypes.POINTER(_vp)]),
    "qmk_engine_destroy": (None, [_vp]),
    "qmk_engine_num_ctas": (_i32, [_vp]),
    "qmk_engine_sync_status": (_i32, [_vp, _vp, ctypes.POINTER(ctypes.c_int32)]),
    "qmk_engine_trace_enable": (_i32, [_vp, _i32]),
    "qmk_engine_trace_read": (_i32, [_vp, _vp, ctypes.POINTER(ctypes.c_longlong), _i64]),
    "qmk_engine_poll_stats": (_i32, [_vp, _vp, ctypes.POINTER(ctypes.c_int32), _i64]),
    "qmk_model_create": (_i32, [_vp, _vp, _i32, _vp, _i32, _vp, ctypes.POINTER(_vp)]),
    "qmk_model_add_head": (_i32, [_vp, _vp, _i32, _vp]),
    "qmk_model_set_group_embedding": (_i32, [_vp, _i32, _vp]),
    "qmk_model_set_mrope": (_i32, [_vp, ctypes.POINTER(ctypes.c_int32), _i32]),
    "qmk_model_destroy": (None, [_vp]),
    "qmk_model_packed_bytes": (_i64, [_vp]),
    "qmk_decode_step": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _i32, _vp]),
    "qmk_decode_step_mrope": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32,
                                     ctypes.POINTER(ctypes.c_int32), _i32, _f32, _vp]),
    "qmk_decode_step_codes": (_i32, [_vp, _i32, _vp, _vp, _i32, ctypes.POINTER(_vp), _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                     _i32, _i32, _f32, _vp]),
    "qmk_generate_nosync": (_i32, [_vp, _vp]),
    "qmk_generate_args_size": (_i32, []),
    "qmk_cp_predict": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _i32, _u64, _u64, _vp,
                              _vp, _vp, _vp, _vp]),
    "qmk_cp_predict_dev": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _i32, _u64, _u64, _vp, _vp]),
    "qmk_batched_create": (_i32, [_i32, _vp, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _i32, ctypes.POINTER(_vp)]),
    "qmk_batched_destroy": (None, [_vp]),
    "qmk_batched_step": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "qmk_batched_last_error": (ctypes.c_char_p, []),
    "qmk_batched_add_head": (_i32, [_vp, _vp, _i32]),
    "qmk_batched_step_ex": (_i32, [_vp, _vp, _vp]),
    "qmk_batched_embed_sum": (_i32, [_i32, _vp, _vp, _i32, ctypes.POINTER(_vp), _i32, _vp, _i32, _vp, _vp]),
    "qmk_batched_is_persistent": (_i32, [_vp]),
    "qmk_batched_trace_read": (_i32, [_vp, _vp, ctypes.POINTER(ctypes.c_longlong), _i32]),
    "qmk_batched_counter_add": (_i32, [_vp, ctypes.c_uint64, _vp]),
    "qmk_batched_chain_trace": (_i32, [_i32, _vp, ctypes.POINTER(ctypes.c_ulonglong), _i32]),
    "qmk_batched_sync_status": (_i32, [_vp, _vp]),
    "qmk_batched_prefill": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "qmk_text_proj_create": (_i32, [_i32, _vp, _i32, _vp, _vp, _vp, _vp, ctypes.POINTER(_vp)]),
    "qmk_text_proj_embed": (_i32, [_vp, _vp, _i32, _vp, _vp]),
    "qmk_text_proj_destroy": (None, [_vp]),
    "launch_ldg_decode_direct": (None, [_i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                        _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _vp]),
    "qmk_legacy_configure": (_i32, [_vp, _i32, _i32]),
    "qmk_legacy_status": (_i32, []),
    "qmk_legacy_sync_status": (_i32, [_vp]),
    "qmk_legacy_invalidate": (None, [_vp]),
    "qmk_legacy_check_blob": (_i32, [_vp, _i32, _vp]),
    "qmk_legacy_release": (None, []),
}


def load_library(path: str = LIB_PATH) -> ctypes.CDLL:
    """dlopen the C-ABI library and type its entry points.  No GPU work happens here."""
    if not os.path.exists(path):
        raise RuntimeError(f"qwen_megakernel: native library {path} not found (run build_tts.build())")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


class NativeError(RuntimeError):
    pass


class BatchedStepArgs(ctypes.Structure):
    """Mirror of ``qmk_batched_step_args`` (include/qmk_b200.h)."""
    _fields_ = [
        ("token_ids", _vp), ("token_table", _vp), ("table_rows", ctypes.c_int32), ("embeds_bf16", _vp), ("embeds_f32", _vp),
        ("positions", _vp), ("k_cache", _vp), ("v_cache", _vp), ("hidden_out", _vp), ("head", ctypes.c_int32),
        ("do_sample", ctypes.c_int32), ("top_k", ctypes.c_int32), ("group", ctypes.c_int32), ("temperature", ctypes.c_float),
        ("seed", ctypes.c_uint64), ("counter", ctypes.c_uint64), ("tokens_out", _vp), ("codes_out", _vp),
        ("codes_stride", ctypes.c_int32), ("codes_col", ctypes.c_int32), ("counter_ptr", _vp), ("depth_hint", ctypes.c_int32),
    ]


class GenerateArgs(ctypes.Structure):
    """Mirror of ``qmk_generate_args`` (include/qmk_b200.h)."""
    _fields_ = [
        ("talker", _vp), ("talker_head", ctypes.c_int32), ("talker_vocab", ctypes.c_int32),
        ("talker_embed_weight", _vp), ("talker_cos", _vp), ("talker_sin", _vp), ("talker_k_cache", _vp),
        ("talker_v_cache", _vp), ("talker_max_seq", ctypes.c_int32), ("position", ctypes.c_int32),
        ("rope_pos", ctypes.POINTER(ctypes.c_int32)), ("hidden_buffer", _vp), ("talker_hidden", _vp),
        ("talker_token", _vp), ("cp", _vp), ("cp_cos", _vp), ("cp_sin", _vp), ("cp_k_cache", _vp), ("cp_v_cache", _vp),
        ("cp_max_seq", ctypes.c_int32), ("cp_vocab", ctypes.c_int32), ("group_embedding_tables", ctypes.POINTER(_vp)),
        ("n_frames", ctypes.c_int32), ("eos_token", ctypes.c_int32), ("trailing_text", _vp),
        ("n_trailing", ctypes.c_int32), ("trailing_offset", ctypes.c_int32), ("pad_embed", _vp),
        ("do_sample", ctypes.c_int32), ("top_k", ctypes.c_int32), ("temperature", ctypes.c_float),
        ("reset_state", ctypes.c_int32), ("seed", ctypes.c_uint64), ("frame_counter", ctypes.c_uint64),
        ("codes_out", _vp), ("tokens_out", _vp), ("gen_state", _vp),
    ]


def check(lib, rc: int, what: str):
    if rc < 0:
        raise NativeError(f"{what}: {lib.qmk_last_error().decode()} (code {rc})")
    return rc


DECODE_SCHEMA = (
    "decode(Tensor output_token, int input_token_id, "
    "Tensor embed_weight, Tensor layer_weights_packed, "
    "Tensor final_norm_weight, Tensor lm_head_weight, "
    "Tensor cos_table, Tensor sin_table, "
    "Tensor k_cache, Tensor v_cache, "
    "Tensor hidden_buffer, Tensor activations, Tensor residual, "
    "Tensor q, Tensor k, Tensor v, Tensor attn_out, "
    "Tensor mlp_intermediate, Tensor normalized, "
    "Tensor block_max_vals, Tensor block_max_idxs, "
    "int num_layers, int position, int max_seq_len, "
    "float attn_scale) -> ()"
)


def _decode_op(output_token, input_token_id, embed_weight, layer_weights_packed, final_norm_weight, lm_head_weight,
               cos_table, sin_table, k_cache, v_cache, hidden_buffer, activations, residual, q, k, v, attn_out,
               mlp_intermediate, normalized, block_max_vals, block_max_idxs, num_layers, position, max_seq_len,
               attn_scale):
    """Same schema and semantics as upstream torch_bindings.cpp:55-81 / :130-141, with the argument
    validation upstream lacks (upstream performs no checks at all)."""
    lib = _lib
    for name, t, dt in (("output_token", output_token, torch.int32), ("embed_weight", embed_weight, torch.bfloat16),
                        ("final_norm_weight", final_norm_weight, torch.bfloat16),
                        ("lm_head_weight", lm_head_weight, torch.bfloat16), ("cos_table", cos_table, torch.bfloat16),
                        ("sin_table", sin_table, torch.bfloat16), ("k_cache", k_cache, torch.bfloat16),
                        ("v_cache", v_cache, torch.bfloat16), ("hidden_buffer", hidden_buffer, torch.bfloat16),
                        ("normalized", normalized, torch.float32),
                        ("layer_weights_packed", layer_weights_packed, torch.uint8)):
        if not t.is_cuda or t.dtype != dt or not t.is_contiguous():
            raise ValueError(f"decode: {name} must be a contiguous CUDA tensor of dtype {dt}")
    if layer_weights_packed.numel() != int(num_layers) * 88:
        raise ValueError("decode: layer_weights_packed must hold num_layers * 88 bytes (11 pointers per layer)")
    if hidden_buffer.numel() != 1024 or normalized.numel() != 1024:
        raise ValueError("decode: hidden_buffer / normalized must have 1024 elements")
    if not (0 <= int(position) < int(max_seq_len)):
        raise ValueError(f"decode: position {position} outside [0, {max_seq_len})")
    _revalidate_blob(lib, layer_weights_packed, int(num_layers))
    with torch.cuda.device(hidden_buffer.device):
        stream = torch.cuda.current_stream().cuda_stream
        lib.launch_ldg_decode_direct(
            int(input_token_id), output_token.data_ptr(), embed_weight.data_ptr(), layer_weights_packed.data_ptr(),
            final_norm_weight.data_ptr(), lm_head_weight.data_ptr(), cos_table.data_ptr(), sin_table.data_ptr(),
            k_cache.data_ptr(), v_cache.data_ptr(), hidden_buffer.data_ptr(), activations.data_ptr(),
            residual.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(), attn_out.data_ptr(),
            mlp_intermediate.data_ptr(), normalized.data_ptr(), block_max_vals.data_ptr(), block_max_idxs.data_ptr(),
            int(num_layers), int(position), int(max_seq_len), float(attn_scale), stream)
    rc = lib.qmk_legacy_status()
    if rc < 0:
        raise NativeError(f"decode: {lib.qmk_last_error().decode()} (code {rc})")


# The C entry caches the re-packed weights by blob ADDRESS; torch's caching allocator re-uses addresses.  Remember which
# tensor (identity + in-place version) was seen at an address; when another one shows up there, compare the pointer table
# it holds with the one the cached model was packed from and drop the model if they differ.
_blob_seen: dict = {}


def _revalidate_blob(lib, blob: torch.Tensor, num_layers: int) -> None:
    key = (blob.device.index, blob.data_ptr(), num_layers)
    ent = _blob_seen.get(key)
    if ent is not None and ent[0]() is blob and ent[1] == blob._version:
        return
    if ent is not None:     # same address, different tensor or modified in place: one D2H copy of 88 bytes per layer
        host = blob.cpu().contiguous()
        lib.qmk_legacy_check_blob(blob.data_ptr(), num_layers, host.data_ptr())
    _blob_seen[key] = (weakref.ref(blob), blob._version)


class _Extension:
    """What ``get_extension()`` returns: the loaded C-ABI library plus the registered torch op."""

    def __init__(self, lib):
        self.lib = lib
        self.decode = torch.ops.qwen_megakernel_C.decode
        self.path = LIB_PATH

    def sync_status(self) -> None:
        """Synchronise the current stream and raise if a ``decode`` launch since the last call failed on the device
        (watchdog); clears the condition so that later calls work again."""
        rc = self.lib.qmk_legacy_sync_status(torch.cuda.current_stream().cuda_stream)
        if rc < 0:
            raise NativeError(f"decode: {self.lib.qmk_last_error().decode()} (code {rc})")


_ext = None


def get_extension():
    """Build (if needed), load, and register ``torch.ops.qwen_megakernel_C.decode`` (upstream build_tts.py:55-71)."""
    global _lib, _op_lib, _ext
    with _lock:
        if _ext is not None:
            return _ext
        path = build()
        _lib = load_library(path)
        if _lib.qmk_abi_version() != 3:
            raise RuntimeError("qwen_megakernel: ABI version mismatch between Python layer and libqmk_b200.so")
        _op_lib = torch.library.Library("qwen_megakernel_C", "DEF")
        _op_lib.define(DECODE_SCHEMA)
        _op_lib.impl("decode", _decode_op, "CUDA")
        _ext = _Extension(_lib)
        return _ext

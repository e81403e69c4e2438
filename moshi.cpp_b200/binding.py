"""ctypes binding of the C ABI in include/moshi_b200.h (libmoshi_b200.so) + in-tree build helper.

Fails loudly when the shared library is missing or a call returns an error: there is no fallback.
"""
from __future__ import annotations
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
SO_PATH = os.path.join(_HERE, "libmoshi_b200.so")
HOST_SO_PATH = os.path.join(_HERE, "libmoshi.so")
CSRC = os.path.join(_HERE, "csrc")
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared",
]

MSX_MAX_CODEBOOKS, MSX_MAX_STEPS = 40, 40
MSX_NO_TOKEN = -(2 ** 31)


STEP_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32))


class MsxError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"msx error {code}: {msg}")
        self.code = code


class MsxConfig(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("num_heads", C.c_int32), ("num_layers", C.c_int32), ("context", C.c_int32), ("max_period", C.c_int32),
        ("n_q", C.c_int32), ("dep_q", C.c_int32), ("card", C.c_int32), ("text_card", C.c_int32),
        ("dep_dim", C.c_int32), ("dep_heads", C.c_int32), ("dep_layers", C.c_int32), ("dep_context", C.c_int32), ("dep_max_period", C.c_int32),
        ("n_delays", C.c_int32), ("delays", C.c_int32 * MSX_MAX_CODEBOOKS),
        ("schedule_len", C.c_int32), ("schedule", C.c_int32 * MSX_MAX_STEPS),
        ("personaplex", C.c_int32), ("extra_heads", C.c_int32),
        ("cross_attention", C.c_int32), ("demux_second_stream", C.c_int32), ("dep_low_rank", C.c_int32),
    ]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/ for sm_100a into moshi.cpp_b200/libmoshi_b200.so (in-tree; nvcc cross-compiles without a GPU)."""
    srcs = sources() + [os.path.join(ROOT, "include", "moshi_b200.h")]
    stale = (not os.path.exists(SO_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(SO_PATH) for s in srcs)
    if force or stale:
        cmd = ["nvcc"] + NVCC_FLAGS + ["-o", SO_PATH, os.path.join(CSRC, "engine.cu"), os.path.join(CSRC, "gguf_file.cpp")]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return SO_PATH


HOST_DIR = os.path.join(_HERE, "host")
STS_BENCH = os.path.join(_HERE, "moshi-sts-bench")


TOOLS_DIR = os.path.join(ROOT, "tools")
BIN_DIR = os.path.join(_HERE, "bin")
TOOLS = ("moshi-sts", "personaplex", "moshi-tts", "moshi-stt")


def build_host(force: bool = False) -> str:
    """g++: the C++ side of the drop-in boundary — libmoshi.so (include/moshi/moshi.h on top of the C ABI), the four LM tools
    of the reference (tools/*.cpp -> moshi.cpp_b200/bin/) and the API test driver (moshi-sts-bench)."""
    srcs = [os.path.join(HOST_DIR, f) for f in ("moshi_api.cpp", "moshi_api.h", "moshi_sts_bench.cpp")]
    hdrs = [os.path.join(ROOT, "include", "moshi", f) for f in ("moshi.h", "ptrs.h", "ggml-backend.h")]
    tool_srcs = [os.path.join(TOOLS_DIR, t + ".cpp") for t in TOOLS] + [os.path.join(TOOLS_DIR, "lm_tool.h")]
    outs = [HOST_SO_PATH, STS_BENCH] + [os.path.join(BIN_DIR, t) for t in TOOLS]
    stale = any(not os.path.exists(o) for o in outs) or any(
        os.path.getmtime(s) > min(os.path.getmtime(o) for o in outs) for s in srcs + hdrs + tool_srcs)
    if force or stale:
        inc = "-I" + os.path.join(ROOT, "include")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-DMOSHI_BUILD", inc, "-o", HOST_SO_PATH,
                               srcs[0], os.path.join(_HERE, "csrc", "gguf_file.cpp"), "-L" + _HERE, "-lmoshi_b200", "-Wl,-rpath,$ORIGIN"])
        subprocess.check_call(["g++", "-O2", "-std=c++17", inc, "-o", STS_BENCH, srcs[2], "-L" + _HERE, "-lmoshi", "-lmoshi_b200", "-Wl,-rpath,$ORIGIN"])
        os.makedirs(BIN_DIR, exist_ok=True)
        for t in TOOLS:
            subprocess.check_call(["g++", "-O2", "-std=c++17", inc, "-o", os.path.join(BIN_DIR, t), os.path.join(TOOLS_DIR, t + ".cpp"),
                                   "-L" + _HERE, "-lmoshi", "-lmoshi_b200", "-Wl,-rpath,$ORIGIN/.."])
    return HOST_SO_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise MsxError(-4, f"{SO_PATH} is missing: run __graft_entry__.build() (nvcc, sm_100a). No CPU fallback exists.")
    L = C.CDLL(os.environ.get("MSX_LIB_EXPERIMENT") or SO_PATH)     # dev: A/B an experimental build of the same ABI
    vp, i32p = C.c_void_p, C.POINTER(C.c_int32)
    L.msx_last_error.restype = C.c_char_p
    L.msx_version.restype = C.c_char_p
    L.msx_device_count.restype = C.c_int
    L.msx_model_load_gguf.argtypes = [C.c_char_p, C.POINTER(MsxConfig), C.c_int, C.POINTER(vp)]
    L.msx_model_free.argtypes = [vp]
    L.msx_model_load_gguf_tp.argtypes = [C.c_char_p, C.POINTER(MsxConfig), C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.msx_tp_unique_id.argtypes = [vp]
    L.msx_model_load_gguf_ex.argtypes = [C.c_char_p, C.POINTER(MsxConfig), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.msx_stream_create_tp.argtypes = [vp, C.c_int, vp, C.POINTER(vp)]
    L.msx_stream_tp_export.argtypes = [vp, vp]
    L.msx_stream_prefill.argtypes = [vp, vp, C.c_int]
    L.msx_stream_tp_connect.argtypes = [vp, vp]
    L.msx_model_config.argtypes = [vp, C.POINTER(MsxConfig)]
    L.msx_model_weight_bytes_per_frame.restype = C.c_int64; L.msx_model_weight_bytes_per_frame.argtypes = [vp]
    L.msx_model_device_bytes.restype = C.c_int64; L.msx_model_device_bytes.argtypes = [vp]
    L.msx_model_device.argtypes = [vp]
    L.msx_stream_create.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.msx_stream_create_ex.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp)]
    L.msx_stream_free.argtypes = [vp]
    L.msx_stream_reset.argtypes = [vp]
    L.msx_stream_offset.argtypes = [vp]
    L.msx_stream_kv_bytes_next.restype = C.c_int64; L.msx_stream_kv_bytes_next.argtypes = [vp]
    L.msx_step_temporal.argtypes = [vp, vp, i32p, vp, vp]
    L.msx_step_depformer.argtypes = [vp, C.c_int32, vp, vp, vp]
    L.msx_step_temporal_embedding.argtypes = [vp, vp, i32p, vp, vp]
    L.msx_gen_prompt_embedding.argtypes = [vp, vp]
    L.msx_gen_set_cache.argtypes = [vp, vp]
    L.msx_gen_cache_rows.argtypes = [vp]
    L.msx_step.argtypes = [vp, vp, vp]
    L.msx_vad.argtypes = [vp, C.POINTER(C.c_float)]
    L.msx_stream_set_condition.argtypes = [vp, vp, vp, C.c_int]
    L.msx_stream_set_voice.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp]
    L.msx_stream_load_voice.argtypes = [vp, C.c_char_p]
    L.msx_model_has_conditioners.argtypes = [vp]
    L.msx_stream_set_sampling.argtypes = [vp, C.c_float, C.c_float, C.c_int, C.c_int]
    L.msx_stream_set_noise.argtypes = [vp, vp, vp]
    L.msx_gen_seed.argtypes = [vp, C.c_uint]
    L.msx_run_resident.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.POINTER(C.c_float)]
    L.msx_stream_launches_per_frame.argtypes = [vp]
    L.msx_run_resident_async.argtypes = [vp, vp, C.c_int, C.c_int]
    L.msx_stream_wait.argtypes = [vp, C.POINTER(C.c_float)]
    L.msx_profile_frame.argtypes = [vp, vp, vp, vp, vp, C.c_int]
    L.msx_run_resident_split.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp]
    L.msx_step_timeline.argtypes = [vp, vp, vp, C.c_int, vp, vp]
    L.msx_family_count.restype = C.c_int
    L.msx_timer_start.argtypes = [vp]
    L.msx_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.msx_family_name.restype = C.c_char_p; L.msx_family_name.argtypes = [C.c_int]
    L.msx_stream_get_kv.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp]
    L.msx_gen_create.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.msx_gen_free.argtypes = [vp]
    L.msx_gen_create_with_callback.argtypes = [C.POINTER(MsxConfig), C.c_int, STEP_FN, vp, C.POINTER(vp)]
    L.msx_gen_step.argtypes = [vp, vp, C.c_int, C.c_int, i32p, vp]
    L.msx_gen_offset.argtypes = [vp]
    L.msx_gen_prefill.argtypes = [vp, vp, C.c_int]
    L.msx_gen_max_delay.argtypes = [vp]
    L.msx_test_gemv.argtypes = [C.c_int, C.c_int, vp, C.c_int64, C.c_int64, vp, vp, C.c_int, vp]
    L.msx_batch_create.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp)]
    L.msx_batch_free.argtypes = [vp]
    L.msx_batch_size.argtypes = [vp]
    L.msx_batch_offset.argtypes = [vp, C.c_int]
    L.msx_batch_launches_per_frame.argtypes = [vp]
    L.msx_batch_kv_bytes_next.restype = C.c_int64; L.msx_batch_kv_bytes_next.argtypes = [vp]
    L.msx_batch_reset_stream.argtypes = [vp, C.c_int]
    L.msx_batch_step.argtypes = [vp, vp, vp]
    L.msx_bgen_create.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.msx_bgen_free.argtypes = [vp]
    L.msx_bgen_offset.argtypes = [vp, C.c_int]
    L.msx_bgen_reset_stream.argtypes = [vp, C.c_int]
    L.msx_bgen_step.argtypes = [vp, vp, C.c_int, vp, vp, vp]
    L.msx_batch_set_sampling.argtypes = [vp, C.c_float, C.c_float, C.c_int, C.c_int]
    L.msx_batch_set_noise.argtypes = [vp, vp, vp]
    L.msx_batch_get_logits.argtypes = [vp, C.c_int, vp, vp]
    L.msx_batch_run_resident.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.POINTER(C.c_float)]
    L.msx_batch_profile_frame.argtypes = [vp, vp, vp, vp, C.c_int]
    L.msx_stream_prefill_profile.argtypes = [vp, vp, vp, vp, C.c_int]
    L.msx_rvq_create.argtypes = [C.c_int] * 6 + [vp] * 7
    L.msx_rvq_free.argtypes = [vp]; L.msx_rvq_free.restype = None
    L.msx_rvq_encode.argtypes = [vp, vp, C.c_int, C.c_int, vp]
    L.msx_rvq_decode.argtypes = [vp, vp, C.c_int, C.c_int, vp]
    L.msx_rvq_bench.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp]
    L.msx_test_gemm_batch.argtypes = [C.c_int, C.c_int, vp, C.c_int64, C.c_int64, vp, C.c_int, vp, vp]
    L.msx_bench_gemm_batch_ex.argtypes = [C.c_int, vp, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), vp]
    L.msx_bench_gemm_batch.argtypes = [C.c_int, vp, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.msx_test_dequant_rows.argtypes = [C.c_int, C.c_int, vp, C.c_int64, C.c_int64, vp, C.c_int, vp]
    L.msx_test_dequant_repacked.argtypes = [C.c_int, C.c_int, vp, C.c_int64, C.c_int64, vp]
    L.msx_gguf_quantize.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    L.msx_safetensors_to_gguf.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    L.msx_test_quantize_rows.argtypes = [C.c_int, C.c_int, C.c_int, vp, C.c_int64, C.c_int64, vp]
    _lib = L
    return L


def _check(rc: int):
    if rc < 0:
        raise MsxError(rc, lib().msx_last_error().decode(errors="replace"))
    return rc


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def make_config(cfg: dict) -> MsxConfig:
    c = MsxConfig()
    c.dim, c.num_heads, c.num_layers, c.context, c.max_period = cfg["dim"], cfg["num_heads"], cfg["num_layers"], cfg["context"], cfg["max_period"]
    c.n_q, c.dep_q, c.card, c.text_card = cfg["n_q"], cfg["dep_q"], cfg["card"], cfg["text_card"]
    c.dep_dim, c.dep_heads, c.dep_layers = cfg["depformer_dim"], cfg["depformer_num_heads"], cfg["depformer_num_layers"]
    c.dep_context, c.dep_max_period = cfg["depformer_context"], cfg["depformer_max_period"]
    c.n_delays = len(cfg["delays"])
    for i, d in enumerate(cfg["delays"]):
        c.delays[i] = d
    c.schedule_len = len(cfg["schedule"])
    for i, s in enumerate(cfg["schedule"]):
        c.schedule[i] = s
    c.personaplex = 1 if cfg["model_type"] == "personaplex" else 0
    c.extra_heads = cfg["extra_heads"]
    c.cross_attention = 1 if cfg.get("cross_attention") else 0
    c.demux_second_stream = 1 if cfg.get("demux") else 0
    c.dep_low_rank = int(cfg.get("dep_low_rank") or 0)
    return c


def tp_unique_id() -> bytes:
    """128-byte NCCL id for one tensor-parallel group: create on one rank, broadcast to the others"""
    buf = np.zeros(128, dtype=np.uint8)
    _check(lib().msx_tp_unique_id(_p(buf)))
    return buf.tobytes()


class Model:
    def __init__(self, gguf_path: str, cfg: dict, device: int = 0, tp_rank: int = 0, tp_world: int = 1, quantize: str | None = None):
        self.cfg = cfg
        self._c = make_config(cfg)
        self.tp_rank, self.tp_world = tp_rank, tp_world
        h = C.c_void_p()
        q = {None: 0, "q8_0": 8, "q4_k": 12}[quantize]
        _check(lib().msx_model_load_gguf_ex(gguf_path.encode(), C.byref(self._c), device, tp_rank, tp_world, q, C.byref(h)))
        self.h = h

    @property
    def weight_bytes_per_frame(self) -> int:
        return int(lib().msx_model_weight_bytes_per_frame(self.h))

    @property
    def device_bytes(self) -> int:
        return int(lib().msx_model_device_bytes(self.h))

    def close(self):
        if getattr(self, "h", None):
            lib().msx_model_free(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Stream:
    def __init__(self, model: Model, context: int = 0, step_kernel: bool = False, nccl_id: bytes | None = None):
        """step_kernel: one persistent cooperative kernel per stack (MSX_STREAM_STEP_KERNEL) instead of PDL-chained launches"""
        self.model = model
        h = C.c_void_p()
        if nccl_id is not None:
            idb = np.frombuffer(nccl_id, dtype=np.uint8).copy()
            assert idb.size == 128
            _check(lib().msx_stream_create_tp(model.h, context, _p(idb), C.byref(h)))
        else:
            _check(lib().msx_stream_create_ex(model.h, context, 1 if step_kernel else 0, C.byref(h)))
        self.h = h

    @property
    def offset(self) -> int:
        return lib().msx_stream_offset(self.h)

    @property
    def launches_per_frame(self) -> int:
        return lib().msx_stream_launches_per_frame(self.h)

    @property
    def kv_bytes_next(self) -> int:
        return int(lib().msx_stream_kv_bytes_next(self.h))

    def reset(self):
        _check(lib().msx_stream_reset(self.h))

    def step_temporal(self, tokens, want_logits=True):
        cfg = self.model.cfg
        tok = np.ascontiguousarray(tokens, dtype=np.int32)
        assert tok.size == cfg["n_q"] + 1
        t = C.c_int32(0)
        logits = np.empty(cfg["text_card"], dtype=np.float32) if want_logits else None
        tout = np.empty(cfg["dim"], dtype=np.float32) if want_logits else None
        _check(lib().msx_step_temporal(self.h, _p(tok), C.byref(t), _p(logits), _p(tout)))
        return int(t.value), logits, tout

    def step_temporal_embedding(self, x, want_logits=True):
        cfg = self.model.cfg
        xx = np.ascontiguousarray(x, dtype=np.float32)
        assert xx.size == cfg["dim"]
        t = C.c_int32(0)
        logits = np.empty(cfg["text_card"], dtype=np.float32) if want_logits else None
        tout = np.empty(cfg["dim"], dtype=np.float32) if want_logits else None
        _check(lib().msx_step_temporal_embedding(self.h, _p(xx), C.byref(t), _p(logits), _p(tout)))
        return int(t.value), logits, tout

    def step_depformer(self, text_token: int, force=None, want_logits=True):
        cfg = self.model.cfg
        toks = np.empty(cfg["dep_q"], dtype=np.int32)
        logits = np.empty((cfg["dep_q"], cfg["card"]), dtype=np.float32) if want_logits else None
        f = np.ascontiguousarray(force, dtype=np.int32) if force is not None else None
        _check(lib().msx_step_depformer(self.h, int(text_token), _p(f), _p(toks), _p(logits)))
        return toks, logits

    def step(self, tokens):
        cfg = self.model.cfg
        tok = np.ascontiguousarray(tokens, dtype=np.int32)
        assert tok.size == cfg["n_q"] + 1
        out = np.empty(1 + cfg["dep_q"], dtype=np.int32)
        _check(lib().msx_step(self.h, _p(tok), _p(out)))
        return out

    def set_sampling(self, temp_text, temp_audio, top_k_text=25, top_k_audio=250):
        _check(lib().msx_stream_set_sampling(self.h, temp_text, temp_audio, top_k_text, top_k_audio))

    def set_noise(self, noise_text, noise_audio):
        nt = np.ascontiguousarray(noise_text, dtype=np.float32); na = np.ascontiguousarray(noise_audio, dtype=np.float32)
        _check(lib().msx_stream_set_noise(self.h, _p(nt), _p(na)))

    def prefill(self, tokens):
        """tokens [T][n_q+1]: batched-T prompt prefill (KV rings + position only)"""
        tk = np.ascontiguousarray(tokens, dtype=np.int32).reshape(-1, self.model.cfg["n_q"] + 1)
        _check(lib().msx_stream_prefill(self.h, _p(tk), tk.shape[0]))

    def prefill_profile(self, tokens):
        """tokens [1 + pass][n_q+1]: one prompt row, then one full prefill pass launched eagerly -> {family: (ms, launches)}"""
        tk = np.ascontiguousarray(tokens, dtype=np.int32).reshape(-1, self.model.cfg["n_q"] + 1)
        n = lib().msx_family_count()
        ms = np.zeros(n, dtype=np.float32); cnt = np.zeros(n, dtype=np.int32)
        _check(lib().msx_stream_prefill_profile(self.h, _p(tk), _p(ms), _p(cnt), n))
        return {lib().msx_family_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i]}

    def tp_export(self) -> bytes:
        buf = np.zeros(64, dtype=np.uint8)
        _check(lib().msx_stream_tp_export(self.h, _p(buf)))
        return buf.tobytes()

    def tp_connect(self, handles):
        """handles: list of 64-byte IPC handles in rank order -> fused GEMV + peer-memory all-reduce"""
        buf = np.frombuffer(b"".join(handles), dtype=np.uint8).copy()
        _check(lib().msx_stream_tp_connect(self.h, _p(buf)))

    def set_condition(self, cond_sum=None, cond_cross=None):
        """TTS conditioning: cond_sum [dim] and / or cond_cross [Tc][dim]"""
        cs = np.ascontiguousarray(cond_sum, dtype=np.float32) if cond_sum is not None else None
        cc = np.ascontiguousarray(cond_cross, dtype=np.float32) if cond_cross is not None else None
        _check(lib().msx_stream_set_condition(self.h, _p(cs), _p(cc), 0 if cc is None else cc.shape[0]))

    def set_voice(self, speaker_wavs):
        """TTS voice conditioners on the GPU; speaker_wavs [channels][frames].  Returns (cond_sum [dim], cond_cross [5T][dim])"""
        w = np.ascontiguousarray(speaker_wavs, dtype=np.float32)
        dim = self.model.cfg["dim"]
        cs = np.empty(dim, dtype=np.float32); cc = np.empty((5 * w.shape[1], dim), dtype=np.float32)
        _check(lib().msx_stream_set_voice(self.h, _p(w), w.shape[0], w.shape[1], _p(cs), _p(cc)))
        return cs, cc

    def load_voice(self, path: str):
        _check(lib().msx_stream_load_voice(self.h, path.encode()))

    def vad(self) -> float:
        v = C.c_float(0)
        _check(lib().msx_vad(self.h, C.byref(v)))
        return float(v.value)

    def run_resident(self, frames, n_steps: int, want_tokens: bool = False):
        cfg = self.model.cfg
        fr = np.ascontiguousarray(frames, dtype=np.int32).reshape(-1, cfg["n_q"] + 1)
        out = np.empty((n_steps, 1 + cfg["dep_q"]), dtype=np.int32) if want_tokens else None
        ms = C.c_float(0)
        _check(lib().msx_run_resident(self.h, _p(fr), fr.shape[0], n_steps, _p(out), C.byref(ms)))
        return float(ms.value), out

    def run_resident_split(self, frames, n_steps: int):
        """-> (temporal_ms, depformer_ms) summed over n_steps frames of the graph-replayed run (one event between the two graphs)"""
        cfg = self.model.cfg
        fr = np.ascontiguousarray(frames, dtype=np.int32).reshape(-1, cfg["n_q"] + 1)
        t = C.c_float(0); d = C.c_float(0)
        _check(lib().msx_run_resident_split(self.h, _p(fr), fr.shape[0], n_steps, C.byref(t), C.byref(d)))
        return float(t.value), float(d.value)

    def profile_frame(self, tokens):
        """-> (out_tokens, {family: (ms, launches)}) for one eagerly-launched frame"""
        cfg = self.model.cfg
        tok = np.ascontiguousarray(tokens, dtype=np.int32)
        out = np.empty(1 + cfg["dep_q"], dtype=np.int32)
        n = lib().msx_family_count()
        ms = np.zeros(n, dtype=np.float32); cnt = np.zeros(n, dtype=np.int32)
        _check(lib().msx_profile_frame(self.h, _p(tok), _p(out), _p(ms), _p(cnt), n))
        fam = {lib().msx_family_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i]}
        return out, fam

    def step_timeline(self, tokens):
        """one frame through the persistent step kernels -> (families [n_phases], stamps [n_phases][n_cta][8] ns =
        start, end, after prologue, after main loop, input loaded, rms scale known, epilogue stored, -)"""
        tok = np.ascontiguousarray(tokens, dtype=np.int32)
        max_rows = 4096 * 160
        rows = np.zeros((max_rows, 9), dtype=np.int64); n = C.c_int(0); nc = C.c_int(0)
        _check(lib().msx_step_timeline(self.h, _p(tok), _p(rows), max_rows, C.byref(n), C.byref(nc)))
        r = rows[: n.value * nc.value].reshape(n.value, nc.value, 9)
        fams = [lib().msx_family_name(int(f)).decode() for f in r[:, 0, 0]]
        return fams, r[:, :, 1:]

    def timer_start(self):
        _check(lib().msx_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        _check(lib().msx_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    def run_resident_async(self, frames, n_steps: int):
        cfg = self.model.cfg
        fr = np.ascontiguousarray(frames, dtype=np.int32).reshape(-1, cfg["n_q"] + 1)
        _check(lib().msx_run_resident_async(self.h, _p(fr), fr.shape[0], n_steps))

    def wait(self) -> float:
        ms = C.c_float(0)
        _check(lib().msx_stream_wait(self.h, C.byref(ms)))
        return float(ms.value)

    def get_kv(self, layer: int, head: int, slot: int):
        cfg = self.model.cfg
        dh = cfg["dim"] // cfg["num_heads"]
        k = np.empty(dh, dtype=np.uint16); v = np.empty(dh, dtype=np.uint16)
        _check(lib().msx_stream_get_kv(self.h, layer, head, slot, _p(k), _p(v)))
        return k, v

    def close(self):
        if getattr(self, "h", None):
            lib().msx_stream_free(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Gen:
    """LMGen over a stream (msx_gen_*): mirrors moshi_lm_send2 / moshi_lm_receive.
    With `step_fn(tokens, replace) -> [text, audio...]` instead of a stream, the same host logic runs
    over a Python model step (CPU-only host-logic tests)."""

    def __init__(self, stream: "Stream | None", delay_steps: int = 0, cfg: dict | None = None, step_fn=None):
        self.stream = stream
        self.cfg = stream.model.cfg if stream is not None else cfg
        h = C.c_void_p()
        if stream is not None:
            _check(lib().msx_gen_create(stream.h, delay_steps, C.byref(h)))
        else:
            n_in, n_out = self.cfg["n_q"] + 1, 1 + self.cfg["dep_q"]

            def _cb(user, tokens, replace, out):
                res = step_fn([tokens[i] for i in range(n_in)], bool(replace))
                for i in range(n_out):
                    out[i] = int(res[i])
                return 0
            self._cb = STEP_FN(_cb)
            self._ccfg = make_config(self.cfg)
            _check(lib().msx_gen_create_with_callback(C.byref(self._ccfg), delay_steps, self._cb, None, C.byref(h)))
        self.h = h

    @property
    def offset(self):
        return lib().msx_gen_offset(self.h)

    def prompt_embedding(self, x):
        """one PersonaPlex voice-prompt frame given as an embedding row [dim] (lm.h:1005-1036)"""
        xx = np.ascontiguousarray(x, dtype=np.float32)
        _check(lib().msx_gen_prompt_embedding(self.h, _p(xx)))

    def set_cache(self, ring):
        """replace the token delay ring [CT][n_q+1] (lm.h:1038-1051)"""
        r = np.ascontiguousarray(ring, dtype=np.int32)
        assert r.shape[0] == lib().msx_gen_cache_rows(self.h)
        _check(lib().msx_gen_set_cache(self.h, _p(r)))

    def prefill(self, rows):
        """rows [T][n_q+1] (all tokens given): batched-T prompt prefill with the generator's ring bookkeeping"""
        r = np.ascontiguousarray(rows, dtype=np.int32).reshape(-1, self.cfg["n_q"] + 1)
        _check(lib().msx_gen_prefill(self.h, _p(r), r.shape[0]))

    def step(self, in_tokens, replace: bool = False):
        cfg = self.cfg
        tok = np.ascontiguousarray(in_tokens, dtype=np.int32)
        text = C.c_int32(0)
        audio = np.full(max(1, cfg["dep_q"]), -7, dtype=np.int32)
        rc = _check(lib().msx_gen_step(self.h, _p(tok), tok.size, int(replace), C.byref(text), _p(audio)))
        return rc, int(text.value), audio[: cfg["dep_q"]].copy()

    def close(self):
        if getattr(self, "h", None):
            lib().msx_gen_free(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """Lock-step batch of independent streams on one GPU (msx_batch_*): weights are read once per frame for all."""

    def __init__(self, model: Model, n_streams: int, context: int = 0):
        self.model = model
        self.n = n_streams
        self.h = C.c_void_p()
        _check(lib().msx_batch_create(model.h, n_streams, context, C.byref(self.h)))

    @property
    def launches_per_frame(self) -> int:
        return lib().msx_batch_launches_per_frame(self.h)

    def offset(self, stream: int) -> int:
        return lib().msx_batch_offset(self.h, stream)

    def kv_bytes_next(self) -> int:
        return int(lib().msx_batch_kv_bytes_next(self.h))

    def reset(self, stream: int = -1):
        _check(lib().msx_batch_reset_stream(self.h, stream))

    def step(self, tokens):
        cfg = self.model.cfg
        tok = np.ascontiguousarray(tokens, dtype=np.int32).reshape(self.n, cfg["n_q"] + 1)
        out = np.empty((self.n, 1 + cfg["dep_q"]), dtype=np.int32)
        _check(lib().msx_batch_step(self.h, _p(tok), _p(out)))
        return out

    def set_sampling(self, temp_text, temp_audio, top_k_text=25, top_k_audio=250):
        _check(lib().msx_batch_set_sampling(self.h, temp_text, temp_audio, top_k_text, top_k_audio))

    def set_noise(self, noise_text, noise_audio):
        nt = np.ascontiguousarray(noise_text, dtype=np.float32); na = np.ascontiguousarray(noise_audio, dtype=np.float32)
        _check(lib().msx_batch_set_noise(self.h, _p(nt), _p(na)))

    def logits(self, stream: int):
        cfg = self.model.cfg
        tl = np.empty(cfg["text_card"], dtype=np.float32)
        al = np.empty((cfg["dep_q"], cfg["card"]), dtype=np.float32) if cfg["dep_q"] > 0 else None
        _check(lib().msx_batch_get_logits(self.h, stream, _p(tl), _p(al)))
        return tl, al

    def run_resident(self, frames, n_steps: int, want_tokens: bool = False):
        """frames [n][n_frames][n_q+1] -> (elapsed ms, tokens [n][n_steps][1+dep_q] or None)"""
        cfg = self.model.cfg
        fr = np.ascontiguousarray(frames, dtype=np.int32).reshape(self.n, -1, cfg["n_q"] + 1)
        out = np.empty((self.n, n_steps, 1 + cfg["dep_q"]), dtype=np.int32) if want_tokens else None
        ms = C.c_float(0)
        _check(lib().msx_batch_run_resident(self.h, _p(fr), fr.shape[1], n_steps, _p(out), C.byref(ms)))
        return float(ms.value), out

    def profile_frame(self, tokens):
        cfg = self.model.cfg
        tok = np.ascontiguousarray(tokens, dtype=np.int32).reshape(self.n, cfg["n_q"] + 1)
        n = lib().msx_family_count()
        ms = np.zeros(n, dtype=np.float32); cnt = np.zeros(n, dtype=np.int32)
        _check(lib().msx_batch_profile_frame(self.h, _p(tok), _p(ms), _p(cnt), n))
        return {lib().msx_family_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i]}

    def close(self):
        if self.h:
            lib().msx_batch_free(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class BatchGen:
    """LMGen host logic (delay ring, delayed emit) for every stream of a Batch around one batched model step"""

    def __init__(self, batch: Batch, delay_steps: int = 0):
        self.batch = batch
        self.h = C.c_void_p()
        _check(lib().msx_bgen_create(batch.h, delay_steps, C.byref(self.h)))

    def offset(self, stream: int) -> int:
        return lib().msx_bgen_offset(self.h, stream)

    def reset(self, stream: int):
        _check(lib().msx_bgen_reset_stream(self.h, stream))

    def step(self, in_tokens):
        """in_tokens [n][n_in] -> (valid [n], text [n], audio [n][dep_q])"""
        cfg = self.batch.model.cfg
        n = self.batch.n
        tok = np.ascontiguousarray(in_tokens, dtype=np.int32).reshape(n, -1)
        text = np.zeros(n, dtype=np.int32); audio = np.full((n, max(1, cfg["dep_q"])), -7, dtype=np.int32); valid = np.zeros(n, dtype=np.int32)
        _check(lib().msx_bgen_step(self.h, _p(tok), tok.shape[1], _p(text), _p(audio), _p(valid)))
        return valid, text, audio[:, : cfg["dep_q"]]

    def close(self):
        if self.h:
            lib().msx_bgen_free(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- unit-level helpers -------------------------------------------------------------------------
def test_gemv(gtype: int, w_raw: np.ndarray, k: int, x: np.ndarray, alpha=None, device: int = 0) -> np.ndarray:
    w_raw = np.ascontiguousarray(w_raw); x = np.ascontiguousarray(x, dtype=np.float32)
    rows = w_raw.shape[0]
    y = np.empty(rows, dtype=np.float32)
    al = np.ascontiguousarray(alpha, dtype=np.float32) if alpha is not None else None
    _check(lib().msx_test_gemv(device, gtype, _p(w_raw), k, rows, _p(x), _p(al), 1 if alpha is not None else 0, _p(y)))
    return y


def test_dequant_rows(gtype: int, table_raw: np.ndarray, k: int, ids, device: int = 0) -> np.ndarray:
    table_raw = np.ascontiguousarray(table_raw)
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    out = np.empty((ids.size, k), dtype=np.float32)
    _check(lib().msx_test_dequant_rows(device, gtype, _p(table_raw), k, table_raw.shape[0], _p(ids), ids.size, _p(out)))
    return out


def test_dequant_repacked(gtype: int, w_raw: np.ndarray, k: int, device: int = 0) -> np.ndarray:
    w_raw = np.ascontiguousarray(w_raw)
    out = np.empty((w_raw.shape[0], k), dtype=np.float32)
    _check(lib().msx_test_dequant_repacked(device, gtype, _p(w_raw), k, w_raw.shape[0], _p(out)))
    return out


def gguf_quantize(in_path: str, out_path: str, quantize: str | None, device: int = 0) -> None:
    """unquantised GGUF -> q8_0 / q4_k GGUF on the GPU (the reference's `-q <quant> -g out.gguf`)"""
    _check(lib().msx_gguf_quantize(in_path.encode(), out_path.encode(), {None: 0, "q8_0": 8, "q4_k": 12}[quantize], device))


def safetensors_to_gguf(in_path: str, out_path: str, quantize: str | None, device: int = 0) -> None:
    """model.safetensors (torch names) -> GGUF with the reference loader's names, splits and quantisation rules"""
    _check(lib().msx_safetensors_to_gguf(in_path.encode(), out_path.encode(), {None: 0, "q8_0": 8, "q4_k": 12}[quantize], device))


def test_quantize_rows(dst_type: int, x: np.ndarray, src_type: int = 0, device: int = 0) -> np.ndarray:
    """GGUF blocks of the on-load quantisers; x is [rows][k] f32 (src_type 0), f16 bits (1) or bf16 bits (30)"""
    x = np.ascontiguousarray(x)
    rows, k = x.shape
    per = {8: (32, 34), 2: (32, 18), 12: (256, 144)}[dst_type]
    out = np.empty((rows, k // per[0] * per[1]), dtype=np.uint8)
    _check(lib().msx_test_quantize_rows(device, src_type, dst_type, _p(x), k, rows, _p(out)))
    return out


def test_gemm_batch(gtype: int, w_raw: np.ndarray, k: int, x: np.ndarray, alpha=None, device: int = 0) -> np.ndarray:
    """y[nb][rows] through the batched quantise + tensor-core GEMM kernels; x is [nb][k]"""
    w_raw = np.ascontiguousarray(w_raw); x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, k)
    rows = w_raw.shape[0]
    y = np.empty((x.shape[0], rows), dtype=np.float32)
    al = np.ascontiguousarray(alpha, dtype=np.float32) if alpha is not None else None
    _check(lib().msx_test_gemm_batch(device, gtype, _p(w_raw), k, rows, _p(x), x.shape[0], _p(al), _p(y)))
    return y


def bench_gemm_batch(w_raw: np.ndarray, k: int, nb: int, n_mats: int, iters: int, epilogue: int = 0, with_quant: bool = True, device: int = 0) -> float:
    w_raw = np.ascontiguousarray(w_raw)
    us = C.c_float(0)
    _check(lib().msx_bench_gemm_batch(device, _p(w_raw), k, w_raw.shape[0], nb, n_mats, iters, epilogue, 1 if with_quant else 0, C.byref(us)))
    return float(us.value)


def bench_gemm_batch_stamps(w_raw: np.ndarray, k: int, nb: int, n_mats: int, iters: int, epilogue: int = 0, with_quant: bool = True, device: int = 0):
    w_raw = np.ascontiguousarray(w_raw)
    us = C.c_float(0)
    st = np.zeros((iters, 148, 8), dtype=np.int64)
    _check(lib().msx_bench_gemm_batch_ex(device, _p(w_raw), k, w_raw.shape[0], nb, n_mats, iters, epilogue, 1 if with_quant else 0, C.byref(us), _p(st)))
    return float(us.value), st


class RVQ:
    """Mimi split residual vector quantiser on the GPU (msx_rvq_*): codes <-> latent, reference quantization/vq.h:69-117.
    cb_first [n_sem][bins][D] f32, cb_rest [n_rest][bins][D] f32; projections as F16 bit patterns: in_* [D][dim], out_* [dim][D]"""

    def __init__(self, cb_first, cb_rest, in_first, in_rest, out_first, out_rest, device: int = 0):
        cb_first = np.ascontiguousarray(cb_first, dtype=np.float32); cb_rest = np.ascontiguousarray(cb_rest, dtype=np.float32)
        ws = [np.ascontiguousarray(w, dtype=np.uint16) for w in (in_first, in_rest, out_first, out_rest)]
        self.n_sem, self.bins, self.D = cb_first.shape
        self.n_rest = cb_rest.shape[0]
        self.dim = ws[0].shape[1]
        self.h = C.c_void_p()
        _check(lib().msx_rvq_create(device, self.n_sem, self.n_rest, self.bins, self.D, self.dim, _p(cb_first), _p(cb_rest),
                                    _p(ws[0]), _p(ws[1]), _p(ws[2]), _p(ws[3]), C.byref(self.h)))

    def encode(self, x, n_q: int):
        x = np.ascontiguousarray(x, dtype=np.float32)
        T = x.shape[0]
        codes = np.zeros((n_q, T), dtype=np.int32)
        _check(lib().msx_rvq_encode(self.h, _p(x), T, n_q, _p(codes)))
        return codes

    def decode(self, codes):
        codes = np.ascontiguousarray(codes, dtype=np.int32)
        K, T = codes.shape
        y = np.zeros((T, self.dim), dtype=np.float32)
        _check(lib().msx_rvq_decode(self.h, _p(codes), K, T, _p(y)))
        return y

    def bench(self, T: int, n_q: int, reps: int = 20):
        """-> (encode_ms, decode_ms) per call, device-timed on resident buffers"""
        e = C.c_float(0); d = C.c_float(0)
        _check(lib().msx_rvq_bench(self.h, T, n_q, reps, C.byref(e), C.byref(d)))
        return float(e.value), float(d.value)

    def close(self):
        if self.h:
            lib().msx_rvq_free(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

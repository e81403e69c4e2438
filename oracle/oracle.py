"""ctypes front-end of the TEST-ONLY CPU oracle (oracle/ggml_ref.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
GGUF files are read with gguf-py's GGUFReader (an implementation independent of the product's C++
GGUF parser), and tensor pointers are handed to the C restatement.
"""
from __future__ import annotations
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libggml_ref.so")

ORC_MAX_CODEBOOKS, ORC_MAX_STEPS = 40, 40


class OrcConfig(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("num_heads", C.c_int32), ("num_layers", C.c_int32), ("context", C.c_int32), ("max_period", C.c_int32),
        ("n_q", C.c_int32), ("dep_q", C.c_int32), ("card", C.c_int32), ("text_card", C.c_int32),
        ("dep_dim", C.c_int32), ("dep_heads", C.c_int32), ("dep_layers", C.c_int32), ("dep_context", C.c_int32), ("dep_max_period", C.c_int32),
        ("n_delays", C.c_int32), ("delays", C.c_int32 * ORC_MAX_CODEBOOKS),
        ("schedule_len", C.c_int32), ("schedule", C.c_int32 * ORC_MAX_STEPS),
        ("personaplex", C.c_int32), ("extra_heads", C.c_int32), ("delay_steps", C.c_int32),
        ("cross_attention", C.c_int32), ("demux_second_stream", C.c_int32), ("dep_low_rank", C.c_int32),
    ]


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("ggml_ref.c", "mimi_rvq_ref.c", "ggml_ref.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def set_threads(n: int) -> int:
    """use n OpenMP threads from now on (torchrun exports OMP_NUM_THREADS=1, which would otherwise pin the CPU legs to one core);
    returns the thread count the runtime reports afterwards"""
    lib().orc_set_threads(int(n))
    return int(lib().orc_max_threads())


def max_threads() -> int:
    return int(lib().orc_max_threads())


def lib():
    global _lib
    if _lib is None:
        try:
            _lib = C.CDLL(build())
        except OSError:
            _lib = C.CDLL(build(force=True))
        L = _lib
        vp, i32p, fp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float)
        L.orc_set_threads.argtypes = [C.c_int]; L.orc_max_threads.restype = C.c_int
        L.orc_row_size.restype = C.c_int64; L.orc_row_size.argtypes = [C.c_int, C.c_int64]
        L.orc_dequantize_row.argtypes = [C.c_int, vp, vp, C.c_int64]
        L.orc_quantize_row_q8_0.argtypes = [vp, vp, C.c_int64]
        L.orc_quantize_row_q8_K.argtypes = [vp, vp, vp, vp, C.c_int64]
        L.orc_quantize_row_q4_0.argtypes = [vp, vp, C.c_int64]
        L.orc_timestep_freq.argtypes = [C.c_int, C.c_int, vp]
        L.orc_quantize_row_q4_K.argtypes = [vp, vp, C.c_int64]
        L.orc_mul_mat_vec.argtypes = [C.c_int, vp, C.c_int64, C.c_int64, vp, vp]
        L.orc_mul_mat_vec_ideal.argtypes = [C.c_int, vp, C.c_int64, C.c_int64, vp, vp]
        L.orc_rms_norm.argtypes = [vp, vp, C.c_float, vp, C.c_int64]
        L.orc_fp32_to_bf16.restype = C.c_uint16; L.orc_fp32_to_bf16.argtypes = [C.c_float]
        L.orc_model_new.restype = vp; L.orc_model_new.argtypes = [C.POINTER(OrcConfig)]
        L.orc_model_free.argtypes = [vp]
        L.orc_model_set_tensor.restype = C.c_int
        L.orc_model_set_tensor.argtypes = [vp, C.c_char_p, C.c_int, C.c_int64, C.c_int64, vp]
        L.orc_model_missing.restype = C.c_int; L.orc_model_missing.argtypes = [vp, C.c_char_p, C.c_int]
        L.orc_model_set_ideal.argtypes = [vp, C.c_int]
        L.orc_state_new.restype = vp; L.orc_state_new.argtypes = [vp]
        L.orc_state_free.argtypes = [vp]; L.orc_state_reset.argtypes = [vp]
        L.orc_state_offset.restype = C.c_int; L.orc_state_offset.argtypes = [vp]
        L.orc_step_temporal.restype = C.c_int; L.orc_step_temporal.argtypes = [vp, vp, vp, vp, vp]
        L.orc_step_depformer.argtypes = [vp, vp, C.c_int, vp, vp, vp]
        L.orc_step_temporal_embedding.restype = C.c_int; L.orc_step_temporal_embedding.argtypes = [vp, vp, vp, vp, vp]
        L.orc_vad.restype = C.c_float; L.orc_vad.argtypes = [vp, vp]
        L.orc_state_set_condition.argtypes = [vp, vp, vp, C.c_int]
        L.orc_state_set_sampling.argtypes = [vp, C.c_float, C.c_float, C.c_int, C.c_int]
        L.orc_state_set_noise.argtypes = [vp, vp, vp]
        L.orc_state_get_kv.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp]
        L.orc_lmgen_new.restype = vp; L.orc_lmgen_new.argtypes = [vp]
        L.orc_lmgen_free.argtypes = [vp]
        L.orc_lmgen_step.restype = C.c_int; L.orc_lmgen_step.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp]
        L.orc_lmgen_state.restype = vp; L.orc_lmgen_state.argtypes = [vp]
        L.orc_lmgen_offset.restype = C.c_int; L.orc_lmgen_offset.argtypes = [vp]
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def make_config(cfg: dict, delay_steps: int = 0) -> OrcConfig:
    c = OrcConfig()
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
    c.delay_steps = delay_steps
    c.cross_attention = 1 if cfg.get("cross_attention") else 0
    c.demux_second_stream = 1 if cfg.get("demux") else 0
    c.dep_low_rank = int(cfg.get("dep_low_rank") or 0)
    return c


# ---- T0 helpers -------------------------------------------------------------------------------
def dequantize(gtype: int, raw: np.ndarray, k: int) -> np.ndarray:
    """raw: uint8 [rows, row_bytes] -> float32 [rows, k]"""
    raw = np.ascontiguousarray(raw)
    rows = raw.shape[0]
    out = np.empty((rows, k), dtype=np.float32)
    rs = lib().orc_row_size(gtype, k)
    assert raw.shape[1] == rs, (raw.shape, rs)
    for r in range(rows):
        lib().orc_dequantize_row(gtype, raw[r].ctypes.data, out[r].ctypes.data, k)
    return out


def quantize_q8_0(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.size // 32 * 34, dtype=np.uint8)
    lib().orc_quantize_row_q8_0(_p(x), _p(out), x.size)
    return out


def quantize_q4_0(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.size // 32 * 18, dtype=np.uint8)
    lib().orc_quantize_row_q4_0(_p(x), _p(out), x.size)
    return out


def quantize_q4_K(x: np.ndarray) -> np.ndarray:
    """quantize_row_q4_K_ref restated (scale / min search per 32-element sub-block); x.size % 256 == 0"""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.size // 256 * 144, dtype=np.uint8)
    lib().orc_quantize_row_q4_K(_p(x), _p(out), x.size)
    return out


def quantize_q8_K(x: np.ndarray):
    x = np.ascontiguousarray(x, dtype=np.float32)
    qs = np.empty(x.size, dtype=np.int8); d = np.empty(x.size // 256, dtype=np.float32); bs = np.empty(x.size // 16, dtype=np.int16)
    lib().orc_quantize_row_q8_K(_p(x), _p(qs), _p(d), _p(bs), x.size)
    return qs, d, bs


def mul_mat_vec(gtype: int, w_raw: np.ndarray, k: int, x: np.ndarray, ideal: bool = False) -> np.ndarray:
    w_raw = np.ascontiguousarray(w_raw); x = np.ascontiguousarray(x, dtype=np.float32)
    rows = w_raw.shape[0]
    y = np.empty(rows, dtype=np.float32)
    fn = lib().orc_mul_mat_vec_ideal if ideal else lib().orc_mul_mat_vec
    fn(gtype, _p(w_raw), k, rows, _p(x), _p(y))
    return y


def rms_norm(x: np.ndarray, alpha: np.ndarray | None, eps: float = 1e-8) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    lib().orc_rms_norm(_p(x), _p(alpha) if alpha is not None else None, eps, _p(y), x.size)
    return y


# ---- Mimi split residual vector quantiser (mimi_rvq_ref.c) ----------------------------------------
class SplitRVQ:
    """codes <-> latent (reference quantization/vq.h:69-117).  cb_first [n_sem][bins][D] f32, cb_rest [n_rest][bins][D] f32,
    projections f16 bit patterns: in_* [D][dim], out_* [dim][D]"""

    def __init__(self, cb_first, cb_rest, in_first, in_rest, out_first, out_rest):
        self.cb_first = np.ascontiguousarray(cb_first, dtype=np.float32); self.cb_rest = np.ascontiguousarray(cb_rest, dtype=np.float32)
        self.in_first = np.ascontiguousarray(in_first, dtype=np.uint16); self.in_rest = np.ascontiguousarray(in_rest, dtype=np.uint16)
        self.out_first = np.ascontiguousarray(out_first, dtype=np.uint16); self.out_rest = np.ascontiguousarray(out_rest, dtype=np.uint16)
        self.n_sem, self.bins, self.D = self.cb_first.shape
        self.n_rest = self.cb_rest.shape[0]
        self.dim = self.in_first.shape[1]

    def encode(self, x, n_q: int):
        """x [T][dim] -> codes [n_q][T]"""
        x = np.ascontiguousarray(x, dtype=np.float32)
        T = x.shape[0]
        codes = np.zeros((n_q, T), dtype=np.int32)
        lib().orc_split_rvq_encode(_p(self.cb_first), _p(self.cb_rest), _p(self.in_first), _p(self.in_rest), self.n_sem, n_q, self.bins, self.D,
                                   self.dim, _p(x), T, _p(codes))
        return codes

    def decode(self, codes):
        """codes [K][T] -> latent [T][dim]"""
        codes = np.ascontiguousarray(codes, dtype=np.int32)
        K, T = codes.shape
        y = np.zeros((T, self.dim), dtype=np.float32)
        lib().orc_split_rvq_decode(_p(self.cb_first), _p(self.cb_rest), _p(self.out_first), _p(self.out_rest), self.n_sem, K, self.bins, self.D,
                                   self.dim, _p(codes), T, _p(y))
        return y


# ---- model ------------------------------------------------------------------------------------
class Model:
    """Oracle model over a GGUF file (tensors stay mmapped by gguf-py)."""

    def __init__(self, gguf_path: str, cfg: dict, delay_steps: int = 0, ideal: bool = False):
        from gguf import GGUFReader
        self.cfg = cfg
        self.reader = GGUFReader(gguf_path)
        self._ccfg = make_config(cfg, delay_steps)
        self.h = lib().orc_model_new(C.byref(self._ccfg))
        self._keep = []
        for t in self.reader.tensors:
            if ".condition_provider." in t.name:    # voice conditioners: used by voice_condition() below, not by the step
                continue
            shape = [int(s) for s in t.shape]       # ggml order: ne0 fastest
            ne0 = shape[0]; ne1 = shape[1] if len(shape) > 1 else 1
            data = t.data
            self._keep.append(data)
            ok = lib().orc_model_set_tensor(self.h, t.name.encode(), int(t.tensor_type), ne0, ne1, data.ctypes.data)
            if not ok:
                raise KeyError(f"oracle: tensor {t.name} not recognised")
        buf = C.create_string_buffer(128)
        n = lib().orc_model_missing(self.h, buf, 128)
        if n:
            raise KeyError(f"oracle: {n} tensors missing (first: {buf.value.decode()})")
        if ideal:
            lib().orc_model_set_ideal(self.h, 1)

    def tensor_raw(self, name: str):
        for t in self.reader.tensors:
            if t.name == name:
                return t
        raise KeyError(name)

    def close(self):
        if self.h:
            lib().orc_model_free(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class State:
    def __init__(self, model: Model):
        self.model = model
        self.h = lib().orc_state_new(model.h)
        self._own = True

    @property
    def offset(self):
        return lib().orc_state_offset(self.h)

    def reset(self):
        lib().orc_state_reset(self.h)

    def step_temporal(self, tokens):
        cfg = self.model.cfg
        tok = np.ascontiguousarray(tokens, dtype=np.int32)
        assert tok.size == cfg["n_q"] + 1
        logits = np.empty(cfg["text_card"], dtype=np.float32)
        tout = np.empty(cfg["dim"], dtype=np.float32)
        t = lib().orc_step_temporal(self.model.h, self.h, _p(tok), _p(logits), _p(tout))
        return t, logits, tout

    def step_temporal_embedding(self, x):
        cfg = self.model.cfg
        xx = np.ascontiguousarray(x, dtype=np.float32)
        assert xx.size == cfg["dim"]
        logits = np.empty(cfg["text_card"], dtype=np.float32)
        tout = np.empty(cfg["dim"], dtype=np.float32)
        t = lib().orc_step_temporal_embedding(self.model.h, self.h, _p(xx), _p(logits), _p(tout))
        return t, logits, tout

    def step_depformer(self, text_token: int, force=None):
        cfg = self.model.cfg
        toks = np.empty(cfg["dep_q"], dtype=np.int32)
        logits = np.empty((cfg["dep_q"], cfg["card"]), dtype=np.float32)
        f = np.ascontiguousarray(force, dtype=np.int32) if force is not None else None
        lib().orc_step_depformer(self.model.h, self.h, int(text_token), _p(f) if f is not None else None, _p(toks), _p(logits))
        return toks, logits

    def vad(self) -> float:
        return float(lib().orc_vad(self.model.h, self.h))

    def set_condition(self, cond_sum=None, cond_cross=None):
        """cond_sum [dim] and / or cond_cross [Tc][dim] (moshi.cpp:851-883)"""
        s = np.ascontiguousarray(cond_sum, dtype=np.float32) if cond_sum is not None else None
        c = np.ascontiguousarray(cond_cross, dtype=np.float32) if cond_cross is not None else None
        lib().orc_state_set_condition(self.h, _p(s) if s is not None else None, _p(c) if c is not None else None, 0 if c is None else c.shape[0])

    def set_sampling(self, temp_text, temp_audio, top_k_text=25, top_k_audio=250):
        lib().orc_state_set_sampling(self.h, temp_text, temp_audio, top_k_text, top_k_audio)

    def set_noise(self, noise_text, noise_audio):
        self._nt = np.ascontiguousarray(noise_text, dtype=np.float32)
        self._na = np.ascontiguousarray(noise_audio, dtype=np.float32)
        lib().orc_state_set_noise(self.h, _p(self._nt), _p(self._na))

    def get_kv(self, layer: int, head: int, slot: int):
        Dh = self.model.cfg["dim"] // self.model.cfg["num_heads"]
        k = np.empty(Dh, dtype=np.uint16); v = np.empty(Dh, dtype=np.uint16)
        lib().orc_state_get_kv(self.h, layer, head, slot, _p(k), _p(v))
        return k, v

    def __del__(self):
        try:
            if self._own and self.h:
                lib().orc_state_free(self.h); self.h = None
        except Exception:
            pass


class LMGen:
    """Greedy LMGen (lm.h:778-979) on the oracle."""

    def __init__(self, model: Model):
        self.model = model
        self.h = lib().orc_lmgen_new(model.h)

    @property
    def offset(self):
        return lib().orc_lmgen_offset(self.h)

    def step(self, in_tokens, replace: bool = False):
        cfg = self.model.cfg
        tok = np.ascontiguousarray(in_tokens, dtype=np.int32)
        text = C.c_int32(0)
        audio = np.full(max(1, cfg["dep_q"]), -7, dtype=np.int32)
        ok = lib().orc_lmgen_step(self.h, _p(tok), tok.size, int(replace), C.byref(text), _p(audio))
        return ok, int(text.value), audio[: cfg["dep_q"]].copy()

    def __del__(self):
        try:
            if self.h:
                lib().orc_lmgen_free(self.h); self.h = None
        except Exception:
            pass


# ---- voice conditioners (one-off per voice) ------------------------------------------------------------------------
def _tensor_f32(t):
    """(float32 [rows, ne0], ggml type) of a GGUFReader tensor stored as f32 / f16 / bf16"""
    raw = np.ascontiguousarray(t.data).view(np.uint8).reshape(-1)
    gt = int(t.tensor_type)
    ne0 = int(t.shape[0])
    if gt == 0:
        v = raw.view(np.float32)
    elif gt == 1:
        v = raw.view(np.float16).astype(np.float32)
    elif gt == 30:
        v = (raw.view(np.uint16).astype(np.uint32) << 16).view(np.float32)
    else:
        raise ValueError(f"{t.name}: conditioner tensors are unquantised, got type {gt}")
    return v.reshape(-1, ne0), gt


def _round_to_type(x, gt):
    """ggml_mul_mat rounds the activation to the weight's type (vec_dot_type): f16 / bf16 RNE, f32 as is"""
    x = np.asarray(x, dtype=np.float32)
    if gt == 1:
        return x.astype(np.float16).astype(np.float32)
    if gt == 30:
        u = x.view(np.uint32)
        return (((u + (0x7FFF + ((u >> 16) & 1))) >> 16).astype(np.uint32) << 16).view(np.float32)
    return x


def voice_condition(gguf_path: str, cfg: dict, speaker_wavs: np.ndarray):
    """voice_condition() of the reference (src/moshi.cpp:296-366) restated with the order-independent arithmetic of the
    rest of the oracle (exact products, double sums, one rounding).  speaker_wavs: [channels][frames] as stored in the
    voice file.  Returns (condition_sum [dim], condition_cross [5 * frames][dim]).  PARITY UNPINNED: the reference ships
    no voice fixtures; ggml's own summation order and libm cosf / sinf differ from this in the last bit."""
    from gguf import GGUFReader
    tens = {t.name: t for t in GGUFReader(gguf_path).tensors}
    cp = "lm.condition_provider.conditioners."

    def proj(wname, x):
        w, gt = _tensor_f32(tens[cp + wname])
        return (w.astype(np.float64) @ _round_to_type(x, gt).astype(np.float64)).astype(np.float32)

    cfg_emb = _tensor_f32(tens[cp + "cfg.embed.weight"])[0][2]            # cfg 2.0 -> row 2
    control_emb = _tensor_f32(tens[cp + "control.embed.weight"])[0][0]    # control "ok" -> row 0
    cond_sum = proj("cfg.output_proj.weight", cfg_emb) + proj("control.output_proj.weight", control_emb)
    wavs = np.ascontiguousarray(speaker_wavs, dtype=np.float32)
    T, dim = wavs.shape[1], cfg["dim"]
    pad = _tensor_f32(tens[cp + "speaker_wavs.learnt_padding"])[0].reshape(-1)
    cross = np.tile(pad, (5 * T, 1)).astype(np.float32)
    for t in range(T):
        cross[t] = proj("speaker_wavs.output_proj.weight", wavs[:, t])
    half = dim // 2                                                      # ggml_timestep_embedding(positions, dim, 10000)
    freq = np.empty(half, dtype=np.float32)
    lib().orc_timestep_freq(half, 10000, _p(freq))
    arg = (np.arange(5 * T, dtype=np.float32)[:, None] * freq[None, :]).astype(np.float32)
    pos = np.concatenate([np.cos(arg.astype(np.float64)), np.sin(arg.astype(np.float64))], axis=1).astype(np.float32)
    return cond_sum.astype(np.float32), (cross + pos).astype(np.float32)

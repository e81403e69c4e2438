"""Random-init GGUF writer for the LM decode step (test / bench input generator).

There is no network, so checkpoints are synthesised: every tensor the reference's loader resolves
for the LM (names / shapes / types: SURVEY.md App. B; reference lm.h:370-395, transformer.h:764-779,
1042-1080, gating.h:39-43, lm_utils.h:131-150) is filled with *valid quantised blocks* whose
dequantised values are zero-mean with std ~ 1/sqrt(K), so activations stay O(1) through 32 layers.
Like the reference's own files the GGUF carries tensors only, no key/value metadata
(loader.h:227-233); the config travels as JSON next to it.

The file format written here is GGUF v3 (magic, version, n_tensors, n_kv, tensor infos, 32-byte
aligned data).  tests/test_gguf_roundtrip.py reads it back with gguf-py's independent GGUFReader.
"""
from __future__ import annotations
import json
import os
import struct
import numpy as np

from . import configs

GGML_F32, GGML_F16, GGML_Q4_0, GGML_Q8_0, GGML_Q4_K, GGML_BF16 = 0, 1, 2, 8, 12, 30
TYPE_NAMES = {"f32": GGML_F32, "f16": GGML_F16, "q4_0": GGML_Q4_0, "q8_0": GGML_Q8_0, "q4_k": GGML_Q4_K, "bf16": GGML_BF16}
BLOCK = {GGML_F32: (1, 4), GGML_F16: (1, 2), GGML_BF16: (1, 2), GGML_Q4_0: (32, 18), GGML_Q8_0: (32, 34), GGML_Q4_K: (256, 144)}
ALIGN = 32


def row_bytes(gtype: int, k: int) -> int:
    bs, nb = BLOCK[gtype]
    assert k % bs == 0, (gtype, k)
    return k // bs * nb


def pack_q4k_scales(sc: np.ndarray, mn: np.ndarray) -> np.ndarray:
    """sc, mn: [..., 8] uint8 in [0,63] -> [..., 12] packed like ggml (see gguf/quants.py Q4_K.get_scale_min)."""
    out = np.empty(sc.shape[:-1] + (12,), dtype=np.uint8)
    out[..., 0:4] = (sc[..., 0:4] & 63) | ((sc[..., 4:8] >> 4) << 6)
    out[..., 4:8] = (mn[..., 0:4] & 63) | ((mn[..., 4:8] >> 4) << 6)
    out[..., 8:12] = (sc[..., 4:8] & 0xF) | ((mn[..., 4:8] & 0xF) << 4)
    return out


def random_q4k(rng: np.random.Generator, rows: int, k: int, std: float) -> np.ndarray:
    """Valid block_q4_K rows: w = d*sc*q - dmin*m with m ~= sc and dmin ~= 7.5 d  => zero-mean weights."""
    nb = rows * (k // 256)
    blk = np.empty((nb, 144), dtype=np.uint8)
    sc = rng.integers(8, 64, size=(nb, 8), dtype=np.uint8)
    mn = np.clip(sc.astype(np.int16) + rng.integers(-2, 3, size=(nb, 8), dtype=np.int16), 0, 63).astype(np.uint8)
    d0 = std / (39.0 * 4.61)
    d = (d0 * rng.uniform(0.5, 1.5, size=nb)).astype(np.float16)
    dmin = (d.astype(np.float32) * 7.5 * rng.uniform(0.97, 1.03, size=nb)).astype(np.float16)
    blk[:, 0:2] = d.view(np.uint8).reshape(nb, 2)
    blk[:, 2:4] = dmin.view(np.uint8).reshape(nb, 2)
    blk[:, 4:16] = pack_q4k_scales(sc, mn)
    blk[:, 16:144] = rng.integers(0, 256, size=(nb, 128), dtype=np.uint8)
    return blk.reshape(rows, -1)


def random_q8_0(rng: np.random.Generator, rows: int, k: int, std: float) -> np.ndarray:
    nb = rows * (k // 32)
    blk = np.empty((nb, 34), dtype=np.uint8)
    d = (std / 73.3 * rng.uniform(0.5, 1.5, size=nb)).astype(np.float16)
    blk[:, 0:2] = d.view(np.uint8).reshape(nb, 2)
    q = rng.integers(-127, 128, size=(nb, 32), dtype=np.int8)
    blk[:, 2:34] = q.view(np.uint8)
    return blk.reshape(rows, -1)


def random_q4_0(rng: np.random.Generator, rows: int, k: int, std: float) -> np.ndarray:
    nb = rows * (k // 32)
    blk = np.empty((nb, 18), dtype=np.uint8)
    d = (std / 4.61 * rng.uniform(0.5, 1.5, size=nb)).astype(np.float16)
    blk[:, 0:2] = d.view(np.uint8).reshape(nb, 2)
    blk[:, 2:18] = rng.integers(0, 256, size=(nb, 16), dtype=np.uint8)
    return blk.reshape(rows, -1)


def random_tensor(rng, gtype: int, rows: int, k: int, std: float) -> np.ndarray:
    """Returns a uint8 array [rows, row_bytes]."""
    if gtype == GGML_Q4_K:
        return random_q4k(rng, rows, k, std)
    if gtype == GGML_Q8_0:
        return random_q8_0(rng, rows, k, std)
    if gtype == GGML_Q4_0:
        return random_q4_0(rng, rows, k, std)
    x = (rng.standard_normal(size=(rows, k), dtype=np.float32) * np.float32(std)).astype(np.float32)
    if gtype == GGML_F32:
        return x.view(np.uint8).reshape(rows, -1)
    if gtype == GGML_F16:
        return x.astype(np.float16).view(np.uint8).reshape(rows, -1)
    if gtype == GGML_BF16:
        u = x.view(np.uint32)
        u = ((u + (0x7FFF + ((u >> 16) & 1))) >> 16).astype(np.uint16)
        return u.view(np.uint8).reshape(rows, -1)
    raise ValueError(gtype)


COND_EMBED, COND_CHANNELS = 64, 96       # conditioner embedding width / speaker_wavs channels of the synthetic TTS preset


def linear_type(qtype: int, k: int) -> int:
    """loader.h:162-173: Q4_K needs K%256==0 else Q4_0, Q4_0/Q8_0 need K%32==0 else source dtype (bf16)."""
    if qtype == GGML_Q4_K and k % 256:
        qtype = GGML_Q4_0
    if qtype in (GGML_Q4_0, GGML_Q8_0) and k % 32:
        qtype = GGML_BF16
    return qtype


def emb_type(qtype: int, k: int) -> int:
    """lm_utils.h:131-147: embedding tables are Q4_0 when the model is q4_k."""
    if qtype == GGML_Q4_K:
        qtype = GGML_Q4_0
    return linear_type(qtype, k)


def manifest(cfg: dict, qtype: int):
    """[(name, gtype, ne0=K, ne1=rows, std)] in file order."""
    d, L, F = cfg["dim"], cfg["num_layers"], cfg["hidden"]
    out = []
    lin = lambda name, k, rows: out.append((name, linear_type(qtype, k), k, rows, 1.0 / np.sqrt(k)))
    emb = lambda name, k, rows: out.append((name, emb_type(qtype, k), k, rows, 0.25))
    f32 = lambda name, k: out.append((name, GGML_F32, k, 1, None))
    bias = lambda name, k: out.append((name, GGML_F32, k, 1, -0.1))        # std < 0: plain N(0, |std|) f32 vector
    cross, demux, lr = cfg.get("cross_attention"), cfg.get("demux"), cfg.get("dep_low_rank") or 0
    emb("lm.text_emb.weight", d, cfg["text_card"] + 1)
    if demux:                                                              # lm_utils.h:14-40
        lin("lm.text_emb.out1.weight", d, d)
        lin("lm.text_emb.out2.weight", d, d)
    for c in range(cfg["n_q"]):
        emb(f"lm.emb.{c}.weight", d, cfg["card"] + 1)
    for i in range(L):
        p = f"lm.transformer.layers.{i}."
        f32(p + "norm1.alpha", d)
        lin(p + "self_attn.in_projs.0.weight", d, 3 * d)
        lin(p + "self_attn.out_projs.0.weight", d, d)
        if cross:                                                          # transformer.h:1053-1056
            f32(p + "norm_cross.weight", d)
            bias(p + "norm_cross.bias", d)
            lin(p + "cross_attention.in_projs.0.weight", d, 3 * d)
            lin(p + "cross_attention.out_projs.0.weight", d, d)
        f32(p + "norm2.alpha", d)
        lin(p + "gating.linear_in.weight", d, 2 * F)
        lin(p + "gating.linear_out.weight", F, d)
    f32("lm.out_norm.alpha", d)
    lin("lm.text_linear.weight", d, cfg["text_card"])
    if cfg["dep_q"] > 0:
        dd, Fd, nw = cfg["depformer_dim"], cfg["dep_hidden"], configs.dep_num_weights(cfg)
        for k in range(nw):
            lin(f"lm.depformer_in.{k}.weight", d, dd)
        de = lr if lr else dd                                              # low-rank tables are [lr, rows] + a [lr -> dd] linear
        emb("lm.depformer_text_emb.weight", de, cfg["text_card"] + 1)
        if demux:
            lin("lm.depformer_text_emb.out1.weight", de, dd)
            lin("lm.depformer_text_emb.out2.weight", de, dd)
        elif lr:
            lin("lm.depformer_text_emb.low_rank.weight", lr, dd)
        for k in range(cfg["dep_q"] - 1):
            emb(f"lm.depformer_emb.{k}.weight", de, cfg["card"] + 1)
            if lr:
                lin(f"lm.depformer_emb.{k}.low_rank.weight", lr, dd)
        for i in range(cfg["depformer_num_layers"]):
            p = f"lm.depformer.layers.{i}."
            f32(p + "norm1.alpha", dd)
            f32(p + "norm2.alpha", dd)
            for k in range(nw):
                lin(p + f"self_attn.in_projs.{k}.weight", dd, 3 * dd)
                lin(p + f"self_attn.out_projs.{k}.weight", dd, dd)
                lin(p + f"gating.{k}.linear_in.weight", dd, 2 * Fd)
                lin(p + f"gating.{k}.linear_out.weight", Fd, dd)
        for k in range(cfg["dep_q"]):
            lin(f"lm.linears.{k}.weight", dd, cfg["card"])
    for j in range(cfg["extra_heads"]):
        lin(f"lm.extra_heads.{j}.weight", d, cfg["extra_heads_dim"])
    if cfg.get("conditioners"):                                            # tts.h:16-35; never quantised (fetched as stored)
        cp, E, C = "lm.condition_provider.conditioners.", COND_EMBED, COND_CHANNELS
        out.append((cp + "cfg.embed.weight", GGML_F32, E, 7, 0.5))
        bias(cp + "cfg.learnt_padding", d)
        out.append((cp + "cfg.output_proj.weight", GGML_BF16, E, d, 1.0 / np.sqrt(E)))
        out.append((cp + "control.embed.weight", GGML_BF16, E, 2, 0.5))
        bias(cp + "control.learnt_padding", d)
        out.append((cp + "control.output_proj.weight", GGML_F32, E, d, 1.0 / np.sqrt(E)))
        bias(cp + "speaker_wavs.learnt_padding", d)
        out.append((cp + "speaker_wavs.output_proj.weight", GGML_F16, C, d, 1.0 / np.sqrt(C)))
    return out


def write_gguf(path: str, cfg: dict, quant: str = "q4_k", seed: int = 1234) -> dict:
    """Write a random-init GGUF + `<path>.json` config. Returns {'bytes': file size, 'tensors': n}."""
    qtype = TYPE_NAMES[quant]
    man = manifest(cfg, qtype)
    rng = np.random.default_rng(seed)
    infos, offs = [], 0
    for name, gtype, k, rows, _ in man:
        nbytes = row_bytes(gtype, k) * rows
        infos.append((name, gtype, k, rows, offs, nbytes))
        offs += (nbytes + ALIGN - 1) // ALIGN * ALIGN
    hdr = bytearray()
    hdr += struct.pack("<IIQQ", 0x46554747, 3, len(man), 0)
    for name, gtype, k, rows, off, _ in infos:
        nb = name.encode()
        hdr += struct.pack("<Q", len(nb)) + nb
        if rows == 1 and gtype == GGML_F32:
            hdr += struct.pack("<IQ", 1, k)
        else:
            hdr += struct.pack("<IQQ", 2, k, rows)
        hdr += struct.pack("<IQ", gtype, off)
    pad = (-len(hdr)) % ALIGN
    hdr += b"\0" * pad
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(hdr)
        for (name, gtype, k, rows, std), (_, _, _, _, off, nbytes) in zip(man, infos):
            if std is None:  # norm alpha: around 1
                data = (1.0 + 0.1 * rng.standard_normal(size=k)).astype(np.float32).view(np.uint8)
                f.write(data.tobytes())
            elif std < 0:    # f32 bias vector
                f.write((-std * rng.standard_normal(size=k)).astype(np.float32).tobytes())
            else:
                # generate in row chunks to bound peak memory on the 7B model
                chunk = max(1, (64 << 20) // max(1, row_bytes(gtype, k)))
                for r0 in range(0, rows, chunk):
                    r1 = min(rows, r0 + chunk)
                    f.write(random_tensor(rng, gtype, r1 - r0, k, std).tobytes())
            f.write(b"\0" * ((-nbytes) % ALIGN))
    os.replace(tmp, path)
    with open(path + ".json", "w") as f:
        json.dump({"preset": cfg, "config": configs.to_config_json(cfg), "quant": quant, "seed": seed}, f)
    return {"bytes": len(hdr) + offs, "tensors": len(man)}


def write_gguf_tensors(path: str, tensors) -> None:
    """Write a GGUF v3 file from explicit data: tensors = [(name, gtype, k, rows, uint8 array of rows*row_bytes)] (rows == 1 and
    F32 -> 1-D tensor like the norm vectors).  Used to build a quantised twin of an unquantised model in the tests."""
    infos, offs = [], 0
    for name, gtype, k, rows, data in tensors:
        nbytes = row_bytes(gtype, k) * rows
        assert data.size == nbytes, (name, data.size, nbytes)
        infos.append((name, gtype, k, rows, offs, nbytes))
        offs += (nbytes + ALIGN - 1) // ALIGN * ALIGN
    hdr = bytearray()
    hdr += struct.pack("<IIQQ", 0x46554747, 3, len(tensors), 0)
    for name, gtype, k, rows, off, _ in infos:
        nb = name.encode()
        hdr += struct.pack("<Q", len(nb)) + nb
        if rows == 1 and gtype == GGML_F32:
            hdr += struct.pack("<IQ", 1, k)
        else:
            hdr += struct.pack("<IQQ", 2, k, rows)
        hdr += struct.pack("<IQ", gtype, off)
    hdr += b"\0" * ((-len(hdr)) % ALIGN)
    with open(path, "wb") as f:
        f.write(hdr)
        for (_, _, _, _, data), (_, _, _, _, _, nbytes) in zip(tensors, infos):
            f.write(np.ascontiguousarray(data).tobytes())
            f.write(b"\0" * ((-nbytes) % ALIGN))


def cached_gguf(preset: str, quant: str = "q4_k", seed: int = 1234, root: str | None = None) -> str:
    """Generate once per (preset, quant, seed) under $MSX_CACHE (default /tmp/msx_cache)."""
    root = root or os.environ.get("MSX_CACHE", "/tmp/msx_cache")
    os.makedirs(root, exist_ok=True)
    path = os.path.join(root, f"{preset}-{quant}-s{seed}.gguf")
    if not os.path.exists(path):
        write_gguf(path, configs.get(preset), quant, seed)
    return path


def write_safetensors(path, tensors):
    """tensors: [(name, dtype string, shape, bytes)] -> the safetensors container (8-byte header length, JSON, data)"""
    import json
    hdr, off = {"__metadata__": {"format": "pt", "note": "written by tests, {nested: [1, 2]}"}}, 0
    for name, dtype, shape, raw in tensors:
        hdr[name] = {"dtype": dtype, "shape": [int(v) for v in shape], "data_offsets": [off, off + len(raw)]}
        off += len(raw)
    js = json.dumps(hdr).encode()
    js += b" " * (-len(js) % 8)
    with open(path, "wb") as f:
        f.write(len(js).to_bytes(8, "little")); f.write(js)
        for _, _, _, raw in tensors:
            f.write(raw)

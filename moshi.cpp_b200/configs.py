"""Model shape presets for the LM decode step.

Field names follow moshi_config_t (reference include/moshi/moshi.h:111-156) and the shipped
tools/moshi-config.json / tools/personaplex-config.json.  `hidden` / `dep_hidden` are NOT config
fields in the reference (the weights carry them, moshi.h:132); they only parameterise the
random-init GGUF writer.  The stt/tts shapes come from upstream model cards quoted in
SURVEY.md §8d (not present in the reference tree) and are marked unverified there.
"""
from __future__ import annotations
import copy

_DELAYS_7B = [0, 0, 1, 1, 1, 1, 1, 1, 1, 0, 1, 1, 1, 1, 1, 1, 1]

MOSHI_7B = dict(
    name="moshi7b", model_type="moshi",
    dim=4096, num_heads=32, num_layers=32, context=3000, max_period=10000,
    n_q=16, dep_q=8, card=2048, text_card=32000, delays=_DELAYS_7B,
    hidden=11264,
    depformer_dim=1024, depformer_num_heads=16, depformer_num_layers=6, depformer_context=8,
    depformer_max_period=0,  # depformer_pos_emb "none"
    dep_hidden=2816, schedule=[], extra_heads=0, extra_heads_dim=0,
)


def _derive(base, **kw):
    c = copy.deepcopy(base)
    c.update(kw)
    return c


PRESETS = {
    "moshi7b": MOSHI_7B,
    # identical shapes, fewer temporal layers: full-size kernels, oracle finishes in seconds
    "moshi7b_l2": _derive(MOSHI_7B, name="moshi7b_l2", num_layers=2),
    "moshi7b_l4": _derive(MOSHI_7B, name="moshi7b_l4", num_layers=4),
    "personaplex7b": _derive(MOSHI_7B, name="personaplex7b", model_type="personaplex", dep_q=16),
    # small end-to-end model: same topology (17 codebooks, delays, depformer per-step weights),
    # ring capacity 24 so a 250-frame run wraps the KV ring many times
    "tiny": dict(
        name="tiny", model_type="moshi",
        dim=512, num_heads=4, num_layers=3, context=24, max_period=10000,
        n_q=16, dep_q=8, card=256, text_card=1000, delays=_DELAYS_7B,
        hidden=768,
        depformer_dim=256, depformer_num_heads=4, depformer_num_layers=2, depformer_context=8,
        depformer_max_period=0, dep_hidden=512, schedule=[], extra_heads=0, extra_heads_dim=0,
    ),
    # personaplex-like tiny: dep_q 16 > depformer_context 8 (depformer ring wraps inside a frame)
    "tiny_pplex": dict(
        name="tiny_pplex", model_type="personaplex",
        dim=512, num_heads=4, num_layers=2, context=24, max_period=10000,
        n_q=16, dep_q=16, card=256, text_card=1000, delays=_DELAYS_7B,
        hidden=768,
        depformer_dim=256, depformer_num_heads=4, depformer_num_layers=2, depformer_context=8,
        depformer_max_period=0, dep_hidden=512, schedule=[], extra_heads=0, extra_heads_dim=0,
    ),
    # depformer_weights_per_step_schedule (lm.h:457-462, transformer.h:74-83): 8 codebook steps share 4 weight sets in a
    # non-identity order; depformer_in / in_projs / out_projs / gating pick schedule[k], linears / embeddings stay per step
    "tiny_sched": dict(
        name="tiny_sched", model_type="moshi",
        dim=512, num_heads=4, num_layers=2, context=24, max_period=10000,
        n_q=16, dep_q=8, card=256, text_card=1000, delays=_DELAYS_7B,
        hidden=768,
        depformer_dim=256, depformer_num_heads=4, depformer_num_layers=2, depformer_context=8,
        depformer_max_period=0, dep_hidden=512, schedule=[0, 1, 2, 3, 3, 1, 0, 2], extra_heads=0, extra_heads_dim=0,
    ),
    # PersonaPlex shapes (dep_q 16), 2 temporal layers, ring of 1100 slots: fills in seconds on the CPU oracle and takes the
    # long-ring attention path (> 1024 slots) at full 7B head / layer shapes
    "pplex7b_l2_c1100": _derive(MOSHI_7B, name="pplex7b_l2_c1100", model_type="personaplex", dep_q=16, num_layers=2, context=1100),
    # STT-like: no depformer, 32 input codebooks, extra heads (VAD)
    "tiny_stt": dict(
        name="tiny_stt", model_type="stt",
        dim=512, num_heads=4, num_layers=2, context=20, max_period=10000,
        n_q=32, dep_q=0, card=256, text_card=500, delays=[0] + [6] * 32,
        hidden=768,
        depformer_dim=0, depformer_num_heads=0, depformer_num_layers=0, depformer_context=0,
        depformer_max_period=0, dep_hidden=0, schedule=[], extra_heads=4, extra_heads_dim=6,
    ),
    "stt1b": dict(
        name="stt1b", model_type="stt",
        dim=2048, num_heads=16, num_layers=16, context=750, max_period=100000,
        n_q=32, dep_q=0, card=2048, text_card=8000, delays=[0] + [6] * 32,
        hidden=5632,
        depformer_dim=0, depformer_num_heads=0, depformer_num_layers=0, depformer_context=0,
        depformer_max_period=0, dep_hidden=0, schedule=[], extra_heads=4, extra_heads_dim=6,
    ),
    # TTS-like: n_q == dep_q (no user stream), cross-attention to a conditioning memory, demuxed two-stream text
    # embedding, low-rank (128 -> Q4_0 in a q4_k model) depformer embeddings (moshi.h:111-156)
    "tiny_tts": dict(
        name="tiny_tts", model_type="tts",
        dim=512, num_heads=4, num_layers=2, context=20, max_period=10000,
        n_q=8, dep_q=8, card=256, text_card=500, delays=[0, 0, 1, 1, 1, 1, 1, 1, 1],
        hidden=768,
        depformer_dim=256, depformer_num_heads=4, depformer_num_layers=2, depformer_context=8,
        depformer_max_period=0, dep_hidden=512, schedule=[], extra_heads=0, extra_heads_dim=0,
        cross_attention=True, demux=True, dep_low_rank=128,
    ),
    # low-rank depformer embeddings without demux (the other moshi_scaled_embedding_t variant)
    "tiny_lowrank": dict(
        name="tiny_lowrank", model_type="moshi",
        dim=512, num_heads=4, num_layers=2, context=24, max_period=10000,
        n_q=16, dep_q=8, card=256, text_card=1000, delays=_DELAYS_7B,
        hidden=768,
        depformer_dim=256, depformer_num_heads=4, depformer_num_layers=2, depformer_context=8,
        depformer_max_period=0, dep_hidden=512, schedule=[], extra_heads=0, extra_heads_dim=0,
        dep_low_rank=64,
    ),
    "tts1_6b": dict(
        name="tts1_6b", model_type="tts",
        dim=2048, num_heads=16, num_layers=16, context=500, max_period=10000,
        n_q=32, dep_q=32, card=2048, text_card=8000, delays=[0] + [1] * 16 + [2] * 16,
        hidden=8448,
        depformer_dim=1024, depformer_num_heads=16, depformer_num_layers=4, depformer_context=32,
        depformer_max_period=0, dep_hidden=2816, schedule=[], extra_heads=0, extra_heads_dim=0,
        cross_attention=True, demux=True, dep_low_rank=128,
    ),
}
# tiny_tts whose GGUF also carries the voice conditioners (tts.h:16-35): voice files go through msx_stream_set_voice
PRESETS["tiny_tts_voice"] = _derive(PRESETS["tiny_tts"], name="tiny_tts_voice", conditioners=True)
for _c in PRESETS.values():
    _c.setdefault("cross_attention", False); _c.setdefault("demux", False); _c.setdefault("dep_low_rank", 0)


def get(name: str) -> dict:
    return copy.deepcopy(PRESETS[name])


def dep_num_weights(cfg: dict) -> int:
    """lm_default.h:72-83"""
    if cfg["dep_q"] <= 0:
        return 0
    if cfg["schedule"]:
        return max(cfg["schedule"]) + 1
    return cfg["dep_q"]


def to_config_json(cfg: dict) -> dict:
    """The config.json the reference tools would read next to the GGUF (tools/moshi-config.json)."""
    return {
        "card": cfg["card"], "n_q": cfg["n_q"], "dep_q": cfg["dep_q"], "delays": cfg["delays"],
        "dim": cfg["dim"], "text_card": cfg["text_card"], "existing_text_padding_id": 3,
        "num_heads": cfg["num_heads"], "num_layers": cfg["num_layers"], "hidden_scale": 4.125,
        "causal": True, "layer_scale": None, "context": cfg["context"], "max_period": cfg["max_period"],
        "gating": "silu", "norm": "rms_norm_f32", "positional_embedding": "rope",
        "depformer_dim": cfg["depformer_dim"], "depformer_num_heads": cfg["depformer_num_heads"],
        "depformer_num_layers": cfg["depformer_num_layers"], "depformer_multi_linear": True,
        "depformer_context": cfg["depformer_context"], "depformer_max_period": cfg["depformer_max_period"],
        "depformer_gating": "silu", "depformer_pos_emb": "rope" if cfg["depformer_max_period"] else "none",
        "depformer_weights_per_step": True,
        "depformer_weights_per_step_schedule": cfg["schedule"] or None,
        "conditioners": {}, "cross_attention": bool(cfg.get("cross_attention")), "model_type": cfg["model_type"],
        "demux_second_stream": bool(cfg.get("demux")), "depformer_low_rank_embeddings": cfg.get("dep_low_rank") or None,
        "tts_config": {"audio_delay": 1.28, "second_stream_ahead": 2} if cfg["model_type"] == "tts" else None,
        "extra_heads_num_heads": cfg["extra_heads"], "extra_heads_dim": cfg["extra_heads_dim"],
    }

"""invertible_cd_b200 — B200-native (sm_100a) implementation of the iCD few-step inversion/generation hot path.

Drop-in Python surface of yandex-research/invertible-cd's `utils/` for that path:
    loading.load_models / load_models_xl, generation.Generator / runner, inversion.invert,
    p2p.{AttentionStore, make_controller, register_attention_control, ...}, generation_sdxl.sample_deterministic /
    inverse_sample_deterministic, dist_utils.init
over hand-written CUDA kernels behind a C ABI (include/icd_b200.h, libicd_b200.so).
"""
__version__ = "0.1.0"

from . import arch  # noqa: F401  (pure-python; safe without a GPU)


def __getattr__(name):
    import importlib
    if name in ("ops", "unet", "loading", "generation", "generation_sdxl", "inversion", "p2p", "seq_aligner",
                "dist_utils", "schedulers", "packing", "_lib"):
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)

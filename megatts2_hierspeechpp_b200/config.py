"""Hyper-parameters of the path's models (the ``model`` section of the reference's config.json files).

``HIER_CFG``: hierspeechpp_libritts960 vocoder (config not shipped with the reference; reconstructed and
cross-checked in SURVEY.md §0.4 / Appendix B.1) — kwargs of ``Generator``.
``SR_CFG``: speechsr24k/config.json:38-46 == speechsr48k/config.json:38-46 — kwargs of ``SpeechSR24/48``.
"""
import json

HIER_CFG = dict(initial_channel=192, resblock_kernel_sizes=[3, 7, 11],
                resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], upsample_rates=[4, 5, 4, 2, 2],
                upsample_initial_channel=512, upsample_kernel_sizes=[8, 11, 8, 4, 4], gin_channels=256)

SR_CFG = dict(resblock="0", resblock_kernel_sizes=[3, 7, 11],
              resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], upsample_rates=[3],
              upsample_initial_channel=32, upsample_kernel_sizes=[3], use_spectral_norm=False)

HOP = 320            # prod(upsample_rates): 50 Hz frames -> 16 kHz samples
SAMPLE_RATE = 16000


def load_model_config(path: str) -> dict:
    """Read the ``model`` section of a reference config.json (utils.get_hparams_from_file, utils.py:210-216)."""
    with open(path, "r") as f:
        return dict(json.load(f)["model"])

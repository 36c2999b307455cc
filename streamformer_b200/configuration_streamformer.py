"""StreamformerConfig — field-for-field mirror of the reference configuration
(models/configuration_streamformer.py:29-137) so that existing ``config.json`` files load unchanged.
Extra, optional keys (ignored by the reference) tune the B200 runtime only.
"""
from transformers.configuration_utils import PretrainedConfig


class StreamformerConfig(PretrainedConfig):
    model_type = "timesformer"

    def __init__(
        self,
        image_size=224,
        patch_size=16,
        num_channels=3,
        num_frames=16,
        hidden_size=768,
        num_hidden_layers=12,
        num_attention_heads=12,
        intermediate_size=3072,
        hidden_act="gelu",
        hidden_dropout_prob=0.0,
        attention_probs_dropout_prob=0.0,
        initializer_range=0.02,
        layer_norm_eps=1e-6,
        qkv_bias=True,
        attention_type="divided_space_time",
        drop_path_rate=0,
        clip_config=None,
        enable_causal_temporal=False,
        add_lora_spatial=False,
        # ---- B200 runtime knobs (not in the reference) ----
        compute_dtype="bfloat16",      # used when the parameters are fp32: "bfloat16" | "float16"
        fold_temporal_proj=True,       # pre-multiply temporal_dense . temporal out-proj at bind time
        kv_cache_max_frames=64,        # capacity of an auto-created streaming cache
        training_recompute="auto",     # backward: "never" keeps every layer's intermediates (35 GB at 32 clips/GPU),
                                       # "always" recomputes them layer by layer, "auto" decides from free memory
        **kwargs,
    ):
        super().__init__(**kwargs)
        self.image_size = image_size
        self.patch_size = patch_size
        self.num_channels = num_channels
        self.num_frames = num_frames
        self.hidden_size = hidden_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.intermediate_size = intermediate_size
        self.hidden_act = hidden_act
        self.hidden_dropout_prob = hidden_dropout_prob
        self.attention_probs_dropout_prob = attention_probs_dropout_prob
        self.initializer_range = initializer_range
        self.layer_norm_eps = layer_norm_eps
        self.qkv_bias = qkv_bias
        self.attention_type = attention_type
        self.drop_path_rate = drop_path_rate
        self.clip_config = clip_config
        self.enable_causal_temporal = enable_causal_temporal
        self.add_lora_spatial = add_lora_spatial
        self.compute_dtype = compute_dtype
        self.fold_temporal_proj = fold_temporal_proj
        self.kv_cache_max_frames = kv_cache_max_frames
        self.training_recompute = training_recompute

"""Hyper-parameters of the hot-path networks (values of the reference's kandinsky2/configs.py:64-163 that the
denoising path consumes: UNet model_config, diffusion_config, MoVQ ddconfig).  Prior / CLIP / text-encoder
sections are omitted: those stages are outside the path."""

_MOVQ_DD = {
    "double_z": False, "z_channels": 4, "resolution": 256, "in_channels": 3, "out_ch": 3, "ch": 128,
    "ch_mult": [1, 2, 2, 4], "num_res_blocks": 2, "attn_resolutions": [32], "dropout": 0.0,
}

_DIFFUSION = {
    "learn_sigma": True, "sigma_small": False, "steps": 1000, "noise_schedule": "linear", "timestep_respacing": "",
    "use_kl": False, "predict_xstart": False, "rescale_timesteps": True, "rescale_learned_sigmas": True,
    "linear_start": 0.00085, "linear_end": 0.012,
}

_UNET = {
    "version": "2.1", "image_size": 64, "num_channels": 384, "num_res_blocks": 3, "channel_mult": "",
    "num_heads": 1, "num_head_channels": 64, "num_heads_upsample": -1, "attention_resolutions": "32,16,8",
    "dropout": 0, "model_dim": 768, "use_scale_shift_norm": True, "resblock_updown": True, "use_fp16": True,
    "cache_text_emb": True, "text_encoder_in_dim1": 1024, "text_encoder_in_dim2": 768, "image_encoder_in_dim": 768,
    "num_image_embs": 10, "pooling_type": "from_model", "in_channels": 4, "out_channels": 8,
    "use_flash_attention": False,
}

CONFIG_2_1 = {
    "image_enc_params": {"name": "MOVQ", "scale": 1, "ckpt_path": "",
                         "params": {"embed_dim": 4, "n_embed": 16384, "ddconfig": _MOVQ_DD}},
    "model_config": _UNET,
    "diffusion_config": _DIFFUSION,
}

# Kandinsky 2.2 decoder (diffusers UNet2DConditionModel config of kandinsky-community/kandinsky-2-2-decoder):
# same backbone; conditioning = 1280-d CLIP-bigG image embedding -> 32 context tokens + time-embedding add.
CONFIG_2_2 = {
    "image_enc_params": CONFIG_2_1["image_enc_params"],
    "model_config": dict(_UNET, version="2.2", image_encoder_in_dim=1280, num_image_embs=32),
    "diffusion_config": dict(_DIFFUSION, rescale_timesteps=False),
}

"""Task pipelines with the reference's call surface, decoder side computed by libk2b200.so.

Mirrors Kandinsky2_1 (kandinsky2/kandinsky2_1_model.py:21-548) and Kandinsky2_2 (kandinsky2/kandinsky2_2_model.py:15-173):
same method names, keyword arguments, defaults and return type (list of PIL images).  What runs here is the
hot path: latent init -> `num_steps` x [CFG-doubled UNet + scheduler update] -> MoVQ decode -> uint8.

The stages BEFORE the path (CLIP text/image towers, the diffusion prior, the XLM-R text encoder) are outside the
scope of this build: they enter through an `embedder` object.  The
default SyntheticEmbedder draws deterministic N(0,1) embeddings keyed by the prompt text, which is what the
benchmark configurations specify (BASELINE.json: "synthetic CLIP embeds"); a real deployment passes an embedder
wrapping its prior / encoders.  img2img / inpainting take a PIL image (encoded by the MoVQ encoder) or an already-encoded latent tensor.
"""
import hashlib
import math

import torch

from . import ops, parallel
from ._native import K2Error
from .model.gaussian_diffusion import DDIMSampler, PLMSSampler, create_ddpm_v22, create_gaussian_diffusion
from .model.model_creation import create_model
from .utils import prepare_image, prepare_mask, q_sample, uint8_to_pil
from .vqgan import MOVQ


class SyntheticEmbedder:
    """Deterministic stand-in for prior + encoders: N(0,1) tensors seeded by sha256(prompt)."""

    def __init__(self, image_dim, text_dim=1024, pooled_dim=768, text_len=77, seed=0):
        self.image_dim, self.text_dim, self.pooled_dim, self.text_len, self.seed = image_dim, text_dim, pooled_dim, text_len, seed

    def _gen(self, key):
        h = int.from_bytes(hashlib.sha256(f"{self.seed}:{key}".encode()).digest()[:7], "little")
        return torch.Generator().manual_seed(h)

    def image_emb(self, prompt, batch_size):
        """[batch_size, image_dim]: what generate_clip_emb / the prior pipeline returns for `prompt`."""
        return torch.randn(1, self.image_dim, generator=self._gen(("img", prompt))).repeat(batch_size, 1)

    def zero_image_emb(self, batch_size):
        """CLIP embedding of a black image (create_zero_img_emb, kandinsky2_1_model.py:295-297) / negative embeds."""
        return torch.randn(1, self.image_dim, generator=self._gen(("img", "<zero>"))).repeat(batch_size, 1)

    def text_emb(self, prompt, batch_size):
        """(full_emb [2B, text_len, text_dim], pooled_emb [2B, pooled_dim]): cond rows then uncond rows
        (encode_text, kandinsky2_1_model.py:115-157)."""
        def one(p):
            g = self._gen(("txt", p))
            return torch.randn(1, self.text_len, self.text_dim, generator=g), torch.randn(1, self.pooled_dim, generator=g)
        fc, pc = one(prompt)
        fu, pu = one("")
        return (torch.cat([fc.repeat(batch_size, 1, 1), fu.repeat(batch_size, 1, 1)]),
                torch.cat([pc.repeat(batch_size, 1), pu.repeat(batch_size, 1)]))

    def interpolate(self, items, weights, batch_size):
        """Weighted mix of the embeddings of prompts / images (mix_images)."""
        acc = None
        for it, w in zip(items, weights):
            key = it if isinstance(it, str) else ("pil", getattr(it, "size", None), hashlib.sha256(
                it.tobytes() if hasattr(it, "tobytes") else repr(it).encode()).hexdigest())
            e = torch.randn(1, self.image_dim, generator=self._gen(("img", key))) * w
            acc = e if acc is None else acc + e
        return acc.repeat(batch_size, 1)


def _new_h_w_latent_21(h, w):  # kandinsky2_1_model.py:106-113 (latent side, /8)
    return math.ceil(h / 64) * 8, math.ceil(w / 64) * 8


class _DecoderBase:
    version = None

    def __init__(self, config, device, task_type="text2img", embedder=None, unet_state_dict=None, movq_state_dict=None,
                 seed=0):
        if not str(device).startswith("cuda"):
            raise K2Error("k2b200 pipelines run on a CUDA sm_100 device only (no CPU fallback)")
        self.config = config
        self.device = torch.device(device)
        self.task_type = task_type
        self.use_fp16 = True
        mc = dict(config["model_config"])
        if task_type == "controlnet":  # Kandinsky 2.2 ControlNet-depth (BASELINE configs[4]): 4 latent + 4 hint-feature channels
            mc.update(in_channels=mc["in_channels"] + 4, hint_channels=4)
        self.model = create_model(**mc, up=False, inpainting=(task_type == "inpainting"), device=self.device,
                                  param_dtype=torch.float16)
        if unet_state_dict is not None:
            self.model.load_state_dict(unet_state_dict)
        else:
            self.model.init_synthetic_(seed)  # no checkpoint offline: random weights of the architecture
        self.model.convert_to_fp16()
        ie = config["image_enc_params"]
        self.scale = ie["scale"]
        self.image_encoder = MOVQ(**ie["params"], device=self.device, param_dtype=torch.float16)
        if movq_state_dict is not None:
            self.image_encoder.load_state_dict(movq_state_dict)
        else:
            self.image_encoder.init_synthetic_(seed + 1)
        self.embedder = embedder or SyntheticEmbedder(mc["image_encoder_in_dim"], mc["text_encoder_in_dim1"],
                                                      mc["text_encoder_in_dim2"])
        self.base_seed = 1234

    # shared tail: decode + crop + uint8 + PIL (kandinsky2_1_model.py:286-292)
    def _finish(self, latents, h, w):
        u8 = self.image_encoder.decode_to_uint8(latents / self.scale, crop_h=h, crop_w=w)
        return uint8_to_pil(u8)

    def _encode_image(self, image, h, w):
        """PIL image (resized to (w, h), utils.py:33-39) or image tensor [1,3,H,W] in [-1,1] -> latent via the MoVQ
        encoder (kandinsky2_1_model.py:458-461); a [1,4,h/8,w/8] tensor is taken as an already-encoded latent."""
        if torch.is_tensor(image):
            if image.shape[1] == self.image_encoder.embed_dim:
                return image.float().to(self.device)
            return self.image_encoder.encode(image.to(self.device))
        return self.image_encoder.encode(prepare_image(image, w=w, h=h).to(self.device))

    def _shard(self, batch_size):
        rank, ws = parallel.world()
        lo, hi = parallel.shard_range(batch_size, rank, ws)
        return rank, ws, lo, hi

    def _latents(self, lo, hi, shape):
        return parallel.sample_noise(range(lo, hi), shape, base_seed=self.base_seed, device=self.device)

    def _generators(self, lo, hi):
        """One device RNG stream per GLOBAL sample index (step noise independent of world size / batch position)."""
        return [torch.Generator(device=self.device).manual_seed(self.base_seed * 7919 + gi) for gi in range(lo, hi)]


class Kandinsky2_1(_DecoderBase):
    version = "2.1"

    def get_new_h_w(self, h, w):
        return _new_h_w_latent_21(h, w)

    @torch.no_grad()
    def generate_img(self, prompt, img_prompt, batch_size=1, diffusion=None, guidance_scale=7, init_step=None,
                     noise=None, init_img=None, img_mask=None, h=512, w=512, sampler="ddim_sampler", num_steps=50):
        """kandinsky2_1_model.py:184-292. img_prompt = cat([cond image emb, zero image emb]) [2B, 768]."""
        if sampler not in ("p_sampler", "ddim_sampler", "plms_sampler"):
            raise ValueError("Only ddim_sampler and plms_sampler is available")
        new_h, new_w = self.get_new_h_w(h, w)
        rank, ws, lo, hi = self._shard(batch_size)
        B = hi - lo
        full_emb, pooled_emb = self.embedder.text_emb(prompt, batch_size)
        cond = {"full_emb": full_emb.to(self.device), "pooled_emb": pooled_emb.to(self.device),
                "image_emb": img_prompt.to(self.device).float()}
        parallel.broadcast_conditioning(cond, src=0)   # the path's only collective
        rows = list(range(lo, hi)) + list(range(batch_size + lo, batch_size + hi))
        kw = {k: v[rows].contiguous() for k, v in cond.items()}
        inpaint = {}
        if self.task_type == "inpainting":
            init = init_img.to(self.device).float()
            mask = img_mask.to(self.device).float()
            # the reference repeats ONE image / mask for the cond and uncond rows (:536-537); same image for every sample here
            kw["inpaint_image"] = (init * mask)[:1].repeat(2 * B, 1, 1, 1)
            kw["inpaint_mask"] = mask[:1].repeat(2 * B, 1, 1, 1)
            inpaint = dict(inpaint_init=init[:1].repeat(B, 1, 1, 1), inpaint_mask=mask[:1].repeat(B, 1, 1, 1))
        if noise is None:
            x = self._latents(lo, hi, (4, new_h, new_w))
            noise = torch.cat([x, x], 0)
        elif noise.shape[0] == 2 * batch_size and ws > 1:
            noise = noise[rows].contiguous()   # a caller-supplied start latent covers the GLOBAL batch: keep this rank's rows
        self.model.del_cache()
        if sampler == "p_sampler":
            samples = diffusion.p_sample_loop(self.model, (2 * B, 4, new_h, new_w), device=self.device, noise=noise,
                                              progress=False, model_kwargs=kw, init_step=init_step,
                                              guidance_scale=guidance_scale, cond_first=True, clip_denoised=True,
                                              sample_generators=self._generators(lo, hi), **inpaint)[:B]
        else:  # kandinsky2_1_model.py:259-284: DDIM / PLMS over the un-respaced schedule, eta 0
            cls = DDIMSampler if sampler == "ddim_sampler" else PLMSSampler
            samples, _ = cls(self.model, diffusion).sample(num_steps, 2 * B, (4, new_h, new_w), conditioning=kw,
                                                                     x_T=noise, init_step=init_step,
                                                                     guidance_scale=guidance_scale, cond_first=True)
            samples = samples[:B]
        self.model.del_cache()
        return self._finish(samples, h, w)

    def _diffusion(self, sampler, num_steps):
        dc = dict(self.config["diffusion_config"])
        if sampler == "p_sampler":
            dc["timestep_respacing"] = str(num_steps)
        return create_gaussian_diffusion(**dc)

    def _image_embs(self, prompt, batch_size, negative_decoder_prompt=""):
        pos = self.embedder.image_emb(prompt, batch_size)
        neg = (self.embedder.zero_image_emb(batch_size) if negative_decoder_prompt == ""
               else self.embedder.image_emb(negative_decoder_prompt, batch_size))
        return torch.cat([pos, neg], 0)

    def generate_text2img(self, prompt, num_steps=100, batch_size=1, guidance_scale=7, h=512, w=512,
                          sampler="ddim_sampler", prior_cf_scale=4, prior_steps="25", negative_prior_prompt="",
                          negative_decoder_prompt=""):
        image_emb = self._image_embs(prompt, batch_size, negative_decoder_prompt)
        return self.generate_img(prompt=prompt, img_prompt=image_emb, batch_size=batch_size,
                                 guidance_scale=guidance_scale, h=h, w=w, sampler=sampler, num_steps=num_steps,
                                 diffusion=self._diffusion(sampler, num_steps))

    def mix_images(self, images_texts, weights, num_steps=100, batch_size=1, guidance_scale=7, h=512, w=512,
                   sampler="ddim_sampler", prior_cf_scale=4, prior_steps="25", negative_prior_prompt="",
                   negative_decoder_prompt=""):
        assert len(images_texts) == len(weights) and len(images_texts) > 0
        pos = self.embedder.interpolate(images_texts, weights, batch_size)
        image_emb = torch.cat([pos, self.embedder.zero_image_emb(batch_size)], 0)
        return self.generate_img(prompt="", img_prompt=image_emb, batch_size=batch_size, guidance_scale=guidance_scale,
                                 h=h, w=w, sampler=sampler, num_steps=num_steps,
                                 diffusion=self._diffusion(sampler, num_steps))


    def generate_img2img(self, prompt, pil_img, strength=0.7, num_steps=100, batch_size=1, guidance_scale=7, h=512,
                         w=512, sampler="ddim_sampler", prior_cf_scale=4, prior_steps="25"):
        """kandinsky2_1_model.py:428-484: encode the image, noise it to step int(T*(1-strength)) and run the remaining steps."""
        diffusion = self._diffusion(sampler, num_steps)
        image = self._encode_image(pil_img, h, w) * self.scale
        start_step = int(diffusion.num_timesteps * (1 - strength))
        g = torch.Generator().manual_seed(self.base_seed)
        noise = torch.randn(image.shape, generator=g).to(self.device)
        dc = self.config["diffusion_config"]
        x = q_sample(image, diffusion.timestep_map[start_step - 1], schedule_name=dc["noise_schedule"],
                     num_steps=dc["steps"], noise=noise)
        x = x.repeat(2 * batch_size, 1, 1, 1)
        image_emb = self._image_embs(prompt, batch_size)
        return self.generate_img(prompt=prompt, img_prompt=image_emb, batch_size=batch_size,
                                 guidance_scale=guidance_scale, h=h, w=w, sampler=sampler, num_steps=num_steps,
                                 diffusion=diffusion, noise=x, init_step=start_step)

    def generate_inpainting(self, prompt, pil_img, img_mask, num_steps=100, batch_size=1, guidance_scale=7, h=512,
                            w=512, sampler="ddim_sampler", prior_cf_scale=4, prior_steps="25",
                            negative_prior_prompt="", negative_decoder_prompt=""):
        """kandinsky2_1_model.py:487-548 (mask: 1 = keep, nearest-resized to the latent grid, then prepare_mask)."""
        image = self._encode_image(pil_img, h, w) * self.scale
        m = torch.as_tensor(img_mask).float()[None, None]
        m = torch.nn.functional.interpolate(m, tuple(image.shape[-2:]), mode="nearest")
        m = prepare_mask(m)
        image_emb = torch.cat([self.embedder.image_emb(prompt, batch_size), self.embedder.zero_image_emb(batch_size)], 0)
        return self.generate_img(prompt=prompt, img_prompt=image_emb, batch_size=batch_size,
                                 guidance_scale=guidance_scale, h=h, w=w, sampler=sampler, num_steps=num_steps,
                                 diffusion=self._diffusion(sampler, num_steps),
                                 init_img=image.repeat(2, 1, 1, 1), img_mask=m.repeat(2, 1, 1, 1))


class Kandinsky2_2(_DecoderBase):
    version = "2.2"

    def get_new_h_w(self, h, w):  # kandinsky2_2_model.py:46-53 (pixels)
        return math.ceil(h / 64) * 64, math.ceil(w / 64) * 64

    @torch.no_grad()
    def _decode_loop(self, image_embeds, negative_embeds, batch_size, steps, guidance, h, w, latents=None,
                     inpaint_latent=None, inpaint_mask=None, init_step=None, hint=None):
        """The body of diffusers KandinskyV22Pipeline.__call__ (reference call sites kandinsky2_2_model.py:78-80,
        106-111,138-141,168-172): uncond rows first, DDPM learned-range step, +-2 clip, no dynamic threshold."""
        H, W = h // 8, w // 8
        rank, ws, lo, hi = self._shard(batch_size)
        B = hi - lo
        cond = {"image_emb": torch.cat([negative_embeds, image_embeds], 0).to(self.device).float()}
        parallel.broadcast_conditioning(cond, src=0)
        rows = list(range(lo, hi)) + list(range(batch_size + lo, batch_size + hi))
        kw = {"image_emb": cond["image_emb"][rows].contiguous()}
        if hint is not None:  # one depth map for the whole batch (cond and uncond rows alike, as the diffusers pipeline does)
            kw["hint"] = hint.to(self.device).float().expand(2 * B, -1, -1, -1).contiguous()
        if latents is None:
            x = self._latents(lo, hi, (4, H, W))
            latents = torch.cat([x, x], 0)
        elif latents.shape[0] == 2 * batch_size and ws > 1:
            latents = latents[rows].contiguous()   # caller-supplied start latents cover the GLOBAL batch
        extra = {}
        if inpaint_latent is not None:
            kw["inpaint_image"] = (inpaint_latent * inpaint_mask).repeat(2 * B, 1, 1, 1).to(self.device)
            kw["inpaint_mask"] = inpaint_mask.repeat(2 * B, 1, 1, 1).to(self.device)
            # diffusers KandinskyV22InpaintPipeline (the reference delegates to it, kandinsky2_2_model.py:143-173): after every
            # scheduler step the known region (mask = 1) is the clean latent noised to the next timestep with the run's initial
            # noise, and the result is blended with the clean latent at the end -- k2_sampler_step's inpaint_noise mode
            extra = dict(inpaint_init=inpaint_latent.repeat(B, 1, 1, 1).to(self.device),
                         inpaint_mask=inpaint_mask.repeat(B, 1, 1, 1).to(self.device), inpaint_renoise=True)
        diffusion = create_ddpm_v22(steps)
        self.model.del_cache()
        out = diffusion.p_sample_loop(self.model, (2 * B, 4, H, W), device=self.device, noise=latents,
                                      model_kwargs=kw, guidance_scale=guidance, cond_first=False, clip_denoised=False,
                                      init_step=init_step, sample_generators=self._generators(lo, hi), **extra)[:B]
        self.model.del_cache()
        return self._finish(out, h, w)

    def _embeds(self, prompt, batch_size, negative_decoder_prompt):
        pos = self.embedder.image_emb(prompt, batch_size)
        neg = (self.embedder.zero_image_emb(batch_size) if negative_decoder_prompt == ""
               else self.embedder.image_emb(negative_decoder_prompt, batch_size))
        return pos, neg

    def generate_text2img(self, prompt, batch_size=1, decoder_steps=50, prior_steps=25, decoder_guidance_scale=4,
                          prior_guidance_scale=4, h=512, w=512, negative_prior_prompt="", negative_decoder_prompt=""):
        h, w = self.get_new_h_w(h, w)
        pos, neg = self._embeds(prompt, batch_size, negative_decoder_prompt)
        return self._decode_loop(pos, neg, batch_size, decoder_steps, decoder_guidance_scale, h, w)

    def mix_images(self, images_texts, weights, batch_size=1, decoder_steps=50, prior_steps=25,
                   decoder_guidance_scale=4, prior_guidance_scale=4, h=512, w=512, negative_prior_prompt="",
                   negative_decoder_prompt=""):
        assert len(images_texts) == len(weights) and len(images_texts) > 0
        pos = self.embedder.interpolate(images_texts, weights, batch_size)
        _, neg = self._embeds("", batch_size, negative_decoder_prompt)
        return self._decode_loop(pos, neg, batch_size, decoder_steps, decoder_guidance_scale, h, w)

    def generate_img2img(self, prompt, image, strength=0.4, batch_size=1, decoder_steps=100, prior_steps=25,
                         decoder_guidance_scale=4, prior_guidance_scale=4, h=512, w=512, negative_prior_prompt="",
                         negative_decoder_prompt=""):
        h, w = self.get_new_h_w(h, w)
        pos, neg = self._embeds(prompt, batch_size, negative_decoder_prompt)
        lat = self._encode_image(image, h, w)
        diffusion = create_ddpm_v22(decoder_steps)
        # diffusers KandinskyV22Img2ImgPipeline: the last int(steps*strength) timesteps, scheduler.add_noise at the first of them
        start = max(min(int(decoder_steps * strength), decoder_steps), 1)
        ac = float(diffusion.alphas_cumprod[start - 1])
        g = torch.Generator().manual_seed(self.base_seed)
        noise = torch.randn(lat.shape, generator=g).to(self.device)
        x = ac ** 0.5 * lat + (1.0 - ac) ** 0.5 * noise
        return self._decode_loop(pos, neg, batch_size, decoder_steps, decoder_guidance_scale, h, w,
                                 latents=x.repeat(2 * batch_size, 1, 1, 1), init_step=start)

    def generate_controlnet(self, prompt, hint, batch_size=1, decoder_steps=50, prior_steps=25, decoder_guidance_scale=4,
                            prior_guidance_scale=4, h=512, w=512, negative_prior_prompt="", negative_decoder_prompt=""):
        """Kandinsky 2.2 ControlNet-depth (BASELINE configs[4]).  The reference package has no method for it -- its
        notebooks/kandinsky2_2_controlnet.ipynb calls diffusers' KandinskyV22ControlnetPipeline(image_embeds=...,
        negative_image_embeds=..., hint=hint, height=h, width=w) directly -- so this follows the sibling methods' signature.
        hint: depth map tensor [1, 3, h, w] in [0, 1] (the pipeline object must be built with task_type="controlnet")."""
        if self.task_type != "controlnet":
            raise ValueError("generate_controlnet needs a pipeline built with task_type='controlnet'")
        h, w = self.get_new_h_w(h, w)
        pos, neg = self._embeds(prompt, batch_size, negative_decoder_prompt)
        hint = torch.as_tensor(hint).float()
        if hint.dim() == 3:
            hint = hint[None]
        if tuple(hint.shape[-2:]) != (h, w):
            hint = torch.nn.functional.interpolate(hint, (h, w), mode="bilinear", align_corners=False)
        return self._decode_loop(pos, neg, batch_size, decoder_steps, decoder_guidance_scale, h, w, hint=hint)

    def generate_inpainting(self, prompt, pil_img, img_mask, batch_size=1, decoder_steps=50, prior_steps=25,
                            decoder_guidance_scale=4, prior_guidance_scale=4, h=512, w=512, negative_prior_prompt="",
                            negative_decoder_prompt=""):
        h, w = self.get_new_h_w(h, w)
        pos, neg = self._embeds(prompt, batch_size, negative_decoder_prompt)
        lat = self._encode_image(pil_img, h, w)
        m = torch.as_tensor(img_mask).float()[None, None]
        m = torch.nn.functional.interpolate(m, (h // 8, w // 8), mode="nearest").to(self.device)
        return self._decode_loop(pos, neg, batch_size, decoder_steps, decoder_guidance_scale, h, w,
                                 inpaint_latent=lat, inpaint_mask=m)

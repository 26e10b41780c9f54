"""ORACLE — test infrastructure only.  Literal CPU restatement of the randomized-blending denoising loop of
`I2VGenXLPipeline.__call__` (reference code/i2v_enhance/pipeline_i2vgen_xl.py:841-909) with the UNet passed in as a
callable, python `random` offsets in the reference's order, and `DDIMScheduler.step` / `scale_model_input` restated
from the published definition of diffusers==0.30.2 (eta = 0, no clipping, no thresholding).  diffusers is not
installed offline and the loop lives inline in a 400-line `__call__` that needs the CLIP / VAE encoders, so this row
is **parity unpinned**: the restatement is checked line by line against the reference source, not against its output."""
from __future__ import annotations

import random

import torch


def ddim_step(model_output, timestep, sample, alphas_cumprod, num_train_timesteps, num_inference_steps,
              prediction_type="v_prediction", final_alpha_cumprod=1.0):
    """diffusers DDIMScheduler.step with eta = 0, use_clipped_model_output False, clip_sample False."""
    prev_timestep = timestep - num_train_timesteps // num_inference_steps
    alpha_prod_t = alphas_cumprod[timestep]
    alpha_prod_t_prev = alphas_cumprod[prev_timestep] if prev_timestep >= 0 else final_alpha_cumprod
    beta_prod_t = 1 - alpha_prod_t
    if prediction_type == "epsilon":
        pred_original_sample = (sample - beta_prod_t ** 0.5 * model_output) / alpha_prod_t ** 0.5
        pred_epsilon = model_output
    elif prediction_type == "v_prediction":
        pred_original_sample = (alpha_prod_t ** 0.5) * sample - (beta_prod_t ** 0.5) * model_output
        pred_epsilon = (alpha_prod_t ** 0.5) * model_output + (beta_prod_t ** 0.5) * sample
    else:
        raise ValueError(prediction_type)
    pred_sample_direction = (1 - alpha_prod_t_prev) ** 0.5 * pred_epsilon          # std_dev_t = 0
    return alpha_prod_t_prev ** 0.5 * pred_original_sample + pred_sample_direction


def blending_loop(unet, latents, timesteps, per_chunk_kwargs, *, chunk_size, overlap_size, guidance_scale,
                  alphas_cumprod, num_train_timesteps, num_inference_steps, prediction_type="v_prediction",
                  rng=random, **unet_kwargs):
    """pipeline_i2vgen_xl.py:841-909 for batch 1.  latents [1, C, F, H, W]."""
    do_cfg = guidance_scale is not None and guidance_scale > 1.0
    for t in timesteps:
        latents_denoised = torch.empty_like(latents)
        CHUNK_START = 0
        for idx in range(len(per_chunk_kwargs)):
            latents_chunk = latents[:, :, CHUNK_START:CHUNK_START + chunk_size]
            latent_model_input = torch.cat([latents_chunk] * 2) if do_cfg else latents_chunk
            noise_pred = unet(latent_model_input, int(t), **per_chunk_kwargs[idx], **unet_kwargs)
            if do_cfg:
                noise_pred_uncond, noise_pred_text = noise_pred.chunk(2)
                noise_pred = noise_pred_uncond + guidance_scale * (noise_pred_text - noise_pred_uncond)
            b, c, f, w, h = latents_chunk.shape
            lc = latents_chunk.permute(0, 2, 1, 3, 4).reshape(b * f, c, w, h)
            npd = noise_pred.permute(0, 2, 1, 3, 4).reshape(b * f, c, w, h)
            lc = ddim_step(npd, int(t), lc, alphas_cumprod, num_train_timesteps, num_inference_steps, prediction_type)
            latents_chunk = lc[None, :].reshape(b, f, c, w, h).permute(0, 2, 1, 3, 4)
            if CHUNK_START == 0:
                random_offset = 0
            else:
                random_offset = rng.randint(0, overlap_size - 1) if overlap_size != 0 else 0
            latents_denoised[:, :, CHUNK_START + random_offset:CHUNK_START + chunk_size] = \
                latents_chunk[:, :, random_offset:]
            CHUNK_START += chunk_size - overlap_size
        latents = latents_denoised
        if CHUNK_START + overlap_size > latents_denoised.shape[2]:
            raise NotImplementedError("not dividable into chunks")
    return latents

"""Host-side image helpers of the pipelines (reference: kandinsky2/utils.py)."""
import numpy as np
import torch


def prepare_mask(mask):
    """Grow the zero (inpaint) region of a keep-mask [1,1,H,W] exactly as the reference's double python loop does
    (utils.py:11-30): every zero pixel (i,j) of the ORIGINAL mask also zeroes (i-1,j), (i,j-1), (i-1,j-1), (i+1,j),
    (i,j+1) and (i+1,j+1) -- note the asymmetric neighbourhood -- evaluated here with shifted copies."""
    m = torch.as_tensor(mask).float()[0]          # [1, H, W]
    zero = (m[0] != 1)                            # pixels the reference does not `continue` over
    H, W = zero.shape
    grown = zero.clone()
    for di, dj in ((-1, 0), (0, -1), (-1, -1), (1, 0), (0, 1), (1, 1)):
        src = zero[max(0, -di):H - max(0, di), max(0, -dj):W - max(0, dj)]
        grown[max(0, di):H - max(0, -di), max(0, dj):W - max(0, -dj)] |= src
    out = m.clone()
    out[:, grown] = 0
    return out.unsqueeze(0)


def prepare_image(pil_image, w=512, h=512):
    """PIL -> float [-1, 1] NCHW (utils.py:33-39)."""
    from PIL import Image
    pil_image = pil_image.resize((w, h), resample=Image.BICUBIC, reducing_gap=1)
    arr = np.array(pil_image.convert("RGB")).astype(np.float32) / 127.5 - 1
    return torch.from_numpy(np.transpose(arr, [2, 0, 1])).unsqueeze(0)


def uint8_to_pil(batch_u8):
    """uint8 NHWC tensor -> list of PIL images (utils.py:57-70 tail)."""
    from PIL import Image
    arr = batch_u8.cpu().numpy()
    return [Image.fromarray(a) for a in arr]


def q_sample(x_start, t, schedule_name="linear", num_steps=1000, noise=None):
    """Forward diffusion of x_start to timestep t (utils.py:42-54).  As in the reference this uses
    get_named_beta_schedule's DEFAULT linear range (1e-4 .. 2e-2 scaled by 1000/num_steps), not the decoder's."""
    if schedule_name != "linear":
        raise NotImplementedError(schedule_name)
    scale = 1000 / num_steps
    betas = np.linspace(scale * 0.0001, scale * 0.02, num_steps, dtype=np.float64)
    ac = np.cumprod(1.0 - betas, axis=0)
    if noise is None:
        noise = torch.randn_like(x_start)
    assert noise.shape == x_start.shape
    tt = torch.as_tensor(t).long().reshape(-1).cpu()           # 0-d (the pipelines' call) or one timestep per sample
    shape = (-1,) + (1,) * (x_start.dim() - 1)                 # _extract_into_tensor: float32 table values, broadcast
    a = torch.from_numpy(np.sqrt(ac)).float()[tt].to(x_start.device).reshape(shape)
    b = torch.from_numpy(np.sqrt(1.0 - ac)).float()[tt].to(x_start.device).reshape(shape)
    return a * x_start + b * noise

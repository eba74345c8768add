"""TEST INFRASTRUCTURE ONLY -- plain-PyTorch fp32 restatement of the reference MMDiT path.

A functional forward over a reference-schema state_dict, the rectified-flow
train step, and the Euler/CFG sampler.  Every function cites the reference lines it restates.
It is pinned against the UNMODIFIED reference modules (imported through oracle/ref_shim.py in
the build container) by tests/golden/*.pt -- see oracle/make_golden.py and
tests/test_oracle_golden.py.  Runs on CPU, or on CUDA tensors in fp32 when the GPU tests want
a bigger case; it never calls the product kernels.

Third-party arithmetic that is not under /root/reference and is restated from published
semantics: xformers 0.0.29.post3 SwiGLU (packed w12, bias, silu(x1)*x2) and flash-attn 2.6.3
(softmax(QK^T/8)V) -- both "parity unpinned" by the reference's own (absent) tests; the golden
vectors pin this file to the reference's eager `softmax` path, which is the same math.
"""
import math

import torch
import torch.nn.functional as F

TEXT_SPLIT = 77      # diff_model.py:324-325
TEXT_DIM = 2304      # diff_model.py:167
RMS_EPS = torch.finfo(torch.float32).eps   # nn.RMSNorm(eps=None) on fp32 input (torch 2.11)
# False: exact fp32 attention (the truth our bf16 kernels are measured against).  True: replay the
# reference's eager `softmax` path bit-for-bit (explicit bf16 casts) -- used to pin this file to the
# golden vectors, which were minted from that path.
BF16_ATTENTION_CORE = False
# "eager" (default): softmax(QK^T/8)V materialised.  "flash" / "sdpa": the library kernel the reference
# itself calls on a GPU (flash_attn_func, Attention.py:293) resp. torch SDPA -- only for timing the
# reference's GPU path beside ours (bench.py `gpu_library_baseline`), never for parity.
ATTENTION_KERNEL = "eager"


# ------------------------------------------------------------------ synthetic weights / data
def _gen(name, salt=0):
    import hashlib
    seed = int.from_bytes(hashlib.sha256(f"{name}/{salt}".encode()).digest()[:4], "little")
    return torch.Generator().manual_seed(seed)


def synth_state_dict(shapes, salt=0):
    """Deterministic weights keyed by parameter NAME (so reference and product modules, whose
    constructors draw random numbers in different orders, load identical values).
    shapes: {key: shape}.  Linear-like weights ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in))."""
    sd = {}
    for k, shp in shapes.items():
        g = _gen(k, salt)
        shp = tuple(shp)
        if k.endswith("rotary_emb.freqs"):
            n = shp[0] * 2   # RotaryEmbedding(dim=head_dim/2): 1/theta^(arange(0,dim,2)/dim)
            sd[k] = 1.0 / (10000 ** (torch.arange(0, n, 2)[: n // 2].float() / n))
        elif k == "time_scale":
            sd[k] = torch.tensor([1000.0])
        elif k in ("learnable_scalar", "learnable_scalar2"):
            sd[k] = torch.tensor([0.01 if k.endswith("r") else 0.013])
        elif "norm" in k.split(".")[-2] and len(shp) == 1:      # RMSNorm weights
            sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif len(shp) == 1:                                      # biases
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * 0.05
        else:
            fan_in = int(torch.tensor(shp[1:]).prod())
            sd[k] = (torch.rand(shp, generator=g) * 2 - 1) / math.sqrt(fan_in)
    return sd


def synth_batch(B, C, h, w, M=154, class_dim=768, seed=1000, p_null=(0.1, 0.316, 0.316)):
    """One synthetic training batch, drawn on the CPU: latents, text, pooled, t
    (TimeSampler.py:14-20), null masks (model_trainer.py:382-387) and epsilon."""
    g = torch.Generator().manual_seed(seed)
    x0 = torch.randn((B, C, h, w), generator=g)
    c = torch.randn((B, M, TEXT_DIM), generator=g)
    pooled = torch.randn((B, class_dim), generator=g)
    t = torch.sigmoid(torch.randn(B, generator=g))
    nulls = [torch.rand(B, generator=g) < p for p in p_null]
    eps = torch.randn((B, C, h, w), generator=g)
    return dict(x0=x0, c=c, pooled=pooled, t=t, null_pooled=nulls[0], null_gemma=nulls[1],
                null_bert=nulls[2], eps=eps)


# ------------------------------------------------------------------------------ the model
def timestep_embedding(t, time_scale, dim):
    """PositionalEncoding.py:15-16,23-30 applied to t*time_scale (diff_model.py:306)."""
    denom = (torch.tensor(10000.0) ** ((2 * torch.arange(dim)) / dim)).to(t.device, torch.float32)
    e = (t.float() * time_scale)[:, None] / denom[None, :]
    return torch.cat((e[:, ::2].sin(), e[:, 1::2].cos()), dim=1)


def adaln(x, y, w_shift, w_scale):
    """Norm.py:16-23."""
    x = F.layer_norm(x, (x.shape[-1],))
    return x * (1 + F.linear(y, w_scale)[:, None, :]) + F.linear(y, w_shift)[:, None, :]


def axial_angles(freqs, h, w):
    """rotary_embedding.py:269-288 (get_axial_freqs) + :316 (repeat n -> (n r), r=2): [h,w,64]."""
    fh = torch.arange(h, device=freqs.device).float()[:, None] * freqs[None]
    fw = torch.arange(w, device=freqs.device).float()[:, None] * freqs[None]
    fh = fh.repeat_interleave(2, -1)[:, None, :].expand(h, w, -1)
    fw = fw.repeat_interleave(2, -1)[None, :, :].expand(h, w, -1)
    return torch.cat([fh, fw], -1)


def rope(x, ang):
    """rotary_embedding.py:36-40,72: interleaved pairs (x0,x1) -> (x0 c - x1 s, x1 c + x0 s)."""
    xr = x.unflatten(-1, (-1, 2))
    rot = torch.stack((-xr[..., 1], xr[..., 0]), -1).flatten(-2)
    return x * ang.cos() + rot * ang.sin()


def joint_attention(P, pre, x, c, H, hw, last):
    """Attention.py:130-135 (QKV + per-head RMSNorm), :174-194 (RoPE2d on image tokens),
    :259-263 (concat image;text), :267-284 softmax(QK^T * hd^-0.5)V, :411-425 (split, out-proj)."""
    B, N, d = x.shape
    M = c.shape[1]
    hd = d // H

    def heads(t, T):
        return t.reshape(B, T, H, hd).permute(0, 2, 1, 3)

    def proj(name, t, T):
        return heads(F.linear(t, P[pre + name + ".weight"]), T)

    qx = F.rms_norm(proj("query_proj_x", x, N), (hd,), P[pre + "q_norm_x.weight"], RMS_EPS)
    kx = F.rms_norm(proj("key_proj_x", x, N), (hd,), P[pre + "k_norm_x.weight"], RMS_EPS)
    vx = proj("value_proj_x", x, N)
    qc = F.rms_norm(proj("query_proj_c", c, M), (hd,), P[pre + "q_norm_c.weight"], RMS_EPS)
    kc = F.rms_norm(proj("key_proj_c", c, M), (hd,), P[pre + "k_norm_c.weight"], RMS_EPS)
    vc = proj("value_proj_c", c, M)
    if (pre + "rotary_emb.freqs") in P:
        h, w = hw
        ang = axial_angles(P[pre + "rotary_emb.freqs"], h, w).reshape(1, 1, N, hd)
        qx, kx = rope(qx, ang), rope(kx, ang)
    q, k, v = torch.cat([qx, qc], 2), torch.cat([kx, kc], 2), torch.cat([vx, vc], 2)
    if ATTENTION_KERNEL == "flash":
        from flash_attn import flash_attn_func
        att = flash_attn_func(q.transpose(1, 2).to(torch.bfloat16), k.transpose(1, 2).to(torch.bfloat16),
                              v.transpose(1, 2).to(torch.bfloat16), softmax_scale=hd ** -0.5).transpose(1, 2)
    elif ATTENTION_KERNEL == "sdpa":
        att = F.scaled_dot_product_attention(q.to(torch.bfloat16), k.to(torch.bfloat16), v.to(torch.bfloat16),
                                             scale=hd ** -0.5)
    elif BF16_ATTENTION_CORE:
        # Attention.py:277-284 verbatim: the eager path casts q, k, v to bf16 even outside autocast
        att = (q.to(torch.bfloat16) @ k.to(torch.bfloat16).mT) * hd ** -0.5
        att = (att.softmax(dim=-1) @ v.to(torch.bfloat16)).to(q.dtype)
    else:
        att = ((q @ k.transpose(-1, -2)) * hd ** -0.5).softmax(-1) @ v
    ax = att[:, :, :N].permute(0, 2, 1, 3).reshape(B, N, d)
    ac = att[:, :, N:].permute(0, 2, 1, 3).reshape(B, M, d)
    ox = F.linear(ax, P[pre + "out_proj_x.weight"])
    oc = ac if last else F.linear(ac, P[pre + "out_proj_c.weight"])
    return ox, oc


def swiglu_mlp(P, pre, x):
    """MLP.py:19,32 -> xformers SwiGLU: w3(silu(x1) * x2), x1,x2 = w12(x).chunk(2)."""
    a, b = F.linear(x, P[pre + "MLP.w12.weight"], P[pre + "MLP.w12.bias"]).chunk(2, -1)
    return F.linear(F.silu(a) * b, P[pre + "MLP.w3.weight"], P[pre + "MLP.w3.bias"])


def block(P, i, x, c, y, H, hw, last):
    """Transformer_Block_Dual.py:56-78."""
    b = f"blocks.{i}."
    yp = F.silu(F.linear(y, P[b + "y_proj.0.weight"], P[b + "y_proj.0.bias"]))
    xn = adaln(x, yp, P[b + "norm1_x.c_shift.weight"], P[b + "norm1_x.c_scale.weight"])
    cn = adaln(c, yp, P[b + "norm1_c.c_shift.weight"], P[b + "norm1_c.c_scale.weight"])
    ax, ac = joint_attention(P, b + "attn.", xn, cn, H, hw, last)
    x = ax * F.linear(yp, P[b + "scale1_x.weight"])[:, None, :] + x
    if not last:
        c = ac * F.linear(yp, P[b + "scale1_c.weight"])[:, None, :] + c
    xn = adaln(x, yp, P[b + "norm2_x.c_shift.weight"], P[b + "norm2_x.c_scale.weight"])
    x = swiglu_mlp(P, b + "MLP_x.", xn) * F.linear(yp, P[b + "scale2_x.weight"])[:, None, :] + x
    if not last:
        cn = adaln(c, yp, P[b + "norm2_c.c_shift.weight"], P[b + "norm2_c.c_scale.weight"])
        c = swiglu_mlp(P, b + "MLP_c.", cn) * F.linear(yp, P[b + "scale2_c.weight"])[:, None, :] + c
    return x, c


def forward(P, cfg, x_t, t, c, pooled, null_pooled=None, null_gemma=None, null_bert=None):
    """diff_model.py:264-346.  P: state_dict (fp32 tensors), cfg: the ctor kwargs dict.
    Does NOT mutate the caller's tensors (the reference does, in place, :281-287)."""
    d, H, p, C, depth = cfg["dim"], cfg["num_heads"], cfg["patch_size"], cfg["inCh"], cfg["num_blocks"]
    B, _, hh, ww = x_t.shape
    c, pooled = c.float().clone(), pooled.float().clone()
    if null_pooled is not None:
        pooled[null_pooled] *= 0
    if null_gemma is not None:
        c[null_gemma, :TEXT_SPLIT] *= 0
    if null_bert is not None:
        c[null_bert, TEXT_SPLIT:] *= 0
    tv = F.linear(timestep_embedding(t, P["time_scale"], d), P["t_emb2.weight"])
    y = tv + F.linear(pooled, P["cond_MLP.weight"])
    c = torch.cat([
        F.linear(P["learnable_scalar"] * F.rms_norm(c[:, :TEXT_SPLIT], (TEXT_DIM,), P["pre_c_norm.weight"], RMS_EPS),
                 P["c_proj.weight"]),
        F.linear(P["learnable_scalar2"] * F.rms_norm(c[:, TEXT_SPLIT:], (TEXT_DIM,), P["pre_c_norm2.weight"], RMS_EPS),
                 P["c_proj2.weight"]),
    ], dim=1)
    x = F.conv2d(x_t.float(), P["pos_enc.proj.weight"], stride=p).flatten(2).transpose(1, 2)
    x = F.linear(x, P["patch_emb.weight"], P["patch_emb.bias"])
    hw = (hh // p, ww // p)
    for i in range(depth):
        x, c = block(P, i, x, c, y, H, hw, last=(i == depth - 1))
    x = F.linear(adaln(x, y, P["out_norm.c_shift.weight"], P["out_norm.c_scale.weight"]),
                 P["out_proj.weight"], P["out_proj.bias"])
    # patchify.py:41-72 (unpatchify): column index c*p*p + i*p + j
    x = x.view(B, hw[0], hw[1], C, p, p).permute(0, 3, 1, 4, 2, 5).reshape(B, C, hh, ww)
    return x


# --------------------------------------------------------------------------- training step
def rf_loss(P, cfg, batch):
    """model_trainer.py:394 (noise_batch -> diff_model.py:235-238), :421 forward, :429-446 loss."""
    t = batch["t"]
    x_t = (1 - t)[:, None, None, None] * batch["x0"] + t[:, None, None, None] * batch["eps"]
    v = forward(P, cfg, x_t, t, batch["c"], batch["pooled"], batch["null_pooled"], batch["null_gemma"],
                batch["null_bert"])
    loss = F.mse_loss(v, batch["eps"] - batch["x0"], reduction="none").flatten(1, -1).mean()
    return loss, v


class TrainOracle:
    """model_trainer.py:260 (AdamW lr, betas, eps 1e-8, wd 0.01 on every parameter), :463-503
    (loss/accum, backward, clip_grad_norm_ 1.0, step, zero_grad) without GradScaler (a 2^16 loss
    scale followed by unscale_ is numerically neutral in fp32/bf16)."""

    def __init__(self, state_dict, cfg, lr=1e-4, device="cpu"):
        self.cfg = cfg
        self.P = {k: v.clone().to(device).float().requires_grad_(not k.endswith("rotary_emb.freqs"))
                  for k, v in state_dict.items()}
        self.params = [v for v in self.P.values() if v.requires_grad]
        self.opt = torch.optim.AdamW(self.params, lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)

    def step(self, batch):
        loss, _ = rf_loss(self.P, self.cfg, batch)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.params, 1.0)
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        return float(loss)


# ------------------------------------------------------------------------------- sampling
@torch.no_grad()
def sample_euler(P, cfg, noise, text_hidden, text_pooled, num_steps, cfg_scale):
    """diff_model.py:377-430 with sampler='euler': batch-2B forward, CFG combine, x -= v*dt."""
    B = noise.shape[0]
    x = noise.clone().float()
    null = torch.tensor([0] * B + [1] * B).bool().to(noise.device)
    th = text_hidden.float().repeat(2 * B, 1, 1)
    tp = text_pooled.float().repeat(2 * B, 1)
    for t in torch.linspace(1, 1.0 / num_steps, num_steps):
        tt = t.repeat(2 * B).to(noise.device)
        v = forward(P, cfg, x.repeat(2, 1, 1, 1), tt, th, tp, null, null, null)
        v = (1 + cfg_scale) * v[:B] - cfg_scale * v[B:]
        x = x - v * (1 / num_steps)
    return x

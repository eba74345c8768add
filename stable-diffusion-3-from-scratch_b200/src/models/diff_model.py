"""MMDiT rectified-flow model with the reference's class API (reference:
src/models/diff_model.py:83-217 ctor, :229-241 noise_batch, :264-346 forward,
:367-480 sample_imgs, :489-536 saveModel, :553-579 loadModel).

Constructor keywords, attribute names, forward/sample signatures, state_dict keys and the
checkpoint file set are the reference's, so train.py / infer.py and
existing checkpoints drive this class unchanged.  All device math runs in hand-written
sm_100a kernels behind libmmdit_b200.so (bf16 storage, fp32 accumulation -- the
reference's numerics under torch.autocast(bf16), model_trainer.py:416).  There is no CPU
path: parameters may live on the CPU (EMA copies, checkpoints) but forward needs a B200.
"""
import json
import os

import numpy as np
import torch
from torch import nn

from mmdit import ops, streams
from mmdit.functional import LinearFn, TextFrontFn, UnpatchifyFn
from mmdit.shadow import packed_weight
from src.blocks.ImagePositionalEncoding import PatchEmbed
from src.blocks.Norm import Norm
from src.blocks.PositionalEncoding import PositionalEncoding
from src.blocks.Transformer_Block_Dual import Transformer_Block_Dual
from src.helpers.VAE_T5_CLIP_inference import VAE_T5_CLIP_inference

BF16 = torch.bfloat16
F32 = torch.float32
# MMDIT_SAMPLE_GRAPH=0: run the Euler loop eagerly instead of replaying one captured step
SAMPLE_GRAPH = os.environ.get("MMDIT_SAMPLE_GRAPH", "1") == "1"


class _GraphCache(dict):
    """Captured sampler steps, keyed by geometry.  Derived state: never copied or pickled with the module
    (the reference deep-copies the model for its EMA, model_trainer.py:256)."""

    def __deepcopy__(self, memo):
        return _GraphCache()

    def __reduce__(self):
        return (_GraphCache, ())


class diff_model(nn.Module):
    def __init__(self, inCh, class_dim, patch_size, dim, hidden_scale, num_heads, attn_type, MLP_type,
                 num_blocks, device, positional_encoding, max_res_orig=256, max_res=256,
                 update_max_res=False, kv_merge_attn=False, qk_half_dim=False, text_loss=False,
                 checkpoint_MLP=True, checkpoint_attn=True, start_step=0, wandb_id=None):
        super(diff_model, self).__init__()
        self.update_max_res = update_max_res
        self.max_res = max_res
        self.RoPE_Scale = max_res_orig / max_res
        self.inCh = inCh
        self.class_dim = class_dim
        self.patch_size = patch_size
        self.start_step = start_step
        self.wandb_id = wandb_id
        self.text_loss = text_loss
        self.dim = dim

        assert positional_encoding in ["absolute", "RoPE", "NoPE", "RoPE2d", "RoPE2dV2"], \
            "positional_encoding must be 'absolute', 'RoPE', or 'NoPE' or 'RoPE2d' or 'RoPE2dV2'"
        assert MLP_type in ["gelu", "swiglu", "swiglu_old"]
        if text_loss:
            raise NotImplementedError("text_loss=True is a research flag outside the MMDiT hot path")
        self.legacy_MLP = MLP_type == "swiglu_old"

        # JSON written next to every checkpoint; loadModel re-invokes __init__(**defaults)
        self.defaults = {
            "inCh": inCh, "class_dim": class_dim, "patch_size": patch_size, "dim": dim,
            "hidden_scale": hidden_scale, "num_heads": num_heads, "attn_type": attn_type,
            "MLP_type": MLP_type, "num_blocks": num_blocks, "positional_encoding": positional_encoding,
            "max_res_orig": max_res_orig, "max_res": max_res, "kv_merge_attn": kv_merge_attn,
            "qk_half_dim": qk_half_dim, "text_loss": text_loss, "device": "cpu",
            "start_step": start_step, "wandb_id": wandb_id,
        }

        if type(device) is str:
            if device.lower() == "gpu" and torch.cuda.is_available():
                dev = "gpu"
                device = torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}")
            else:
                if device.lower() == "gpu":
                    print("GPU not available, defaulting to CPU. Please ignore this message if you do "
                          "not wish to use a GPU\n")
                dev = "cpu"
                device = torch.device("cpu")
            self.device, self.dev = device, dev
        else:
            self.device = device
            self.dev = "cpu" if device.type == "cpu" else "gpu"

        self.blocks = nn.ModuleList([
            Transformer_Block_Dual(dim, c_dim=dim, hidden_scale=hidden_scale, num_heads=num_heads,
                                   attn_type=attn_type, MLP_type=MLP_type,
                                   positional_encoding=positional_encoding, RoPE_Scale=self.RoPE_Scale,
                                   kv_merge_attn=kv_merge_attn, qk_half_dim=qk_half_dim,
                                   checkpoint_MLP=checkpoint_MLP, checkpoint_attn=checkpoint_attn,
                                   layer_idx=i, last=(i == num_blocks - 1 and not self.text_loss)).to(device)
            for i in range(num_blocks)
        ])
        self.t_emb = PositionalEncoding(dim, device=device).to(device)
        self.t_emb2 = nn.Linear(dim, dim, bias=False).to(device)
        self.cond_MLP = nn.Linear(self.class_dim, dim, bias=False).to(device)
        self.text_hidden_shape = 2304
        self.c_proj = nn.Linear(self.text_hidden_shape, dim, bias=False).to(device)
        self.c_proj2 = nn.Linear(self.text_hidden_shape, dim, bias=False).to(device)
        self.pre_c_norm = nn.RMSNorm(self.text_hidden_shape).to(device)
        self.pre_c_norm2 = nn.RMSNorm(self.text_hidden_shape).to(device)
        self.learnable_scalar = nn.Parameter(torch.tensor([0.01], dtype=torch.float, device=device),
                                             requires_grad=True).to(device)
        self.learnable_scalar2 = nn.Parameter(torch.tensor([0.01], dtype=torch.float, device=device),
                                              requires_grad=True).to(device)
        self.patch_emb = nn.Linear(dim, dim).to(device)
        self.pos_enc = PatchEmbed(height=256, width=256, patch_size=self.patch_size, in_channels=inCh,
                                  embed_dim=dim, layer_norm=False, flatten=True, bias=False,
                                  interpolation_scale=1, pos_embed_type=positional_encoding,
                                  pos_embed_max_size=256).to(device)
        self.out_norm = Norm(dim, dim).to(device)
        self.out_proj = nn.Linear(dim, inCh * patch_size * patch_size).to(device)
        self.time_scale = nn.Parameter(torch.tensor([1000.0], dtype=torch.float, device=device),
                                       requires_grad=True).to(device)

    # ----------------------------------------------------------------- noise
    def noise_batch(self, X, t):
        """Rectified-flow noising (reference :229-241): returns (x_t fp32, epsilon)."""
        X = X.to(self.device)
        t = t.to(self.device)
        epsilon = torch.randn_like(X, device=self.device)
        X_t = ops.rf_noise(X.contiguous(), epsilon, t.to(F32).contiguous())
        return X_t, epsilon

    def load_text_encoders(self):
        self.text_encoders = VAE_T5_CLIP_inference(self.device)

    # --------------------------------------------------------------- forward
    def _linear(self, lin, x, owner_key="w"):
        params = [lin.weight] + ([lin.bias] if lin.bias is not None else [])
        return LinearFn.apply(x, packed_weight(lin, owner_key, [lin.weight]),
                              None if lin.bias is None else lin.bias.detach(), 0, 1, *params)

    def forward(self, x_t, t, c, c_pooled, nullCls_pooled=None, nullCls_gemma=None, nullCls_bert=None):
        dev = self.device
        x_t, c, c_pooled = x_t.to(dev), c.to(dev), c_pooled.to(dev)
        B = x_t.shape[0]

        # null-conditioning masks, applied IN PLACE on the caller's tensors like the reference
        # (:278-287) but as a multiply by {0,1} (no boolean-index host sync -> graph capturable)
        with torch.no_grad():
            if nullCls_pooled is not None:
                c_pooled.mul_((~nullCls_pooled.to(dev).bool()).to(c_pooled.dtype)[:, None])
            if nullCls_gemma is not None:
                c[:, :77].mul_((~nullCls_gemma.to(dev).bool()).to(c.dtype)[:, None, None])
            if nullCls_bert is not None:
                c[:, 77:].mul_((~nullCls_bert.to(dev).bool()).to(c.dtype)[:, None, None])

        if isinstance(t, (int, float)):
            t = torch.full((B,), float(t), device=dev)
        else:
            t = torch.as_tensor(t, device=dev)
            if t.dim() == 0:
                t = t.repeat(B)
        # time embedding (:306): t_emb2(PositionalEncoding(t * time_scale))
        t_vec = self._linear(self.t_emb2, self.t_emb.embed(t.to(dev).float(), self.time_scale))
        # pooled text embedding (:310,313)
        y = t_vec + self._linear(self.cond_MLP, c_pooled if c_pooled.dtype == BF16 else c_pooled.to(BF16))

        orig_shape = x_t.shape
        # text front-end (:323-326): per-encoder RMSNorm * scalar, projection, concat over tokens
        cb = c if c.dtype == BF16 else c.to(BF16)
        cseq = TextFrontFn.apply(cb, self.pre_c_norm.weight, self.pre_c_norm2.weight,
                                 self.learnable_scalar, self.learnable_scalar2, 77,
                                 packed_weight(self.c_proj, "w", [self.c_proj.weight]),
                                 packed_weight(self.c_proj2, "w", [self.c_proj2.weight]),
                                 self.c_proj.weight, self.c_proj2.weight)

        # patch embedding (:329,332)
        x = self.pos_enc(x_t if x_t.dtype in (BF16, F32) else x_t.float())
        x = self._linear(self.patch_emb, x)

        # y' = SiLU(y_proj_i(y)) of ALL blocks in one GEMM (they depend only on y), instead of one
        # 64-row GEMM + SiLU + their backward chain per block (Transformer_Block_Dual.py:57)
        yws = [blk.y_proj[0].weight for blk in self.blocks]
        ybs = [blk.y_proj[0].bias for blk in self.blocks]
        yp_all = LinearFn.apply(y, packed_weight(self, "y_proj_all", yws),
                                torch.cat([b_.detach() for b_ in ybs]), ops.EPI_SILU, len(yws), *yws, *ybs)
        yps = yp_all.unflatten(1, (len(yws), self.dim)).unbind(1)
        for block, yp in zip(self.blocks, yps):
            x, cseq = block(x, cseq, y, orig_shape, yp=yp)
        if streams.active(x):   # two-stream block schedule: the text branch must not outlive the forward
            torch.cuda.current_stream().wait_stream(streams.side(x.device))
            streams.infer_keepalive.clear()

        # output head (:339,342)
        x = self._linear(self.out_proj, self.out_norm(x, y))
        p = self.patch_size
        N = x.shape[1]
        return UnpatchifyFn.apply(x.reshape(B * N, -1), B, self.inCh, orig_shape[-2], orig_shape[-1], p)

    # -------------------------------------------------------------- sampling
    @torch.no_grad()
    def sample_imgs(self, batchSize, num_steps, text_input, cfg_scale=0.0, width=256, height=256,
                    save_intermediate=False, use_tqdm=False, sampler="euler", generator=None):
        """Euler / stochastic-Euler / Heun sampler with classifier-free guidance (reference :367-480).
        The velocity combine + Euler update is one fused kernel."""
        if sampler not in ("euler", "euler_stochastic", "heun"):
            raise ValueError("Invalid sampler specified. Choose 'euler', 'euler_stochastic', or 'heun'.")
        self.eval()
        enc = self.text_encoders
        h, w = width, height  # (sic) the reference swaps them (:375-376)
        output = torch.randn((batchSize, enc.VAE.config.latent_channels, h // 8, w // 8),
                             generator=generator).to(self.device).float().contiguous()
        text_hidden, text_pooled = enc.text_to_embedding(text_input)
        nullCls = torch.tensor([0] * batchSize + [1] * batchSize).bool().to(self.device)
        text_hidden = text_hidden.repeat(2 * batchSize, 1, 1).to(self.device)
        text_pooled = text_pooled.repeat(2 * batchSize, 1).to(self.device)
        imgs = []

        def decode(z):
            return enc.VAE.decode((z.to(enc.VAE.dtype) - enc.VAE.config.shift_factor)
                                  / enc.VAE.config.scaling_factor).sample.clamp(-1, 1)

        def velocity(x, t):
            return self.forward(x.repeat(2, 1, 1, 1), t, text_hidden, text_pooled, nullCls, nullCls, nullCls)

        timesteps = torch.linspace(1, 0 + (1.0 / num_steps), num_steps).to(self.device)
        dt = 1 / num_steps
        it = timesteps
        if use_tqdm:
            from tqdm import tqdm
            it = tqdm(timesteps, total=num_steps)
        euler_graph = None
        if sampler == "euler" and SAMPLE_GRAPH and output.is_cuda:
            # one Euler step (batch-2B forward + fused CFG combine / update) captured once per geometry and
            # replayed num_steps times: ~1700 kernel launches per step leave the host's critical path
            euler_graph = self._euler_step_graph(output, text_hidden, text_pooled, nullCls, cfg_scale, dt)
        for t in it:
            t = t.repeat(2 * batchSize)
            if euler_graph is not None:
                graph, static_x, static_t = euler_graph
                static_t.copy_(t)
                graph.replay()
                output = static_x
                if save_intermediate:
                    imgs.append(decode(output)[0].float().cpu().detach())
                continue
            v = velocity(output, t)
            if sampler == "euler":
                ops.cfg_euler_step(output, v.contiguous(), cfg_scale, dt)
            elif sampler == "euler_stochastic":
                sigma = (t * (1 - t) / (1 - t + 0.008))[:batchSize, None, None, None]
                noise = torch.randn(output.shape, generator=generator).to(output.device)
                ops.cfg_euler_step(output, v.contiguous(), cfg_scale, dt)
                output = (output + sigma * noise * np.sqrt(dt)).contiguous()
            else:  # heun
                x_pred = output.clone()
                ops.cfg_euler_step(x_pred, v.contiguous(), cfg_scale, dt)
                v2 = velocity(x_pred, t - dt)
                ops.cfg_euler_step(output, v.contiguous(), cfg_scale, dt / 2)
                ops.cfg_euler_step(output, v2.contiguous(), cfg_scale, dt / 2)
            if save_intermediate:
                imgs.append(decode(output)[0].float().cpu().detach())
        if save_intermediate:
            imgs.append(decode(output)[0].float().cpu().detach())
        output = decode(output).float()
        return (output, imgs) if save_intermediate else output

    def _euler_step_graph(self, x, text_hidden, text_pooled, nullCls, cfg_scale, dt):
        """(graph, static x, static t) for one CFG Euler step at this geometry (diff_model.py:407-430 body).
        The capture is dropped and redone when a parameter was re-assigned or written in place since
        (weights updated by the fused optimizer keep their bf16 shadows fresh on the device)."""
        cache = self.__dict__.setdefault("_sample_graphs", _GraphCache())
        wkey = tuple((p.data_ptr(), p._version) for p in self.parameters())
        key = (tuple(x.shape), tuple(text_hidden.shape), text_hidden.dtype, float(cfg_scale), float(dt), x.device)
        ent = cache.get(key)
        if ent is not None and ent["wkey"] == wkey:
            # every tensor the captured kernels read lives in the cache entry; refresh the inputs
            ent["x"].copy_(x)
            ent["th"].copy_(text_hidden)
            ent["tp"].copy_(text_pooled)
            ent["null"].copy_(nullCls)
            return ent["graph"], ent["x"], ent["t"]
        B = x.shape[0]
        static_x = x.clone()
        static_t = torch.ones(2 * B, device=x.device)
        th, tp, null = text_hidden.clone(), text_pooled.clone(), nullCls.clone()

        def step():
            v = self.forward(static_x.repeat(2, 1, 1, 1), static_t, th, tp, null, null, null)
            ops.cfg_euler_step(static_x, v.contiguous(), cfg_scale, dt)

        s = torch.cuda.Stream(device=x.device)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):       # warm-up (lazy initialisation, allocator) off the capture
            step()
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        static_x.copy_(x)                # the warm-up and the capture pass moved the scratch state
        th.copy_(text_hidden)            # (the forward masks the null half in place: start from the caller's values)
        tp.copy_(text_pooled)
        # the graph reads these tensors on every replay: they must live as long as the graph does
        cache[key] = dict(wkey=wkey, graph=graph, x=static_x, t=static_t, th=th, tp=tp, null=null)
        return graph, static_x, static_t

    # ------------------------------------------------------------ checkpoint
    def saveModel(self, saveDir, EMA_state_dict=None, optimizer=None, scheduler=None, grad_scalar=None,
                  step=None):
        """Same file set and names as the reference (:489-536)."""
        suffix = f"_{step}s" if step else ""
        if step:
            self.defaults["start_step"] = step
        self.defaults["wandb_id"] = self.wandb_id
        os.makedirs(saveDir, exist_ok=True)
        torch.save(self.state_dict(), os.path.join(saveDir, f"model{suffix}.pkl"))
        if EMA_state_dict:
            torch.save(EMA_state_dict, os.path.join(saveDir, f"model_ema{suffix}.pkl"))
        if optimizer:
            torch.save(optimizer.state_dict(), os.path.join(saveDir, f"optim{suffix}.pkl"))
        if scheduler:
            torch.save(scheduler.state_dict(), os.path.join(saveDir, f"scheduler{suffix}.pkl"))
        if grad_scalar:
            torch.save(grad_scalar.state_dict(), os.path.join(saveDir, f"scaler{suffix}.pkl"))
        with open(os.path.join(saveDir, f"model_params{suffix}.json"), "w") as f:
            json.dump(self.defaults, f)

    def loadModel(self, loadDir, loadFile, loadDefFile=None, wandb_id=None):
        """Rebuilds the module from the saved JSON, then loads the state dict strictly (:553-579)."""
        if loadDefFile:
            device_, dev_ = self.device, self.dev
            with open(os.path.join(loadDir, loadDefFile), "r") as f:
                self.defaults = json.load(f)
            D = self.defaults
            D.setdefault("MLP_type", "swiglu_old")
            D.setdefault("text_loss", False)
            if self.update_max_res:
                D["max_res"] = self.max_res
            self.__init__(**D)
            self.to(device_)
            self.device, self.dev = device_, dev_
        self.load_state_dict(torch.load(os.path.join(loadDir, loadFile), map_location=self.device,
                                        weights_only=False), strict=True)

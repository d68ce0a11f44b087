"""Stand-in networks with the reference's LoRA-relevant topology (HARNESS, not product).

diffusers is not installed here and the wrapped UNet/DiT is out of scope by design
(BASELINE.json north_star: the path OneTrainer owns itself, "as opposed to the diffusers UNet/DiT
it wraps").  To exercise the path at the true shapes, these modules reproduce the *hooked leaves* of
the real networks — same diffusers module names (so LoRAModuleWrapper's substring filter selects the
same leaves: 192 for SD 1.5, 722 for SDXL), same (M, K, N) per leaf, same
attention/GEGLU dataflow between them — with seeded random-init frozen weights.  ResNet blocks and
down/up-sampling between the attention stages are replaced by frozen strided 2x2 / 1x1 convolutions:
they are not hooked by the `attn-mlp` presets and are not part of the measured path.
Hyper-parameters: /root/reference/resources/model_config/stable_diffusion_xl/sd_xl_base.yaml:19-37,
/root/reference/resources/model_config/stable_diffusion/v1-inference.yaml:29-44.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn


class _Attention(nn.Module):
    def __init__(self, dim: int, ctx_dim: int | None, head_dim: int = 64):
        super().__init__()
        self.heads = dim // head_dim
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(ctx_dim or dim, dim, bias=False)
        self.to_v = nn.Linear(ctx_dim or dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        B, L, D = x.shape
        q = self.to_q(x).view(B, L, self.heads, -1).transpose(1, 2)
        k = self.to_k(ctx).view(B, ctx.shape[1], self.heads, -1).transpose(1, 2)
        v = self.to_v(ctx).view(B, ctx.shape[1], self.heads, -1).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, L, D)
        return self.to_out[0](o)


class _GEGLU(nn.Module):
    def __init__(self, dim: int, inner: int):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class _FeedForward(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([_GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, ctx_dim: int):
        super().__init__()
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.attn1 = _Attention(dim, None)
        self.attn2 = _Attention(dim, ctx_dim)
        self.ff = _FeedForward(dim)

    def forward(self, x, ctx):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), ctx)
        return x + self.ff(self.norm3(x))


class Transformer2DModel(nn.Module):
    """`use_linear_projection=True` (SDXL): proj_in/out are Linear; SD 1.5: 1x1 Conv2d."""

    def __init__(self, dim: int, depth: int, ctx_dim: int, linear_proj: bool):
        super().__init__()
        self.linear_proj = linear_proj
        self.norm = nn.GroupNorm(32, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim) if linear_proj else nn.Conv2d(dim, dim, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, ctx_dim) for _ in range(depth)])
        self.proj_out = nn.Linear(dim, dim) if linear_proj else nn.Conv2d(dim, dim, 1)

    def forward(self, x, ctx):
        B, C, H, W = x.shape
        res = x
        h = self.norm(x)
        if self.linear_proj:
            h = self.proj_in(h.permute(0, 2, 3, 1).reshape(B, H * W, C))
        else:
            h = self.proj_in(h).permute(0, 2, 3, 1).reshape(B, H * W, C)
        for blk in self.transformer_blocks:
            h = blk(h, ctx)
        if self.linear_proj:
            h = self.proj_out(h).view(B, H, W, C).permute(0, 3, 1, 2)
        else:
            h = self.proj_out(h.view(B, H, W, C).permute(0, 3, 1, 2))
        return h + res


class _Stage(nn.Module):
    def __init__(self, dim, n_attn, depth, ctx_dim, linear_proj):
        super().__init__()
        self.attentions = nn.ModuleList([Transformer2DModel(dim, depth, ctx_dim, linear_proj) for _ in range(n_attn)])

    def forward(self, x, ctx):
        for a in self.attentions:
            x = a(x, ctx)
        return x


class _Output:
    def __init__(self, sample):
        self.sample = sample


class StandInUNet(nn.Module):
    """UNet2DConditionModel-shaped stand-in: `unet(sample, timestep, encoder_hidden_states,
    added_cond_kwargs=None).sample`."""

    def __init__(self, in_ch: int, stages: list[tuple[int, int, int, int]], mid: tuple[int, int] | None, ctx_dim: int,
                 linear_proj: bool, pooled_dim: int = 0):
        """stages: (dim, downsample_factor_vs_latent, n_attn_down, depth); the up path mirrors it with n_attn_down+1."""
        super().__init__()
        self.in_ch = in_ch
        dims = [s[0] for s in stages]
        self.time_embed = nn.Linear(256, dims[-1])
        self.add_embed = nn.Linear(pooled_dim + 6 * 256, dims[-1]) if pooled_dim else None
        self.down_convs = nn.ModuleList()
        prev_c, prev_f = in_ch, 1
        for (dim, factor, _, _) in stages:
            k = factor // prev_f
            self.down_convs.append(nn.Conv2d(prev_c, dim, k, stride=k))
            prev_c, prev_f = dim, factor
        first = 4 - len(stages)          # diffusers numbering: attention stages are the LAST down blocks
        self.down_blocks = nn.ModuleList([nn.Module() for _ in range(first)] +
                                         [_Stage(d, n, dep, ctx_dim, linear_proj) for (d, _, n, dep) in stages])
        self.mid_block = _Stage(mid[0], 1, mid[1], ctx_dim, linear_proj) if mid else None
        self.up_blocks = nn.ModuleList([_Stage(d, n + 1, dep, ctx_dim, linear_proj) for (d, _, n, dep) in reversed(stages)])
        self.up_convs = nn.ModuleList()
        rev = list(reversed(stages))
        for i, (dim, factor, _, _) in enumerate(rev):
            nxt = rev[i + 1] if i + 1 < len(rev) else (in_ch, 1, 0, 0)
            k = factor // nxt[1]
            self.up_convs.append(nn.ConvTranspose2d(dim, nxt[0], k, stride=k))
        self.time_projs = nn.ModuleList([nn.Linear(dims[-1], d) for d in dims])

    @staticmethod
    def _sinusoid(t, dim=256):
        half = dim // 2
        freqs = torch.exp(-math.log(10000.0) * torch.arange(half, device=t.device, dtype=torch.float32) / half)
        args = t.float().reshape(-1, 1) * freqs
        return torch.cat([args.cos(), args.sin()], dim=-1)

    def forward(self, sample, timestep, encoder_hidden_states, added_cond_kwargs=None):
        B = sample.shape[0]
        dt = sample.dtype
        temb = self.time_embed(self._sinusoid(timestep.expand(B) if timestep.numel() == 1 else timestep).to(dt))
        if self.add_embed is not None and added_cond_kwargs is not None:
            ids = self._sinusoid(added_cond_kwargs["time_ids"].flatten()).view(B, -1).to(dt)
            temb = temb + self.add_embed(torch.cat([added_cond_kwargs["text_embeds"].to(dt), ids], dim=-1))
        x = sample
        skips = []
        n_pass = len(self.down_blocks) - len(self.down_convs)
        for i, conv in enumerate(self.down_convs):
            x = conv(x) + self.time_projs[i](temb)[:, :, None, None]
            x = self.down_blocks[n_pass + i](x, encoder_hidden_states)
            skips.append(x)
        if self.mid_block is not None:
            x = self.mid_block(x, encoder_hidden_states)
        for i, (blk, conv) in enumerate(zip(self.up_blocks, self.up_convs)):
            x = blk(x + skips[-1 - i], encoder_hidden_states)
            x = conv(x)
        return _Output(x)


def sdxl_unet() -> StandInUNet:
    """SDXL base: attention at 64x64 (C=640, depth 2) and 32x32 (C=1280, depth 10), mid depth 10, ctx 2048.
    722 hooked leaves / 85.1 M LoRA parameters at rank 32."""
    return StandInUNet(4, [(640, 2, 2, 2), (1280, 4, 2, 10)], (1280, 10), ctx_dim=2048, linear_proj=True, pooled_dim=1280)


def sd15_unet() -> StandInUNet:
    """SD 1.5: attention at 64x64 (320), 32x32 (640), 16x16 (1280), depth 1, mid 1280, ctx 768, conv projections.
    192 hooked leaves."""
    return StandInUNet(4, [(320, 1, 2, 1), (640, 2, 2, 1), (1280, 4, 2, 1)], (1280, 1), ctx_dim=768, linear_proj=False)


def tiny_unet(dim: int = 64, ctx_dim: int = 64) -> StandInUNet:
    """A few-leaf network of the same shape for smoke tests."""
    return StandInUNet(4, [(dim, 2, 1, 1)], (dim, 1), ctx_dim=ctx_dim, linear_proj=True)


# ---------------------------------------------------------------------------------------------
# FLUX.1 DiT stand-in (config C5)
# ---------------------------------------------------------------------------------------------
class _FluxAttention(nn.Module):
    """diffusers `Attention` as used by FluxTransformerBlock (joint, `added_kv`) and FluxSingleTransformerBlock."""

    def __init__(self, dim: int, head_dim: int, joint: bool):
        super().__init__()
        self.heads = dim // head_dim
        self.to_q, self.to_k, self.to_v = nn.Linear(dim, dim), nn.Linear(dim, dim), nn.Linear(dim, dim)
        if joint:
            self.add_q_proj, self.add_k_proj, self.add_v_proj = nn.Linear(dim, dim), nn.Linear(dim, dim), nn.Linear(dim, dim)
            self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])
            self.to_add_out = nn.Linear(dim, dim)

    def _heads(self, t):
        B, L, _ = t.shape
        return t.view(B, L, self.heads, -1).transpose(1, 2)

    def forward(self, x, ctx=None):
        B, L, D = x.shape
        q, k, v = self._heads(self.to_q(x)), self._heads(self.to_k(x)), self._heads(self.to_v(x))
        if ctx is not None:
            Lc = ctx.shape[1]
            q = torch.cat([self._heads(self.add_q_proj(ctx)), q], dim=2)
            k = torch.cat([self._heads(self.add_k_proj(ctx)), k], dim=2)
            v = torch.cat([self._heads(self.add_v_proj(ctx)), v], dim=2)
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, -1, D)
        if ctx is not None:
            return self.to_out[0](o[:, Lc:]), self.to_add_out(o[:, :Lc])
        return o


class _GELUProj(nn.Module):
    def __init__(self, dim: int, inner: int):
        super().__init__()
        self.proj = nn.Linear(dim, inner)

    def forward(self, x):
        return F.gelu(self.proj(x), approximate="tanh")


class _FluxFeedForward(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList([_GELUProj(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class _AdaNorm(nn.Module):
    """AdaLayerNormZero-style modulation: `linear` is NOT matched by the attn-mlp filter."""

    def __init__(self, dim: int, n: int):
        super().__init__()
        self.linear = nn.Linear(dim, n * dim)
        self.norm = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.n = n

    def forward(self, x, vec):
        return self.norm(x), self.linear(F.silu(vec))[:, None, :].chunk(self.n, dim=-1)


class FluxTransformerBlock(nn.Module):
    def __init__(self, dim: int, head_dim: int):
        super().__init__()
        self.norm1, self.norm1_context = _AdaNorm(dim, 6), _AdaNorm(dim, 6)
        self.attn = _FluxAttention(dim, head_dim, joint=True)
        self.norm2 = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.ff = _FluxFeedForward(dim)
        self.norm2_context = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.ff_context = _FluxFeedForward(dim)

    def forward(self, x, ctx, vec):
        xn, (sh1, sc1, g1, sh2, sc2, g2) = self.norm1(x, vec)
        cn, (csh1, csc1, cg1, csh2, csc2, cg2) = self.norm1_context(ctx, vec)
        ao, co = self.attn(xn * (1 + sc1) + sh1, cn * (1 + csc1) + csh1)
        x = x + g1 * ao
        x = x + g2 * self.ff(self.norm2(x) * (1 + sc2) + sh2)
        ctx = ctx + cg1 * co
        ctx = ctx + cg2 * self.ff_context(self.norm2_context(ctx) * (1 + csc2) + csh2)
        return x, ctx


class FluxSingleTransformerBlock(nn.Module):
    def __init__(self, dim: int, head_dim: int):
        super().__init__()
        self.norm = _AdaNorm(dim, 3)
        self.proj_mlp = nn.Linear(dim, dim * 4)
        self.attn = _FluxAttention(dim, head_dim, joint=False)
        self.proj_out = nn.Linear(dim * 5, dim)

    def forward(self, x, vec):
        xn, (sh, sc, g) = self.norm(x, vec)
        h = xn * (1 + sc) + sh
        return x + g * self.proj_out(torch.cat([self.attn(h), F.gelu(self.proj_mlp(h), approximate="tanh")], dim=-1))


class StandInFluxTransformer(nn.Module):
    """FluxTransformer2DModel-shaped stand-in: `transformer(hidden_states, timestep, guidance, pooled_projections,
    encoder_hidden_states, txt_ids, img_ids, joint_attention_kwargs, return_dict).sample` on PACKED latents
    [B, (H/2)(W/2), 64].  Same hooked-leaf names, counts and shapes as FLUX.1-dev under the reference's `attn-mlp`
    preset (`["attn", "ff.net"]`, FluxLoRASetup.py:12-16): per double block attn.{to_q,to_k,to_v,to_out.0,add_q_proj,
    add_k_proj,add_v_proj,to_add_out} + ff.net.{0.proj,2} (ff_context does not match the filter), per single block
    attn.{to_q,to_k,to_v}: 19*10 + 38*3 = 304 leaves, 35.5 M LoRA parameters at rank 16.
    Rotary embeddings are omitted (they do not touch the hooked leaves)."""

    def __init__(self, dim: int = 3072, head_dim: int = 128, n_double: int = 19, n_single: int = 38, in_ch: int = 64,
                 ctx_dim: int = 4096, pooled_dim: int = 768, guidance_embeds: bool = True):
        super().__init__()
        self.config = {"guidance_embeds": guidance_embeds, "in_channels": in_ch}
        self.x_embedder = nn.Linear(in_ch, dim)
        self.context_embedder = nn.Linear(ctx_dim, dim)
        self.time_text_embed = nn.ModuleDict({"time": nn.Linear(256, dim), "guidance": nn.Linear(256, dim), "pooled": nn.Linear(pooled_dim, dim)})
        self.transformer_blocks = nn.ModuleList([FluxTransformerBlock(dim, head_dim) for _ in range(n_double)])
        self.single_transformer_blocks = nn.ModuleList([FluxSingleTransformerBlock(dim, head_dim) for _ in range(n_single)])
        self.norm_out = _AdaNorm(dim, 2)
        self.proj_out = nn.Linear(dim, in_ch)

    def forward(self, hidden_states, timestep, guidance=None, pooled_projections=None, encoder_hidden_states=None,
                txt_ids=None, img_ids=None, joint_attention_kwargs=None, return_dict=True):
        dt = hidden_states.dtype
        vec = self.time_text_embed["time"](StandInUNet._sinusoid(timestep * 1000.0).to(dt)) + self.time_text_embed["pooled"](pooled_projections.to(dt))
        if guidance is not None:
            vec = vec + self.time_text_embed["guidance"](StandInUNet._sinusoid(guidance.float() * 1000.0).to(dt))
        x = self.x_embedder(hidden_states)
        ctx = self.context_embedder(encoder_hidden_states)
        for blk in self.transformer_blocks:
            x, ctx = blk(x, ctx, vec)
        Lc = ctx.shape[1]
        h = torch.cat([ctx, x], dim=1)
        for blk in self.single_transformer_blocks:
            h = blk(h, vec)
        x = h[:, Lc:]
        xn, (sh, sc) = self.norm_out(x, vec)
        return _Output(self.proj_out(xn * (1 + sc) + sh))


def flux_transformer() -> StandInFluxTransformer:
    """FLUX.1-dev: 19 double + 38 single blocks, D = 3072, 24 heads x 128, T5 context 4096, 304 hooked leaves."""
    return StandInFluxTransformer()


def tiny_flux(dim: int = 128, ctx_dim: int = 64) -> StandInFluxTransformer:
    """A 2 + 2 block network of the same shape for tests."""
    return StandInFluxTransformer(dim=dim, head_dim=32, n_double=2, n_single=2, in_ch=64, ctx_dim=ctx_dim, pooled_dim=32)


class StandInFlowScheduler:
    """What the hot path reads from diffusers' FlowMatchEulerDiscreteScheduler: only the LENGTH of `timesteps` / `sigmas`
    (ModelSetupFlowMatchingMixin.py:22-25, ModelSetupDiffusionLossMixin.py:297-300)."""

    def __init__(self, num_train_timesteps: int = 1000, device=None):
        self.config = {"num_train_timesteps": num_train_timesteps}
        self.timesteps = torch.arange(num_train_timesteps, 0, -1, dtype=torch.float32, device=device)
        self.sigmas = self.timesteps / num_train_timesteps


class StandInScheduler:
    """The three things the hot path reads from a diffusers DDIMScheduler."""

    def __init__(self, betas: torch.Tensor, prediction_type: str = "epsilon"):
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.config = {"num_train_timesteps": betas.shape[0], "prediction_type": prediction_type}

    @property
    def prediction_type(self):
        return self.config["prediction_type"]

    def get_velocity(self, sample, noise, timesteps):
        """diffusers DDIMScheduler.get_velocity (not vendored in the reference; call site
        BaseStableDiffusionXLSetup.py:285): v = sqrt(abar) * eps - sqrt(1 - abar) * x0."""
        ac = self.alphas_cumprod.to(device=sample.device)
        a = (ac[timesteps] ** 0.5).flatten()
        b = ((1 - ac[timesteps]) ** 0.5).flatten()
        while a.dim() < sample.dim():
            a, b = a.unsqueeze(-1), b.unsqueeze(-1)
        return a * noise - b * sample


def enable_checkpointing(net: nn.Module, preserve_rng_state: bool = True) -> int:
    """Gradient checkpointing as the reference enables it by default (TrainConfig.py:786): every transformer block's
    forward goes through non-reentrant `torch.utils.checkpoint.checkpoint`
    (checkpointing_util.py:150-166, enable_checkpointing_for_basic_transformer_blocks) — activations inside a block are
    dropped after the forward and the block is run a second time in backward, i.e. one extra forward of every hooked
    leaf (55.2 instead of 37.0 TFLOP per SDXL step).  `preserve_rng_state=False` is for steps captured in
    a CUDA graph with dropout off (stashing the generator state is a host read that cannot be captured).
    Returns the number of blocks wrapped."""
    from torch.utils.checkpoint import checkpoint
    n = 0
    for m in net.modules():
        if isinstance(m, (BasicTransformerBlock, FluxTransformerBlock, FluxSingleTransformerBlock)) and not hasattr(m, "_plain_forward"):
            m._plain_forward = m.forward

            def fwd(*args, _m=m, **kwargs):
                if torch.is_grad_enabled():
                    return checkpoint(_m._plain_forward, *args, use_reentrant=False, preserve_rng_state=preserve_rng_state, **kwargs)
                return _m._plain_forward(*args, **kwargs)
            m.forward = fwd
            n += 1
    return n


def disable_checkpointing(net: nn.Module) -> None:
    for m in net.modules():
        if hasattr(m, "_plain_forward"):
            m.forward = m._plain_forward
            del m._plain_forward


def count_lora_leaves(net: nn.Module, filters: list[str]) -> int:
    return sum(1 for name, m in net.named_modules()
               if isinstance(m, (nn.Linear, nn.Conv2d)) and any(f in name for f in filters))

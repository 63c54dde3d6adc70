"""Training side of the DiT hot path: the backward of DGSDenoiser.image_to_gaussians as ONE
autograd node over libdgs_b200.so (dgs_dit_backward), flat fp32 master parameters / gradients / Adam moments, the
fused AdamW update (diffusionGS/configs/diffusionGS_rel.yaml:57-62) and the gradient all-reduce (dist.GradArena).

The reference trains through torch autograd with `torch.utils.checkpoint` around every block (denoiser.py:348-354),
i.e. it runs the forward twice.  Here the forward keeps its activations in one big HBM buffer (4 GB per sample at
N = 4098 tokens; a B200 has 180 GB) and the backward consumes them: fwd + 2x bwd FLOPs instead of 2x fwd + 2x bwd.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import DitGrads, DitOutGrads, DitWeightsT, check
from .dist import GradArena


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class DitTrainer:
    """Owns the flat training state of one DGSDenoiser (one process per GPU).

    * `master`: every parameter, fp32, in module.parameters() order; `p.data` are views of it;
    * `arena`:  the gradients in the same layout (dist.GradArena: `p.grad` are views; one bucket per block);
    * `exp_avg`, `exp_avg_sq`: AdamW moments, same layout -> the optimizer is ONE kernel over the whole model;
    * bf16 / transposed-bf16 GEMM weights are re-derived from `master` after each step (refresh_weights)."""

    def __init__(self, model, lr=1e-5, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.01, clip=0.5):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay, self.clip = lr, betas, eps, weight_decay, clip
        params = [p for p in model.parameters()]
        dev = params[0].device
        if dev.type != "cuda":
            raise _lib.DgsError("DitTrainer needs the model on a CUDA device (no CPU fallback)")
        total = sum(p.numel() for p in params)
        self.master = torch.empty(total, dtype=torch.float32, device=dev)
        off = 0
        for p in params:
            n = p.numel()
            self.master[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.master[off:off + n].view_as(p)
            off += n
        self.arena = GradArena(model)
        assert self.arena.total == total
        self.exp_avg = torch.zeros_like(self.master)
        self.exp_avg_sq = torch.zeros_like(self.master)
        self.steps = 0
        self._state = None
        self._grads = self._grad_struct()
        self.anchor = torch.zeros(1, device=dev, requires_grad=True)  # makes autograd call our backward
        model._trainer = self
        self.refresh_weights()

    # -- pointers into the gradient arena, in the layout dgs_dit_grads wants --
    def _grad_struct(self):
        m = self.model
        T = m.transformer
        g = DitGrads()
        ptr = lambda p: p.grad.data_ptr()  # noqa: E731
        g.tokenizer_w, g.pos_embed, g.in_ln_w = ptr(m.image_tokenizer[1].weight), ptr(m.gaussians_pos_embedding), \
            ptr(m.transformer_input_layernorm.weight)
        g.t0_w, g.t0_b = ptr(m.t_embedder.mlp[0].weight), ptr(m.t_embedder.mlp[0].bias)
        g.t2_w, g.t2_b = ptr(m.t_embedder.mlp[2].weight), ptr(m.t_embedder.mlp[2].bias)
        fields = dict(qkv_w=lambda b: b.attn.qkv.weight, qkv_b=lambda b: b.attn.qkv.bias,
                      proj_w=lambda b: b.attn.proj.weight, proj_b=lambda b: b.attn.proj.bias,
                      fc1_w=lambda b: b.mlp.fc1.weight, fc1_b=lambda b: b.mlp.fc1.bias,
                      fc2_w=lambda b: b.mlp.fc2.weight, fc2_b=lambda b: b.mlp.fc2.bias,
                      adaln_w=lambda b: b.adaLN_modulation[1].weight, adaln_b=lambda b: b.adaLN_modulation[1].bias)
        stride = None
        for name, get in fields.items():
            setattr(g, name, ptr(get(T[0])))
            if len(T) > 1:
                st = (ptr(get(T[1])) - ptr(get(T[0]))) // 4
                assert stride in (None, st), "transformer blocks are not laid out with one common stride"
                stride = st
                assert all(ptr(get(T[i])) == ptr(get(T[0])) + 4 * st * i for i in range(len(T)))
        g.layer_stride = stride or 0
        u, d = m.upsampler, m.image_token_decoder
        g.ups_ln_w, g.ups_w = ptr(u.layernorm.weight), ptr(u.linear.weight)
        g.ups_adaln_w, g.ups_adaln_b = ptr(u.adaLN_modulation[1].weight), ptr(u.adaLN_modulation[1].bias)
        g.dec_ln_w, g.dec_w = ptr(d.layernorm.weight), ptr(d.linear.weight)
        g.dec_adaln_w, g.dec_adaln_b = ptr(d.adaLN_modulation[1].weight), ptr(d.adaLN_modulation[1].bias)
        return g

    _BIG = dict(qkv_w=lambda b: b.attn.qkv.weight, proj_w=lambda b: b.attn.proj.weight,
                fc1_w=lambda b: b.mlp.fc1.weight, fc2_w=lambda b: b.mlp.fc2.weight)

    def refresh_weights(self):
        """fp32 master -> the bf16 stacks the forward reads + the transposed bf16 stacks the dgrad GEMMs read.
        The packed tensors are allocated once and updated IN PLACE (the dgs_dit_weights struct keeps its pointers):
        the four big per-block matrices by one fused cast+transpose launch each (dgs_cast_transpose_f32, batched over
        the blocks through the master arena's block stride), the small vectors / split-bf16 end matrices by torch copies."""
        m = self.model
        T = m.transformer
        dev = self.master.device
        first = m._packed is None or getattr(self, "_wT_keep", None) is None
        if first:
            m.packed_weights(force=True)
            self._wT_keep = {k + "T": torch.empty(len(T), get(T[0]).shape[1], get(T[0]).shape[0], dtype=torch.bfloat16,
                                                  device=dev) for k, get in self._BIG.items()}
        w, t = m._packed
        L = _lib.lib()
        stride = int(self._grads.layer_stride)  # master and gradient arenas share one layout
        with torch.cuda.device(dev):
            for k, get in self._BIG.items():
                p0 = get(T[0])
                check(L.dgs_cast_transpose_f32(p0.data_ptr(), stride, len(T), p0.shape[0], p0.shape[1], t[k].data_ptr(),
                                               self._wT_keep[k + "T"].data_ptr(), _stream(dev)))
        if not first:
            for k, v in m._pack_dict(skip=tuple(self._BIG)).items():
                t[k].copy_(v)
        m._packed_key = m._pack_key()
        self._wT_keep["dec_wT"] = m.image_token_decoder.linear.weight.detach().t().to(torch.bfloat16).contiguous()
        self._wT_keep["ups_w"] = m.upsampler.linear.weight.detach().float().contiguous()
        wT = DitWeightsT()
        for k, v in self._wT_keep.items():
            setattr(wT, k, v.data_ptr())
        self._wT = wT

    def train_state(self, B, V, H, W):
        w, _ = self.model.packed_weights()
        n = _lib.lib().dgs_dit_train_state_bytes(C.byref(w), B, V, H, W)
        if n == 0:
            raise _lib.DgsError(_lib.lib().dgs_last_error().decode())
        if self._state is None or self._state.numel() < n:
            self._state = None
            self._state = torch.empty(n, dtype=torch.uint8, device=self.master.device)
        return self._state

    def zero_grad(self):
        self.arena.zero_()

    def optimizer_step(self, allreduce=True):
        """all-reduce (mean) -> clip at `clip` (Lightning gradient_clip_val) -> fused AdamW -> refresh bf16 weights."""
        if allreduce:
            self.arena.allreduce_mean_()
        norm = scale = None
        if self.clip:  # torch clip_grad_norm_ semantics; the factor stays on the device and is applied inside AdamW
            norm = torch.linalg.vector_norm(self.arena.flat)
            scale = torch.clamp(self.clip / (norm + 1e-6), max=1.0).reshape(1).float()
        self.steps += 1
        dev = self.master.device
        with torch.cuda.device(dev):
            check(_lib.lib().dgs_adamw_step(self.master.data_ptr(), self.arena.flat.data_ptr(), self.exp_avg.data_ptr(),
                                            self.exp_avg_sq.data_ptr(), self.master.numel(), self.lr, self.betas[0],
                                            self.betas[1], self.eps, self.weight_decay, self.steps, 1.0,
                                            None if scale is None else scale.data_ptr(), _stream(dev)))
        self.refresh_weights()
        return norm


class _DitFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, images, ray_o, ray_d, t, anchor):
        tr = model._trainer
        B, V, _, H, W = images.shape
        state = tr.train_state(B, V, H, W)
        out, img_xyz, _, keep = model._run_dit(images, ray_o, ray_d, t, train_state=state)
        ctx.model, ctx.keep = model, keep
        ctx.mark_non_differentiable(img_xyz)
        return out.xyz, out.features, out.scaling, out.rotation, out.opacity, img_xyz

    @staticmethod
    def backward(ctx, d_xyz, d_features, d_scaling, d_rotation, d_opacity, _d_img):
        model = ctx.model
        tr = model._trainer
        io, ws, nbytes, images, *_ = ctx.keep
        w, _k = ctx.keep[7], ctx.keep[8]
        dev = images.device
        B, V, _, H, W = images.shape
        P = model.cfg.n_gaussians + V * H * W
        z = lambda g, *s: (torch.zeros(*s, device=dev) if g is None else g.float().contiguous())  # noqa: E731
        gs = [z(d_xyz, B, P, 3), z(d_features, B, P, 1, 3), z(d_scaling, B, P, 3), z(d_rotation, B, P, 4),
              z(d_opacity, B, P, 1)]
        dout = DitOutGrads(*(g.data_ptr() for g in gs))
        with torch.cuda.device(dev):
            check(_lib.lib().dgs_dit_backward(C.byref(w), C.byref(tr._wT), C.byref(io), C.byref(dout),
                                              C.byref(tr._grads), ws.data_ptr(), nbytes, _stream(dev)))
        # parameter gradients were written straight into the arena (p.grad views); nothing to hand to autograd
        return None, None, None, None, None, torch.zeros_like(tr.anchor)


def dit_train_forward(model, images, ray_o, ray_d, t):
    from .denoiser import AttrDict
    xyz, features, scaling, rotation, opacity, img_xyz = _DitFunction.apply(model, images, ray_o, ray_d, t,
                                                                           model._trainer.anchor)
    return AttrDict(xyz=xyz, features=features, scaling=scaling, rotation=rotation, opacity=opacity), img_xyz

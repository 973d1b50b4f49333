"""Stage-1 fine-tuning step of the reference recipe on B200 kernels.

Reference (src/main.py:210-223): ``model.layers[0].trainable = False`` freezes the whole Wav2Vec2 body, so the only
trainable variables are ``lm_head/kernel`` and ``lm_head/bias`` (24 608 parameters for the base model);
``optimizer = Adam(1e-3)``; ``loss = CTCLoss(config, input_shape, division_factor=global_batch)``; Keras ``fit`` then runs,
per step: forward (training=True) -> CTC loss -> gradients -> cross-replica SUM all-reduce (MirroredStrategy/TPUStrategy,
main.py:141-156) -> Adam.  Here every arithmetic step is a kernel behind the C ABI:
forward (all the inference kernels), ``w2v2_ctc_loss`` (loss + d loss / d logits), ``w2v2_lm_head_wgrad``, ONE
``torch.distributed.all_reduce`` over the flat gradient buffer (NCCL on GPUs), ``w2v2_adam``.

Stage 2 (main.py:234-250: ``freeze_feature_extractor()``, everything else trains) is ``Stage2Trainer`` below: the full
backward through the encoder on the same kernels, dropout at the reference's six sites from a stateless counter-based
generator, SpecAugment and StochasticDepth.  Limits (both trainers): the training forward with dropout / StochasticDepth
covers the base architecture (group-norm extractor, post-norm encoder) without an attention mask - which is what
main.py:128-133 fine-tunes; other combinations raise ``NotImplementedError`` instead of silently skipping a regulariser.
"""
import os

import numpy as np
import torch
import torch.distributed as dist

from . import ops
from .losses import CTCLoss
from .modeling import Wav2Vec2ForCTC


class Stage1Trainer:
    def __init__(self, model: Wav2Vec2ForCTC, loss_fn: CTCLoss, learning_rate=1e-3, beta_1=0.9, beta_2=0.999,
                 epsilon=1e-7):
        self.model, self.loss_fn = model, loss_fn
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta_1, beta_2, epsilon
        self.t = 0
        k, b = model.variables["lm_head/kernel"], model.variables["lm_head/bias"]
        self.d, self.V = k.shape
        # trainable variables live in ONE flat fp32 buffer (a single all-reduce message, a single Adam launch)
        self.flat_w = torch.cat([k.reshape(-1), b.reshape(-1)]).contiguous()
        model.variables["lm_head/kernel"] = self.flat_w[: self.d * self.V].view(self.d, self.V)
        model.variables["lm_head/bias"] = self.flat_w[self.d * self.V:]
        self.flat_g = torch.zeros_like(self.flat_w)
        self.m = torch.zeros_like(self.flat_w)
        self.v = torch.zeros_like(self.flat_w)
        for name in model.trainable:
            model.trainable[name] = name.startswith("lm_head/")
        model._invalidate_graphs()          # lm_head variables were re-bound to views of flat_w

    @torch.no_grad()
    def step(self, speech, labels, attention_mask=None):
        """One optimisation step on this rank's shard; returns this rank's (already 1/division_factor-scaled) loss."""
        model = self.model
        logits, hidden = model.forward_with_hidden(speech, attention_mask=attention_mask, training=True)
        loss, grad_logits = self.loss_fn(labels, logits, return_grad=True)
        B, T, V = logits.shape
        gk = self.flat_g[: self.d * self.V].view(self.d, self.V)
        gb = self.flat_g[self.d * self.V:]
        ops.lm_head_wgrad(hidden, grad_logits.view(B * T, V), gk, gb)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM)          # the step's single collective
        self.t += 1
        lr_t = self.lr * (1.0 - self.b2 ** self.t) ** 0.5 / (1.0 - self.b1 ** self.t)
        ops.adam(self.flat_w, self.flat_g, self.m, self.v, lr_t, self.b1, self.b2, self.eps)
        model._repack_lm_head()
        return loss


# ============================================================================================ stage 2
from .modeling import _PRECISIONS, pack_posconv_kernel  # noqa: E402
from .ops import Pair  # noqa: E402


def pack_posconv(kern: torch.Tensor, groups: int) -> torch.Tensor:
    """bf16 posconv-kernel layout of a TF grouped-conv kernel [k, cin/groups, cout]."""
    return pack_posconv_kernel(kern, groups).to(torch.bfloat16)


def transposed_conv_kernel(kern: torch.Tensor, groups: int) -> torch.Tensor:
    """TF grouped-conv kernel [k, cin/groups, cout] -> the kernel of its input gradient in the same layout: taps flipped,
    in/out channels swapped inside each group.  A forward convolution with it, the tap window shifted by one frame
    (``shift=1``), computes dx[u] = sum_j W_j^T dpre[u - j + k/2]."""
    k, cpg, d = kern.shape
    return kern.reshape(k, cpg, groups, cpg).flip(0).permute(0, 3, 2, 1).reshape(k, cpg, d)


def pack_posconv_transposed(kern: torch.Tensor, groups: int) -> torch.Tensor:
    """``transposed_conv_kernel`` in the posconv kernel's bf16 layout (fed to w2v2_posconv with linear=1, shift=1)."""
    return pack_posconv(transposed_conv_kernel(kern, groups), groups)


def weight_norm_backward(d_kernel: torch.Tensor, weight_v: torch.Tensor, weight_g: torch.Tensor):
    """Chain rule of Conv1DWithWeightNorm (tensorflow_addons.py:16-21): kernel = g * v / ||v||, the norm taken over axes
    (1, 2) of every tap with l2_normalize's 1e-12 floor.  Returns (d weight_v, d weight_g)."""
    nrm = torch.sqrt(torch.clamp(weight_v.pow(2).sum(dim=(1, 2), keepdim=True), min=1e-12))
    dot = (d_kernel * weight_v).sum(dim=(1, 2), keepdim=True)
    return weight_g / nrm * (d_kernel - weight_v * dot / (nrm * nrm)), dot / nrm


class Stage2Trainer:
    """Stage-2 fine-tune step of the reference recipe (src/main.py:234-250): ``model.freeze_feature_extractor()``, then
    Keras ``fit`` differentiates everything else - feature projection, positional conv (through its weight norm),
    encoder LayerNorms, all transformer layers, ``masked_spec_embed`` and ``lm_head`` (90 195 872 parameters for base) -
    sums the gradients over replicas and applies Adam.

    Every arithmetic step of forward and backward is a kernel behind the C ABI: the inference kernels (forward, with the
    activations the backward needs kept in the arena), ``w2v2_ctc_loss`` (loss + dlogits), the tcgen05 GEMM for every
    dgrad / wgrad product (wgrad with MN-major operands: no transposed copies), ``w2v2_ln_bwd``, ``w2v2_dact_colsum``,
    ``w2v2_attn_bwd``, ``w2v2_posconv`` (transposed-conv mode) + ``w2v2_posconv_wgrad``, then ONE ``all_reduce(SUM)``
    over the flat fp32 gradient buffer and ONE ``w2v2_adam`` launch.  Parameter-sized algebra (weight-norm chain rule,
    bf16 casts of the updated kernels) is host-side torch on the parameter tensors.

    Dropout (config.py:9, rate 0.1 by default) is applied at the reference's six sites - after the feature projection, on
    the attention probabilities (inside the attention kernel), after the attention output projection, on the GELU
    intermediate, after the encoder LayerNorm and before ``lm_head`` - with masks from a stateless counter-based generator
    (``include/w2v2.h``), so the backward pass regenerates them instead of storing them.
    Limits: post-norm encoder with the group-norm extractor (the base architecture of BASELINE config 3), no attention
    mask; SpecAugment masks and the StochasticDepth draw of the FFN branch (``survival_prob``) use the host RNG like the
    reference's numpy / Keras RNG.
    Backward products run single-pass bf16 with fp32 accumulation; LayerNorm / softmax / GELU derivatives in fp32.
    """

    # dropout sites (tf.keras.layers.Dropout in the reference): id of the counter-based mask stream
    SITE_PROJ, SITE_ENC, SITE_HEAD = 1, 2, 3                    # feature_extractor.py:95, encoder.py:270, modeling.py:253

    @staticmethod
    def site_attn_probs(i):                                      # encoder.py:41-43
        return 16 + 4 * i

    @staticmethod
    def site_attn_out(i):                                        # encoder.py:118
        return 17 + 4 * i

    @staticmethod
    def site_ffn_mid(i):                                         # encoder.py:128
        return 18 + 4 * i

    def _drop(self, site):
        """(rate, seed, site) of this step's mask stream; rate 0 = off."""
        return (float(self.model.config.dropout), self.seed + 0x9E3779B1 * (self.t + 1), site)

    @classmethod
    def forward_only(cls, model, seed=0):
        """The training-mode forward (dropout + SpecAugment) without optimizer state: what ``model(x, training=True)`` runs."""
        self = cls.__new__(cls)
        self.model, self.loss_fn, self.seed, self.t, self.saved, self._wt = model, None, int(seed), 0, None, None
        return self

    @staticmethod
    def supports(cfg):
        return cfg.attention_norm_type == "postnorm" and cfg.feature_extractor_norm_type == "group" and not cfg.is_gelu_approx

    def __init__(self, model: Wav2Vec2ForCTC, loss_fn: CTCLoss, learning_rate=5e-5, beta_1=0.9, beta_2=0.999,
                 epsilon=1e-7, seed=0):
        cfg = model.config
        if cfg.attention_norm_type != "postnorm" or cfg.feature_extractor_norm_type != "group":
            raise NotImplementedError("Stage2Trainer covers the base architecture (group-norm extractor, post-norm encoder)")
        if cfg.is_gelu_approx:
            raise NotImplementedError("Stage2Trainer differentiates the erf GELU (config.is_gelu_approx=False, the reference default)")
        if model.precision not in ("bf16", "bf16x3"):
            raise NotImplementedError("Stage2Trainer runs its forward in precision 'bf16' or 'bf16x3' (backward products are bf16)")
        self.model, self.loss_fn = model, loss_fn
        self.seed = int(seed)
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta_1, beta_2, epsilon
        self.t = 0
        model.freeze_feature_extractor()                               # main.py:236-237
        self.names = [n for n in model.variables if model.trainable[n]]
        sizes = [model.variables[n].numel() for n in self.names]
        # all trainable variables in ONE flat fp32 buffer: a single all-reduce message, a single Adam launch
        self.flat_w = torch.cat([model.variables[n].reshape(-1) for n in self.names]).contiguous()
        self.flat_g = torch.zeros_like(self.flat_w)
        self.m = torch.zeros_like(self.flat_w)
        self.v = torch.zeros_like(self.flat_w)
        self.G = {}
        off = 0
        for n, sz in zip(self.names, sizes):
            shape = model.variables[n].shape
            model.variables[n] = self.flat_w[off: off + sz].view(shape)
            self.G[n] = self.flat_g[off: off + sz].view(shape)
            off += sz
        model._packed = None
        self._wt = None
        self.saved = None

    # ------------------------------------------------------------------ operand packs of the backward GEMMs
    def _pack_backward(self):
        """dgrad weight operands: the TF kernels themselves ([in, out] row-major == W^T, K-major) as bf16."""
        m, v, cfg = self.model, self.model.variables, self.model.config
        bf = torch.bfloat16
        W = {"proj": Pair(v["wav2vec2/feature_projection/projection/kernel"].to(bf))}
        pc = "wav2vec2/encoder/pos_conv_embed/conv/"
        wv, wg = v[pc + "weight_v"], v[pc + "weight_g"]
        kern = wv * torch.rsqrt(torch.clamp(wv.pow(2).sum(dim=(1, 2), keepdim=True), min=1e-12)) * wg
        W["pos.T"] = Pair(pack_posconv_transposed(kern, cfg.num_conv_pos_embedding_groups))
        for i in range(cfg.num_layers):
            a = f"wav2vec2/encoder/layers/{i}/attention/"
            f = f"wav2vec2/encoder/layers/{i}/feed_forward/"
            W[f"l{i}.qkv"] = Pair(torch.cat([v[a + "q_proj/kernel"], v[a + "k_proj/kernel"], v[a + "v_proj/kernel"]], 1).to(bf))
            W[f"l{i}.out"] = Pair(v[a + "out_proj/kernel"].to(bf))
            W[f"l{i}.ff1"] = Pair(v[f + "intermediate_dense/kernel"].to(bf))
            W[f"l{i}.ff2"] = Pair(v[f + "output_dense/kernel"].to(bf))
        self._wt = W
        return W

    # ------------------------------------------------------------------ one-launch re-packing after the optimizer step
    def _build_pack_jobs(self):
        """Job / tile tables of w2v2_pack_weights: every trainable Dense kernel -> its forward operand (transposed, q rows
        scaled, q/k/v fused) and its dgrad operand (the TF kernel as bf16).  The tables are static: variables are views
        of ``flat_w`` and the packed operands are persistent tensors that are overwritten in place."""
        import ctypes as C

        import numpy as np

        from ._lib import PackJob
        m, v, cfg = self.model, self.model.variables, self.model.config
        P, W = m._packed, self._wt
        d = cfg.hidden_size
        scale = cfg.head_size ** (-0.5)
        jobs = []

        def job(src, dst, dst_off, rows, cols, dst_ld, transpose, s=1.0, f32=False):
            esz = 4 if f32 else 2
            jobs.append(PackJob(src.data_ptr(), dst.data_ptr() + dst_off * esz, rows, cols, cols, dst_ld, int(transpose),
                                int(f32), float(s), 0))
        fp = "wav2vec2/feature_projection/projection/kernel"
        Cl = v[fp].shape[0]
        job(v[fp], P["proj.w"].hi, 0, Cl, d, Cl, True)
        job(v[fp], W["proj"].hi, 0, Cl, d, d, False)
        for i in range(cfg.num_layers):
            a = f"wav2vec2/encoder/layers/{i}/attention/"
            f = f"wav2vec2/encoder/layers/{i}/feed_forward/"
            ff = v[f + "intermediate_dense/kernel"].shape[1]
            for j, n in enumerate(("q", "k", "v")):
                sc = scale if n == "q" else 1.0
                job(v[a + f"{n}_proj/kernel"], P[f"l{i}.qkv.w"].hi, j * d * d, d, d, d, True, sc)
                job(v[a + f"{n}_proj/bias"], P[f"l{i}.qkv.b"], j * d, 1, d, d, False, sc, f32=True)
                job(v[a + f"{n}_proj/kernel"], W[f"l{i}.qkv"].hi, j * d, d, d, 3 * d, False)
            job(v[a + "out_proj/kernel"], P[f"l{i}.out.w"].hi, 0, d, d, d, True)
            job(v[a + "out_proj/kernel"], W[f"l{i}.out"].hi, 0, d, d, d, False)
            job(v[f + "intermediate_dense/kernel"], P[f"l{i}.ff1.w"].hi, 0, d, ff, d, True)
            job(v[f + "intermediate_dense/kernel"], W[f"l{i}.ff1"].hi, 0, d, ff, ff, False)
            job(v[f + "output_dense/kernel"], P[f"l{i}.ff2.w"].hi, 0, ff, d, ff, True)
            job(v[f + "output_dense/kernel"], W[f"l{i}.ff2"].hi, 0, ff, d, d, False)
        V = v["lm_head/kernel"].shape[1]
        job(v["lm_head/kernel"], P["lm.w"].hi, 0, d, V, d, True)
        tiles = []
        for ji, jb in enumerate(jobs):
            rr, cc = np.meshgrid(np.arange(0, jb.rows, 64), np.arange(0, jb.cols, 64), indexing="ij")
            tiles.append(np.stack([np.full(rr.size, ji), rr.ravel(), cc.ravel()], 1))
        tiles = np.concatenate(tiles).astype(np.int32)
        raw = (PackJob * len(jobs))(*jobs)
        jobs_dev = torch.frombuffer(bytearray(bytes(raw)), dtype=torch.uint8).to(m.device)
        self._pack_tables = (jobs_dev, torch.from_numpy(tiles).to(m.device), int(tiles.shape[0]))
        self._pack_refs = (P, W)                      # the packed tensors the job table points into stay alive

    def _repack(self):
        """Kernel-layout copies follow the optimizer step: one w2v2_pack_weights launch for every Dense kernel, host torch
        algebra for the weight-normalised positional conv (4.7 M parameters)."""
        model = self.model
        if _PRECISIONS[model.precision] != 1 or os.environ.get("W2V2_TORCH_REPACK", "0") == "1":
            model._packed = None                      # split-bf16 (hi + lo) operands: re-packed by the host at the next forward
            self._wt = None
            return
        if model._packed is None:
            model._pack()
        if self._wt is None:
            self._pack_backward()
        if getattr(self, "_pack_tables", None) is None or self._pack_refs[0] is not model._packed or self._pack_refs[1] is not self._wt:
            self._build_pack_jobs()
        ops.pack_weights(*self._pack_tables)
        model._fold_stale = True                      # the LayerNorm-folded inference copies are re-derived at the next eval call
        cfg, v = model.config, model.variables
        pc = "wav2vec2/encoder/pos_conv_embed/conv/"
        wv, wg = v[pc + "weight_v"], v[pc + "weight_g"]
        kern = wv * torch.rsqrt(torch.clamp(wv.pow(2).sum(dim=(1, 2), keepdim=True), min=1e-12)) * wg
        G_ = cfg.num_conv_pos_embedding_groups
        model._packed["pos.w"].hi.copy_(pack_posconv(kern, G_))
        self._wt["pos.T"].hi.copy_(pack_posconv_transposed(kern, G_))

    # ------------------------------------------------------------------ forward keeping what the backward needs
    def _forward(self, speech, spec_mask=None, layer_keep=None):
        model = self.model
        cfg, v = model.config, model.variables
        pre = self.__dict__.pop("_prefetched", None)
        if pre is not None and pre[0] is speech:
            last_f32, B, T = pre[1]                       # extractor output computed during the previous step's all-reduce
        else:
            last_f32, B, T = model._features(speech)      # frozen extractor: the inference kernels, nothing kept
        P, A = model._packed, model._arena
        passes = _PRECISIONS[model.precision]
        lo = passes == 3
        f32, eps = torch.float32, cfg.layer_norm_eps
        Cl, d, ffn = cfg.filter_sizes[-1], cfg.hidden_size, cfg.intermediate_size
        H, dh = cfg.num_heads, cfg.head_size
        M = B * T
        S = {"B": B, "T": T, "last_f32": last_f32}
        fp = "wav2vec2/feature_projection/"
        pn = A.pair("proj.in", (M, Cl), lo)
        ops.ln_rows(last_f32, v[fp + "layer_norm/gamma"], v[fp + "layer_norm/beta"], eps, M, Cl, out_hi=pn.hi, out_lo=pn.lo)
        h_f32 = A.get("h.f32", (M, d), f32)
        h = A.pair("h", (M, d), lo)
        p_drop = float(cfg.dropout)
        S["dropout"] = p_drop

        def resplit(t_f32, pair):
            sp = ops.split_bf16(t_f32, lo)
            pair.hi.copy_(sp.hi)
            if lo:
                pair.lo.copy_(sp.lo)
        # dropout after the projection (feature_extractor.py:95) and the SpecAugment replacement of masked frames by
        # masked_spec_embed (modeling.py:193-199; span starts from the host numpy RNG like the reference) both happen in the
        # projection GEMM's epilogue: no extra pass over h
        S["spec_mask"], row_replace = None, None
        if cfg.apply_spec_augment:
            if spec_mask is None:
                from .spec_augment import _compute_mask_indices
                spec_mask = _compute_mask_indices((B, T), cfg.mask_time_prob, cfg.mask_time_length, min_masks=2)
            mask = torch.as_tensor(spec_mask).to(model.device).bool().reshape(M)
            S["spec_mask"] = mask
            row_replace = (mask.to(torch.uint8).contiguous(), v["wav2vec2/masked_spec_embed"])
        ops.gemm(pn, P["proj.w"], K=Cl, N=d, rows_per_batch=T, batch=B, bias=v[fp + "projection/bias"], out_f32=h_f32,
                 out_hi=h.hi, out_lo=h.lo, passes=passes, row_replace=row_replace,
                 drop=self._drop(self.SITE_PROJ) if p_drop else None)
        S["pn"], S["h"] = pn, h
        enc = "wav2vec2/encoder/"
        y0 = A.get("t.y0", (M, d), f32)
        pos_pre = A.get("t.pos.pre", (M, d), f32)
        ops.posconv_train(h, P["pos.w"], v[enc + "pos_conv_embed/conv/bias"], h_f32, y0, B, T, d,
                          cfg.num_conv_pos_embedding_groups, cfg.num_conv_pos_embeddings, passes, pre_out=pos_pre)
        S["y0"], S["pos_pre"] = y0, pos_pre
        xs_f32 = A.get("x.f32", (M, d), f32)
        x1_f32 = A.get("x1.f32", (M, d), f32)
        xs = A.pair("t.xs.0", (M, d), lo)
        ops.ln_rows(y0, v[enc + "layer_norm/gamma"], v[enc + "layer_norm/beta"], eps, M, d, out_f32=xs_f32, out_hi=xs.hi,
                    out_lo=xs.lo)
        if p_drop:                                          # encoder.py:270
            ops.dropout_rows(xs_f32, self._drop(self.SITE_ENC), out_f32=xs_f32)
            resplit(xs_f32, xs)
        L = []
        for i in range(cfg.num_layers):
            lb = f"{enc}layers/{i}/"
            qkv = A.pair(f"t.qkv.{i}", (M, 3 * d), lo)
            ops.gemm(xs, P[f"l{i}.qkv.w"], K=d, N=3 * d, rows_per_batch=M, bias=P[f"l{i}.qkv.b"], out_hi=qkv.hi, out_lo=qkv.lo,
                     passes=passes)
            ctx = A.pair(f"t.ctx.{i}", (M, d), lo)
            y1 = A.get(f"t.y1.{i}", (M, d), f32)
            if p_drop:
                ops.attn_fwd_train(qkv, B, T, H, dh, None, ctx, passes, self._drop(self.site_attn_probs(i)))
                # y1 = x + dropout(out_proj(ctx)) (encoder.py:118-119): the dropout draws (same element stream as w2v2_dropout_rows:
                # index = row * d + column) and the residual add happen in the GEMM epilogue
                ops.gemm(ctx, P[f"l{i}.out.w"], K=d, N=d, rows_per_batch=M, bias=v[lb + "attention/out_proj/bias"],
                         residual=xs_f32, out_f32=y1, passes=passes, drop=self._drop(self.site_attn_out(i)))
            else:
                ops.attn_fwd(qkv, B, T, H, dh, None, ctx, passes)
                ops.gemm(ctx, P[f"l{i}.out.w"], K=d, N=d, rows_per_batch=M, bias=v[lb + "attention/out_proj/bias"],
                         residual=xs_f32, out_f32=y1, passes=passes)
            x1 = A.pair(f"t.x1.{i}", (M, d), lo)
            ops.ln_rows(y1, v[lb + "layer_norm/gamma"], v[lb + "layer_norm/beta"], eps, M, d, out_f32=x1_f32, out_hi=x1.hi,
                        out_lo=x1.lo)
            # StochasticDepth on the FFN branch (encoder.py:130, tensorflow_addons.py:374-394): ONE Bernoulli(survival_prob)
            # draw per layer call decides whether the branch is added at all (host RNG, numpy like SpecAugment)
            keep = True
            if layer_keep is not None:
                keep = bool(layer_keep[i])
            elif cfg.survival_prob < 1.0:
                keep = bool(np.random.rand() < cfg.survival_prob)
            y2 = A.get(f"t.y2.{i}", (M, d), f32)
            pre = mid = None
            if keep:
                pre = A.get(f"t.pre.{i}", (M, ffn), f32)
                ops.gemm(x1, P[f"l{i}.ff1.w"], K=d, N=ffn, rows_per_batch=M, bias=v[lb + "feed_forward/intermediate_dense/bias"],
                         out_f32=pre, passes=passes)
                mid = A.pair(f"t.mid.{i}", (M, ffn), lo)
                ops.gelu_rows(pre, mid.hi, fast=(passes == 1), out_lo=mid.lo,
                              drop=self._drop(self.site_ffn_mid(i)) if p_drop else ops.NO_DROP)
                ops.gemm(mid, P[f"l{i}.ff2.w"], K=ffn, N=d, rows_per_batch=M, bias=v[lb + "feed_forward/output_dense/bias"],
                         residual=x1_f32, out_f32=y2, passes=passes)
            else:
                y2.copy_(x1_f32)
            nxt = A.pair(f"t.xs.{i + 1}", (M, d), lo)
            ops.ln_rows(y2, v[lb + "final_layer_norm/gamma"], v[lb + "final_layer_norm/beta"], eps, M, d, out_f32=xs_f32,
                        out_hi=nxt.hi, out_lo=nxt.lo)
            L.append(dict(xs=xs, qkv=qkv, ctx=ctx, y1=y1, x1=x1, pre=pre, mid=mid, y2=y2, ffn_on=keep))
            xs = nxt
        S["layers"], S["hidden_f32"] = L, xs_f32
        self.saved = S
        if not model.with_head:                             # Wav2Vec2Model: hidden states, no head dropout
            return None
        V = cfg.vocab_size
        logits = A.get("t.logits", (B, T, V), f32)
        hidden_f32 = xs_f32
        if p_drop:                                          # modeling.py:253
            hidden_f32 = A.get("t.hidden.drop", (M, d), f32)
            ops.dropout_rows(xs_f32, self._drop(self.SITE_HEAD), out_f32=hidden_f32)
            xs = A.pair("t.hidden.drop.op", (M, d), lo)
            resplit(hidden_f32, xs)
        ops.gemm(xs, P["lm.w"], K=d, N=V, rows_per_batch=M, bias=v["lm_head/bias"], out_f32=logits, passes=passes, block_n=32)
        S["layers"], S["hidden_f32"] = L, hidden_f32
        self.saved = S
        return logits

    # ------------------------------------------------------------------ backward
    def _wgrad(self, x_hi, dy_hi, M, n_in, n_out, out, dy_ld=None):
        """out[n_in][n_out] = x^T dy (the TF Dense kernel layout) straight from the row-major activations: MN-major operands,
        the reduction runs over the M rows (W2V2_GEMM_MN_MAJOR), no transposed copies."""
        # few output tiles, long reduction: split the M rows over up to 8 slices so that ~all SMs work; the slices add into the
        # (zeroed) gradient buffer with fp32 atomics
        tiles = ((n_in + 127) // 128) * ((n_out + 127) // 128)
        kb = (M + 63) // 64
        splits = max(1, min(8, 148 // tiles, kb))
        kb_per = (kb + splits - 1) // splits
        splits = (kb + kb_per - 1) // kb_per
        ops.gemm(Pair(x_hi), Pair(dy_hi), K=kb_per * 64, N=n_out, rows_per_batch=n_in, batch=splits, a_rows=M, a_row_stride=n_in,
                 a_batch_stride=0, out_f32=out, mn_major=True, w_row_stride=n_out if dy_ld is None else dy_ld, cluster=1, block_n=128)

    def _backward(self, dlogits):
        model, G, S = self.model, self.G, self.saved
        cfg, v, A = model.config, model.variables, model._arena
        W = self._wt or self._pack_backward()
        B, T = S["B"], S["T"]
        M = B * T
        f32, bf, eps = torch.float32, torch.bfloat16, cfg.layer_norm_eps
        Cl, d, ffn = cfg.filter_sizes[-1], cfg.hidden_size, cfg.intermediate_size
        H, dh, V = cfg.num_heads, cfg.head_size, cfg.vocab_size
        self.flat_g.zero_()
        ops.lm_head_wgrad(S["hidden_f32"], dlogits.view(M, V), G["lm_head/kernel"], G["lm_head/bias"])
        g = A.get("b.g", (M, d), f32)
        ops.lm_head_dgrad(dlogits, v["lm_head/kernel"], g)
        p_drop = S["dropout"]
        if p_drop:
            ops.dropout_rows(g, self._drop(self.SITE_HEAD), out_f32=g)
        dy, dyh = A.get("b.dy", (M, d), f32), A.get("b.dyh", (M, d), bf)
        g1 = A.get("b.g1", (M, d), f32)
        dmid, dpre = A.get("b.dmid", (M, ffn), bf), A.get("b.dpre", (M, ffn), bf)
        dctx, dqkv = A.get("b.dctx", (M, d), bf), A.get("b.dqkv", (M, 3 * d), bf)
        qkv_bias = A.get("b.qkvb", (3 * d,), f32)
        ws = A.get("b.attn.ws", (2 * B * H * T,), f32)
        enc = "wav2vec2/encoder/"
        for i in reversed(range(cfg.num_layers)):
            Li = S["layers"][i]
            lb = f"{enc}layers/{i}/"
            ff, at = lb + "feed_forward/", lb + "attention/"
            # x_{i+1} = LN2(y2),  y2 = x1 + mid W2 + b2
            on = Li["ffn_on"]
            ops.ln_bwd(Li["y2"], v[lb + "final_layer_norm/gamma"], g, eps, M, d, dx_f32=dy, dx_hi=dyh,
                       dgamma=G[lb + "final_layer_norm/gamma"], dbeta=G[lb + "final_layer_norm/beta"],
                       colsum=G[ff + "output_dense/bias"] if on else None)
            if on:
                ops.gemm(Pair(dyh), W[f"l{i}.ff2"], K=d, N=ffn, rows_per_batch=M, out_hi=dmid)
                self._wgrad(Li["mid"].hi, dyh, M, ffn, d, G[ff + "output_dense/kernel"])
                # mid = dropout(gelu(pre)),  pre = x1 W1 + b1
                ops.dact_colsum(dmid, Li["pre"], M, ffn, out_hi=dpre, colsum=G[ff + "intermediate_dense/bias"],
                                drop=self._drop(self.site_ffn_mid(i)) if p_drop else ops.NO_DROP)
                ops.gemm(Pair(dpre), W[f"l{i}.ff1"], K=ffn, N=d, rows_per_batch=M, residual=dy, out_f32=g1)
                self._wgrad(Li["x1"].hi, dpre, M, d, ffn, G[ff + "intermediate_dense/kernel"])
            else:
                g1.copy_(dy)                # the branch was dropped by StochasticDepth: only the shortcut carries gradient
            # x1 = LN1(y1),  y1 = x + ctx Wo + bo
            if p_drop:
                # the attention branch sees the gradient through its dropout mask, the residual branch sees all of it
                ops.ln_bwd(Li["y1"], v[lb + "layer_norm/gamma"], g1, eps, M, d, dx_f32=dy, dx_hi=dyh,
                           dgamma=G[lb + "layer_norm/gamma"], dbeta=G[lb + "layer_norm/beta"])
                ops.dact_colsum(dyh, None, M, d, out_hi=dyh, colsum=G[at + "out_proj/bias"], drop=self._drop(self.site_attn_out(i)))
            else:
                ops.ln_bwd(Li["y1"], v[lb + "layer_norm/gamma"], g1, eps, M, d, dx_f32=dy, dx_hi=dyh,
                           dgamma=G[lb + "layer_norm/gamma"], dbeta=G[lb + "layer_norm/beta"], colsum=G[at + "out_proj/bias"])
            ops.gemm(Pair(dyh), W[f"l{i}.out"], K=d, N=d, rows_per_batch=M, out_hi=dctx)
            self._wgrad(Li["ctx"].hi, dyh, M, d, d, G[at + "out_proj/kernel"])
            # ctx = softmax(q k^T) v  (q carries dh^-1/2: encoder.py:28 folded into the packed q projection)
            ops.attn_bwd(Li["qkv"].hi, Li["ctx"].hi, dctx, B, T, H, dh, None, dh ** -0.5, dqkv, workspace=ws,
                         drop=self._drop(self.site_attn_probs(i)) if p_drop else ops.NO_DROP)
            qkv_bias.zero_()
            ops.dact_colsum(dqkv, None, M, 3 * d, colsum=qkv_bias)
            for j, n in enumerate(("q", "k", "v")):
                G[at + f"{n}_proj/bias"].copy_(qkv_bias[j * d:(j + 1) * d])
            ops.gemm(Pair(dqkv), W[f"l{i}.qkv"], K=3 * d, N=d, rows_per_batch=M, residual=dy, out_f32=g)
            for j, n in enumerate(("q", "k", "v")):
                self._wgrad(Li["xs"].hi, dqkv[:, j * d:(j + 1) * d], M, d, d, G[at + f"{n}_proj/kernel"], dy_ld=3 * d)
        # x_0 = dropout(LN_enc(y0)),  y0 = h + gelu(pos_pre),  pos_pre = conv(h) + b
        if p_drop:
            ops.dropout_rows(g, self._drop(self.SITE_ENC), out_f32=g)
        ops.ln_bwd(S["y0"], v[enc + "layer_norm/gamma"], g, eps, M, d, dx_f32=dy, dx_hi=dyh,
                   dgamma=G[enc + "layer_norm/gamma"], dbeta=G[enc + "layer_norm/beta"])
        pc = enc + "pos_conv_embed/conv/"
        dpc = dctx
        ops.dact_colsum(dyh, S["pos_pre"], M, d, out_hi=dpc, colsum=G[pc + "bias"])
        groups, ktaps = cfg.num_conv_pos_embedding_groups, cfg.num_conv_pos_embeddings
        dh_f32 = g1
        ops.posconv_train(Pair(dpc), W["pos.T"], None, dy, dh_f32, B, T, d, groups, ktaps, 1, shift=1, linear=True)
        dwn = A.get("b.dwn", (ktaps, d // groups, d), f32)
        ops.posconv_wgrad(S["h"].hi, dpc, B, T, d, groups, ktaps, dwn)
        # weight-norm chain rule (tensorflow_addons.py:16-21: W = g * v / ||v||, norm over axes (1, 2) of every tap)
        dv, dg = weight_norm_backward(dwn, v[pc + "weight_v"], v[pc + "weight_g"])
        G[pc + "weight_v"].copy_(dv)
        G[pc + "weight_g"].copy_(dg)
        if S["spec_mask"] is not None:                      # masked frames were replaced by masked_spec_embed (modeling.py:193-199)
            msk = S["spec_mask"]
            G["wav2vec2/masked_spec_embed"].copy_((dh_f32 * msk[:, None]).sum(0))
            dh_f32.mul_((~msk)[:, None])
        # h = dropout(pn Wp + bp),  pn = LN_fp(extractor output)
        fp = "wav2vec2/feature_projection/"
        if p_drop:
            ops.dropout_rows(dh_f32, self._drop(self.SITE_PROJ), out_f32=dh_f32)
        dhh = ops.split_bf16(dh_f32, False).hi
        ops.dact_colsum(dhh, None, M, d, colsum=G[fp + "projection/bias"])
        self._wgrad(S["pn"].hi, dhh, M, Cl, d, G[fp + "projection/kernel"])
        dpn = A.get("b.dpn", (M, Cl), f32)
        ops.gemm(Pair(dhh), W["proj"], K=d, N=Cl, rows_per_batch=M, out_f32=dpn)
        ops.ln_bwd(S["last_f32"], v[fp + "layer_norm/gamma"], dpn, eps, M, Cl, dgamma=G[fp + "layer_norm/gamma"],
                   dbeta=G[fp + "layer_norm/beta"])

    # ------------------------------------------------------------------ one optimisation step
    @torch.no_grad()
    def loss_and_gradients(self, speech, labels, spec_mask=None, layer_keep=None):
        """Forward + CTC loss + backward on this rank's shard; gradients land in ``self.G`` (views of ``flat_g``).
        ``spec_mask`` [B, T'] overrides the sampled SpecAugment mask (tests)."""
        logits = self._forward(speech.to(self.model.device), spec_mask, layer_keep)
        loss, dlogits = self.loss_fn(labels, logits, return_grad=True)
        self._backward(dlogits)
        return loss

    @torch.no_grad()
    def step(self, speech, labels, spec_mask=None, next_speech=None):
        """One optimisation step.  ``next_speech``: the NEXT step's waveform batch (already on the device).  The conv extractor is
        frozen in stage 2 (main.py:236-237), so its forward does not depend on this step's update: it runs on the compute stream
        WHILE the gradient all-reduce is in flight on NCCL's stream and is handed to the next ``step`` call (which must receive
        the same tensor object) - the step's single collective is hidden behind ~40 % of the next forward."""
        loss = self.loss_and_gradients(speech, labels, spec_mask)
        work = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            work = dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, async_op=True)   # the step's single collective
        if next_speech is not None:
            nxt = next_speech.to(self.model.device)
            self._prefetched = (next_speech, self.model._features(nxt))
        if work is not None:
            work.wait()                                   # the compute stream waits for the reduced gradients
        self.t += 1
        lr_t = self.lr * (1.0 - self.b2 ** self.t) ** 0.5 / (1.0 - self.b1 ** self.t)
        ops.adam(self.flat_w, self.flat_g, self.m, self.v, lr_t, self.b1, self.b2, self.eps)
        self._repack()                     # kernel-layout copies follow the update
        return loss

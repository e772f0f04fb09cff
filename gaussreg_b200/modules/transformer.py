"""Geometric transformer (reference: geotransformer/modules/geotransformer/geotransformer.py,
geotransformer/modules/transformer/{rpe_transformer,vanilla_transformer,output_layer,positional_embedding,
conditional_transformer}.py).  Batch dimension is always 1 in the reference's forward (model.py:137-142);
the modules accept (1,N,C) or (N,C)."""
import numpy as np
import torch
import torch.nn as nn

from .. import ops


def _squeeze(x):
    return (x[0], True) if x.dim() == 3 else (x, False)


class SinusoidalPositionalEmbedding(nn.Module):
    """positional_embedding.py:8-35."""

    def __init__(self, d_model):
        super().__init__()
        if d_model % 2 != 0:
            raise ValueError(f"Sinusoidal positional encoding with odd d_model: {d_model}")
        self.d_model = d_model
        div_indices = torch.arange(0, d_model, 2).float()
        self.register_buffer("div_term", torch.exp(div_indices * (-np.log(10000.0) / d_model)))

    @torch.no_grad()
    def forward(self, emb_indices):
        out = ops.sinusoid_rows(emb_indices.contiguous().view(-1), self.div_term)
        return out.view(*emb_indices.shape, self.d_model)


class GeometricStructureEmbedding(nn.Module):
    """geotransformer.py:9-72."""

    # rows of the (N*N, C) embedding processed per chunk: bounds the sinusoid / projection temporaries
    CHUNK_ROWS = 1 << 18
    use_fused = True  # tensor-core kernel with in-kernel sinusoid generation (csrc/embedding_tc.cu)

    def __init__(self, hidden_dim, sigma_d, sigma_a, angle_k, reduction_a="max"):
        super().__init__()
        self.sigma_d, self.sigma_a, self.angle_k = sigma_d, sigma_a, angle_k
        self.factor_a = 180.0 / (self.sigma_a * np.pi)
        self.embedding = SinusoidalPositionalEmbedding(hidden_dim)
        self.proj_d = nn.Linear(hidden_dim, hidden_dim)
        self.proj_a = nn.Linear(hidden_dim, hidden_dim)
        self.reduction_a = reduction_a
        if reduction_a != "max":
            raise ValueError(f"Unsupported reduction mode: {reduction_a} (the GaussReg config uses 'max').")

    @torch.no_grad()
    def get_embedding_indices(self, points):
        pts, _ = _squeeze(points)
        d_idx, a_idx, _ = ops.embedding_indices(pts, self.sigma_d, self.sigma_a, self.angle_k)
        return d_idx, a_idx

    @torch.no_grad()
    def forward(self, points):
        pts, batched = _squeeze(points)
        N = pts.shape[0]
        C = self.embedding.d_model
        k = self.angle_k
        if self.use_fused and C % 64 == 0 and k <= 3 and ops._t1_mode() == "table":
            # indices + tabulated embedding in one C-ABI call (csrc/embedding_tab.cu)
            out = ops.structure_embedding_points(pts, self.embedding.div_term, self.proj_d.weight, self.proj_d.bias,
                                                 self.proj_a.weight, self.proj_a.bias, self.sigma_d, self.sigma_a, k)
            if out is not None:
                return out.unsqueeze(0) if batched else out
        d_idx, a_idx = self.get_embedding_indices(pts)
        if self.use_fused and C % 64 == 0 and k <= 3 and ops._t1_mode() == "table":
            # proj(sinusoid(x)) tabulated per channel: no projection in the hot path (csrc/embedding_tab.cu)
            out = ops.structure_embedding_tabulated(d_idx, a_idx, self.embedding.div_term, self.proj_d.weight, self.proj_d.bias,
                                                    self.proj_a.weight, self.proj_a.bias, self.sigma_a)
            if out is not None:
                return out.unsqueeze(0) if batched else out
        if self.use_fused and C == 256 and k <= 3 and ops._lib.lib().gr_get_gemm_mode() == 1:
            out = ops.structure_embedding_fused(d_idx, a_idx, self.embedding.div_term, self.proj_d.weight, self.proj_d.bias,
                                                self.proj_a.weight, self.proj_a.bias)
            return out.unsqueeze(0) if batched else out
        d_flat, a_flat = d_idx.view(-1), a_idx.view(-1)
        out = torch.empty((N * N, C), dtype=torch.float32, device=pts.device)
        rows = N * N
        chunk = min(rows, self.CHUNK_ROWS)
        E = torch.empty((chunk * k, C), dtype=torch.float32, device=pts.device)
        PD = torch.empty((chunk, C), dtype=torch.float32, device=pts.device)
        PA = torch.empty((chunk * k, C), dtype=torch.float32, device=pts.device)
        div = self.embedding.div_term
        for r0 in range(0, rows, chunk):
            r1 = min(rows, r0 + chunk)
            n = r1 - r0
            ops.sinusoid_rows(d_flat[r0:r1], div, out=E[:n])
            ops.linear(E[:n], self.proj_d.weight, self.proj_d.bias, out=PD[:n])
            ops.sinusoid_rows(a_flat[r0 * k:r1 * k], div, out=E[:n * k])
            ops.linear(E[:n * k], self.proj_a.weight, self.proj_a.bias, out=PA[:n * k])
            ops.embedding_combine(PD[:n], PA[:n * k], k, out[r0:r1])
        out = out.view(N, N, C)
        return out.unsqueeze(0) if batched else out


class AttentionOutput(nn.Module):
    """output_layer.py:6-21."""

    def __init__(self, d_model, dropout=None, activation_fn="ReLU"):
        super().__init__()
        if activation_fn != "ReLU" or dropout is not None:
            raise NotImplementedError
        self.expand = nn.Linear(d_model, d_model * 2)
        self.squeeze = nn.Linear(d_model * 2, d_model)
        self.norm = nn.LayerNorm(d_model)

    @torch.no_grad()
    def forward(self, x):
        h = ops.linear(x, self.expand.weight, self.expand.bias, act="relu")
        h = ops.linear(h, self.squeeze.weight, self.squeeze.bias)
        return ops.layer_norm_add(x, h, self.norm.weight, self.norm.bias, self.norm.eps)


class _AttentionCore(nn.Module):
    def __init__(self, d_model, num_heads, rpe):
        super().__init__()
        if d_model % num_heads != 0:
            raise ValueError("`d_model` ({}) must be a multiple of `num_heads` ({}).".format(d_model, num_heads))
        self.d_model, self.num_heads, self.d_model_per_head = d_model, num_heads, d_model // num_heads
        self.proj_q = nn.Linear(d_model, d_model)
        self.proj_k = nn.Linear(d_model, d_model)
        self.proj_v = nn.Linear(d_model, d_model)
        if rpe:
            self.proj_p = nn.Linear(d_model, d_model)

    def _pv(self, P, v, N):
        """hidden[:, h*dh:(h+1)*dh] = P[h] @ v[:, h*dh:(h+1)*dh]  ('b h n c -> b n (h c)')."""
        H, dh, C = self.num_heads, self.d_model_per_head, self.d_model
        M = v.shape[0]
        hidden = torch.empty((N, C), dtype=torch.float32, device=v.device)
        ops.gemm_batched(P.data_ptr(), M, N * M, v.data_ptr(), C, dh, False, hidden.data_ptr(), C, dh, N, dh, M, H)
        return hidden


class RPEMultiHeadAttention(_AttentionCore):
    """rpe_transformer.py:18-70; the (N,N,C) proj_p product is reassociated (see csrc/attention.cu)."""

    def __init__(self, d_model, num_heads, dropout=None):
        super().__init__(d_model, num_heads, rpe=True)

    @torch.no_grad()
    def forward(self, input_q, input_k, input_v, embed_qk):
        H, dh, C = self.num_heads, self.d_model_per_head, self.d_model
        N = input_q.shape[0]
        q = ops.linear(input_q, self.proj_q.weight, self.proj_q.bias)
        k = ops.linear(input_k, self.proj_k.weight, self.proj_k.bias)
        v = ops.linear(input_v, self.proj_v.weight, self.proj_v.bias)
        # U[h] = q_h (N,dh) @ W_p[h*dh:(h+1)*dh, :] (dh,C) ;  qb[h] = q_h @ b_p[h*dh:(h+1)*dh]
        U = torch.empty((H, N, C), dtype=torch.float32, device=q.device)
        ops.gemm_batched(q.data_ptr(), C, dh, self.proj_p.weight.data_ptr(), C, dh * C, False, U.data_ptr(), C, N * C, N, C, dh, H)
        qb = torch.empty((H, N), dtype=torch.float32, device=q.device)
        ops.gemm_batched(q.data_ptr(), C, dh, self.proj_p.bias.data_ptr(), dh, dh, True, qb.data_ptr(), 1, N, N, 1, dh, H)
        P = ops.rpe_attention_probs(q, k, U, qb, embed_qk, H)
        return self._pv(P, v, N), P


class MultiHeadAttention(_AttentionCore):
    """vanilla_transformer.py:15-73."""

    def __init__(self, d_model, num_heads, dropout=None):
        super().__init__(d_model, num_heads, rpe=False)

    @torch.no_grad()
    def forward(self, input_q, input_k, input_v):
        H, dh, C = self.num_heads, self.d_model_per_head, self.d_model
        N, M = input_q.shape[0], input_k.shape[0]
        q = ops.linear(input_q, self.proj_q.weight, self.proj_q.bias)
        k = ops.linear(input_k, self.proj_k.weight, self.proj_k.bias)
        v = ops.linear(input_v, self.proj_v.weight, self.proj_v.bias)
        P = torch.empty((H, N, M), dtype=torch.float32, device=q.device)
        ops.gemm_batched(q.data_ptr(), C, dh, k.data_ptr(), C, dh, True, P.data_ptr(), M, N * M, N, M, dh, H,
                         alpha=1.0 / dh ** 0.5)
        ops.softmax_rows_(P)
        return self._pv(P, v, N), P


class RPEAttentionLayer(nn.Module):
    """rpe_transformer.py:73-104."""

    def __init__(self, d_model, num_heads, dropout=None):
        super().__init__()
        self.attention = RPEMultiHeadAttention(d_model, num_heads, dropout=dropout)
        self.linear = nn.Linear(d_model, d_model)
        self.norm = nn.LayerNorm(d_model)

    @torch.no_grad()
    def forward(self, input_states, memory_states, position_states):
        hidden, scores = self.attention(input_states, memory_states, memory_states, position_states)
        hidden = ops.linear(hidden, self.linear.weight, self.linear.bias)
        return ops.layer_norm_add(hidden, input_states, self.norm.weight, self.norm.bias, self.norm.eps), scores


class AttentionLayer(nn.Module):
    """vanilla_transformer.py:76-107."""

    def __init__(self, d_model, num_heads, dropout=None):
        super().__init__()
        self.attention = MultiHeadAttention(d_model, num_heads, dropout=dropout)
        self.linear = nn.Linear(d_model, d_model)
        self.norm = nn.LayerNorm(d_model)

    @torch.no_grad()
    def forward(self, input_states, memory_states):
        hidden, scores = self.attention(input_states, memory_states, memory_states)
        hidden = ops.linear(hidden, self.linear.weight, self.linear.bias)
        return ops.layer_norm_add(hidden, input_states, self.norm.weight, self.norm.bias, self.norm.eps), scores


class RPETransformerLayer(nn.Module):
    """rpe_transformer.py:107-131."""

    def __init__(self, d_model, num_heads, dropout=None, activation_fn="ReLU"):
        super().__init__()
        self.attention = RPEAttentionLayer(d_model, num_heads, dropout=dropout)
        self.output = AttentionOutput(d_model, dropout=dropout, activation_fn=activation_fn)

    @torch.no_grad()
    def forward(self, input_states, memory_states, position_states, memory_masks=None):
        if memory_masks is not None:
            raise NotImplementedError("key masks are never passed by the GaussReg forward (model.py:137-142)")
        hidden, scores = self.attention(input_states, memory_states, position_states)
        return self.output(hidden), scores


class TransformerLayer(nn.Module):
    """vanilla_transformer.py:110-129."""

    def __init__(self, d_model, num_heads, dropout=None, activation_fn="ReLU"):
        super().__init__()
        self.attention = AttentionLayer(d_model, num_heads, dropout=dropout)
        self.output = AttentionOutput(d_model, dropout=dropout, activation_fn=activation_fn)

    @torch.no_grad()
    def forward(self, input_states, memory_states, memory_masks=None):
        if memory_masks is not None:
            raise NotImplementedError("key masks are never passed by the GaussReg forward (model.py:137-142)")
        hidden, scores = self.attention(input_states, memory_states)
        return self.output(hidden), scores


class RPEConditionalTransformer(nn.Module):
    """conditional_transformer.py:73-117 (sequential cross blocks, parallel=False)."""

    def __init__(self, blocks, d_model, num_heads, dropout=None, activation_fn="ReLU", return_attention_scores=False,
                 parallel=False):
        super().__init__()
        self.blocks = blocks
        layers = []
        for block in blocks:
            if block not in ("self", "cross"):
                raise ValueError('Unsupported block type "{}".'.format(block))
            cls = RPETransformerLayer if block == "self" else TransformerLayer
            layers.append(cls(d_model, num_heads, dropout=dropout, activation_fn=activation_fn))
        self.layers = nn.ModuleList(layers)
        self.return_attention_scores, self.parallel = return_attention_scores, parallel
        self.fused = True

    @torch.no_grad()
    def forward(self, feats0, feats1, embeddings0, embeddings1, masks0=None, masks1=None):
        f0, b0 = _squeeze(feats0)
        f1, _ = _squeeze(feats1)
        e0, _ = (embeddings0[0], True) if embeddings0.dim() == 4 else (embeddings0, False)
        e1, _ = (embeddings1[0], True) if embeddings1.dim() == 4 else (embeddings1, False)
        if self.fused and not self.return_attention_scores and not self.parallel and masks0 is None and masks1 is None:
            # one C-ABI call for all layers (csrc/transformer.cu): same kernels, issued from C++
            f0, f1 = ops.conditional_transformer(list(self.layers), self.blocks, f0, f1, e0, e1, self.layers[0].attention.attention.num_heads)
            if b0:
                f0, f1 = f0.unsqueeze(0), f1.unsqueeze(0)
            return f0, f1
        scores = []
        for i, block in enumerate(self.blocks):
            if block == "self":
                f0, s0 = self.layers[i](f0, f0, e0, memory_masks=masks0)
                f1, s1 = self.layers[i](f1, f1, e1, memory_masks=masks1)
            elif self.parallel:
                n0, s0 = self.layers[i](f0, f1, memory_masks=masks1)
                n1, s1 = self.layers[i](f1, f0, memory_masks=masks0)
                f0, f1 = n0, n1
            else:
                f0, s0 = self.layers[i](f0, f1, memory_masks=masks1)
                f1, s1 = self.layers[i](f1, f0, memory_masks=masks0)
            if self.return_attention_scores:
                scores.append([s0, s1])
        if b0:
            f0, f1 = f0.unsqueeze(0), f1.unsqueeze(0)
        return (f0, f1, scores) if self.return_attention_scores else (f0, f1)


def _squeeze_emb(e):
    return (e[0], True) if e.dim() == 4 else (e, False)


class GeometricTransformer(nn.Module):
    """geotransformer.py:75-155."""

    def __init__(self, input_dim, output_dim, hidden_dim, num_heads, blocks, sigma_d, sigma_a, angle_k, dropout=None,
                 activation_fn="ReLU", reduction_a="max"):
        super().__init__()
        self.embedding = GeometricStructureEmbedding(hidden_dim, sigma_d, sigma_a, angle_k, reduction_a=reduction_a)
        self.in_proj = nn.Linear(input_dim, hidden_dim)
        self.transformer = RPEConditionalTransformer(blocks, hidden_dim, num_heads, dropout=dropout, activation_fn=activation_fn)
        self.out_proj = nn.Linear(hidden_dim, output_dim)

    @torch.no_grad()
    def forward(self, ref_points, src_points, ref_feats, src_feats, ref_masks=None, src_masks=None, embeddings=None):
        """`embeddings` (extension): the two structure embeddings if the caller already evaluated them."""
        rp, batched = _squeeze(ref_points)
        sp, _ = _squeeze(src_points)
        rf, _ = _squeeze(ref_feats)
        sf, _ = _squeeze(src_feats)
        if embeddings is not None:
            ref_emb, _ = _squeeze_emb(embeddings[0])
            src_emb, _ = _squeeze_emb(embeddings[1])
        else:
            ref_emb = self.embedding(rp)
            src_emb = self.embedding(sp)
        n0 = rf.shape[0]
        both = ops.stacked_rows(rf, sf)  # the model passes row slices of one (N0+N1, C) matrix: project them in one product
        if both is not None:
            both = ops.linear(both, self.in_proj.weight, self.in_proj.bias)
            rf, sf = both[:n0], both[n0:]
        else:
            rf = ops.linear(rf, self.in_proj.weight, self.in_proj.bias)
            sf = ops.linear(sf, self.in_proj.weight, self.in_proj.bias)
        rf, sf = self.transformer(rf, sf, ref_emb, src_emb, masks0=ref_masks, masks1=src_masks)
        both = ops.stacked_rows(rf, sf)
        if both is not None:
            both = ops.linear(both, self.out_proj.weight, self.out_proj.bias)
            rf, sf = both[:n0], both[n0:]
        else:
            rf = ops.linear(rf, self.out_proj.weight, self.out_proj.bias)
            sf = ops.linear(sf, self.out_proj.weight, self.out_proj.bias)
        if batched:
            rf, sf = rf.unsqueeze(0), sf.unsqueeze(0)
        return rf, sf

"""The score U-Net of ``fbs/nn/unet.py`` (``UNet(dt, dim=64, upsampling='pixel_shuffle')``, the configuration of every
image experiment: experiments/imgs/inpainting.py:85, experiments/sb_imgs/supr.py:67) as a fixed schedule of sm_100a
kernels, and the NN-score closures built on it (experiments/imgs/inpainting.py:94-147).

* every 3x3 / 1x1 / stride-2 convolution with >= 64 input channels is an implicit GEMM on the tcgen05 tensor cores
  (bf16 operands, fp32 accumulation in tensor memory): ``fbs_nn_conv_bf16``; skip concatenations are never
  materialised (two-source K loop), pixel-shuffle is an epilogue addressing mode;
* weight standardisation (unet.py:110-118, recomputed on every call upstream) is a one-time transform here: the
  weights are constants at sampling time;
* GroupNorm + time scale/shift + swish (+ residual), LayerNorm (+ residual), both attention cores and the time MLP are
  one fused kernel each (nn_ops.cu);
* the whole evaluation is captured in a CUDA graph per batch size and replayed (the network time lives in a device
  scalar), so one score evaluation costs one graph launch from the host.

Parameters: a flat dict ``name -> float32 array`` with flax's HWIO kernels, the naming of oracle/unet.py's
``unet_param_shapes`` (this module does not import the oracle; the names are the interface).
"""
import math
import numpy as np
import torch
from . import ops
from .._tensor import cuda_device, dev
from .. import random as frandom

BF16, F32 = torch.bfloat16, torch.float32


def _standardize(w, eps=1e-5):
    """unet.py:110-118 in float32: per output filter over (kh, kw, Cin)."""
    w = np.asarray(w, dtype=np.float32)
    mean = w.mean(axis=(0, 1, 2), keepdims=True, dtype=np.float32)
    var = w.var(axis=(0, 1, 2), keepdims=True, dtype=np.float32)
    return ((w - mean) / np.sqrt(var + np.float32(eps))).astype(np.float32)


def _pack(w):
    """flax HWIO kernel [kh, kw, Cin, Cout] -> bf16 [Cout, kh * kw * Cin], K ordered (ty, tx, channel)."""
    kh, kw, cin, cout = w.shape
    return np.ascontiguousarray(w.reshape(kh * kw * cin, cout).T)


def _pack_stride2(w):
    """4x4 stride-2 kernel -> the 2x2 kernel over the space-to-depth copy: k2[ty, tx, (r, s, c), o] = w[2 ty + r, 2 tx + s, c, o]."""
    _, _, cin, cout = w.shape
    k2 = w.reshape(2, 2, 2, 2, cin, cout).transpose(0, 2, 1, 3, 4, 5)  # [ty, r, tx, s, c, o] -> [ty, tx, r, s, c, o]
    return _pack(np.ascontiguousarray(k2).reshape(2, 2, 4 * cin, cout))


def unet_param_shapes(in_ch, dim=64, dim_mults=(1, 2, 4), heads=4, dim_head=32):
    """name -> shape of every parameter of ``UNet(dim, dim_mults, upsampling='pixel_shuffle')`` (fbs/nn/unet.py:279-368),
    flax layouts (HWIO convolution kernels, [in, out] dense kernels)."""
    sh = {}

    def conv(name, k, cin, cout, bias=True):
        sh[name + '.kernel'] = (k, k, cin, cout)
        if bias:
            sh[name + '.bias'] = (cout,)

    def res(name, cin, d):
        conv(name + '.conv_0', 3, cin, d)
        conv(name + '.conv_1', 3, d, d)
        for nm in ('norm_0', 'norm_1'):
            sh[f'{name}.{nm}.scale'] = (d,)
            sh[f'{name}.{nm}.bias'] = (d,)
        sh[name + '.time_mlp.dense_0.kernel'] = (4 * dim, 2 * d)
        sh[name + '.time_mlp.dense_0.bias'] = (2 * d,)
        if cin != d:
            conv(name + '.res_conv_0', 1, cin, d)

    def attn(name, c, linear=True):
        sh[name + '.norm.scale'] = (c,)
        conv(name + '.attn.to_qkv.conv_0', 1, c, 3 * heads * dim_head, bias=False)
        conv(name + '.attn.to_out.conv_0', 1, heads * dim_head, c)
        if linear:
            sh[name + '.attn.to_out.norm_0.scale'] = (c,)

    nres = len(dim_mults)
    conv('init.conv_0', 7, in_ch, dim)
    for i, (a, b) in enumerate(((dim, 4 * dim), (4 * dim, 4 * dim))):
        sh[f'time.dense_{i}.kernel'] = (a, b)
        sh[f'time.dense_{i}.bias'] = (b,)
    c = dim
    for ind in range(nres):
        res(f'down_{ind}.resblock_0', c, c)
        res(f'down_{ind}.resblock_1', c, c)
        attn(f'down_{ind}.attnblock_0', c)
        if ind < nres - 1:
            conv(f'down_{ind}.downsample_0', 4, c, dim * dim_mults[ind])
            c = dim * dim_mults[ind]
    mid = dim * dim_mults[-1]
    conv(f'down_{nres - 1}.conv_0', 3, c, mid)
    res('mid.resblock_0', mid, mid)
    attn('mid.attenblock_0', mid, linear=False)
    res('mid.resblock_1', mid, mid)
    for ind in reversed(range(nres)):
        d_in = dim * dim_mults[ind]
        d_out = dim * dim_mults[ind - 1] if ind > 0 else dim
        res(f'up_{ind}.resblock_0', d_in + d_out, d_in)
        res(f'up_{ind}.resblock_1', d_in + d_out, d_in)
        attn(f'up_{ind}.attnblock_0', d_in)
        if ind > 0:
            conv(f'up_{ind}.upsample_0.conv_0', 3, d_in, 4 * d_in)
            conv(f'up_{ind}.upsample_0.conv_1', 3, d_in, d_out)
    conv('up_0.conv_0', 3, dim, dim)
    res('final.resblock_0', 2 * dim, dim)
    conv('final.conv_0', 1, dim, in_ch)
    return sh


def random_unet_params(seed, in_ch, dim=64, dim_mults=(1, 2, 4)):
    """Random-init weights of the architecture (there is no network access for checkpoints): kernels N(0, 1 / fan_in)
    as flax's default lecun_normal, biases 0, norm scales 1 -- what ``nn.init`` gives (fbs/nn/base.py:30)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in unet_param_shapes(in_ch, dim, dim_mults).items():
        if name.endswith('.kernel'):
            out[name] = (rng.standard_normal(shape) / math.sqrt(int(np.prod(shape[:-1])))).astype(np.float32)
        elif name.endswith('.scale'):
            out[name] = np.ones(shape, np.float32)
        else:
            out[name] = np.zeros(shape, np.float32)
    return out


class ScoreUNet:
    def __init__(self, params, image_shape, dt, dim=64, dim_mults=(1, 2, 4), groups=8, heads=4, dim_head=32, device=None):
        self.H, self.W, self.Cimg = (int(s) for s in image_shape)
        self.dt, self.dim, self.dim_mults, self.groups = float(dt), int(dim), tuple(dim_mults), int(groups)
        self.heads, self.dim_head = heads, dim_head
        self.device = device or cuda_device()
        if dim % 64:
            raise NotImplementedError('ScoreUNet: dim must be a multiple of 64 (tensor-core K blocks of 64 channels)')
        nres = len(self.dim_mults)
        if self.H % (1 << (nres - 1)) or self.W % (1 << (nres - 1)) or self.W > 128:
            raise NotImplementedError('ScoreUNet: H, W must be divisible by 2^(levels-1) and W <= 128')
        self._w = {}
        self._bufs = {}
        self._graphs = {}
        self._gn_slots = {}
        self.gn_input_bf16 = False    # GroupNorm inputs (convolution outputs) kept in fp32
        self._time_blocks = []       # resnet block name -> offset into the time table
        self._prepare(params)
        self.tval = torch.zeros((1,), dtype=F32, device=self.device)
        self._side = torch.cuda.Stream(device=self.device)

    # ------------------------------------------------------------------ weights
    def _put(self, name, arr, dtype=F32):
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).to(self.device)
        self._w[name] = t.to(dtype).contiguous()

    def _prepare(self, P):
        off = 0
        wcat, bcat = [], []
        for name in sorted(P):
            if name.endswith('.kernel'):
                base = name[:-len('.kernel')]
                w = np.asarray(P[name], dtype=np.float32)
                if base == 'init.conv_0' or base == 'final.conv_0':
                    self._put(base + '.w', w if base == 'init.conv_0' else w.reshape(w.shape[2], w.shape[3]))
                elif base.startswith('time.'):
                    self._put(base + '.w', w)
                elif base.endswith('time_mlp.dense_0'):
                    pass  # concatenated below, in execution order
                elif base.endswith('.conv_0') and '.resblock' in base or base.endswith('.conv_1') and '.resblock' in base:
                    self._put(base + '.w', _pack(_standardize(w)), BF16)
                elif 'downsample_0' in base:
                    self._put(base + '.w', _pack_stride2(w), BF16)
                else:
                    self._put(base + '.w', _pack(w), BF16)
            elif name.endswith('.bias') or name.endswith('.scale'):
                if not name.endswith('time_mlp.dense_0.bias'):
                    self._put(name, P[name])
        for blk in self._resblock_order():
            k = np.asarray(P[blk + '.time_mlp.dense_0.kernel'], dtype=np.float32)
            wcat.append(k)
            bcat.append(np.asarray(P[blk + '.time_mlp.dense_0.bias'], dtype=np.float32))
            self._time_blocks.append((blk, off, k.shape[1] // 2))
            off += k.shape[1]
        self._put('time.Wcat', np.concatenate(wcat, axis=1))
        self._put('time.bcat', np.concatenate(bcat))
        self._toff = {blk: (o, d) for blk, o, d in self._time_blocks}
        self.table = torch.empty((off,), dtype=F32, device=self.device)

    def _resblock_order(self):
        nres = len(self.dim_mults)
        names = []
        for ind in range(nres):
            names += [f'down_{ind}.resblock_0', f'down_{ind}.resblock_1']
        names += ['mid.resblock_0', 'mid.resblock_1']
        for ind in reversed(range(nres)):
            names += [f'up_{ind}.resblock_0', f'up_{ind}.resblock_1']
        names.append('final.resblock_0')
        return names

    # ------------------------------------------------------------------ buffers
    def _buf(self, B, name, shape, dtype):
        key = (B, name)
        t = self._bufs.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(tuple(shape), dtype=dtype, device=self.device)
            self._bufs[key] = t
        return t

    # ------------------------------------------------------------------ blocks
    def _resblock(self, B, name, srcs, x_f32, H, W, d, tag, ln=None):
        """ResnetBlock (unet.py:127-172).  srcs: bf16 tensors whose channel concatenation is the block input.
        ``ln = (gamma, out_bf16)``: the attention block that follows wants ``LayerNorm(output) * gamma`` (unet.py:258); where the
        one-pass GroupNorm runs (and C <= 128) it writes that too and the method returns ``(out_f32, out_bf16, True)``."""
        w = self._w
        in0, in1 = srcs[0], (srcs[1] if len(srcs) > 1 else None)
        cin = sum(s.shape[-1] for s in srcs)
        t_bf = self._buf(B, f'tmp_bf16_{H}x{d}', (B, H, W, d), BF16)
        o, dd = self._toff[name]
        # conv -> GroupNorm pairs: the convolution's epilogue hands the GroupNorm its statistics (sum / sum of squares per row
        # tile and channel quad), so the normalisation is one streaming pass; where the tiling cannot (several samples per tile:
        # the 7x7 level) the stand-alone GroupNorm kernel computes them itself
        t_in = self._buf(B, f'tmp_gnin_{H}x{d}', (B, H, W, d), BF16 if self.gn_input_bf16 else F32)
        tkw = {'out_bf16': t_in} if self.gn_input_bf16 else {'out_f32': t_in}

        def conv_gn(src0, src1, wname, bname, gamma, beta, tss, residual, o32, o16, ln=None):
            args = (src0, w[wname], d, 3, 3, -1, H, W)
            key = (B, wname)
            slots = self._gn_slots.get(key)
            if slots is None:
                slots = self._gn_slots[key] = ops.conv_gn_slots(*args, in1=src1, bias=w[bname], **tkw)
            if slots and 256 % (d // 4) == 0:
                part = self._buf(B, f'gn_part_{H}x{d}', (B, slots, d // 4, 2), F32)
                ops.conv(*args, in1=src1, bias=w[bname], gn_partials=part, **tkw)
                fuse_ln = ln is not None and d <= 128
                ops.groupnorm_swish_stats(t_in, part, gamma, beta, self.groups, tss=tss, residual=residual, out_f32=o32, out_bf16=o16,
                                          ln_gamma=ln[0] if fuse_ln else None, ln_out_bf16=ln[1] if fuse_ln else None)
                return fuse_ln
            else:
                t_f32 = self._buf(B, f'tmp_f32_{H}x{d}', (B, H, W, d), F32)
                ops.conv(*args, in1=src1, bias=w[bname], out_f32=t_f32)
                ops.groupnorm_swish(t_f32, gamma, beta, self.groups, tss=tss, residual=residual, out_f32=o32, out_bf16=o16)
                return False

        if cin != d:
            res = self._buf(B, f'res_f32_{H}x{d}', (B, H, W, d), F32)
            ops.conv(in0, w[name + '.res_conv_0.w'], d, 1, 1, 0, H, W, in1=in1, bias=w[name + '.res_conv_0.bias'], out_f32=res)
        else:
            res = x_f32
        conv_gn(in0, in1, name + '.conv_0.w', name + '.conv_0.bias', w[name + '.norm_0.scale'], w[name + '.norm_0.bias'],
                self.table[o:o + 2 * dd], None, None, t_bf)
        out_f32 = self._buf(B, tag + '_f32', (B, H, W, d), F32)
        out_bf = self._buf(B, tag + '_bf16', (B, H, W, d), BF16)
        ln_done = conv_gn(t_bf, None, name + '.conv_1.w', name + '.conv_1.bias', w[name + '.norm_1.scale'], w[name + '.norm_1.bias'],
                          None, res, out_f32, out_bf, ln=ln)
        return (out_f32, out_bf, ln_done) if ln is not None else (out_f32, out_bf)

    def _attn_ln(self, B, name, H, W, C):
        """(scale, output buffer) of the LayerNorm in front of attention block ``name`` -- the buffer `_attnblock` reads."""
        return self._w[name + '.norm.scale'], self._buf(B, f'attn_norm_{H}x{C}', (B, H, W, C), BF16)

    def _attnblock(self, B, name, x_f32, H, W, C, tag, linear=True, normed=False):
        """AttnBlock (unet.py:248-264) around LinearAttention (:209-245) or Attention (:175-206)."""
        w = self._w
        hd = self.heads * self.dim_head
        n_bf = self._buf(B, f'attn_norm_{H}x{C}', (B, H, W, C), BF16)
        qkv = self._buf(B, f'attn_qkv_{H}', (B, H, W, 3 * hd), BF16)
        ao = self._buf(B, f'attn_out_{H}', (B, H, W, hd), BF16)
        out_f32 = self._buf(B, tag + '_f32', (B, H, W, C), F32)
        out_bf = self._buf(B, tag + '_bf16', (B, H, W, C), BF16)
        if not normed:   # (else the preceding ResnetBlock's GroupNorm kernel has written attn_norm already)
            ops.layernorm(x_f32, w[name + '.norm.scale'], out_bf16=n_bf)
        ops.conv(n_bf, w[name + '.attn.to_qkv.conv_0.w'], 3 * hd, 1, 1, 0, H, W, out_bf16=qkv)
        if linear:
            ops.linear_attention(qkv, ao, self.heads, self.dim_head)
            proj = self._buf(B, f'attn_proj_{H}x{C}', (B, H, W, C), F32)
            ops.conv(ao, w[name + '.attn.to_out.conv_0.w'], C, 1, 1, 0, H, W, bias=w[name + '.attn.to_out.conv_0.bias'], out_f32=proj)
            ops.layernorm(proj, w[name + '.attn.to_out.norm_0.scale'], residual=x_f32, out_f32=out_f32, out_bf16=out_bf)
        else:
            ops.attention(qkv, ao, self.heads, self.dim_head, 10.0)
            ops.conv(ao, w[name + '.attn.to_out.conv_0.w'], C, 1, 1, 0, H, W, bias=w[name + '.attn.to_out.conv_0.bias'],
                     residual=x_f32, out_f32=out_f32, out_bf16=out_bf)
        return out_f32, out_bf

    # ------------------------------------------------------------------ forward
    def _forward(self, x, B):
        """x: fp32 [B, H, W, Cimg] (device); self.tval holds the network time.  Returns fp32 [B, H, W, Cimg]."""
        w, dim = self._w, self.dim
        H, W = self.H, self.W
        nres = len(self.dim_mults)
        # the time-embedding table does not depend on x: it runs beside the first convolution (a fork / join of the stream,
        # which a CUDA-graph capture records as a parallel branch)
        cur = torch.cuda.current_stream(self.device)
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            ops.time_mlp(self.tval, self.dt, dim, w['time.dense_0.w'], w['time.dense_0.bias'], w['time.dense_1.w'],
                         w['time.dense_1.bias'], w['time.Wcat'], w['time.bcat'], self.table)
        h_f32 = self._buf(B, 'h0_f32', (B, H, W, dim), F32)
        h_bf = self._buf(B, 'h0_bf16', (B, H, W, dim), BF16)
        ops.stem_conv(x, w['init.conv_0.w'], w['init.conv_0.bias'], out_f32=h_f32, out_bf16=h_bf)
        cur.wait_stream(self._side)
        skips = [h_bf]
        c = dim
        for ind in range(nres):
            h_f32, h_bf = self._resblock(B, f'down_{ind}.resblock_0', [h_bf], h_f32, H, W, c, f'd{ind}r0')
            skips.append(h_bf)
            h_f32, h_bf, nd = self._resblock(B, f'down_{ind}.resblock_1', [h_bf], h_f32, H, W, c, f'd{ind}r1',
                                             ln=self._attn_ln(B, f'down_{ind}.attnblock_0', H, W, c))
            h_f32, h_bf = self._attnblock(B, f'down_{ind}.attnblock_0', h_f32, H, W, c, f'd{ind}a', normed=nd)
            skips.append(h_bf)
            if ind < nres - 1:
                cout = dim * self.dim_mults[ind]
                s2d = self._buf(B, f's2d_{ind}', (B, H // 2 + 1, W // 2 + 1, 4 * c), BF16)
                ops.space_to_depth(h_bf, s2d)
                H, W = H // 2, W // 2
                h_f32 = self._buf(B, f'd{ind}ds_f32', (B, H, W, cout), F32)
                h_bf = self._buf(B, f'd{ind}ds_bf16', (B, H, W, cout), BF16)
                ops.conv(s2d, w[f'down_{ind}.downsample_0.w'], cout, 2, 2, 0, H, W, bias=w[f'down_{ind}.downsample_0.bias'],
                         out_f32=h_f32, out_bf16=h_bf)
                c = cout
        mid = dim * self.dim_mults[-1]
        m_f32 = self._buf(B, 'midin_f32', (B, H, W, mid), F32)
        m_bf = self._buf(B, 'midin_bf16', (B, H, W, mid), BF16)
        ops.conv(h_bf, w[f'down_{nres - 1}.conv_0.w'], mid, 3, 3, -1, H, W, bias=w[f'down_{nres - 1}.conv_0.bias'], out_f32=m_f32,
                 out_bf16=m_bf)
        h_f32, h_bf = self._resblock(B, 'mid.resblock_0', [m_bf], m_f32, H, W, mid, 'm0')
        h_f32, h_bf = self._attnblock(B, 'mid.attenblock_0', h_f32, H, W, mid, 'ma', linear=False)
        h_f32, h_bf = self._resblock(B, 'mid.resblock_1', [h_bf], h_f32, H, W, mid, 'm1')
        for ind in reversed(range(nres)):
            dim_in = dim * self.dim_mults[ind]
            dim_out = dim * self.dim_mults[ind - 1] if ind > 0 else dim
            h_f32, h_bf = self._resblock(B, f'up_{ind}.resblock_0', [h_bf, skips.pop()], None, H, W, dim_in, f'u{ind}r0')
            h_f32, h_bf, nd = self._resblock(B, f'up_{ind}.resblock_1', [h_bf, skips.pop()], None, H, W, dim_in, f'u{ind}r1',
                                             ln=self._attn_ln(B, f'up_{ind}.attnblock_0', H, W, dim_in))
            h_f32, h_bf = self._attnblock(B, f'up_{ind}.attnblock_0', h_f32, H, W, dim_in, f'u{ind}a', normed=nd)
            if ind > 0:
                ps = self._buf(B, f'u{ind}ps_bf16', (B, 2 * H, 2 * W, dim_in), BF16)
                ops.conv(h_bf, w[f'up_{ind}.upsample_0.conv_0.w'], 4 * dim_in, 3, 3, -1, H, W,
                         bias=w[f'up_{ind}.upsample_0.conv_0.bias'], out_bf16=ps, pixel_shuffle=True)
                H, W = 2 * H, 2 * W
                h_f32 = self._buf(B, f'u{ind}us_f32', (B, H, W, dim_out), F32)
                h_bf = self._buf(B, f'u{ind}us_bf16', (B, H, W, dim_out), BF16)
                ops.conv(ps, w[f'up_{ind}.upsample_0.conv_1.w'], dim_out, 3, 3, -1, H, W,
                         bias=w[f'up_{ind}.upsample_0.conv_1.bias'], out_f32=h_f32, out_bf16=h_bf)
        last_bf = self._buf(B, 'u0c_bf16', (B, H, W, dim), BF16)
        ops.conv(h_bf, w['up_0.conv_0.w'], dim, 3, 3, -1, H, W, bias=w['up_0.conv_0.bias'], out_bf16=last_bf)
        f_f32, _ = self._resblock(B, 'final.resblock_0', [last_bf, skips.pop()], None, H, W, dim, 'fin')
        out = self._buf(B, 'score', (B, H, W, self.Cimg), F32)
        ops.head_conv(f_f32, w['final.conv_0.w'], w['final.conv_0.bias'], out)
        return out

    def __call__(self, x, time, use_graph=True):
        """Score network at ``x`` fp32 [B, H, W, Cimg] (CUDA) and scalar ``time`` -> fp32 [B, H, W, Cimg] (a buffer owned
        by the network, overwritten by the next call)."""
        B = x.shape[0]
        if isinstance(time, torch.Tensor):
            self.tval.copy_(time.reshape(1).to(F32))
        else:
            self.tval.fill_(float(time))
        xin = self._buf(B, 'x_in', (B, self.H, self.W, self.Cimg), F32)
        xin.copy_(x.reshape(xin.shape))
        if not use_graph:
            return self._forward(xin, B)
        g = self._graphs.get(B)
        if g is None:
            self._forward(xin, B)      # allocates every buffer outside the capture
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._forward(xin, B)
            g = self._graphs[B] = (graph, out)
        g[0].replay()
        return g[1]


class ScoreNetModel:
    """The closures of experiments/imgs/inpainting.py:94-147 (identical in supr.py) over a :class:`ScoreUNet`.

    ``unobs_idx`` / ``obs_idx`` are the mask's ravelled pixel index lists (fbs/data/images.py:258-303); the SDE is a
    scalar linear SDE (``drift(x, t) = a(t) x``).  One network evaluation serves both the transition and the weight
    (:meth:`step`); the bound methods keep the reference's closure signatures.

    ``drift_mode=True`` gives the closures of the Schroedinger-bridge image runs instead (experiments/sb_imgs/supr.py:84-129):
    the network output IS the reverse drift, ``reverse_drift(uv, t) = nn_drift(uv, T - t, param_bwd)`` (:84-85) -- no
    ``-a x + g^2 score`` wrapping -- while the dispersion stays the SDE's (:100-101).  With ``fwd_unet`` (a second network,
    ``param_fwd``) ``fwd_sampler`` is the Euler--Maruyama simulation of the learnt FORWARD drift
    (``euler_maruyama(key, xy0, ts, nn_drift(., t, param_fwd), sde.dispersion, integration_nsteps=1)``, :132-137) instead of
    the closed-form noising of a linear SDE.
    """

    def __init__(self, unet: ScoreUNet, sde, ts, T, unobs_idx, obs_idx, drift_mode: bool = False, fwd_unet: ScoreUNet = None):
        self.unet, self.sde, self.T = unet, sde, float(T)
        self.drift_mode, self.fwd_unet = bool(drift_mode), fwd_unet
        if fwd_unet is not None and (fwd_unet.H, fwd_unet.W, fwd_unet.Cimg) != (unet.H, unet.W, unet.Cimg):
            raise ValueError('the forward-drift network must have the image shape of the backward one')
        self.ts = np.asarray(ts, dtype=np.float64)
        self.K = self.ts.shape[0] - 1
        self.dt = self.T / self.K                                   # inpainting.py:58
        d = unet.device
        self.unobs = torch.as_tensor(np.asarray(unobs_idx), dtype=torch.int32, device=d).contiguous()
        self.obs = torch.as_tensor(np.asarray(obs_idx), dtype=torch.int32, device=d).contiguous()
        self.p, self.q, self.c = int(self.unobs.numel()), int(self.obs.numel()), unet.Cimg
        if self.p + self.q != unet.H * unet.W:
            raise ValueError('mask index lists must partition the image')

    def _coef(self, t_prev):
        s = self.T - float(t_prev)
        g = float(self.sde.dispersion(s))
        sd = float(np.float32(math.sqrt(self.dt)) * np.float32(g))
        if self.drift_mode:                                         # sb_imgs/supr.py:84-85: rd = network output
            return s, 0.0, 1.0, sd
        return s, float(self.sde.drift_coef(s)), g * g, sd

    def _score(self, us_prev, v_prev, t_prev):
        B = us_prev.shape[0]
        img = self.unet._buf(B, 'closure_img', (B, self.unet.H, self.unet.W, self.c), F32)
        ops.assemble_image(us_prev, v_prev, self.unobs, self.obs, img)
        s, a, g2, sd = self._coef(t_prev)
        return img, self.unet(img, s), a, g2, sd

    def step(self, us_prev, v_prev, v_next, t_prev, key, row_offset=0, rows_total=None, pin_row=None, pin_value=None, out=None):
        """(us_new [N, p, c], log_w [N]) from ONE score evaluation (transition_sampler + likelihood_logpdf).

        ``row_offset`` / ``rows_total``: the N particles are rows [row_offset, row_offset + N) of a larger (sharded)
        particle set; the transition noise is the matching slice of ``normal(key, (rows_total, p, c))``.
        ``pin_row`` (device int32 [1], global row) / ``pin_value`` ``[p, c]``: the reference particle written by the same
        kernel (csmc.py:143).  ``out``: where the new particles go (e.g. straight into a peer-visible buffer)."""
        us_prev = dev(us_prev, F32).reshape(-1, self.p, self.c)
        v_prev = dev(v_prev, F32).reshape(self.q, self.c)
        v_next = dev(v_next, F32).reshape(self.q, self.c)
        key = dev(key, torch.uint32).reshape(2)
        B = us_prev.shape[0]
        img, score, a, g2, sd = self._score(us_prev, v_prev, t_prev)
        us_new = torch.empty_like(us_prev) if out is None else out.view(us_prev.shape)
        lw = torch.empty((B,), dtype=F32, device=us_prev.device)
        pv = None if pin_value is None else dev(pin_value, F32).reshape(self.p * self.c)
        ops.em_step(img, score, self.unobs, self.obs, B, self.p, self.q, self.c, a, g2, self.dt, sd, v_next=v_next, key=key,
                    us_new=us_new, lw=lw, row_offset=row_offset, rows_total=rows_total, pin_row=pin_row, pin_value=pv)
        return us_new, lw

    def step_chains(self, us_prev, v_prev, v_next, t_prev, keys, pin_rows=None, pin_values=None):
        """:meth:`step` for C independent chains (conditioning targets) at once: ``us_prev [C, N, p, c]``, ``v_prev`` /
        ``v_next [C, q, c]``, ``keys [C, 2]`` -> ``(us_new [C, N, p, c], log_w [C, N])``.  The C x N images go through ONE
        score evaluation (the network is launch / latency bound at N = 101: batching targets is what fills the GPU); image
        assembly, noise and weights stay per chain, so every chain gets exactly the numbers of its own :meth:`step`."""
        us_prev = dev(us_prev, F32)
        C_, N = us_prev.shape[0], us_prev.shape[1]
        us_prev = us_prev.reshape(C_, N, self.p, self.c)
        v_prev = dev(v_prev, F32).reshape(C_, self.q, self.c)
        v_next = dev(v_next, F32).reshape(C_, self.q, self.c)
        keys = dev(keys, torch.uint32).reshape(C_, 2).contiguous()
        B = C_ * N
        img = self.unet._buf(B, 'closure_img', (B, self.unet.H, self.unet.W, self.c), F32)
        for ci in range(C_):
            ops.assemble_image(us_prev[ci], v_prev[ci], self.unobs, self.obs, img[ci * N:(ci + 1) * N])
        s, a, g2, sd = self._coef(t_prev)
        score = self.unet(img, s)
        us_new = torch.empty_like(us_prev)
        lw = torch.empty((C_, N), dtype=F32, device=us_prev.device)
        for ci in range(C_):
            ops.em_step(img[ci * N:(ci + 1) * N], score[ci * N:(ci + 1) * N], self.unobs, self.obs, N, self.p, self.q, self.c,
                        a, g2, self.dt, sd, v_next=v_next[ci], key=keys[ci], us_new=us_new[ci], lw=lw[ci],
                        pin_row=None if pin_rows is None else pin_rows[ci:ci + 1],
                        pin_value=None if pin_values is None else pin_values[ci])
        return us_new, lw

    def mean_and_logw(self, us_prev, v_prev, v_next, t_prev):
        """(transition mean [N, p, c], log-weight [N], sd) from ONE score evaluation: what pmcmc_filter_step needs, since the
        transition of the RESAMPLED particles is the gathered mean plus fresh noise (smc.py:144-150)."""
        us_prev = dev(us_prev, F32).reshape(-1, self.p, self.c)
        v_prev = dev(v_prev, F32).reshape(self.q, self.c)
        v_next = dev(v_next, F32).reshape(self.q, self.c)
        B = us_prev.shape[0]
        img, score, a, g2, sd = self._score(us_prev, v_prev, t_prev)
        mean = torch.empty_like(us_prev)
        lw = torch.empty((B,), dtype=F32, device=us_prev.device)
        ops.em_step(img, score, self.unobs, self.obs, B, self.p, self.q, self.c, a, g2, self.dt, sd, v_next=v_next, mean_out=mean,
                    lw=lw)
        return mean, lw, sd

    def transition_sampler(self, us_prev, v_prev, t_prev, key, **kwargs):
        """inpainting.py:122-128."""
        us_prev = dev(us_prev, F32).reshape(-1, self.p, self.c)
        v_prev = dev(v_prev, F32).reshape(self.q, self.c)
        key = dev(key, torch.uint32).reshape(2)
        B = us_prev.shape[0]
        img, score, a, g2, sd = self._score(us_prev, v_prev, t_prev)
        us_new = torch.empty_like(us_prev)
        ops.em_step(img, score, self.unobs, self.obs, B, self.p, self.q, self.c, a, g2, self.dt, sd, key=key, us_new=us_new)
        return us_new

    def likelihood_logpdf(self, v, us_prev, v_prev, t_prev, **kwargs):
        """inpainting.py:141-147."""
        us_prev = dev(us_prev, F32).reshape(-1, self.p, self.c)
        v_prev = dev(v_prev, F32).reshape(self.q, self.c)
        v = dev(v, F32).reshape(self.q, self.c)
        B = us_prev.shape[0]
        img, score, a, g2, sd = self._score(us_prev, v_prev, t_prev)
        lw = torch.empty((B,), dtype=F32, device=us_prev.device)
        ops.em_step(img, score, self.unobs, self.obs, B, self.p, self.q, self.c, a, g2, self.dt, sd, v_next=v, lw=lw)
        return lw

    def transition_logpdf(self, u, us_prev, v_prev, t_prev, **kwargs):
        """inpainting.py:131-138: sum logN(u; mean(us_prev), sd) per particle."""
        us_prev = dev(us_prev, F32).reshape(-1, self.p, self.c)
        v_prev = dev(v_prev, F32).reshape(self.q, self.c)
        u = dev(u, F32).reshape(1, self.p * self.c)
        B = us_prev.shape[0]
        img, score, a, g2, sd = self._score(us_prev, v_prev, t_prev)
        mean = torch.empty_like(us_prev)
        ops.em_step(img, score, self.unobs, self.obs, B, self.p, self.q, self.c, a, g2, self.dt, sd, mean_out=mean)
        z = (u - mean.reshape(B, -1)) / sd
        return (-0.5 * z * z - math.log(sd) - 0.5 * math.log(2 * math.pi)).sum(dim=-1)

    @property
    def du(self):
        return self.p * self.c

    @property
    def dv(self):
        return self.q * self.c

    def concat(self, x, y, **kwargs):
        """dataset.concat (fbs/data/images.py:352-363) for one image: x [p, c], y [q, c] -> [H, W, c]."""
        x = dev(x, F32).reshape(1, self.p, self.c)
        y = dev(y, F32).reshape(self.q, self.c)
        img = torch.empty((1, self.unet.H, self.unet.W, self.c), dtype=F32, device=x.device)
        ops.assemble_image(x, y, self.unobs, self.obs, img)
        return img[0]

    def _fwd_sampler_em(self, key, xy0):
        """sb_imgs/supr.py:132-137 = simulators.py:53-106 with ``integration_nsteps = 1`` and the forward-drift network: per
        interval k one network evaluation at (x, ts[k]) and ``x += drift ddt + dispersion(ts[k]) sqrt(ddt) normal(keys[k])``."""
        ts32 = self.ts.astype(np.float32)
        keys = frandom.split(dev(key, torch.uint32).reshape(2), self.K)                # simulators.py:81
        x = xy0.reshape(1, self.unet.H, self.unet.W, self.c).contiguous()
        path = torch.empty((self.K + 1,) + tuple(x.shape[1:]), dtype=F32, device=x.device)
        path[0].copy_(x[0])
        for k in range(self.K):
            ddt = np.float32(np.abs(ts32[k + 1] - ts32[k]))                            # simulators.py:90 (m = 1)
            gs = np.float32(self.sde.dispersion(float(ts32[k]))) * np.sqrt(ddt)        # simulators.py:87
            drift = self.fwd_unet(x, float(ts32[k]))
            ops.em_drift_step(keys[k:k + 1], x, drift, ddt, gs, path[k + 1:k + 2])
            x = path[k + 1:k + 2]
        return path

    def fwd_sampler(self, key, x0, y0, **kwargs):
        """inpainting.py:150-152: simulate_cond_forward(key, concat(x0, y0), ts) -> path [K + 1, H, W, c]
        (sb_imgs/supr.py:132-137 when a forward-drift network was given)."""
        if self.fwd_unet is not None:
            return self._fwd_sampler_em(key, self.concat(x0, y0))
        if self.drift_mode:
            raise NotImplementedError('a drift-mode model needs the forward-drift network (fwd_unet) to sample forward paths')
        from ..sdes.linear import step_coefficients, forward_path
        if not hasattr(self, '_fwd_coef'):
            self._fwd_coef = step_coefficients(self.sde, self.ts)
        xy0 = self.concat(x0, y0)
        path = forward_path(dev(key, torch.uint32).reshape(2), xy0.reshape(-1), *self._fwd_coef)
        return path.reshape(self.K + 1, self.unet.H, self.unet.W, self.c)

    def fwd_ys_sampler(self, key, y0, **kwargs):
        """inpainting.py:155-157: simulate_cond_forward(key, y0, ts) on the observed pixels alone -> [K + 1, q, c]."""
        if self.drift_mode:
            raise NotImplementedError('the learnt forward drift of a Schroedinger bridge couples x and y: no y-only sampler '
                                      '(sb_imgs/supr.py defines none)')
        from ..sdes.linear import step_coefficients, forward_path
        if not hasattr(self, '_fwd_coef'):
            self._fwd_coef = step_coefficients(self.sde, self.ts)
        k = dev(key, torch.uint32)
        path = forward_path(k.reshape(2), dev(y0, F32).reshape(-1), *self._fwd_coef)
        return path.reshape(self.K + 1, self.q, self.c)

    def ref_sampler(self, key, _, n, **kwargs):
        """inpainting.py:160-161: N(0, I) reference draw of the unobserved pixels -> [n, p, c]."""
        return frandom.normal(dev(key, torch.uint32).reshape(2), (int(n), self.p, self.c))

    def fwd_sampler_reversed(self, key, x0, y0):
        """(us, vs) = (path_x[::-1], path_y[::-1]) of one forward-noising draw (gibbs.py:127-130), flattened per step.
        With keys ``[C, 2]`` (several conditioning targets): ``x0 [C, p c]``, ``y0 [C, q c]`` or shared -> ``[C, K + 1, .]``."""
        k = dev(key, torch.uint32).reshape(-1, 2)
        if k.shape[0] > 1:
            C_ = k.shape[0]
            x0c = dev(x0, F32).reshape(C_, self.p * self.c)
            y0c = dev(y0, F32).reshape(-1, self.q * self.c)
            parts = [self.fwd_sampler_reversed(k[ci], x0c[ci], y0c[ci if y0c.shape[0] > 1 else 0]) for ci in range(C_)]
            return torch.cat([a for a, _ in parts]).contiguous(), torch.cat([b for _, b in parts]).contiguous()
        path = self.fwd_sampler(key, dev(x0, F32).reshape(self.p, self.c), dev(y0, F32).reshape(self.q, self.c))
        px, py = self.unpack(path)
        return (torch.flip(px, dims=[0]).reshape(1, self.K + 1, self.du).contiguous(),
                torch.flip(py, dims=[0]).reshape(1, self.K + 1, self.dv).contiguous())

    def unpack(self, xy, **kwargs):
        """dataset.unpack (fbs/data/images.py:333-350)."""
        xy = dev(xy, F32)
        flat = xy.reshape(*xy.shape[:-3], self.unet.H * self.unet.W, self.c)
        return flat[..., self.unobs.long(), :], flat[..., self.obs.long(), :]

"""Thin torch-tensor wrappers over the score-network C-ABI entry points (include/fbs_b200.h, "Score network").

Tensors are CUDA, contiguous, NHWC; torch only owns the memory and the stream."""
import ctypes as C
import torch
from .. import _native as nat
from .._tensor import ptr, stream

BF16, F32 = torch.bfloat16, torch.float32


def _conv_args(in0, weight, Cout, kh, kw, off, H, W, in1=None, bias=None, residual=None, out_f32=None, out_bf16=None,
               pixel_shuffle=False, gn_partials=None):
    a = nat.NNConvStruct()
    B, Hin, Win, C0 = in0.shape
    a.B, a.H, a.W, a.Hin, a.Win = B, H, W, Hin, Win
    a.C0, a.C1, a.Cout = C0, (in1.shape[-1] if in1 is not None else 0), Cout
    a.kh, a.kw, a.off_h, a.off_w = kh, kw, off, off
    a.pixel_shuffle, a.reserved = int(pixel_shuffle), 0
    a.in0, a.in1, a.weight = ptr(in0), ptr(in1), ptr(weight)
    a.bias, a.residual, a.out_f32, a.out_bf16 = ptr(bias), ptr(residual), ptr(out_f32), ptr(out_bf16)
    a.gn_partials = ptr(gn_partials)
    return a


def conv(in0, weight, Cout, kh, kw, off, H, W, **kw_args):
    """in0 / in1: bf16 [B, Hin, Win, C]; weight: bf16 [Cout, kh * kw * (C0 + C1)]; H, W: output pixels.
    ``gn_partials`` (fp32 [B, slots, Cout / 4, 2], slots from :func:`conv_gn_slots`): GroupNorm statistics of the output."""
    a = _conv_args(in0, weight, Cout, kh, kw, off, H, W, **kw_args)
    nat.call('fbs_nn_conv_bf16', stream(), C.byref(a))


def conv_gn_slots(in0, weight, Cout, kh, kw, off, H, W, **kw_args):
    """Slots per sample of the ``gn_partials`` this convolution call would write (0: it cannot)."""
    a = _conv_args(in0, weight, Cout, kh, kw, off, H, W, **kw_args)
    n = C.c_int32(0)
    nat.call('fbs_nn_conv_gn_layout', C.byref(a), C.byref(n))
    return int(n.value)


def groupnorm_swish_stats(x, partials, gamma, beta, groups=8, tss=None, residual=None, out_f32=None, out_bf16=None, eps=1e-6,
                          ln_gamma=None, ln_out_bf16=None, ln_eps=1e-5):
    """GroupNorm + swish of ``x`` (fp32 or bf16 [B, ..., C]) with the statistics from the producing convolution's partials;
    ``ln_gamma`` / ``ln_out_bf16``: also ``LayerNorm(result) * ln_gamma`` (C <= 128)."""
    B, Cc = x.shape[0], x.shape[-1]
    P = x.numel() // (B * Cc)
    x32, x16 = (ptr(x), None) if x.dtype == F32 else (None, ptr(x))
    nat.call('fbs_nn_groupnorm_swish_stats', stream(), x32, x16, ptr(partials), partials.shape[1], B, P, Cc, groups, ptr(gamma),
             ptr(beta), ptr(tss), ptr(residual), float(eps), ptr(out_f32), ptr(out_bf16), ptr(ln_gamma), float(ln_eps),
             ptr(ln_out_bf16))


def groupnorm_swish(x, gamma, beta, groups=8, tss=None, residual=None, out_f32=None, out_bf16=None, eps=1e-6):
    B, C = x.shape[0], x.shape[-1]
    P = x.numel() // (B * C)
    nat.call('fbs_nn_groupnorm_swish_f32', stream(), ptr(x), B, P, C, groups, ptr(gamma), ptr(beta), ptr(tss), ptr(residual),
             float(eps), ptr(out_f32), ptr(out_bf16))


def layernorm(x, gamma, residual=None, out_f32=None, out_bf16=None, eps=1e-5):
    Cc = x.shape[-1]
    nat.call('fbs_nn_layernorm_f32', stream(), ptr(x), x.numel() // Cc, Cc, ptr(gamma), ptr(residual), float(eps), ptr(out_f32),
             ptr(out_bf16))


def linear_attention(qkv, out_bf16, heads=4, dim_head=32):
    B = qkv.shape[0]
    P = qkv.numel() // (B * qkv.shape[-1])
    nat.call('fbs_nn_linear_attention_bf16', stream(), ptr(qkv), B, P, heads, dim_head, ptr(out_bf16))


def attention(qkv, out_bf16, heads=4, dim_head=32, scale=10.0):
    B = qkv.shape[0]
    P = qkv.numel() // (B * qkv.shape[-1])
    nat.call('fbs_nn_attention_bf16', stream(), ptr(qkv), B, P, heads, dim_head, float(scale), ptr(out_bf16))


def time_mlp(tval, dt, dim, W0, b0, W1, b1, Wcat, bcat, table):
    nat.call('fbs_nn_time_mlp_f32', stream(), ptr(tval), float(dt), dim, ptr(W0), ptr(b0), ptr(W1), ptr(b1), ptr(Wcat), ptr(bcat),
             Wcat.shape[1], ptr(table))


def stem_conv(x, weight, bias, out_f32=None, out_bf16=None):
    B, H, W, Cin = x.shape
    nat.call('fbs_nn_stem_conv_f32', stream(), ptr(x), B, H, W, Cin, weight.shape[-1], ptr(weight), ptr(bias), ptr(out_f32),
             ptr(out_bf16))


def head_conv(x, weight, bias, out):
    Cc = x.shape[-1]
    nat.call('fbs_nn_head_conv_f32', stream(), ptr(x), x.numel() // Cc, Cc, weight.shape[-1], ptr(weight), ptr(bias), ptr(out))


def space_to_depth(x, out):
    B, H, W, Cc = x.shape
    nat.call('fbs_nn_space_to_depth_bf16', stream(), ptr(x), B, H, W, Cc, ptr(out))


def assemble_image(us, v, unobs, obs, img):
    B, p, c = us.shape
    nat.call('fbs_nn_assemble_image_f32', stream(), ptr(us), ptr(v), ptr(unobs), ptr(obs), B, p, v.shape[0], c, ptr(img))


def em_step(img, score, unobs, obs, B, p, q, c, a, g2, dt, sd, v_next=None, key=None, us_new=None, mean_out=None, lw=None,
            row_offset=0, rows_total=None, pin_row=None, pin_value=None):
    nat.call('fbs_nn_em_step_f32', stream(), ptr(img), ptr(score), ptr(unobs), ptr(obs), ptr(v_next), ptr(key), B, p, q, c,
             float(a), float(g2), float(dt), float(sd), int(row_offset), int(B if rows_total is None else rows_total),
             ptr(pin_row), ptr(pin_value), ptr(us_new), ptr(mean_out), ptr(lw))


def normalise_logw(lw, log_w=None, w=None):
    """``log_w = lw - logsumexp(lw)``, ``w = exp(log_w)`` for ``lw [B, N]`` (or ``[N]``) in one launch."""
    N = lw.shape[-1]
    nat.call('fbs_normalise_logw_f32', stream(), ptr(lw), lw.numel() // N, N, ptr(log_w), ptr(w))


def em_drift_step(keys, x, drift, ddt, gs, out):
    B = keys.shape[0] if keys.dim() == 2 else 1
    nat.call('fbs_em_drift_step_f32', stream(), ptr(keys), ptr(x), ptr(drift), B, x.numel() // B, float(ddt), float(gs), ptr(out))


def gather_rows(src, idx, dst):
    B = idx.shape[0]
    nat.call('fbs_gather_rows_f32', stream(), ptr(src), ptr(idx), B, src.numel() // src.shape[0], src.shape[0], ptr(dst))


def to_bf16(x, y):
    nat.call('fbs_nn_f32_to_bf16', stream(), ptr(x), x.numel(), ptr(y))

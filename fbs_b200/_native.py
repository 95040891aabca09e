"""ctypes binding of the C ABI declared in include/fbs_b200.h.

There is NO fallback: if the shared library is missing, or a call returns non-zero, this
raises.  Nothing here touches ``oracle/``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_lib', 'libfbs_b200.so')

RESAMPLE_MULTINOMIAL, RESAMPLE_KILLING, RESAMPLE_SYSTEMATIC, RESAMPLE_STRATIFIED = 0, 1, 2, 3
INIT_DEGENERATE, INIT_NORMAL = 0, 1

_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_f32 = C.c_float
_f64 = C.c_double
_int = C.c_int


class AffineModelStruct(C.Structure):
    _fields_ = [('K', _i32), ('du', _i32), ('dv', _i32), ('reserved', _i32),
                ('MT', _p), ('m', _p), ('dt', _p), ('sd', _p), ('lognorm', _p), ('MTp', _p), ('MTc', _p)]


_M = C.POINTER(AffineModelStruct)


class NNConvStruct(C.Structure):
    _fields_ = [('B', _i32), ('H', _i32), ('W', _i32), ('Hin', _i32), ('Win', _i32), ('C0', _i32), ('C1', _i32), ('Cout', _i32),
                ('kh', _i32), ('kw', _i32), ('off_h', _i32), ('off_w', _i32), ('pixel_shuffle', _i32), ('reserved', _i32),
                ('in0', _p), ('in1', _p), ('weight', _p), ('bias', _p), ('residual', _p), ('out_f32', _p), ('out_bf16', _p), ('gn_partials', _p)]


_CV = C.POINTER(NNConvStruct)

# name -> (argtypes, restype); every symbol include/fbs_b200.h declares (checked by tests/test_abi.py)
SIGNATURES = {
    'fbs_version': ([], _int),
    'fbs_last_error': ([], C.c_char_p),
    'fbs_launch_count': ([], _i64),
    'fbs_reset_launch_count': ([], None),
    'fbs_debug_set_option': ([C.c_char_p, _int], _int),
    'fbs_random_bits_u32': ([_p, _p, _i64, _i64, _p], _int),
    'fbs_random_split': ([_p, _p, _i64, _i64, _p], _int),
    'fbs_random_uniform_f32': ([_p, _p, _i64, _i64, _f32, _f32, _p], _int),
    'fbs_random_normal_f32': ([_p, _p, _i64, _i64, _p], _int),
    'fbs_random_randint_i32': ([_p, _p, _i64, _i64, _i32, _i32, _p], _int),
    'fbs_random_choice_f32': ([_p, _p, _p, _i64, _i64, _i64, _p], _int),
    'fbs_cond_resample_f32': ([_p, _int, _p, _p, _p, _p, _int, _i64, _i64, _p], _int),
    'fbs_resample_f32': ([_p, _int, _p, _p, _i64, _i64, _p], _int),
    'fbs_ou_forward_path_f32': ([_p, _p, _p, _int, _p, _p, _i64, _i64, _i64, _i64, _int, _p, _p], _int),
    'fbs_em_affine_path_f32': ([_p, _p, _p, _int, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _int, _p, _p], _int),
    'fbs_csmc_step_affine_f32': ([_p, _M, _i32, _int, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _p, _p, _p], _int),
    'fbs_affine_eval_f32': ([_p, _M, _i32, _p, _p, _p, _p, _p, _i64, _i64, _p, _p, _p], _int),
    'fbs_debug_umma_gemm': ([_p, _p, _p, _i32, _i32, _p], _int),
    'fbs_debug_step_tc_timers': ([_p], _int),
    'fbs_debug_v3_timeline': ([_p], _int),
    'fbs_debug_conv_timeline': ([_p], _int),
    'fbs_sweep_workspace_bytes': ([_M, _i64], C.c_size_t),
    'fbs_csmc_forward_affine_f32': ([_p, _M, _p, _p, _p, _p, _int, _f32, _int, _i64, _i64, _p, _p, _p, _p, _p, _p,
                                     C.c_size_t], _int),
    'fbs_backward_scan_f32': ([_p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _p, _p], _int),
    'fbs_pmcmc_filter_affine_f32': ([_p, _M, _p, _p, _p, _int, _i64, _i64, _p, _p, _p, _p, _p, _p, C.c_size_t], _int),
    'fbs_bootstrap_filter_affine_f32': ([_p, _M, _p, _p, _p, _int, _i64, _i64, _p, _p, _p, _p, _p], _int),
    'fbs_backward_sample_affine_f32': ([_p, _M, _int, _p, _p, _p, _p, _int, _i64, _i64, _p, _p], _int),
    'fbs_twisted_smc_affine_f32': ([_p, _p, _p, _p, _p, _p, _f32, _f32, _i64, _i64, _p, _p, _int, _p, _int, _i64, _i64, _p, _p,
                                    _p, _p, _p], _int),
    'fbs_force_move_f32': ([_p, _p, _p, _int, _p, _p, _i64, _i64, _i64, _p, _p, _p], _int),
    'fbs_pcn_combine_f32': ([_p, _f64, _p, _p, _p, _p, _i64, _i64, _p], _int),
    'fbs_mh_accept_f32': ([_p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i32, _p, _p, _p, _p, _p], _int),
    'fbs_gaussian_ref_sample_f32': ([_p, _p, _p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _p], _int),
    'fbs_nn_conv_bf16': ([_p, _CV], _int),
    'fbs_nn_conv_gn_layout': ([_CV, _p], _int),
    'fbs_nn_groupnorm_swish_stats': ([_p, _p, _p, _p, _i32, _i64, _i32, _i32, _i32, _p, _p, _p, _p, _f32, _p, _p, _p, _f32, _p], _int),
    'fbs_nn_groupnorm_swish_f32': ([_p, _p, _i64, _i32, _i32, _i32, _p, _p, _p, _p, _f32, _p, _p], _int),
    'fbs_nn_layernorm_f32': ([_p, _p, _i64, _i32, _p, _p, _f32, _p, _p], _int),
    'fbs_nn_linear_attention_bf16': ([_p, _p, _i64, _i32, _i32, _i32, _p], _int),
    'fbs_nn_attention_bf16': ([_p, _p, _i64, _i32, _i32, _i32, _f32, _p], _int),
    'fbs_nn_time_mlp_f32': ([_p, _p, _f32, _i32, _p, _p, _p, _p, _p, _p, _i32, _p], _int),
    'fbs_nn_stem_conv_f32': ([_p, _p, _i64, _i32, _i32, _i32, _i32, _p, _p, _p, _p], _int),
    'fbs_nn_head_conv_f32': ([_p, _p, _i64, _i32, _i32, _p, _p, _p], _int),
    'fbs_nn_space_to_depth_bf16': ([_p, _p, _i64, _i32, _i32, _i32, _p], _int),
    'fbs_nn_assemble_image_f32': ([_p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _p], _int),
    'fbs_nn_em_step_f32': ([_p, _p, _p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _f32, _f32, _f32, _f32, _i64, _i64, _p, _p, _p, _p, _p],
                           _int),
    'fbs_normalise_logw_f32': ([_p, _p, _i64, _i64, _p, _p], _int),
    'fbs_em_drift_step_f32': ([_p, _p, _p, _p, _i64, _i64, _f32, _f32, _p], _int),
    'fbs_gather_rows_f32': ([_p, _p, _p, _i64, _i64, _i64, _p], _int),
    'fbs_gather_rows_peer_f32': ([_p, _p, _p, _i64, _i64, _i64, _i64, _p], _int),
    'fbs_ipc_export': ([_p, _p, _p], _int),
    'fbs_ipc_import': ([_p, _i64, _p], _int),
    'fbs_ipc_release': ([_p, _i64], _int),
    'fbs_nn_f32_to_bf16': ([_p, _p, _i64, _p], _int),
}

_lib = None


class NativeError(RuntimeError):
    pass


def lib():
    """Load the CUDA library; raise loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(f'{LIB_PATH} is missing: run `python -m fbs_b200.build` (or __graft_entry__.build()). '
                              'fbs_b200 has no CPU fallback.')
        handle = C.CDLL(LIB_PATH)
        for name, (argtypes, restype) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = handle
    return _lib


# Implementation pins for tests / A-B measurements.  The C library never reads the environment; the Python layer maps the
# FBS_* variables onto fbs_debug_set_option whenever their values change (checked per call: a few dict lookups).
_ENV_OPTS = {
    'FBS_SWEEP_IMPL': ('sweep_impl', {'v1': 1, 'v2': 2, 'v3': 3, 'v4': 4}),
    'FBS_SWEEP_VERBOSE': ('sweep_verbose', None),
    'FBS_STEP_IMPL': ('step_impl', {'cuda': 1}),
    'FBS_STEP_TC_WARPS': ('step_tc_warps', None),
    'FBS_STEPVEC_IMPL': ('stepvec_impl', {'old': 1}),
    'FBS_SWEEP_G': ('sweep_g', None),
    'FBS_V3_TWOPASS': ('v3_twopass', None),
    'FBS_EM_IMPL': ('em_impl', {'cta': 1, 'tpc': 2}),
    'FBS_V3_VARIANT': ('v3_variant', None),
    'FBS_CONV_IMPL': ('conv_impl', {'taps': 1}),
}
_env_seen = {}


def _sync_env_options(handle):
    for env, (name, table) in _ENV_OPTS.items():
        raw = os.environ.get(env)
        if _env_seen.get(env) == raw:
            continue
        _env_seen[env] = raw
        if raw is None or raw == '':
            val = 0
        elif table is not None:
            if raw not in table:
                raise ValueError(f'{env}={raw!r}: expected one of {sorted(table)}')
            val = table[raw]
        else:
            val = int(raw)
        if handle.fbs_debug_set_option(name.encode(), val) != 0:
            raise NativeError(handle.fbs_last_error().decode('utf-8', 'replace'))


# name -> list of (start_event, end_event): filled when a caller asks for per-kernel device timing
# (bench.py uses it to time the dominant kernel inside a whole step without a profiler).
TIMED = {}


# FBS_NVTX=1: every C-ABI call is wrapped in an NVTX range named after the entry point (per-phase ranges for
# `ncu --nvtx` / Nsight timelines); off by default -- the push / pop are host-side calls on the launch path.
NVTX = os.environ.get('FBS_NVTX', '') not in ('', '0')


def call(name, *args):
    handle = lib()
    _sync_env_options(handle)
    if NVTX:
        import torch
        torch.cuda.nvtx.range_push(name)
        try:
            return _call(handle, name, args)
        finally:
            torch.cuda.nvtx.range_pop()
    return _call(handle, name, args)


def _call(handle, name, args):
    rec = TIMED.get(name)
    if rec is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(handle, name)(*args)
        e1.record()
        rec.append((e0, e1))
    else:
        rc = getattr(handle, name)(*args)
    if rc != 0:
        msg = handle.fbs_last_error().decode('utf-8', 'replace')
        cls = NotImplementedError if rc == 3 else (ValueError if rc == 1 else NativeError)
        raise cls(f'{name} failed (code {rc}): {msg}')


def launch_count() -> int:
    return int(lib().fbs_launch_count())


def reset_launch_count():
    lib().fbs_reset_launch_count()

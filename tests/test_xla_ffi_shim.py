"""Compile-only check of the XLA-FFI shim (fbs_b200/csrc/xla_ffi_shim.cc).

jaxlib is not installable here, so the real ``xla/ffi/api/ffi.h`` is absent and the shim cannot be linked into the
product or run.  This test keeps it honest anyway: it is compiled by g++ against tests/xla_ffi_stub/ -- a minimal stand-in
for the binding types whose XLA_FFI_DEFINE_HANDLER_SYMBOL statically asserts that every implementation is invocable with
exactly the argument list its ``Bind()`` chain decodes (what the real header checks) -- and the object must define one
handler symbol per hot-path entry point of include/fbs_b200.h, each of which must call that entry point.
"""
import os
import re
import shutil
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, 'fbs_b200', 'csrc', 'xla_ffi_shim.cc')
STUB = os.path.join(ROOT, 'tests', 'xla_ffi_stub')
CUDA_INC = '/usr/local/cuda/include'

# every entry point a jitted reference driver needs -> must be wrapped (host-side IPC plumbing, debug hooks and the
# bookkeeping calls are not XLA custom calls)
NOT_WRAPPED = {'fbs_version', 'fbs_last_error', 'fbs_launch_count', 'fbs_reset_launch_count', 'fbs_debug_set_option',
               'fbs_debug_umma_gemm', 'fbs_debug_step_tc_timers', 'fbs_debug_v3_timeline', 'fbs_debug_conv_timeline', 'fbs_nn_conv_gn_layout', 'fbs_sweep_workspace_bytes', 'fbs_ipc_export', 'fbs_ipc_import',
               'fbs_ipc_release', 'fbs_gather_rows_peer_f32', 'fbs_nn_f32_to_bf16'}


def _declared_entry_points():
    text = open(os.path.join(ROOT, 'include', 'fbs_b200.h')).read()
    return set(re.findall(r'\b(fbs_[a-z0-9_]+)\s*\(', text))


@pytest.mark.skipif(shutil.which('g++') is None or not os.path.exists(os.path.join(CUDA_INC, 'cuda_runtime.h')),
                    reason='needs g++ and the CUDA headers')
def test_shim_compiles_against_the_stand_in_and_defines_every_handler(tmp_path):
    obj = tmp_path / 'xla_ffi_shim.o'
    res = subprocess.run(['g++', '-std=c++17', '-Wall', '-Werror=return-type', '-c', SHIM, '-I', STUB, '-I', CUDA_INC, '-o', str(obj)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-4000:]
    syms = subprocess.run(['nm', '--defined-only', str(obj)], capture_output=True, text=True).stdout
    handlers = set(re.findall(r'\bT (fbs_xla_[a-z0-9_]+)', syms))
    src = open(SHIM).read()
    assert handlers == set(re.findall(r'XLA_FFI_DEFINE_HANDLER_SYMBOL\((fbs_xla_[a-z0-9_]+),', src))
    called = set(re.findall(r'\b(fbs_[a-z0-9_]+)\(', src)) - {h for h in handlers}
    missing = _declared_entry_points() - NOT_WRAPPED - called
    assert not missing, f'ABI entry points without an XLA-FFI handler: {sorted(missing)}'
    # the headline sweeps must reach the tcgen05 kernel: the model descriptor carries MTc
    assert 'mod.MTc = opt(MTc)' in src
    assert len(handlers) >= 30


def test_stub_rejects_a_mismatched_binding(tmp_path):
    """The stand-in's static check is real: an implementation whose signature disagrees with its Bind() chain fails to compile."""
    if shutil.which('g++') is None:
        pytest.skip('needs g++')
    bad = tmp_path / 'bad.cc'
    bad.write_text('#include "xla/ffi/api/ffi.h"\nnamespace ffi = xla::ffi;\n'
                   'static ffi::Error Impl(ffi::Buffer<ffi::F32> a, int32_t k, ffi::ResultBuffer<ffi::F32> out) { return ffi::Error::Success(); }\n'
                   'XLA_FFI_DEFINE_HANDLER_SYMBOL(h, Impl, ffi::Ffi::Bind().Arg<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>());\n')
    res = subprocess.run(['g++', '-std=c++17', '-fsyntax-only', str(bad), '-I', STUB], capture_output=True, text=True)
    assert res.returncode != 0 and 'not invocable' in res.stderr

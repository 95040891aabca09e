"""Host-side logic of the product package that needs no GPU: model precompute, SDE coefficients, chain
partitioning (incl. a real 2-process gloo run)."""
import os
import subprocess
import sys
import numpy as np
import pytest
from oracle import jax_random as jr
from oracle import sdes as osdes
from helpers import gp_problem, oracle_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_affine_model_precompute_matches_oracle_drift():
    import fbs_b200
    from fbs_b200 import sdes
    for kind in ('const', 'lin'):
        p = gp_problem(6, K=12, sde_kind=kind)
        om64 = oracle_model(p, np.float64)
        sde = sdes.StationaryLinLinearSDE(0.02, 4., 0., 1.) if kind == 'lin' else sdes.StationaryConstLinearSDE(-0.5, 1.)
        pm = fbs_b200.AffineGaussianModel.from_linear_sde(sde, p['jm'], p['jc'], 6, p['ts'], T=1.)
        uv = np.random.default_rng(0).normal(size=(9, 12))
        for k in (0, 5, 11):
            M = pm.host['MT'][k].T.astype(np.float64)
            want = om64.reverse_drift(uv, om64.ts[k])
            np.testing.assert_allclose(uv @ M.T + pm.host['m'][k], want, rtol=2e-5, atol=2e-5)
        # packed image used by the tiled kernel: [K][du][dup | dvp]
        assert pm.host['MTp'].shape == (12, 6, 16)
        np.testing.assert_array_equal(pm.host['MTp'][:, :, :6], pm.host['MT'][:, :6, :6])
        np.testing.assert_array_equal(pm.host['MTp'][:, :, 8:14], pm.host['MT'][:, :6, 6:])
        assert (pm.host['MTp'][:, :, 6:8] == 0).all() and (pm.host['MTp'][:, :, 14:] == 0).all()


def test_step_coefficients_match_float32_reference_arithmetic():
    from fbs_b200 import sdes
    from fbs_b200.sdes.linear import step_coefficients
    ts = np.linspace(0., 1., 201)
    for psde, osde in ((sdes.StationaryConstLinearSDE(-0.5, 1.), osdes.StationaryConstLinearSDE(-0.5, 1.)),
                       (sdes.StationaryLinLinearSDE(0.02, 4., 0., 1.), osdes.StationaryLinLinearSDE(0.02, 4., 0., 1.))):
        F, sq = step_coefficients(psde, ts)
        disc, _, _ = osdes.make_linear_sde(osde, np.float32)
        ts32 = ts.astype(np.float32)
        Fo, Qo = disc(ts32[1:], ts32[:-1])
        np.testing.assert_allclose(F, Fo, rtol=2e-7)
        np.testing.assert_allclose(sq, np.sqrt(Qo), rtol=2e-6)


def test_gaussian_sb_affine_drift_matches_oracle():
    from fbs_b200 import sdes
    rng = np.random.default_rng(0)
    m0, m1 = rng.normal(size=3), rng.normal(size=3)
    a, b = rng.normal(size=(3, 3)), rng.normal(size=(3, 3))
    c0, c1 = a @ a.T + np.eye(3), b @ b.T + np.eye(3)
    _, _, drift = sdes.make_gaussian_bw_sb(m0, c0, m1, c1, sig=1.)
    _, _, odrift = osdes.make_gaussian_bw_sb(m0, c0, m1, c1, sig=1.)
    x = rng.normal(size=(5, 3))
    for t in (0.1, 0.5, 0.93):
        np.testing.assert_allclose(drift(x, t), odrift(x, t), rtol=1e-8, atol=1e-9)


def test_chain_partition_covers_everything_once():
    from fbs_b200.parallel import chain_slice, chain_keys, threefry_split_host
    for total, world in [(10, 4), (8192, 8), (3, 8), (4096, 1)]:
        seen = []
        for r in range(world):
            lo, hi = chain_slice(total, r, world)
            seen += list(range(lo, hi))
        assert seen == list(range(total))
    key = jr.PRNGKey(11)
    np.testing.assert_array_equal(threefry_split_host(key, 37), jr.split(key, 37))
    full = jr.split(key, 64)
    got = np.concatenate([chain_keys(key, 64, r, 4) for r in range(4)])
    np.testing.assert_array_equal(got, full)          # a chain's key does not depend on the GPU count


_GLOO_SCRIPT = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from fbs_b200.parallel import chain_slice, chain_keys
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
total = 11
key = np.array([0, 5], dtype=np.uint32)
keys = chain_keys(key, total, rank, world)
lo, hi = chain_slice(total, rank, world)
# every rank "processes" its chains (checksum of its keys) -- no data-path collective; only the final gather
local = torch.tensor([int(keys.astype(np.uint64).sum()), hi - lo], dtype=torch.int64)
gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
dist.all_gather(gathered, local)
t = torch.tensor([float(rank + 1)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)           # the max-over-ranks timing reduction bench.py uses
if rank == 0:
    full = chain_keys(key, total, 0, 1)
    assert sum(int(g[1]) for g in gathered) == total
    assert sum(int(g[0]) for g in gathered) == int(full.astype(np.uint64).sum())
    assert t.item() == world
    print('GLOO_OK')
dist.destroy_process_group()
'''


def test_two_process_gloo_partition(tmp_path):
    script = tmp_path / 'gloo_part.py'
    script.write_text(_GLOO_SCRIPT)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
           '--master-port', '29531', str(script), ROOT]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout + res.stderr
    assert 'GLOO_OK' in res.stdout


_GLOO_SHARD_SCRIPT = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from fbs_b200.sharded import ParticleShard, exchange_rows, all_gather_rows, exchange_plan
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
g = torch.Generator().manual_seed(0)          # identical on every rank
N, D = 12, 5
rows = torch.randn(N, D, generator=g)
for trial in range(6):
    A = torch.randint(0, N, (N,), generator=g)
    if trial == 0:
        A = torch.arange(N)                   # nothing moves
    if trial == 1:
        A = torch.full((N,), N - 1)           # everything comes from the last rank
    sh = ParticleShard(N, rank, world)
    got, moved = exchange_rows(rows[sh.lo:sh.hi].clone(), A, sh)
    assert torch.equal(got, rows[A[sh.lo:sh.hi]]), (trial, rank)
    want_moved = int((A[sh.lo:sh.hi] // sh.n != rank).sum())
    assert moved == want_moved, (moved, want_moved)
    full = all_gather_rows(rows[sh.lo:sh.hi, 0].contiguous(), sh)
    assert torch.equal(full, rows[:, 0])
try:
    ParticleShard(7, rank, world)
    raise SystemExit('expected ValueError')
except ValueError:
    pass
dist.barrier()
if rank == 0:
    print('GLOO_SHARD_OK')
dist.destroy_process_group()
'''


@pytest.mark.parametrize('world', [2, 3])
def test_particle_shard_exchange_gloo(tmp_path, world):
    """fbs_b200/sharded.py (config 5: particle set sharded over ranks): the all-gather of the weights and the
    point-to-point exchange of resampled particle rows reproduce rows_global[A] on every rank, for any ancestor vector."""
    script = tmp_path / 'gloo_shard.py'
    script.write_text(_GLOO_SHARD_SCRIPT)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}', '--master-addr', '127.0.0.1',
           '--master-port', str(29540 + world), str(script), ROOT]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert 'GLOO_SHARD_OK' in res.stdout


def test_bench_clock_sampler_windows_samples_to_the_timed_region():
    """bench.ClockSampler.stop keeps the nvidia-smi samples that fall inside the timed region and says when it had to fall
    back to the warm-up samples (a 0.3 s region can end before nvidia-smi's first sample)."""
    import bench

    class FakeProc:
        def terminate(self):
            pass

    line = '0, {sm}, 1965, 3996, 700.0, Not Active, Not Active, Not Active, {cap}'
    s = bench.ClockSampler(0)
    s.proc = FakeProc()
    s.lines = [(10.0, line.format(sm=1200, cap='Not Active')), (11.0, line.format(sm=1965, cap='Active')),
               (11.2, line.format(sm=1950, cap='Not Active')), (13.0, line.format(sm=300, cap='Not Active'))]
    r = s.stop(10.9, 11.3)
    assert r['samples'] == 2 and r['sm_mhz'] == 1957.5 and r['sm_max_mhz'] == 1965.0
    assert r['reasons'] == ['sw_power_cap'] and r['window'] == 'timed region'
    s2 = bench.ClockSampler(0)
    s2.proc = FakeProc()
    s2.lines = [(10.0, line.format(sm=1965, cap='Not Active'))]
    r2 = s2.stop(20.0, 20.3)
    assert r2['samples'] == 1 and r2['window'].startswith('warm-up')
    rb = bench.rng_bound(1.0e11)
    assert rb is not None and abs(rb['frac'] - 1.0e11 / 423.5e9) < 1e-9

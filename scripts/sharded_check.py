"""2..8-GPU check + timing of the particle-sharded CSMC sweep (fbs_b200/sharded.py) against the unsharded sweep.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node=G --master-addr 127.0.0.1 scripts/sharded_check.py [HxWxC] [N] [K]
Every rank runs the unsharded sweep too (small sizes) and compares its shard bit for bit."""
import json, os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fbs_b200 import sdes
from fbs_b200.nn import ScoreUNet, ScoreNetModel
from fbs_b200.samplers.csmc import csmc, resamplings as R
from fbs_b200.sharded import forward_pass_sharded
from fbs_b200 import random as fr
from oracle import unet as ou      # random checkpoint only

shape = tuple(int(a) for a in sys.argv[1].split('x')) if len(sys.argv) > 1 else (28, 28, 1)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16
K = int(sys.argv[3]) if len(sys.argv) > 3 else 4
check = (len(sys.argv) <= 4) or sys.argv[4] != 'nocheck'
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rank, world = dist.get_rank(), dist.get_world_size()
H, W, C = shape
T = 2.0
ts = np.linspace(0., T, K + 1)
sde = sdes.StationaryLinLinearSDE(beta_min=0.02, beta_max=5., t0=0., T=T)
half = H // 2
rect = np.array([(i + H // 4) * W + (j + W // 4) for i in range(half) for j in range(half)], dtype=np.int32)
obs = np.setdiff1d(np.arange(H * W, dtype=np.int32), rect)
net = ScoreUNet(ou.init_unet_params(0, C), shape, dt=T / 200)
model = ScoreNetModel(net, sde, ts, T, rect, obs)
rng = np.random.default_rng(1)                       # identical inputs on every rank
us_star = rng.standard_normal((K + 1, rect.size, C)).astype(np.float32)
vs = np.cumsum(0.05 * rng.standard_normal((K + 1, obs.size, C)), axis=0).astype(np.float32)
bs_star = rng.integers(0, N, size=K + 1).astype(np.int32)
key = fr.PRNGKey(7)
init = csmc.DegenerateInit(N)
for rep in range(2):                                 # first pass warms the CUDA graphs up
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    r = forward_pass_sharded(key, us_star, bs_star, vs, model, init, R.killing, N, history=check)
    torch.cuda.synchronize(); dist.barrier()
    t_sh = time.perf_counter() - t0
out = {'world': world, 'shape': shape, 'N': N, 'K': K, 'sharded_s': t_sh, 'particle_steps_per_s': N * K / t_sh,
       'moved_rows_per_step': float(np.mean(r['moved']))}
if check:
    full = csmc.forward_pass_nn(key, us_star, bs_star, vs, model, init, R.killing.scheme, N, history=True)
    lo, hi = r['lo'], r['hi']
    assert torch.equal(r['As'], full['As'][0]), 'ancestors differ'
    uss_full = full['uss'].reshape(K + 1, N, rect.size, C)
    err_u = float((r['uss'] - uss_full[:, lo:hi]).abs().max())
    err_w = float((r['log_wss'] - full['log_wss'][0]).abs().max())
    assert err_u == 0.0 and err_w == 0.0, (err_u, err_w)
    out['bitwise_equal_to_unsharded'] = True
t = torch.tensor([t_sh], device='cuda', dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    out['sharded_s'] = float(t.item())
    print(json.dumps(out), flush=True)
from fbs_b200.sharded import close_peer_buffers
close_peer_buffers()
dist.destroy_process_group()

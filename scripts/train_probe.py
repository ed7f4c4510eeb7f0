"""A few fused NGP training steps (for ncu launch lists / quick timing).  python scripts/train_probe.py [steps]"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xrnerf_b200 import synth
from xrnerf_b200.ngp import NgpField
from xrnerf_b200.train import NgpTrainer

dev = torch.device('cuda')
N = 65536
K = int(sys.argv[1]) if len(sys.argv) > 1 else 6
bf = torch.from_numpy(synth.bitfield_from_grid_numpy(synth.lego_like_density_grid(0))[0]).to(dev)
batches = [tuple(torch.from_numpy(x).to(dev) for x in synth.ray_batch(N, seed=b)[:2]) for b in range(4)]
f = NgpField(n_packed_levels=int(os.environ.get('XRB_PACKED_LEVELS', '6'))).to(dev)
tr = NgpTrainer(f, bf, N, target_batch_size=1 << 20)
tgt = torch.rand((N, 3), device=dev); bg = torch.zeros((N, 3), device=dev)
for i in range(3):
    tr.step(*batches[i % 4], tgt, bg, next_rays=batches[(i + 1) % 4])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(K):
    tr.step(*batches[i % 4], tgt, bg, next_rays=batches[(i + 1) % 4])
e1.record(); torch.cuda.synchronize()
print('train step %.3f ms, compacted samples %d, trained rays %d' % (e0.elapsed_time(e1) / K, int(tr.compacted_samples()), int(tr.trained_rays())))

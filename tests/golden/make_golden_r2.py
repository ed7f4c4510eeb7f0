"""Generates tests/golden/nerf_golden_r2.npz (round-2 additions) by running the REFERENCE's own code (imported unmodified from /root/reference through
oracle/ref_import.py) on seeded inputs. Run in the build container only:   python tests/golden/make_golden_r2.py

  batch.*   the NGP training-ray source: load_rays_hash (datasets/load_data/get_rays.py:72-98) -> a fixed row permutation (np.random.shuffle's role,
            hashnerf_dataset.py:41-44) -> HashBatchSample (datasets/pipelines/create.py:153-190, two consecutive batches) -> RandomBGColor
            (augment.py:290-313, np.random.rand replaced by recorded uniforms)
  select.*  SelectRays (augment.py:12-76), full image and precrop window, np.random.choice replaced by recorded indices
  zrand.*   GetZvals(randomized=True) (create.py:518-525), torch.rand replaced by recorded uniforms
  mipr.*    resample_along_rays(randomized=True) (networks/utils/mip.py:7-63,:146-176), torch.rand replaced by recorded uniforms
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import as R  # noqa: E402


def main():
    torch.manual_seed(7)
    rng = np.random.default_rng(7)
    torch.set_num_threads(1)
    create, augment, get_rays = R.load_pipelines()
    mip = R.load('networks.utils.mip')
    out = {}
    # ------------------------------------------------------------------ NGP batch source
    H, W, foc, I = 6, 10, 13.5, 3
    K = np.array([[foc, 0, 0.5 * W], [0, foc, 0.5 * H], [0, 0, 1]], np.float32)
    poses = rng.normal(0, 1, (I, 4, 3)).astype(np.float32)                       # [I,4,3] as poses_nerf2ngp leaves them (utils/hashnerf.py:4-23)
    images = rng.random((I, H, W, 4)).astype(np.float32)
    table = get_rays.load_rays_hash(H, W, K, poses, images)                       # [I*H*W, 11], pixel order
    perm = rng.permutation(table.shape[0])
    shuffled = table[perm]
    hb = create.HashBatchSample(N_rand=16)
    bgc = augment.RandomBGColor()
    u_bg = rng.random((2, 16, 3))
    orig_rand = np.random.rand
    for b in range(2):
        res = hb({'rays_rgb': shuffled})
        np.random.rand = lambda *shape, _b=b: u_bg[_b]
        res = bgc(res)
        np.random.rand = orig_rand
        for k in ('rays_o', 'rays_d', 'target_s', 'alpha', 'img_ids', 'bg_color'):
            out[f'batch.{b}.{k}'] = np.asarray(res[k], np.float32)
    out.update({'batch.poses': poses, 'batch.images': images, 'batch.K': K, 'batch.perm': perm.astype(np.int64), 'batch.u_bg': u_bg.astype(np.float32), 'batch.table_head': table[:24]})
    # ------------------------------------------------------------------ SelectRays
    Hs, Ws, fs = 12, 20, 27.5
    Ks = np.array([[fs, 0, 0.5 * Ws], [0, fs, 0.5 * Hs], [0, 0, 1]], np.float32)
    pose = torch.tensor([[0.36, -0.48, 0.8, 1.5], [0.8, 0.6, 0.0, -0.7], [-0.48, 0.64, 0.6, 2.2], [0, 0, 0, 1]], dtype=torch.float32)
    kw = dict(H=Hs, W=Ws, K=Ks)
    full = create.GetRays(include_radius=True, **kw)({'pose': pose})
    img = torch.rand(Hs, Ws, 3)
    orig_choice = np.random.choice
    for tag, it, frac in (('full', 100, 0.5), ('crop', 3, 0.5)):
        n_all = Hs * Ws if tag == 'full' else (2 * int(Hs // 2 * frac)) * (2 * int(Ws // 2 * frac))
        inds = rng.permutation(n_all)[:24]
        np.random.choice = lambda n, size=None, replace=True, _i=inds: _i
        sr = augment.SelectRays(sel_n=24, precrop_iters=10, precrop_frac=frac, include_radius=True, **kw)
        res = sr({'rays_o': full['rays_o'].clone(), 'rays_d': full['rays_d'].clone(), 'target_s': img.clone(), 'radii': full['radii'].clone(), 'iter_n': it})
        np.random.choice = orig_choice
        out.update({f'select.{tag}.inds': inds.astype(np.int64), f'select.{tag}.rays_o': res['rays_o'], f'select.{tag}.rays_d': res['rays_d'], f'select.{tag}.target_s': res['target_s'],
                    f'select.{tag}.radii': res['radii']})
    out.update({'select.pose': pose, 'select.K': Ks, 'select.image': img})
    # ------------------------------------------------------------------ GetZvals(randomized=True)
    n = 40
    near, far = torch.full((n, 1), 2.0), torch.full((n, 1), 6.0)
    uz = torch.rand(n, 129)
    orig = torch.rand
    torch.rand = lambda *a, **k: uz
    zr = create.GetZvals(N_samples=129, lindisp=False, randomized=True)(dict(rays_o=torch.zeros(n, 3), near=near, far=far))
    torch.rand = orig
    out.update({'zrand.u': uz, 'zrand.z': zr['z_vals']})
    # ------------------------------------------------------------------ resample_along_rays(randomized=True)
    S1 = 33
    zm, _ = torch.sort(torch.linspace(2., 6., S1).expand(n, S1).contiguous() + torch.rand(n, S1) * 0.05, -1)
    w = torch.rand(n, S1 - 1) ** 3
    w[5] = 0.0                                                                    # an all-zero weight row exercises the eps padding (mip.py:12-16)
    rays_o, rays_d, radii = torch.rand(n, 3), torch.randn(n, 3), torch.rand(n, 1) * 0.002 + 0.0005
    ur = torch.rand(n, S1)
    torch.rand = lambda *a, **k: ur
    d4 = mip.resample_along_rays({'rays_o': rays_o, 'rays_d': rays_d, 'radii': radii, 'z_vals': zm, 'weights': w.clone()}, True, 'cone', 0.01)
    torch.rand = orig
    out.update({'mipr.z_vals': zm, 'mipr.weights': w, 'mipr.u': ur, 'mipr.z_resampled': d4['z_vals']})
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'nerf_golden_r2.npz'),
                        **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()})
    print('wrote nerf_golden_r2.npz with', len(out), 'arrays')


if __name__ == '__main__':
    main()

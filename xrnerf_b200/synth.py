"""Synthetic Blender-shaped inputs for the NGP hot path (no dataset, no checkpoint on the box).

Host-side numpy mirrors of the reference's ray-table construction so that the synthetic rays have exactly the
layout/convention the kernels see in production:
  * pose_spherical            /root/reference/xrnerf/datasets/load_data/load_blender.py:22-29
  * poses_nerf2ngp            /root/reference/xrnerf/datasets/utils/hashnerf.py:4-23 (scale .33, offset .5: hashnerf_dataset.py:36-40)
  * get_rays_np_hash          /root/reference/xrnerf/datasets/load_data/get_rays.py:35-69 (+0.5 pixel centre, unit dirs)
  * get_alldata metadata      /root/reference/xrnerf/datasets/hashnerf_dataset.py:56-74
and a seeded "lego-like" occupancy grid (box + sphere-shell union inside [0,1]^3, few % fill).
"""
import numpy as np

H = W = 800
FOCAL = np.float32(0.5 * 800 / np.tan(0.5 * 0.6911112070083618))  # camera_angle_x of the lego fixture -> 1111.111


def pose_spherical(theta, phi, radius):
    def trans_t(t):
        return np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, t], [0, 0, 0, 1]], np.float32)

    def rot_phi(p):
        return np.array([[1, 0, 0, 0], [0, np.cos(p), -np.sin(p), 0], [0, np.sin(p), np.cos(p), 0], [0, 0, 0, 1]], np.float32)

    def rot_theta(t):
        return np.array([[np.cos(t), 0, -np.sin(t), 0], [0, 1, 0, 0], [np.sin(t), 0, np.cos(t), 0], [0, 0, 0, 1]], np.float32)

    c2w = trans_t(radius)
    c2w = rot_phi(phi / 180. * np.pi) @ c2w
    c2w = rot_theta(theta / 180. * np.pi) @ c2w
    c2w = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], np.float32) @ c2w
    return c2w.astype(np.float32)


def poses_nerf2ngp(poses, correct_pose=(1, -1, -1), scale=0.33, offset=(0.5, 0.5, 0.5)):
    out = []
    for i in range(poses.shape[0]):
        m = poses[i, :-1, :].copy()
        m[:, 0] *= correct_pose[0]
        m[:, 1] *= correct_pose[1]
        m[:, 2] *= correct_pose[2]
        m[:, 3] = m[:, 3] * scale + np.asarray(offset, np.float32)
        out.append(m[[1, 2, 0]])
    return np.array(out).astype(np.float32).transpose(0, 2, 1)  # [I,4,3]


def spiral_poses_ngp(n=40):
    poses = np.stack([pose_spherical(a, -30.0, 4.0) for a in np.linspace(-180, 180, n + 1)[:-1]], 0)
    return poses_nerf2ngp(poses)


def get_rays_ngp(pose43, h=H, w=W, focal=FOCAL):
    """One image worth of NGP-convention rays: rays_o, rays_d [h*w,3] float32, unit-norm dirs."""
    c2w = pose43.transpose(1, 0)  # [3,4]
    i, j = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32), indexing='xy')
    i, j = i + 0.5, j + 0.5
    dirs = np.stack([(i - 0.5 * w) / focal, (j - 0.5 * h) / focal, np.ones_like(i)], -1)
    rays_d = np.matmul(c2w[:3, :3], dirs[:, :, :, None])[..., 0]
    rays_d = rays_d / np.linalg.norm(rays_d, axis=-1, keepdims=True)
    rays_o = np.broadcast_to(c2w[:3, -1], rays_d.shape)
    # C-contiguous on purpose: `broadcast_to(...).reshape(...).astype(...)` keeps the broadcast's zero strides ('K' order) and yields a
    # Fortran-ordered [n,3] array, which torch.from_numpy(...).to(device) preserves
    return np.ascontiguousarray(rays_o.reshape(-1, 3), dtype=np.float32), np.ascontiguousarray(rays_d.reshape(-1, 3), dtype=np.float32)


def metadata_for(n_img, focal=FOCAL):
    m = np.array([0, 0, 0, 0, 0.5, 0.5, focal, focal, 0, 0, 0], np.float32)
    return np.tile(m[None], (n_img, 1))


def ray_batch(n_rays, seed=0, n_img=40):
    """n_rays drawn uniformly from n_img spiral 800x800 views (the reference's shuffled [N_img*H*W] table, hashnerf_dataset.py:41-44)."""
    rng = np.random.default_rng(seed)
    poses = spiral_poses_ngp(n_img)
    img = rng.integers(0, n_img, n_rays)
    px = rng.integers(0, W, n_rays).astype(np.float32) + 0.5
    py = rng.integers(0, H, n_rays).astype(np.float32) + 0.5
    dirs = np.stack([(px - 0.5 * W) / FOCAL, (py - 0.5 * H) / FOCAL, np.ones_like(px)], -1).astype(np.float32)
    R = poses.transpose(0, 2, 1)[img][:, :3, :3]  # [n,3,3]
    rays_d = np.einsum('nij,nj->ni', R, dirs)
    rays_d = (rays_d / np.linalg.norm(rays_d, axis=-1, keepdims=True)).astype(np.float32)
    rays_o = poses[img][:, 3, :].astype(np.float32)
    return np.ascontiguousarray(rays_o), np.ascontiguousarray(rays_d), img.astype(np.int32), poses


def morton3d(x, y, z):
    def expand(v):
        v = v.astype(np.uint32)
        v = (v * np.uint32(0x00010001)) & np.uint32(0xFF0000FF)
        v = (v * np.uint32(0x00000101)) & np.uint32(0x0F00F00F)
        v = (v * np.uint32(0x00000011)) & np.uint32(0xC30C30C3)
        v = (v * np.uint32(0x00000005)) & np.uint32(0x49249249)
        return v
    return expand(x) | (expand(y) << np.uint32(1)) | (expand(z) << np.uint32(2))


def lego_like_density_grid(seed=0, fill_value=0.5, speckle=0.002):
    """float32[8*128^3] density grid in the reference's Morton/cascade layout: level 0 occupied inside a box + sphere shell
    (+ seeded speckle), coarser levels empty (the bitfield update OR-pools them). Values are multiples of 2^-10 so every
    summation order gives the same mean (see tests)."""
    rng = np.random.default_rng(seed)
    g = np.arange(128, dtype=np.float32)
    X, Y, Z = np.meshgrid(g, g, g, indexing='ij')
    c = (np.stack([X, Y, Z], -1) + 0.5) / 128.0 - 0.5
    box = (np.abs(c[..., 0]) < 0.22) & (np.abs(c[..., 1]) < 0.12) & (np.abs(c[..., 2]) < 0.10)
    r = np.linalg.norm(c - np.array([0.0, 0.0, 0.12], np.float32), axis=-1)
    shell = (r > 0.20) & (r < 0.23) & (c[..., 2] > 0.0)
    speck = rng.random(box.shape) < speckle   # floaters; 0 = a converged grid without any
    occ = box | shell | (speck & (np.abs(c).max(-1) < 0.4))
    grid = np.zeros(8 * 128 ** 3, np.float32)
    xs, ys, zs = np.nonzero(occ)
    idx = morton3d(xs, ys, zs)
    grid[idx] = np.float32(fill_value)
    return grid


def bitfield_from_grid_numpy(grid):
    """Plain numpy occupancy bitfield builder used to SEED benchmarks/tests with a scene (not a parity oracle)."""
    lvl0 = grid[:128 ** 3]
    mean = np.float32(np.maximum(lvl0, 0).sum(dtype=np.float64) / 128 ** 3)
    thresh = min(np.float32(0.01), mean)
    bits = (grid > thresh).reshape(-1, 8)
    bf = np.zeros(bits.shape[0], np.uint8)
    for j in range(8):
        bf |= (bits[:, j].astype(np.uint8) << j)
    n = 128 ** 3 // 8
    for level in range(1, 8):
        prev = bf[(level - 1) * n: level * n].reshape(-1, 8)
        b = np.zeros(prev.shape[0], np.uint8)
        for j in range(8):
            b |= ((prev[:, j] > 0).astype(np.uint8) << j)
        i = np.arange(prev.shape[0], dtype=np.uint32)

        def inv(x):
            x = x & np.uint32(0x49249249)
            x = (x | (x >> np.uint32(2))) & np.uint32(0xc30c30c3)
            x = (x | (x >> np.uint32(4))) & np.uint32(0x0f00f00f)
            x = (x | (x >> np.uint32(8))) & np.uint32(0xff0000ff)
            x = (x | (x >> np.uint32(16))) & np.uint32(0x0000ffff)
            return x
        x, y, z = inv(i) + 16, inv(i >> np.uint32(1)) + 16, inv(i >> np.uint32(2)) + 16
        bf[level * n + morton3d(x, y, z)] |= b
    return bf, mean


def ngp_weights(seed=0, n_hash_params=12196240, width=64, dens_hidden=1, color_hidden=2, hash_range=1e-4, mlp_gain=1.0):
    """tcnn default initialisation: hash table U(-1e-4,1e-4); MLP Xavier-uniform per weight matrix."""
    rng = np.random.default_rng(seed)
    table = rng.uniform(-hash_range, hash_range, n_hash_params).astype(np.float32)

    def xavier(out_w, in_w):
        s = mlp_gain * np.sqrt(6.0 / (in_w + out_w))
        return rng.uniform(-s, s, (out_w, in_w)).astype(np.float32).reshape(-1)
    dens = [xavier(width, 32)] + [xavier(width, width) for _ in range(dens_hidden - 1)] + [xavier(16, width)]
    color = [xavier(width, 32)] + [xavier(width, width) for _ in range(color_hidden - 1)] + [xavier(16, width)]
    return table, np.concatenate(dens), np.concatenate(color)

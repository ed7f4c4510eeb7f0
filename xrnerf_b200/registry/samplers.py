"""NGPGridSampler with the reference's kwargs, state (density grid / bitfield / adaptive batch) and call protocol
(/root/reference/xrnerf/models/samplers/ngp_grid_sampler.py:12-284), on the xrnerf_b200.raymarch_cuda kernels."""
import torch
from torch import nn

from .. import raymarch_cuda
from .builder import SAMPLERS


@SAMPLERS.register_module()
class NGPGridSampler(nn.Module):
    def __init__(self, update_grid_freq=16, update_block_size=5000000, n_rays_per_batch=4096, cone_angle_constant=0.00390625, near_distance=0.2, target_batch_size=1 << 18,
                 rgb_activation=2, density_activation=3, max_samples_per_ray=1024):
        super().__init__()
        self.update_grid_freq, self.update_block_size = update_grid_freq, update_block_size
        self.n_rays_per_batch, self.target_batch_size = n_rays_per_batch, target_batch_size
        self.rgb_activation, self.density_activation = rgb_activation, density_activation
        self.density_mlp_padded_density_output_width = 1
        self.density_grid_ema_step = 0
        self.NERF_CASCADES, self.NERF_GRIDSIZE, self.MAX_STEP = 8, 128, 1024
        self.near_distance = 0.05                  # ctor args near_distance / cone_angle_constant are ignored by the reference too (Q5)
        self.cone_angle_constant = 0.00390625
        self.ema_grid_decay = 0.95
        self.NERF_MIN_OPTICAL_THICKNESS = 0.01
        # the reference sizes its coords buffer n_rays*1024 and zero-fills it every call (Q11); same upper bound, but allocated once
        self.max_samples_per_ray = max_samples_per_ray
        self.num_coords_elements = self.n_rays_per_batch * self.max_samples_per_ray
        cells = self.NERF_GRIDSIZE ** 3
        self.density_n_elements = self.NERF_CASCADES * cells
        self.density_grid_tmp = torch.zeros([self.density_n_elements], dtype=torch.float32)
        self.density_grid_mean = torch.zeros([(cells + 127) // 128], dtype=torch.float32)
        self.register_buffer('density_grid_bitfield', torch.zeros([self.density_n_elements // 8], dtype=torch.uint8))
        self.measured_batch_size = torch.zeros((1,), dtype=torch.int32)
        self.iter_n = 0
        self._coords_buf = None

    def set_data(self, alldata, datainfo):
        self.resolutions = [datainfo['H'], datainfo['W']]
        self.transforms = torch.as_tensor(alldata['poses'], dtype=torch.float32)
        self.focal = torch.as_tensor(alldata['focal'], dtype=torch.float32)
        self.aabb_scale, self.aabb_range = alldata['aabb_scale'], alldata['aabb_range']
        self.metadata = torch.as_tensor(alldata['metadata'], dtype=torch.float32)
        self.n_img = self.transforms.shape[0]
        self.max_cascade = 0
        while (1 << self.max_cascade) < self.aabb_scale:
            self.max_cascade += 1

    def set_iter(self, iter_n):
        self.iter_n = iter_n

    def check_device(self, data):
        device = data['rays_o'].device
        self.device = device
        for attr in ['transforms', 'focal', 'metadata', 'density_grid_mean', 'density_grid_bitfield', 'density_grid_tmp', 'measured_batch_size']:
            v = getattr(self, attr)
            if v.device != device:
                setattr(self, attr, v.to(device).contiguous())

    # ---- occupancy grid (ngp_grid_sampler.py:90-174)
    def _gen(self, n, thresh):
        pos = torch.empty((n, 3), dtype=torch.float32, device=self.device)
        idx = torch.empty((n,), dtype=torch.int32, device=self.device)
        raymarch_cuda.generate_grid_samples_nerf_nonuniform_api(self.density_grid, self.density_grid_ema_step, n, self.max_cascade, thresh, self.aabb_range[0], self.aabb_range[1], pos, idx)
        return pos, idx

    def update_density_grid_func(self, n_uniform, n_nonuniform, mlp):
        n_elements = self.density_n_elements
        if not hasattr(self, 'density_grid'):
            self.density_grid = torch.empty([n_elements], dtype=torch.float32, device=self.device)
            raymarch_cuda.mark_untrained_density_grid_api(self.focal, self.transforms, n_elements, self.n_img, self.resolutions[0], self.resolutions[1], self.density_grid)
        parts = [self._gen(n_uniform, -0.01)]
        if n_nonuniform > 0:
            parts.append(self._gen(n_nonuniform, self.NERF_MIN_OPTICAL_THICKNESS))
        positions = torch.cat([p for p, _ in parts]) if len(parts) > 1 else parts[0][0]
        indices = torch.cat([i for _, i in parts]) if len(parts) > 1 else parts[0][1]
        with torch.no_grad():
            density = torch.cat([mlp.run_density(positions[i:i + self.update_block_size]) for i in range(0, positions.shape[0], self.update_block_size)], 0)
        self.density_grid_tmp.zero_()
        raymarch_cuda.splat_grid_samples_nerf_max_nearest_neighbor_api(density, indices, self.density_mlp_padded_density_output_width, positions.shape[0], self.density_grid_tmp)
        raymarch_cuda.ema_grid_samples_nerf_api(self.density_grid_tmp, n_elements, self.ema_grid_decay, self.density_grid)
        self.density_grid_ema_step += 1
        raymarch_cuda.update_bitfield_api(self.density_grid, self.density_grid_mean, self.density_grid_bitfield)

    def update_density_grid(self, mlp):
        M = self.NERF_GRIDSIZE ** 3 * (self.max_cascade + 1)
        if self.iter_n < 256:
            self.update_density_grid_func(M, 0, mlp)
        else:
            self.update_density_grid_func(M // 4, M // 4, mlp)

    # ---- per-batch sampling (ngp_grid_sampler.py:189-266)
    def sample(self, data, mlp, is_test=False):
        is_training = not is_test
        self.check_device(data)
        if is_training and (self.iter_n % self.update_grid_freq == 0 or not hasattr(self, 'density_grid')):
            self.update_density_grid(mlp)
        rays_o, rays_d = data['rays_o'].contiguous().float(), data['rays_d'].contiguous().float()
        img_ids = data['img_ids'].to(torch.int32).contiguous() if 'img_ids' in data else None
        if 'bg_color' in data:
            data['bg_color'] = data['bg_color'].to(torch.float32).contiguous()
        n = rays_o.shape[0]
        cap = max(self.num_coords_elements, n * 64)
        if self._coords_buf is None or self._coords_buf.shape[0] < cap or self._coords_buf.device != self.device:
            self._coords_buf = torch.empty((cap, 7), dtype=torch.float32, device=self.device)
        coords = self._coords_buf
        rays_index = torch.zeros((n, 1), dtype=torch.int32, device=self.device)
        rays_numsteps = torch.zeros((n, 2), dtype=torch.int32, device=self.device)
        counter = torch.zeros((2,), dtype=torch.int32, device=self.device)
        raymarch_cuda.rays_sampler_api(rays_o, rays_d, self.density_grid_bitfield, self.metadata, img_ids, self.transforms, self.aabb_range[0], self.aabb_range[1], self.near_distance,
                                       self.cone_angle_constant, coords, rays_index, rays_numsteps, counter)
        samples = min(int(counter[1].item()), coords.shape[0])   # the reference's one host sync per batch (rays_sampler.py:72)
        coords = coords[:samples]
        if not is_training:
            self.coords, self.rays_numsteps = coords, rays_numsteps
            data['pts'], data['viewdirs'] = coords[..., :3], coords[..., 4:]
            return data
        with torch.no_grad():  # pre-pass on all raw samples (ngp_grid_sampler.py:229-230); only consumed by the compaction's dead transmittance loop
            nerf_outputs = mlp({'pts': coords[..., :3], 'viewdirs': coords[..., 4:]})['raw'].detach().float()
        coords_c = torch.zeros((self.target_batch_size, 7), dtype=torch.float32, device=self.device)
        numsteps_c = torch.zeros_like(rays_numsteps)
        rays_counter = torch.zeros((1,), dtype=torch.int32, device=self.device)
        step_counter = torch.zeros((1,), dtype=torch.int32, device=self.device)
        raymarch_cuda.compacted_coord_api(nerf_outputs, coords, rays_numsteps, torch.ones(3), self.rgb_activation, self.density_activation, self.aabb_range[0], self.aabb_range[1],
                                          coords_c, numsteps_c, rays_counter, step_counter)
        self.measured_batch_size += step_counter
        self.update_batch_rays(is_training)
        self.coords, self.rays_numsteps, self.rays_numsteps_compacted = coords_c, rays_numsteps, numsteps_c
        data['pts'], data['viewdirs'] = coords_c[..., :3], coords_c[..., 4:]
        return data

    def update_batch_rays(self, is_training):
        if is_training and self.iter_n % self.update_grid_freq == (self.update_grid_freq - 1):
            measured = max(self.measured_batch_size.item() / 16, 1)
            rays_per_batch = int(self.n_rays_per_batch * self.target_batch_size / measured)
            self.n_rays_per_batch = int(min(self.div_round_up(int(rays_per_batch), 128) * 128, self.target_batch_size))
            self.measured_batch_size.zero_()

    @staticmethod
    def div_round_up(val, divisor):
        return (val + divisor - 1) // divisor

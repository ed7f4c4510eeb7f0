"""Drop-in mirror of the reference's pybind module `raymarch_cuda`
(/root/reference/extensions/ngp_raymarch/src/pybind_api.cu:6-17, include/pybind_api.h:3-95).

Same 10 function names, same positional arguments, same in-place-output convention (the caller allocates every
tensor), so the reference's thin wrappers (xrnerf/models/samplers/utils/*.py, renders/hashnerf_render.py:98,133,167)
work against it unchanged:  `import xrnerf_b200.raymarch_cuda as raymarch_cuda`.

Differences: launches go to torch's CURRENT stream and never synchronise; errors raise; the
hidden `static pcg32 rng{9121}` of each reference translation unit is module state here (`reset_rng`, `rng_calls`),
advanced once per call exactly like the reference (ray_sampler.cu:198, generate_grid_samples...cu:84); sample bases are
assigned in ray order.
"""
import torch

from . import _C

SEED = 9121
rng_calls = {'ray_sampler': 0, 'generate_grid_samples': 0}
_ws = {}


def reset_rng(ray_sampler=0, generate_grid_samples=0):
    rng_calls['ray_sampler'] = int(ray_sampler)
    rng_calls['generate_grid_samples'] = int(generate_grid_samples)


def _workspace(device, nbytes):
    key = (device.type, device.index)
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf


def generate_grid_samples_nerf_nonuniform_api(density_grid, density_grid_ema_step, n_elements, max_cascade, thresh, aabb0, aabb1,
                                              density_grid_positions_uniform, density_grid_indices_uniform):
    _C.require_cuda(density_grid, density_grid_positions_uniform, density_grid_indices_uniform)
    _C.check(_C.lib.xrb_rm_generate_grid_samples(_C.ptr(density_grid), int(density_grid_ema_step), int(n_elements), int(max_cascade), float(thresh),
                                                 float(aabb0), float(aabb1), SEED, rng_calls['generate_grid_samples'],
                                                 _C.ptr(density_grid_positions_uniform), _C.ptr(density_grid_indices_uniform), _C.stream()),
             'generate_grid_samples_nerf_nonuniform_api')
    rng_calls['generate_grid_samples'] += 1


def mark_untrained_density_grid_api(focal_lengths, transforms, n_elements, n_images, img_resolution0, img_resolution1, density_grid):
    _C.require_cuda(focal_lengths, transforms, density_grid)
    _C.check(_C.lib.xrb_rm_mark_untrained_density_grid(_C.ptr(focal_lengths), _C.ptr(transforms), int(n_elements), int(n_images), int(img_resolution0),
                                                       int(img_resolution1), _C.ptr(density_grid), _C.stream()), 'mark_untrained_density_grid_api')


def splat_grid_samples_nerf_max_nearest_neighbor_api(mlp_out, density_grid_indices, padded_output_width, n_density_grid_samples, density_grid_tmp):
    _C.require_cuda(mlp_out, density_grid_indices, density_grid_tmp)
    _C.check(_C.lib.xrb_rm_splat_grid_samples(_C.ptr(mlp_out), _C.ptr(density_grid_indices), int(padded_output_width), int(n_density_grid_samples),
                                              _C.ptr(density_grid_tmp), _C.stream()), 'splat_grid_samples_nerf_max_nearest_neighbor_api')


def ema_grid_samples_nerf_api(density_grid_tmp, n_elements, decay, density_grid):
    _C.require_cuda(density_grid_tmp, density_grid)
    _C.check(_C.lib.xrb_rm_ema_grid_samples(_C.ptr(density_grid_tmp), int(n_elements), float(decay), _C.ptr(density_grid), _C.stream()),
             'ema_grid_samples_nerf_api')


def update_bitfield_api(density_grid, density_grid_mean, density_grid_bitfield):
    _C.require_cuda(density_grid, density_grid_mean, density_grid_bitfield)
    # scratch for the partial sums of the fixed-order mean: one small buffer per (device, stream), owned by this module
    key = ('bitfield', density_grid.device.index, torch.cuda.current_stream().cuda_stream)
    ws = _ws.get(key)
    if ws is None:
        ws = _ws[key] = torch.empty(_C.lib.xrb_rm_update_bitfield_workspace(), dtype=torch.uint8, device=density_grid.device)
    _C.check(_C.lib.xrb_rm_update_bitfield(_C.f32(density_grid), _C.f32(density_grid_mean), _C.u8(density_grid_bitfield), _C.ptr(ws), _C.stream()), 'update_bitfield_api')


def rays_sampler_api(rays_o, rays_d, density_grid_bitfield, metadata, imgs_id, xforms, aabb0, aabb1, near_distance, cone_angle_constant, coords_out,
                     rays_index, rays_numsteps, ray_numstep_counter):
    _C.require_cuda(rays_o, rays_d, density_grid_bitfield, coords_out, rays_index, rays_numsteps, ray_numstep_counter)
    n_rays = rays_o.shape[0]
    ws = _workspace(rays_o.device, _C.lib.xrb_rm_rays_sampler_workspace(n_rays))
    _C.check(_C.lib.xrb_rm_rays_sampler(_C.ptr(rays_o), _C.ptr(rays_d), _C.ptr(density_grid_bitfield), _C.ptr(metadata), _C.ptr(imgs_id), _C.ptr(xforms),
                                        n_rays, coords_out.shape[0], float(aabb0), float(aabb1), float(near_distance), float(cone_angle_constant), SEED,
                                        rng_calls['ray_sampler'], _C.ptr(coords_out), _C.ptr(rays_index), _C.ptr(rays_numsteps), _C.ptr(ray_numstep_counter),
                                        _C.ptr(ws), _C.stream()), 'rays_sampler_api')
    rng_calls['ray_sampler'] += 1


def compacted_coord_api(network_output, coords_in, rays_numsteps, bg_color_in, rgb_activation_i, density_activation_i, aabb0, aabb1, coords_out,
                        rays_numsteps_compacted, compacted_rays_counter, compacted_numstep_counter):
    _C.require_cuda(coords_in, rays_numsteps, coords_out, rays_numsteps_compacted, compacted_rays_counter, compacted_numstep_counter)
    n_rays = rays_numsteps.shape[0]
    ws = _workspace(coords_in.device, _C.lib.xrb_rm_compacted_coord_workspace(n_rays))
    _C.check(_C.lib.xrb_rm_compacted_coord(_C.ptr(network_output), _C.ptr(coords_in), _C.ptr(rays_numsteps), n_rays, coords_out.shape[0], _C.ptr(coords_out),
                                           _C.ptr(rays_numsteps_compacted), _C.ptr(compacted_rays_counter), _C.ptr(compacted_numstep_counter), _C.ptr(ws),
                                           _C.stream()), 'compacted_coord_api')


def calc_rgb_forward_api(network_output, coords_in, rays_numsteps, rays_numsteps_compacted, training_background_color, rgb_activation_i,
                         density_activation_i, aabb0, aabb1, rgb_output):
    _C.require_cuda(network_output, coords_in, rays_numsteps, rays_numsteps_compacted, training_background_color, rgb_output)
    _C.check(_C.lib.xrb_rm_calc_rgb_forward(_C.ptr(network_output), _C.ptr(coords_in), _C.ptr(rays_numsteps), _C.ptr(rays_numsteps_compacted),
                                            _C.ptr(training_background_color), rays_numsteps.shape[0], int(rgb_activation_i), int(density_activation_i),
                                            _C.ptr(rgb_output), _C.stream()), 'calc_rgb_forward_api')


def calc_rgb_backward_api(network_output, rays_numsteps_compacted, coords_in, grad_x, rgb_output, density_grid_mean, rgb_activation_i,
                          density_activation_i, aabb0, aabb1, dloss_doutput):
    _C.require_cuda(network_output, rays_numsteps_compacted, coords_in, grad_x, rgb_output, density_grid_mean, dloss_doutput)
    grad_x = grad_x.contiguous()
    _C.check(_C.lib.xrb_rm_calc_rgb_backward(_C.ptr(network_output), _C.ptr(rays_numsteps_compacted), _C.ptr(coords_in), _C.ptr(grad_x), _C.ptr(rgb_output),
                                             _C.ptr(density_grid_mean), rays_numsteps_compacted.shape[0], int(rgb_activation_i), int(density_activation_i),
                                             _C.ptr(dloss_doutput), _C.stream()), 'calc_rgb_backward_api')


def calc_rgb_influence_api(network_output, coords_in, rays_numsteps, bg_color_cpu, rgb_activation_i, density_activation_i, aabb0, aabb1, rgb_output,
                           alpha_output):
    _C.require_cuda(network_output, coords_in, rays_numsteps, rgb_output, alpha_output)
    bg = _C.float3(bg_color_cpu.detach().cpu().reshape(-1).tolist())
    _C.check(_C.lib.xrb_rm_calc_rgb_inference(_C.ptr(network_output), _C.ptr(coords_in), _C.ptr(rays_numsteps), bg, rays_numsteps.shape[0], int(rgb_activation_i),
                                              int(density_activation_i), _C.ptr(rgb_output), _C.ptr(alpha_output), _C.stream()), 'calc_rgb_influence_api')

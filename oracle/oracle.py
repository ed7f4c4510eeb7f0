"""TEST INFRASTRUCTURE ONLY — ctypes front-end for the CPU oracles.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product package
(``xrnerf_b200``) never does and fails loudly without its CUDA library instead.

Two back-ends with the same call shapes (numpy in, numpy out):

* ``Port``  — ``oracle/liboracle.so``: the plain-C restatement (``ngp_oracle.c``,
  ``tcnn_oracle.c``), each function citing the reference file:line it follows.
* ``Ref``   — ``oracle/_ref/libraymarch_ref.so``: the reference's own
  ``extensions/ngp_raymarch`` kernels and ``*_api`` wrappers compiled for CPU
  (``oracle/Makefile``, ``oracle/ref_shim/``). Covers the 10 ``raymarch_cuda``
  ops; there is no reference source for tcnn (hash grid / SH / MLP), so those
  exist only in ``Port`` ("parity unpinned", see tcnn_oracle.c header).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
GRID_CELLS = 128 ** 3
NERF_CASCADES = 8
SQRT3 = np.float32(1.73205080757)
MIN_CONE_STEPSIZE = np.float32(SQRT3 / np.float32(1024))

f32p = np.ctypeslib.ndpointer(np.float32, flags='C_CONTIGUOUS')
i32p = np.ctypeslib.ndpointer(np.int32, flags='C_CONTIGUOUS')
u8p = np.ctypeslib.ndpointer(np.uint8, flags='C_CONTIGUOUS')
u32p = np.ctypeslib.ndpointer(np.uint32, flags='C_CONTIGUOUS')


def build(force=False):
    """Compile liboracle.so (and _ref/ when /root/reference is present). Building the checker is not using it."""
    if force or not os.path.exists(os.path.join(_HERE, 'liboracle.so')):
        subprocess.check_call(['make', '-C', _HERE, 'liboracle.so'], stdout=subprocess.DEVNULL)
    if os.path.isdir('/root/reference/extensions/ngp_raymarch/src') and (
            force or not os.path.exists(os.path.join(_HERE, '_ref', 'libraymarch_ref.so'))):
        subprocess.check_call(['make', '-C', _HERE, 'ref'], stdout=subprocess.DEVNULL)


def have_ref():
    return os.path.exists(os.path.join(_HERE, '_ref', 'libraymarch_ref.so'))


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class _Base:
    """Shared call shapes. Sub-classes bind `self.L` and the symbol prefix."""

    def mark_untrained(self, focal, xforms, n_images, res, grid=None):
        n = GRID_CELLS * NERF_CASCADES
        grid = np.zeros(n, np.float32) if grid is None else _c(grid, np.float32).copy()
        self._mark(_c(focal, np.float32), _c(xforms, np.float32), n, int(n_images), int(res[0]), int(res[1]), grid)
        return grid

    def ema(self, grid_tmp, grid, decay=0.95):
        grid = _c(grid, np.float32).copy()
        self._ema(_c(grid_tmp, np.float32), grid.size, float(decay), grid)
        return grid

    def splat(self, mlp_out, indices, grid_tmp):
        grid_tmp = _c(grid_tmp, np.float32).copy()
        mlp_out = _c(mlp_out, np.float32).reshape(len(indices), -1)
        self._splat(mlp_out, _c(indices, np.int32), mlp_out.shape[1], len(indices), grid_tmp)
        return grid_tmp

    def update_bitfield(self, grid):
        mean = np.zeros(1, np.float32)
        bitfield = np.zeros(GRID_CELLS * NERF_CASCADES // 8, np.uint8)
        self._bitfield(_c(grid, np.float32), mean, bitfield)
        return bitfield, mean

    def calc_rgb_forward(self, raw, coords, numsteps, numsteps_c, bg, rgb_act=2, dens_act=3):
        n_rays = numsteps.shape[0]
        out = np.zeros((n_rays, 3), np.float32)
        self._fwd(*self._fwd_args(_c(raw, np.float32), _c(coords, np.float32), _c(numsteps, np.int32), _c(numsteps_c, np.int32),
                                  _c(bg, np.float32), n_rays, rgb_act, dens_act, out))
        return out

    def calc_rgb_backward(self, raw, numsteps_c, coords, grad_rgb, rgb, grid_mean, rgb_act=2, dens_act=3):
        n_rays = numsteps_c.shape[0]
        out = np.zeros_like(_c(raw, np.float32))
        self._bwd(*self._bwd_args(_c(raw, np.float32), _c(numsteps_c, np.int32), _c(coords, np.float32), _c(grad_rgb, np.float32),
                                  _c(rgb, np.float32), _c(grid_mean, np.float32), n_rays, rgb_act, dens_act, out))
        return out

    def calc_rgb_inference(self, raw, coords, numsteps, bg3, rgb_act=2, dens_act=3):
        n_rays = numsteps.shape[0]
        rgb = np.zeros((n_rays, 3), np.float32)
        alpha = np.zeros((n_rays, 1), np.float32)
        self._inf(*self._inf_args(_c(raw, np.float32), _c(coords, np.float32), _c(numsteps, np.int32), _c(bg3, np.float32), n_rays,
                                  rgb_act, dens_act, rgb, alpha))
        return rgb, alpha


class Port(_Base):
    kind = 'port'

    def __init__(self):
        build()
        L = self.L = C.CDLL(os.path.join(_HERE, 'liboracle.so'))
        L.oracle_rays_sampler.argtypes = [f32p, f32p, u8p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint64,
                                          C.c_int64, f32p, i32p, i32p, i32p]
        L.oracle_compacted_coord.argtypes = [f32p, i32p, C.c_int, C.c_int, f32p, i32p, i32p, i32p]
        L.oracle_calc_rgb_forward.argtypes = [f32p, f32p, i32p, i32p, f32p, C.c_int, C.c_int, C.c_int, f32p]
        L.oracle_calc_rgb_backward.argtypes = [f32p, i32p, f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, f32p]
        L.oracle_calc_rgb_inference.argtypes = [f32p, f32p, i32p, f32p, C.c_int, C.c_int, C.c_int, f32p, f32p]
        L.oracle_mark_untrained.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
        L.oracle_generate_grid_samples.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_uint64, C.c_int64,
                                                   f32p, i32p]
        L.oracle_splat.argtypes = [f32p, i32p, C.c_int, C.c_int, f32p]
        L.oracle_ema.argtypes = [f32p, C.c_int, C.c_float, f32p]
        L.oracle_update_bitfield.argtypes = [f32p, f32p, u8p]
        L.oracle_pcg32_floats.argtypes = [C.c_uint64, C.c_int64, C.c_uint64, C.c_int, f32p]
        L.oracle_hashgrid_layout.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, u32p, f32p, u32p]
        L.oracle_hashgrid_layout.restype = C.c_int64
        L.oracle_hashgrid_forward.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, f32p]
        L.oracle_hashgrid_backward.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, f32p]
        L.oracle_sh4_forward.argtypes = [f32p, C.c_int, f32p]
        L.oracle_mlp_forward.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
        L.oracle_ngp_mlp_forward.argtypes = [f32p, f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                             C.c_int, C.c_int, f32p]
        L.oracle_ngp_density_forward.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                                 f32p]
        L.oracle_ngp_mlp_backward.argtypes = [f32p, f32p, f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                              C.c_int, C.c_int, C.c_int, f32p, f32p, f32p]
        self._mark, self._ema, self._splat, self._bitfield = L.oracle_mark_untrained, L.oracle_ema, L.oracle_splat, L.oracle_update_bitfield
        self._fwd, self._bwd, self._inf = L.oracle_calc_rgb_forward, L.oracle_calc_rgb_backward, L.oracle_calc_rgb_inference

    # the port needs no sample counts / aabb for compositing
    def _fwd_args(self, raw, coords, ns, nsc, bg, n_rays, ra, da, out):
        return raw, coords, ns, nsc, bg, n_rays, ra, da, out

    def _bwd_args(self, raw, nsc, coords, g, rgb, mean, n_rays, ra, da, out):
        return raw, nsc, coords, g, rgb, mean, n_rays, ra, da, out

    def _inf_args(self, raw, coords, ns, bg3, n_rays, ra, da, rgb, alpha):
        return raw, coords, ns, bg3, n_rays, ra, da, rgb, alpha

    def rays_sampler(self, rays_o, rays_d, bitfield, max_samples, aabb=(0., 1.), near=0.05, cone=1. / 256, seed=9121, n_prior_calls=0,
                     metadata=None, img_ids=None, xforms=None):
        n = rays_o.shape[0]
        coords = np.zeros((max_samples, 7), np.float32)
        ridx = np.zeros((n, 1), np.int32)
        ns = np.zeros((n, 2), np.int32)
        cnt = np.zeros(2, np.int32)
        self.L.oracle_rays_sampler(_c(rays_o, np.float32), _c(rays_d, np.float32), _c(bitfield, np.uint8), n, int(max_samples), aabb[0], aabb[1],
                                   near, cone, seed, n_prior_calls, coords, ridx, ns, cnt)
        return coords, ridx, ns, cnt

    def compacted_coord(self, raw, coords_in, numsteps, max_compacted, **_):
        n = numsteps.shape[0]
        out = np.zeros((max_compacted, 7), np.float32)
        nsc = np.zeros((n, 2), np.int32)
        rc = np.zeros(1, np.int32)
        sc = np.zeros(1, np.int32)
        self.L.oracle_compacted_coord(_c(coords_in, np.float32), _c(numsteps, np.int32), n, int(max_compacted), out, nsc, rc, sc)
        return out, nsc, rc, sc

    def generate_grid_samples(self, grid, step, n_elements, max_cascade, thresh, aabb=(0., 1.), seed=9121, n_prior_calls=0):
        pos = np.zeros((n_elements, 3), np.float32)
        idx = np.zeros(n_elements, np.int32)
        self.L.oracle_generate_grid_samples(_c(grid, np.float32), int(step), int(n_elements), int(max_cascade), thresh, aabb[0], aabb[1], seed,
                                            n_prior_calls, pos, idx)
        return pos, idx

    def pcg32_floats(self, n, advance=0, seed=9121, n_prior_calls=0):
        out = np.zeros(n, np.float32)
        self.L.oracle_pcg32_floats(seed, n_prior_calls, advance, n, out)
        return out

    # ---- tcnn-shaped arithmetic (port only)
    def hashgrid_layout(self, n_levels=16, n_feat=2, log2_hashmap=19, base_res=16, per_level_scale=1.3819128800):
        off = np.zeros(n_levels + 1, np.uint32)
        sc = np.zeros(n_levels, np.float32)
        res = np.zeros(n_levels, np.uint32)
        n = self.L.oracle_hashgrid_layout(n_levels, n_feat, log2_hashmap, base_res, per_level_scale, off, sc, res)
        return int(n), off, sc, res

    def hashgrid_forward(self, table, x, n_levels=16, n_feat=2, log2_hashmap=19, base_res=16, per_level_scale=1.3819128800):
        n = x.shape[0]
        enc = np.zeros((n, n_levels * n_feat), np.float32)
        self.L.oracle_hashgrid_forward(_c(table, np.float32), _c(x, np.float32), n, n_levels, n_feat, log2_hashmap, base_res, per_level_scale, enc)
        return enc

    def hashgrid_backward(self, x, denc, n_params, n_levels=16, n_feat=2, log2_hashmap=19, base_res=16, per_level_scale=1.3819128800):
        dt = np.zeros(n_params, np.float32)
        self.L.oracle_hashgrid_backward(_c(x, np.float32), _c(denc, np.float32), x.shape[0], n_levels, n_feat, log2_hashmap, base_res,
                                        per_level_scale, dt)
        return dt

    def sh4(self, dirs01):
        out = np.zeros((dirs01.shape[0], 16), np.float32)
        self.L.oracle_sh4_forward(_c(dirs01, np.float32), dirs01.shape[0], out)
        return out

    def mlp_forward(self, params, x, width, n_hidden, out_pad=16):
        n, in_w = x.shape
        y = np.zeros((n, out_pad), np.float32)
        self.L.oracle_mlp_forward(_c(params, np.float32), _c(x, np.float32), n, in_w, width, n_hidden, out_pad, y)
        return y

    def ngp_mlp_forward(self, table, dens_params, color_params, pts, dirs, n_levels=16, n_feat=2, log2_hashmap=19, base_res=16,
                        per_level_scale=1.3819128800, width=64, dens_hidden=1, color_hidden=2):
        n = pts.shape[0]
        raw = np.zeros((n, 4), np.float32)
        self.L.oracle_ngp_mlp_forward(_c(table, np.float32), _c(dens_params, np.float32), _c(color_params, np.float32), _c(pts, np.float32),
                                      _c(dirs, np.float32), n, n_levels, n_feat, log2_hashmap, base_res, per_level_scale, width, dens_hidden,
                                      color_hidden, raw)
        return raw

    def ngp_density_forward(self, table, dens_params, pts, n_levels=16, n_feat=2, log2_hashmap=19, base_res=16,
                            per_level_scale=1.3819128800, width=64, dens_hidden=1):
        n = pts.shape[0]
        out = np.zeros(n, np.float32)
        self.L.oracle_ngp_density_forward(_c(table, np.float32), _c(dens_params, np.float32), _c(pts, np.float32), n, n_levels, n_feat,
                                          log2_hashmap, base_res, per_level_scale, width, dens_hidden, out)
        return out

    def ngp_mlp_backward(self, table, dens_params, color_params, pts, dirs, draw, n_levels=16, n_feat=2, log2_hashmap=19, base_res=16,
                         per_level_scale=1.3819128800, width=64, dens_hidden=1, color_hidden=2):
        dt = np.zeros_like(_c(table, np.float32))
        dd = np.zeros_like(_c(dens_params, np.float32))
        dc = np.zeros_like(_c(color_params, np.float32))
        self.L.oracle_ngp_mlp_backward(_c(table, np.float32), _c(dens_params, np.float32), _c(color_params, np.float32), _c(pts, np.float32),
                                       _c(dirs, np.float32), _c(draw, np.float32), pts.shape[0], n_levels, n_feat, log2_hashmap, base_res,
                                       per_level_scale, width, dens_hidden, color_hidden, dt, dd, dc)
        return dt, dd, dc


class Ref(_Base):
    """The reference's own kernels on CPU. REF_SERIAL=1 (default) runs threads in index order, which turns the
    reference's atomic-arrival layout into the ray-order layout; REF_SERIAL=0 uses all host cores (timing)."""
    kind = 'reference'

    def __init__(self, serial=True):
        build()
        if not have_ref():
            raise RuntimeError('oracle/_ref/libraymarch_ref.so missing: build it in the container that has /root/reference')
        os.environ['REF_SERIAL'] = '1' if serial else '0'
        L = self.L = C.CDLL(os.path.join(_HERE, '_ref', 'libraymarch_ref.so'))
        L.ref_rays_sampler.argtypes = [f32p, f32p, u8p, f32p, i32p, f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                                       C.c_float, f32p, i32p, i32p, i32p]
        L.ref_compacted_coord.argtypes = [f32p, f32p, i32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, f32p,
                                          i32p, i32p, i32p]
        L.ref_calc_rgb_forward.argtypes = [f32p, f32p, i32p, i32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, f32p]
        L.ref_calc_rgb_backward.argtypes = [f32p, i32p, f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, f32p]
        L.ref_calc_rgb_inference.argtypes = [f32p, f32p, i32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, f32p, f32p]
        L.ref_generate_grid_samples.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, f32p, i32p]
        L.ref_mark_untrained.argtypes = [f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int, f32p]
        L.ref_splat.argtypes = [f32p, i32p, C.c_int, C.c_int, f32p]
        L.ref_ema.argtypes = [f32p, C.c_int, C.c_float, f32p]
        L.ref_update_bitfield.argtypes = [f32p, f32p, u8p]
        L.ref_ray_sampler_rng_reset.argtypes = [C.c_int64]
        L.ref_generate_grid_samples_rng_reset.argtypes = [C.c_int64]
        self._mark, self._ema, self._splat, self._bitfield = L.ref_mark_untrained, L.ref_ema, L.ref_splat, L.ref_update_bitfield
        self._fwd, self._bwd, self._inf = L.ref_calc_rgb_forward, L.ref_calc_rgb_backward, L.ref_calc_rgb_inference
        self.aabb = (0., 1.)

    def _fwd_args(self, raw, coords, ns, nsc, bg, n_rays, ra, da, out):
        return raw, coords, ns, nsc, bg, raw.shape[0], n_rays, ra, da, self.aabb[0], self.aabb[1], out

    def _bwd_args(self, raw, nsc, coords, g, rgb, mean, n_rays, ra, da, out):
        return raw, nsc, coords, g, rgb, mean, raw.shape[0], n_rays, ra, da, self.aabb[0], self.aabb[1], out

    def _inf_args(self, raw, coords, ns, bg3, n_rays, ra, da, rgb, alpha):
        return raw, coords, ns, bg3, raw.shape[0], n_rays, ra, da, self.aabb[0], self.aabb[1], rgb, alpha

    def rays_sampler(self, rays_o, rays_d, bitfield, max_samples, aabb=(0., 1.), near=0.05, cone=1. / 256, seed=9121, n_prior_calls=0,
                     metadata=None, img_ids=None, xforms=None):
        assert seed == 9121, 'the reference hard-codes seed 9121 (raymarch_shared.h:38)'
        n = rays_o.shape[0]
        if metadata is None:
            metadata = np.tile(np.array([0, 0, 0, 0, .5, .5, 1111., 1111., 0, 0, 0], np.float32), (1, 1))
            xforms = np.zeros((1, 4, 3), np.float32)
            img_ids = np.zeros(n, np.int32)
        coords = np.zeros((max_samples, 7), np.float32)
        ridx = np.zeros((n, 1), np.int32)
        ns = np.zeros((n, 2), np.int32)
        cnt = np.zeros(2, np.int32)
        self.L.ref_ray_sampler_rng_reset(n_prior_calls)
        self.L.ref_rays_sampler(_c(rays_o, np.float32), _c(rays_d, np.float32), _c(bitfield, np.uint8), _c(metadata, np.float32),
                                _c(img_ids, np.int32).reshape(-1), _c(xforms, np.float32), n, metadata.shape[0], int(max_samples), aabb[0], aabb[1],
                                near, cone, coords, ridx, ns, cnt)
        return coords, ridx, ns, cnt

    def compacted_coord(self, raw, coords_in, numsteps, max_compacted, rgb_act=2, dens_act=3):
        n = numsteps.shape[0]
        out = np.zeros((max_compacted, 7), np.float32)
        nsc = np.zeros((n, 2), np.int32)
        rc = np.zeros(1, np.int32)
        sc = np.zeros(1, np.int32)
        raw = _c(raw, np.float32)
        self.L.ref_compacted_coord(raw, _c(coords_in, np.float32), _c(numsteps, np.int32), np.ones(3, np.float32), raw.shape[0], n,
                                   int(max_compacted), rgb_act, dens_act, self.aabb[0], self.aabb[1], out, nsc, rc, sc)
        return out, nsc, rc, sc

    def generate_grid_samples(self, grid, step, n_elements, max_cascade, thresh, aabb=(0., 1.), seed=9121, n_prior_calls=0):
        assert seed == 9121
        pos = np.zeros((n_elements, 3), np.float32)
        idx = np.zeros(n_elements, np.int32)
        self.L.ref_generate_grid_samples_rng_reset(n_prior_calls)
        self.L.ref_generate_grid_samples(_c(grid, np.float32), int(step), int(n_elements), int(max_cascade), thresh, aabb[0], aabb[1], pos, idx)
        return pos, idx

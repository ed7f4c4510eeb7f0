#!/usr/bin/env python
"""bench.py — rays/sec of the Instant-NGP hot path (BASELINE.json configs[1]: lego-shaped scene, 65 536 rays per batch).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path (ray march -> hash/SH encode -> tiny MLPs -> alpha compositing) over one batch of
65 536 synthetic Blender-shaped rays (800x800 spiral views, seeded lego-like occupancy grid, tcnn-default random weights).
  value      device-timed rays/s with the ray batches already resident in HBM (one CUDA-event pair around the K steps, max over
             ranks); inputs larger than L2: 96 distinct ray batches (151 MB) are cycled, nothing is flushed (the 24.4 MB fp16 hash
             table and its cell image are meant to stay L2-resident, as in production)
  e2e        the same metric through the public API with HOST (pinned) ray buffers: H2D of the rays and D2H of rgb+alpha
             inside the timed region
  roofline   dominant kernel (ngp_field_tc_kernel: hash gather + tcgen05 MLPs) timed live with CUDA events recorded inside the
             step; algorithmic bytes = 512 B gathered per sample (16 levels x 8 corners x 2 x fp16) + 28 B coords in
             + 16 B raw out per sample
  cpu_baseline / --impl reference   the reference's own ngp_raymarch kernels compiled for CPU (oracle/_ref) for march and
             compositing + the C restatement of tcnn (oracle port) for the field, all host cores, on a bounded ray sample.
Multi-GPU (torchrun): rays shard over ranks with no data-path collective (weak scaling: 65 536 rays per rank per step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RAYS = 65536
BUDGET = 64          # max samples per ray the workspace is sized for (reference: 1024; measured mean is ~11-30)
N_BATCHES = 96       # distinct ray batches cycled through: 96 x 1.5 MiB of rays = 151 MB > 126 MB L2 (inputs larger than L2)
METRIC = 'rays/sec (inference render, Instant-NGP lego-shaped 800x800, 65536 rays/batch)'
WORKLOAD = 'instant-ngp lego-like synthetic (configs[1]): 65536 rays/batch from 40 spiral 800x800 views, occupancy-grid march + hash(16x2,T=2^19) + SH4 + MLP(64;1+2 hidden) + composite, forward'
BYTES_PER_SAMPLE = 512 + 28 + 16


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            p = json.load(fh)
        return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = 'index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-lms', '200', '-i', str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith('active')})
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons, 'samples': len(sm)}


def make_scene(rank, n_batches=N_BATCHES):
    from xrnerf_b200 import synth
    grid = synth.lego_like_density_grid(0)
    bf, _ = synth.bitfield_from_grid_numpy(grid)
    batches = [synth.ray_batch(N_RAYS, seed=1000 * rank + b)[:2] for b in range(n_batches)]
    table, dens, color = synth.ngp_weights(seed=0)
    return bf, batches, (table, dens, color)


# ------------------------------------------------------------------------------------------- CPU arm (oracle; timing only)
def cpu_reference_rate(sample_rays, reps=1):
    """rays/s of the reference path on the host cores: reference march/composite kernels compiled for CPU (oracle/_ref, all
    cores) + C restatement of the tcnn field (oracle port, OpenMP)."""
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core (libgomp reads the variable when it is loaded)
    os.environ['OMP_NUM_THREADS'] = str(os.cpu_count() or 1)
    from oracle import oracle as O
    O.build()
    port = O.Port()
    use_ref = O.have_ref()
    ref = O.Ref(serial=False) if use_ref else None
    bf, batches, (table, dens, color) = make_scene(0, 1)
    o, d = batches[0][0][:sample_rays], batches[0][1][:sample_rays]
    m = ref or port
    best = None
    n_samples = 0
    for _ in range(reps):
        t0 = time.perf_counter()
        c, _, ns, cnt = m.rays_sampler(o, d, bf, sample_rays * BUDGET)
        if use_ref:  # parallel run: atomic-order layout, still (count, base) consistent
            pass
        coords = c[:cnt[1]]
        raw = port.ngp_mlp_forward(table, dens, color, np.ascontiguousarray(coords[:, :3]), np.ascontiguousarray(coords[:, 4:]))
        rgb_cpu, alpha_cpu = m.calc_rgb_inference(raw, coords, ns, np.zeros(3, np.float32))
        dt = time.perf_counter() - t0
        cpu_reference_rate.last = (np.asarray(rgb_cpu).copy(), np.asarray(alpha_cpu).copy(), np.asarray(ns)[:, 0].copy())
        best = dt if best is None else min(best, dt)
        n_samples = int(cnt[1])
    cores = os.cpu_count() or 1
    kind = 'reference' if use_ref else 'port'
    sample = (f'{sample_rays} rays of batch 0 ({n_samples} samples): march+composite = reference ngp_raymarch kernels compiled for CPU '
              f'(oracle/_ref, OpenMP), field = C restatement of tcnn (oracle port, OpenMP); best of {reps}') if use_ref else \
             f'{sample_rays} rays of batch 0 ({n_samples} samples), plain-C oracle port, OpenMP; best of {reps}'
    return sample_rays / best, cores, kind, sample, best


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    rates = []
    t_start = time.time()
    sample_rays = 8192
    for _ in range(args.warmup):
        cpu_reference_rate(1024)
    per = []
    for _ in range(args.steps):
        r, cores, kind, sample, dt = cpu_reference_rate(sample_rays)
        rates.append(r); per.append(dt)
        if time.time() - t_start > 240:
            break
    value = float(np.median(rates))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'rays/s', 'n_gpus': args.gpus, 'steps': len(rates), 'warmup': args.warmup,
        'ms_per_step': float(np.median(per) * 1e3), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f16', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'note': f'each step = a bounded sample of {sample_rays} rays of the 65536-ray batch on the host cores'},
        'cpu_baseline': {'value': value, 'unit': 'rays/s', 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from xrnerf_b200 import _C
    from xrnerf_b200.ngp import NgpField, NgpRenderer

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; xrnerf_b200 has no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)

    bf_np, batches, (table, dens, color) = make_scene(rank)
    bf = torch.from_numpy(bf_np).to(dev)
    field = NgpField().to(dev)
    with torch.no_grad():
        field.hash_params.copy_(torch.from_numpy(table).to(dev)); field.density_params.copy_(torch.from_numpy(dens).to(dev)); field.color_params.copy_(torch.from_numpy(color).to(dev))
    P = max(1, args.pipeline)   # batches in flight: step i runs on stream i % P with its own workspace, so the (latency-bound) march of
    renderers = [NgpRenderer(field, samples_per_ray_budget=BUDGET) for _ in range(P)]   # batch i+1 overlaps the field kernel of batch i
    streams = [torch.cuda.Stream(device=dev) for _ in range(P)]
    dev_batches = [(torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)) for (o, d) in batches]
    host_batches = [(torch.from_numpy(o).pin_memory(), torch.from_numpy(d).pin_memory()) for (o, d) in batches[:8]]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    main = torch.cuda.current_stream()

    def fork():
        e = torch.cuda.Event(); e.record(main)
        for st in streams:
            st.wait_event(e)

    def join():
        for st in streams:
            e = torch.cuda.Event(); e.record(st); main.wait_event(e)

    # ---- device-resident arm
    K, W = args.steps, args.warmup
    WU = max(W, 2 * P)      # untimed warm-up steps actually run: at least two per stream, so that every stream's renderer has allocated its workspace / outputs
    evf = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    fork()
    for i in range(WU):
        with torch.cuda.stream(streams[i % P]):
            renderers[i % P].render(*dev_batches[i % N_BATCHES], bf)
    join()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    counters_log = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    fork()
    for i in range(K):
        st = streams[i % P]
        with torch.cuda.stream(st):
            for e in evf[i]:
                e.record(st)                          # materialise the cudaEvent_t handles
            out = renderers[i % P].render(*dev_batches[(W + i) % N_BATCHES], bf, profile_events=evf[i])
            counters_log.append(out[3].clone())
    join()
    e1.record(main)
    barrier()
    field_ms = [a.elapsed_time(b) for a, b in evf]
    samples = [int(c[1].item()) for c in counters_log]
    total_ms = float(e0.elapsed_time(e1))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    clk = clocks.stop() if rank == 0 else None

    # ---- single-launch arm: the same K steps through xrb_ngp_render_fused (march + encode + tcgen05 MLPs + composite in ONE kernel per batch)
    evk = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    fork()
    for i in range(WU):
        with torch.cuda.stream(streams[i % P]):
            renderers[i % P].render_fused(*dev_batches[i % N_BATCHES], bf)
    join()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(main)
    fork()
    ns_log = []
    kev = not os.environ.get('XRB_BENCH_NO_KEVENTS')
    th0 = time.perf_counter()
    for i in range(K):
        st = streams[i % P]
        with torch.cuda.stream(st):
            if kev:
                evk[i][0].record(st)
            outf = renderers[i % P].render_fused(*dev_batches[(W + i) % N_BATCHES], bf)
            if kev:
                evk[i][1].record(st)
            if i < P:
                ns_log.append(outf[2])   # per-ray sample counts of this batch; reduced AFTER the timed region (a first-use torch reduction costs ~20 ms of lazy module loading)
    th1 = time.perf_counter()
    join()
    f1.record(main)
    barrier()
    if os.environ.get('XRB_BENCH_DEBUG') and kev:
        gaps = [evk[i][1].elapsed_time(evk[i + 1][0]) for i in range(K - 1)] if P == 1 else []
        kms = [a.elapsed_time(b) for a, b in evk]
        print('[bench] kernel ms: first 8', [round(x, 3) for x in kms[:8]], 'p50', round(float(np.median(kms)), 3), '| gaps ms first 8', [round(x, 3) for x in gaps[:8]], 'p50', round(float(np.median(gaps)), 3) if gaps else None, 'max', round(max(gaps), 3) if gaps else None, file=sys.stderr, flush=True)
    if os.environ.get('XRB_BENCH_DEBUG'):
        print(f'[bench] fused arm: host issue {1e3 * (th1 - th0) / K:.3f} ms/step, total wall {1e3 * (time.perf_counter() - th0) / K:.3f} ms/step', file=sys.stderr, flush=True)
    fused_total = torch.tensor([float(f0.elapsed_time(f1))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(fused_total, op=dist.ReduceOp.MAX)
    fused_total_ms = float(fused_total.item())
    fused_kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in evk])) if kev else fused_total_ms / K
    fused_samples = float(np.mean([float(x.sum().item()) for x in ns_log]))
    use_fused = args.path == 'fused' or (args.path == 'auto' and fused_total_ms < total_ms_max)

    # ---- end-to-end arm: host (pinned) rays in, rgb+alpha out, every step, through the public API; P batches in flight
    slots = [dict(rgb_h=torch.empty((N_RAYS, 3), dtype=torch.float32).pin_memory(), alpha_h=torch.empty((N_RAYS, 1), dtype=torch.float32).pin_memory(),
                  o_d=torch.empty((N_RAYS, 3), dtype=torch.float32, device=dev), d_d=torch.empty((N_RAYS, 3), dtype=torch.float32, device=dev), done=None) for _ in range(P)]

    def e2e_step(i):
        sl, st = slots[i % P], streams[i % P]
        if sl['done'] is not None:
            sl['done'].synchronize()                 # the caller consumes the pixels of the batch that used this slot before reusing it
        o_h, d_h = host_batches[i % len(host_batches)]
        with torch.cuda.stream(st):
            sl['o_d'].copy_(o_h, non_blocking=True); sl['d_d'].copy_(d_h, non_blocking=True)
            rgb, alpha = (renderers[i % P].render_fused if use_fused else renderers[i % P].render)(sl['o_d'], sl['d_d'], bf)[:2]
            sl['rgb_h'].copy_(rgb, non_blocking=True); sl['alpha_h'].copy_(alpha, non_blocking=True)
            sl['done'] = torch.cuda.Event(); sl['done'].record(st)

    def e2e_drain():
        for sl in slots:
            if sl['done'] is not None:
                sl['done'].synchronize(); sl['done'] = None
    fork()
    for i in range(WU):
        e2e_step(i)
    e2e_drain()
    barrier()
    t0 = time.perf_counter()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(main)
    fork()
    for i in range(K):
        e2e_step(W + i)
    e2e_drain()
    join()
    g1.record(main)
    barrier()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = torch.tensor([max(g0.elapsed_time(g1), 0.0)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms.item())

    # ---- whole-image inference (SURVEY 8d C2 "full 640 000-ray images"): 800x800 spiral views rendered in pixel order, one call per image.
    # Placed before the training arm so that it renders the same (initial) weights as the batch arms.
    image_arm = None
    if not args.no_image:
        try:
            from xrnerf_b200 import synth as _synth
            poses = _synth.spiral_poses_ngp(40)
            n_views = 4
            views = []
            for v in range(n_views):
                o_np, d_np = _synth.get_rays_ngp(poses[(rank * n_views + v * 7) % 40])
                views.append((torch.from_numpy(o_np).to(dev), torch.from_numpy(d_np).to(dev)))
            n_img = views[0][0].shape[0]
            img_r = NgpRenderer(field, samples_per_ray_budget=BUDGET)
            res = {}
            for path in ('chain', 'fused'):
                fn = (lambda o_, d_: img_r.render_fused(o_, d_, bf)) if path == 'fused' else (lambda o_, d_: img_r.render(o_, d_, bf))
                for v in range(2):
                    fn(*views[v])
                barrier()
                i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                KI = 12
                ns_acc = torch.zeros((), dtype=torch.float64, device=dev)
                ns_acc += 0.0                                  # first-use kernels loaded before the timed region
                i0.record()
                for i in range(KI):
                    out_i = fn(*views[i % n_views])
                i1.record()
                barrier()
                im = torch.tensor([i0.elapsed_time(i1)], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(im, op=dist.ReduceOp.MAX)
                for v in range(n_views):                       # samples of EVERY timed image (the views are cycled; counts re-measured outside the timed region)
                    ns_v = fn(*views[v])[2]
                    ns_acc += (ns_v[:, 0] if ns_v.dim() == 2 else ns_v).double().sum() * (KI // n_views)
                spr = float(ns_acc.item()) / (n_img * KI)
                ms_img = float(im.item()) / KI
                img_bytes = n_img * spr * (512 if path == 'fused' else BYTES_PER_SAMPLE) + n_img * 44          # gather (+ coords/raw round trip on the chain path) + ray I/O
                pk, _ = peaks()
                res[path] = {'value': world * n_img * KI / (float(im.item()) * 1e-3), 'unit': 'rays/s', 'ms_per_image': ms_img, 'samples_per_ray_mean': spr,
                             'roofline': {'bound': 'hbm', 'achieved': img_bytes / (ms_img * 1e-3) / 1e9, 'peak': pk, 'unit': 'GB/s', 'frac': img_bytes / (ms_img * 1e-3) / 1e9 / pk,
                                          'algorithmic_bytes_per_image': img_bytes, 'note': 'whole call(s) of one image; gather served by L1/L2 (coherent rays)'}}
            image_arm = dict(res, what='800x800 spiral views in pixel order (coherent rays), 640 000 rays per call, sequential calls on one stream; 4 distinct views cycled (61 MB of rays)')
        except Exception as e:   # an auxiliary arm must never take the headline line down
            image_arm = {'error': repr(e)[:300]}

    # ---- parity sample (N=1): this arm's render of the 4096 rays the CPU reference arm renders below, taken BEFORE the training arm updates the weights
    parity_gpu = {}
    if world == 1:
        prn = NgpRenderer(field, samples_per_ray_budget=BUDGET, bg=(0., 0., 0.))
        po, pd = dev_batches[0][0][:4096].contiguous(), dev_batches[0][1][:4096].contiguous()
        for path in ('chain', 'fused'):
            prn.calls = 0                                   # same jitter stream as the CPU arm's call 0 (pcg32 seed 9121)
            if path == 'fused':
                rgb_g, _, ns_g = prn.render_fused(po, pd, bf)
            else:
                rgb_g, _, ns_g, _ = prn.render(po, pd, bf)
            torch.cuda.synchronize()
            ns_np = ns_g.cpu().numpy()
            parity_gpu[path] = (rgb_g.cpu().numpy().copy(), (ns_np[:, 0] if ns_np.ndim == 2 else ns_np).copy())

    # ---- training arm (BASELINE configs[4] shape: 65 536 rays per rank per step; march + compaction + field fwd/bwd (tcgen05) + composite fwd/bwd + loss + gradient exchange +
    # fused Adam; device-resident batches, synthetic targets). T = 2^20 compacted samples so that NO ray of the batch is truncated away (the reference's 2^18 is sized for its
    # adaptive ~24 K-ray batches); `trained_rays_per_step` is counted on the device (rays whose every marched sample took part).
    train = None
    if not args.no_train:
        try:
            from xrnerf_b200.train import NgpTrainer
            T_TRAIN = 1 << 20
            tr = NgpTrainer(field, bf, N_RAYS, target_batch_size=T_TRAIN, grad_comm=args.grad_comm)
            tgt = torch.rand((N_RAYS, 3), device=dev); bgc = torch.zeros((N_RAYS, 3), device=dev)
            nb = lambda i: dev_batches[i % N_BATCHES]
            for i in range(max(W, 3)):
                tr.step(*nb(i), tgt, bgc, next_rays=nb(i + 1))
            barrier()
            t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            trained = torch.zeros((), dtype=torch.int64, device=dev)
            t0e.record()
            for i in range(K):
                tr.step(*nb(W + i), tgt, bgc, next_rays=nb(W + i + 1))
                trained += tr.trained_rays()
            t1e.record()
            barrier()
            tm = torch.tensor([t0e.elapsed_time(t1e)], dtype=torch.float64, device=dev)
            tsum = trained.double().reshape(1)
            if world > 1:
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            train = {'value': float(tsum.item()) / (float(tm.item()) * 1e-3), 'unit': 'rays/s (trained rays only)', 'ms_per_step': float(tm.item()) / K,
                     'rays_per_step': world * N_RAYS, 'trained_rays_per_step': float(tsum.item()) / K, 'target_batch_size': T_TRAIN,
                     'compacted_samples_per_step_rank0': int(tr.compacted_samples().item()), 'grad_comm': tr.grad_comm, 'field_backward': 'tcgen05' if tr.bwd_impl == 1 else 'cuda cores',
                     'what': 'march + compaction (aux stream, one step ahead) | field fwd (tcgen05) + composite fwd + Huber x5 + composite bwd + field bwd (tcgen05 dX/dW) + gradient exchange '
                             '(world>1: bf16 reduce-scatter -> sharded Adam -> fp16 all-gather; MLP weights fp32 all-reduce) + fused Adam over 12.2M params + cell-image refresh'}
            del tr
        except Exception as e:   # an auxiliary arm must never take the headline line down
            import traceback
            train = {'error': repr(e)[:300], 'trace': traceback.format_exc()[-600:]}

    # ---- occupancy-grid update (ngp_grid_sampler.py:90-166; every 16 training steps): candidate cells -> density query -> splat -> EMA -> bitfield + mean
    grid_upd = None
    if not args.no_grid:
        try:
            from xrnerf_b200 import registry as R, synth
            from xrnerf_b200.registry.mlps import HashNerfMLP
            smp = R.build_sampler(dict(type='NGPGridSampler', update_grid_freq=16, update_block_size=5000000, n_rays_per_batch=4096, cone_angle_constant=0.00390625, near_distance=0.2,
                                       target_batch_size=1 << 18, rgb_activation=2, density_activation=3))
            poses = synth.spiral_poses_ngp(40)
            smp.set_data(dict(poses=poses, focal=np.full((40, 2), synth.FOCAL), aabb_scale=1, aabb_range=(0.0, 1.0), metadata=synth.metadata_for(40)), dict(H=800, W=800))
            smp.check_device({'rays_o': dev_batches[0][0]})

            class _Density:   # the sampler only needs run_density (hashnerf_mlp.py:107-111)
                def run_density(self, pts):
                    return field.run_density(pts)
            dm = _Density()
            M = 128 ** 3
            modes = {}
            for name, (nu, nn_) in (('warmup_phase_uniform_M', (M, 0)), ('steady_quarter_plus_quarter', (M // 4, M // 4))):
                for _ in range(2):
                    smp.update_density_grid_func(nu, nn_, dm)
                barrier()
                u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                KU = 5
                u0.record()
                for _ in range(KU):
                    smp.update_density_grid_func(nu, nn_, dm)
                u1.record()
                barrier()
                ms = u0.elapsed_time(u1) / KU
                n_q = nu + nn_
                # algorithmic HBM stream: tmp zero-fill 64 MB + EMA read grid+tmp, write grid (192 MB) + bitfield pass reads level grids (64 MB) writes 2 MB
                # + per candidate 12+4 B written and read, 4 B density, 4 B splat atomic; the density query's gathers (512 B/cell) are L2 traffic
                stream_bytes = (64 + 192 + 66) * 2 ** 20 + n_q * (2 * 16 + 4 + 4)
                modes[name] = {'ms_per_update': ms, 'ms_per_training_step_amortised': ms / 16, 'candidate_cells': n_q, 'hbm_stream_bytes': stream_bytes,
                               'stream_gbs': stream_bytes / (ms * 1e-3) / 1e9}
            grid_upd = dict(modes, what='NGPGridSampler.update_density_grid_func: generate_grid_samples x2, density-only field (tcgen05), splat (atomicMax), EMA, bitfield + cascade pooling + mean')
        except Exception as e:   # an auxiliary arm must never take the headline line down
            grid_upd = {'error': repr(e)[:300]}

    # ---- NeRF arm (BASELINE configs[2]: hierarchical 64 + 128, 800x800-shaped rays): fused tcgen05 NerfMLP path, device-resident rays
    nerf = None
    if not args.no_nerf:
        try:
            from xrnerf_b200 import registry as R
            from xrnerf_b200.nerf import NerfRenderer
            mlp_cfg = dict(type='NerfMLP', skips=[4], netdepth=8, netwidth=256, netchunk=1024 * 32, output_ch=5, use_viewdirs=True, embedder=dict(type='BaseEmbedder', i_embed=0, multires=10, multires_dirs=4))
            net = R.build_network(dict(type='NerfNetwork', cfg=dict(phase='test', N_importance=128, is_perturb=False, chunk=1024 * 32, bs_data='rays_o'), mlp=mlp_cfg, mlp_fine=mlp_cfg,
                                       render=dict(type='NerfRender', white_bkgd=True, raw_noise_std=0))).to(dev)
            nr = NerfRenderer(net, near=2.0, far=6.0, n_samples=64)
            n_nerf = 32768
            ro, rd = dev_batches[0][0][:n_nerf].contiguous(), dev_batches[0][1][:n_nerf].contiguous()
            for _ in range(3):
                nr.render(ro, rd, rd)
            barrier()
            n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            KN = max(3, min(K, 20))
            n0.record()
            for i in range(KN):
                o_i = dev_batches[i % N_BATCHES][0][:n_nerf]; d_i = dev_batches[i % N_BATCHES][1][:n_nerf]
                nr.render(o_i, d_i, d_i)
            n1.record()
            barrier()
            nm = torch.tensor([n0.elapsed_time(n1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(nm, op=dist.ReduceOp.MAX)
            rps = world * n_nerf * KN / (float(nm.item()) * 1e-3)
            flop_per_ray = (64 + 192) * 593408 * 2
            try:
                with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
                    tpeak = float(json.load(fh)['bf16_tflops_sustained'])
            except Exception:
                tpeak = 1400.0
            # CPU baseline leg of this arm (rank 0, N=1): BASELINE configs[0] shape - 1024 rays x 64 samples, coarse network only - through the numpy oracle
            # (embed -> 12-layer NerfMLP -> composite; multi-threaded BLAS on the box's host cores). Checker code, timed only.
            nerf_cpu = None
            if world == 1:
                try:
                    from oracle import nerf_oracle as NO
                    sd = {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}
                    o_c = dev_batches[0][0][:1024].cpu().numpy(); d_c = dev_batches[0][1][:1024].cpu().numpy()
                    z_c = np.broadcast_to(np.linspace(2, 6, 64, dtype=np.float32), (1024, 64)).copy()
                    best = None
                    for _ in range(2):
                        tc0 = time.perf_counter()
                        pts_c = o_c[:, None] + d_c[:, None] * z_c[..., None]
                        raw_c = NO.nerf_mlp(sd, NO.embed(pts_c, d_c), 63, 27, prefix='mlp.').reshape(1024, 64, 4)
                        NO.nerf_render(raw_c, z_c, d_c, white_bkgd=True)
                        dtc = time.perf_counter() - tc0
                        best = dtc if best is None else min(best, dtc)
                    nerf_cpu = {'value': 1024 / best, 'unit': 'rays/s', 'cores': os.cpu_count(), 'kind': 'port',
                                'sample': 'configs[0]: 1024 rays x 64 samples, coarse network only, numpy oracle (fp32, multi-threaded BLAS); best of 2'}
                except Exception as e:
                    nerf_cpu = {'error': repr(e)[:200]}
            nerf = {'value': rps, 'unit': 'rays/s', 'cpu_baseline': nerf_cpu, 'workload': 'vanilla NeRF hierarchical 64 coarse + 192 fine evaluations per ray (configs[2]), 32768-ray batches, inference',
                    'ms_per_batch': float(nm.item()) / KN, 'roofline': {'bound': 'tensor', 'achieved': rps * flop_per_ray / 1e12 / world, 'peak': tpeak, 'unit': 'TFLOP/s',
                                                                         'frac': rps * flop_per_ray / 1e12 / world / tpeak, 'flop_per_ray': flop_per_ray,
                                                                         'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a multi-kernel step)'}}
        except Exception as e:   # an auxiliary arm must never take the headline line down
            nerf = {'error': repr(e)[:300]}

    # ---- NeRF training step (configs[2], N_rand 4096 as configs/nerf/nerf_blender_base01.py): our encoders / composite fwd+bwd / sample_pdf kernels under autograd;
    # the 8x256 GEMM chain and its backward run on library GEMMs (fp32) - the tensor-core backward is not written yet
    nerf_train = None
    if not args.no_nerf:
        try:
            from xrnerf_b200 import registry as R
            mlp_cfg_t = dict(type='NerfMLP', skips=[4], netdepth=8, netwidth=256, netchunk=1024 * 32, output_ch=5, use_viewdirs=True, embedder=dict(type='BaseEmbedder', i_embed=0, multires=10, multires_dirs=4))
            tnet = R.build_network(dict(type='NerfNetwork', cfg=dict(phase='train', N_importance=128, is_perturb=False, chunk=1024 * 32, bs_data='rays_o'), mlp=mlp_cfg_t, mlp_fine=mlp_cfg_t,
                                        render=dict(type='NerfRender', white_bkgd=True, raw_noise_std=0))).to(dev)
            topt = torch.optim.Adam(tnet.parameters(), lr=5e-4, betas=(0.9, 0.999))
            n_t = 4096
            o_t, d_t = dev_batches[0][0][:n_t].contiguous(), dev_batches[0][1][:n_t].contiguous()
            tt = torch.linspace(0., 1., 64, device=dev)
            z_t = (2.0 * (1. - tt) + 6.0 * tt).expand(n_t, 64).contiguous()
            tdata = {'rays_o': o_t[None], 'rays_d': d_t[None], 'viewdirs': d_t[None], 'z_vals': z_t[None], 'pts': (o_t[:, None, :] + d_t[:, None, :] * z_t[:, :, None])[None],
                     'target_s': torch.rand((1, n_t, 3), device=dev)}

            def tstep():
                out = tnet.train_step(dict(tdata), topt)
                topt.zero_grad(set_to_none=True); out['loss'].backward(); topt.step()
            for _ in range(2):
                tstep()
            barrier()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            KT = 5
            q0.record()
            for _ in range(KT):
                tstep()
            q1.record()
            barrier()
            qm = torch.tensor([q0.elapsed_time(q1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(qm, op=dist.ReduceOp.MAX)
            nerf_train = {'value': world * n_t * KT / (float(qm.item()) * 1e-3), 'unit': 'rays/s', 'ms_per_step': float(qm.item()) / KT, 'rays_per_step_per_gpu': n_t,
                          'what': 'NerfNetwork.train_step + Adam (per-rank, no gradient all-reduce in this arm): fused encoders / composite fwd+bwd / sample_pdf kernels, dense layers on library fp32 GEMMs under autograd'}
            del tnet, topt, tdata
        except Exception as e:   # an auxiliary arm must never take the headline line down
            nerf_train = {'error': repr(e)[:300]}

    # ---- Mip-NeRF arm (BASELINE configs[3]: 2 levels x 128 cone samples, IPE): the same tcgen05 NerfMLP on IPE tile images, device-resident rays
    mip = None
    if not args.no_mip:
        try:
            from xrnerf_b200 import registry as R
            from xrnerf_b200.nerf import MipNerfRenderer
            mnet = R.build_network(dict(type='MipNerfNetwork', cfg=dict(phase='test', ray_shape='cone', resample_padding=0.01, use_multiscale=False, coarse_loss_mult=0.1, num_levels=2,
                                                                        chunk=1024 * 32, bs_data='rays_o'),
                                        mlp=dict(type='NerfMLP', skips=[4], netdepth=8, netwidth=256, netchunk=1024 * 32, use_viewdirs=True,
                                                 embedder=dict(type='MipNerfEmbedder', min_deg_point=0, max_deg_point=16, min_deg_view=0, max_deg_view=4, use_viewdirs=True, append_identity=True)),
                                        render=dict(type='MipNerfRender', white_bkgd=True, raw_noise_std=0, rgb_padding=0.001, density_bias=-1, density_activation='softplus'))).to(dev)
            mr = MipNerfRenderer(mnet, near=2.0, far=6.0, n_samples=128)
            n_mip = 32768
            radii = torch.full((n_mip,), 2.0 / (1111.111 * 12 ** 0.5), device=dev)        # GetRays radii of an 800x800 f=1111 camera: |dx| * 2/sqrt(12) (create.py:237-243)
            for i in range(3):
                mr.render(dev_batches[0][0][:n_mip], dev_batches[0][1][:n_mip], dev_batches[0][1][:n_mip], radii)
            barrier()
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            KM = max(3, min(K, 20))
            m0.record()
            for i in range(KM):
                o_i = dev_batches[i % N_BATCHES][0][:n_mip]; d_i = dev_batches[i % N_BATCHES][1][:n_mip]
                mr.render(o_i, d_i, d_i, radii)
            m1.record()
            barrier()
            mm = torch.tensor([m0.elapsed_time(m1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(mm, op=dist.ReduceOp.MAX)
            mrps = world * n_mip * KM / (float(mm.item()) * 1e-3)
            mflop = 256 * 610304 * 2
            try:
                with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
                    tpeak = float(json.load(fh)['bf16_tflops_sustained'])
            except Exception:
                tpeak = 1400.0
            mip = {'value': mrps, 'unit': 'rays/s', 'workload': 'Mip-NeRF 2 levels x 128 conical-frustum samples per ray, IPE 96 + 27 (configs[3]), 32768-ray batches, inference',
                   'ms_per_batch': float(mm.item()) / KM, 'roofline': {'bound': 'tensor', 'achieved': mrps * mflop / 1e12 / world, 'peak': tpeak, 'unit': 'TFLOP/s',
                                                                       'frac': mrps * mflop / 1e12 / world / tpeak, 'flop_per_ray': mflop,
                                                                       'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a multi-kernel step)'}}
        except Exception as e:   # an auxiliary arm must never take the headline line down
            mip = {'error': repr(e)[:300]}

    if rank == 0:
        peak, peak_src = peaks()
        f_ms = float(np.mean(field_ms))
        s_mean = float(np.mean(samples))
        achieved = s_mean * BYTES_PER_SAMPLE / (f_ms * 1e-3) / 1e9
        cpu_rate, cores, kind, sample, _ = cpu_reference_rate(4096, reps=2) if world == 1 else (None, os.cpu_count(), 'reference', 'measured at N=1 only', None)
        parity = None
        if world == 1:   # the CPU arm just rendered 4096 rays of batch 0 with the reference arithmetic: compare this arm's render of the same rays (checker only)
            rgb_cpu, alpha_cpu, ns_cpu = cpu_reference_rate.last
            parity = {'rays': 4096, 'against': kind + ' CPU arm (oracle), fp32 march/composite + fp16 tcnn-shaped field'}
            for path, (rgb_g, ns_g) in parity_gpu.items():
                err = np.abs(rgb_g - rgb_cpu)
                mse = float((err.astype(np.float64) ** 2).mean())
                parity[path] = {'max_abs_rgb_err': float(err.max()), 'psnr_vs_ref_db': float(-10.0 * np.log10(max(mse, 1e-20))), 'sample_counts_bit_exact': bool(np.array_equal(ns_g, ns_cpu))}
        chain = {'value': world * N_RAYS * K / (total_ms_max * 1e-3), 'unit': 'rays/s', 'ms_per_step': total_ms_max / K, 'gpu_launches_per_step': 5,
                 'what': '5 launches per batch on one stream (march count / scan / emit, field, composite), P batches in flight on P streams',
                 'roofline': {'kernel': 'xrb::ngp_field_tc_kernel<false>', 'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                              'kernel_ms': f_ms, 'kernel_share_of_step': f_ms / (total_ms_max / K), 'algorithmic_bytes_per_launch': s_mean * BYTES_PER_SAMPLE}}
        f_bytes = fused_samples * 512 + N_RAYS * 44   # no coords[S,7] / raw[S,4] round trip: gather bytes + 44 B of ray I/O
        f_ach = f_bytes / (fused_kernel_ms * 1e-3) / 1e9
        fused = {'value': world * N_RAYS * K / (fused_total_ms * 1e-3), 'unit': 'rays/s', 'ms_per_step': fused_total_ms / K, 'gpu_launches_per_step': 1,
                 'what': 'ONE launch per batch (xrb_ngp_render_fused: warp-specialised march + hash encode + tcgen05 MLPs + segmented-scan composite), P batches in flight on P streams',
                 'roofline': {'kernel': 'xrb::ngp_render_fused_kernel', 'bound': 'hbm', 'achieved': f_ach, 'peak': peak, 'unit': 'GB/s', 'frac': f_ach / peak,
                              'kernel_ms': fused_kernel_ms, 'kernel_share_of_step': 1.0, 'algorithmic_bytes_per_launch': f_bytes}}
        head = fused if use_fused else chain
        line = {
            'metric': METRIC, 'value': head['value'], 'unit': 'rays/s', 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': head['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f16', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'rays_per_step_per_gpu': N_RAYS, 'samples_per_ray_mean': s_mean / N_RAYS, 'parallelism': f'ray-sharded x{world}, no data-path collective',
                       'l2': f'inputs larger than L2: {N_BATCHES} distinct ray batches = {N_BATCHES * N_RAYS * 24 / 1e6:.0f} MB cycled (L2 126 MB); the 24.4 MB fp16 hash table stays L2-resident as in production',
                       'timing': 'one CUDA-event pair around the K steps on the launching stream, max over ranks', 'batches_in_flight': P, 'untimed_warmup_steps_run': WU,
                       'path': 'fused single launch' if use_fused else 'chain of 5 launches', 'path_selection': args.path},
            'clocks': clk,
            'e2e': {'value': world * N_RAYS * K / (e2e_ms * 1e-3), 'unit': 'rays/s', 'h2d_bytes_per_step': N_RAYS * 24, 'd2h_bytes_per_step': N_RAYS * 16,
                    'ms_per_step': e2e_ms / K, 'host_wall_ms_per_step': e2e_wall_ms / K},
            'gpu_launches': head['gpu_launches_per_step'] * K,
            # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of that kernel at this workload, from the committed ncu --set full captures
            # (profiles/r01b_ngp_field_tc_ncu.md: 41.88 MB + 1.64 MB; profiles/r01b_ngp_render_fused_ncu.md: 27.80 MB + 0.04 MB) - far below the algorithmic
            # bytes because the gather is served by L2/L1
            'roofline': dict(head['roofline'], traffic=(27.80e6 + 0.04e6) if use_fused else (41.88e6 + 1.64e6), traffic_source='ncu --set full capture committed under profiles/ (r01b), bytes per launch', peak_source=peak_src,
                             note='hash table (24.4 MB fp16) is L2-resident by design: the gather is served by L1/L2, so DRAM traffic (ncu, profiles/) is far below the algorithmic bytes'),
            'paths': {'chain': chain, 'fused': fused},
            'cpu_baseline': {'value': cpu_rate, 'unit': 'rays/s', 'cores': cores, 'kind': kind, 'sample': sample},
            'parity': parity,
            'image': image_arm,
            'train': train,
            'grid_update': grid_upd,
            'nerf': nerf,
            'nerf_train': nerf_train,
            'mip': mip,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-train', dest='no_train', action='store_true', help='skip the training arm')
    ap.add_argument('--no-nerf', dest='no_nerf', action='store_true', help='skip the vanilla-NeRF arm')
    ap.add_argument('--no-image', dest='no_image', action='store_true', help='skip the whole-image inference arm')
    ap.add_argument('--no-grid', dest='no_grid', action='store_true', help='skip the occupancy-grid update arm')
    ap.add_argument('--no-mip', dest='no_mip', action='store_true', help='skip the Mip-NeRF arm')
    ap.add_argument('--path', default='auto', choices=['auto', 'chain', 'fused'], help='inference path of the headline/e2e numbers: 5-launch chain, single-launch fused kernel, or the faster of the two (both are always measured)')
    ap.add_argument('--grad-comm', dest='grad_comm', default='sharded', choices=['sharded', 'allreduce'], help='gradient exchange of the training arm at world > 1')
    ap.add_argument('--pipeline', type=int, default=4, help='ray batches in flight (CUDA streams); 1 = strictly sequential steps')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()

"""Developer timing probe: single-launch fused NGP render vs the 5-launch path (sequential and with batches in flight on several streams)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from xrnerf_b200 import synth, _C
from xrnerf_b200.ngp import NgpField, NgpRenderer


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    nb = int(os.environ.get('BF_NB', '24'))
    grid = synth.lego_like_density_grid(0)
    bf, _ = synth.bitfield_from_grid_numpy(grid)
    bf = torch.from_numpy(bf).cuda()
    f = NgpField().cuda()
    if os.environ.get('BF_WEIGHTS'):
        table, dens, color = synth.ngp_weights(seed=0)
        with torch.no_grad():
            f.hash_params.copy_(torch.from_numpy(table).cuda()); f.density_params.copy_(torch.from_numpy(dens).cuda()); f.color_params.copy_(torch.from_numpy(color).cuda())
    batches = []
    for b in range(nb):
        o, d, _, _ = synth.ray_batch(N, seed=b + 1)
        batches.append((torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()))
    r = NgpRenderer(f)
    rgb_u, alpha_u, ns_u, _ = r.render(*batches[0], bf)
    r2 = NgpRenderer(f)
    rgb, alpha, ns = r2.render_fused(*batches[0], bf)
    torch.cuda.synchronize()
    print('samples/ray %.2f  ns equal %s  max|drgb| %.2e' % (ns.float().mean().item(), torch.equal(ns, ns_u[:, 0]), (rgb - rgb_u).abs().max().item()), flush=True)

    timers = False

    def run(fn, P, steps=96):
        timers = globals().get('timers', False)
        streams = [torch.cuda.Stream() for _ in range(P)]
        rs = [NgpRenderer(f) for _ in range(P)]
        for w in range(2):
            for k in range(P):
                with torch.cuda.stream(streams[k]):
                    fn(rs[k], *batches[k % nb], bf)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(P):
            streams[k].wait_event(e0)
        for i in range(steps):
            with torch.cuda.stream(streams[i % P]):
                fn(rs[i % P], *batches[i % nb], bf)
        for k in range(P):
            torch.cuda.current_stream().wait_stream(streams[k])
        e1.record(); torch.cuda.synchronize()
        if timers and getattr(rs[0], '_ws_fused', None) is not None:
            dump_timers(rs[0]._ws_fused, 'last launch of stream 0 at P=%d:' % P)
        return e0.elapsed_time(e1) / steps

    def dump_timers(ws, label):
        torch.cuda.synchronize()
        tail = ws[-296 * 128:].view(torch.int64).reshape(296, 16).cpu().numpy()
        tail = tail[tail[:, 0] > 0]
        t0 = tail[:, 0].min()
        st, en = (tail[:, 0] - t0) / 1e3, (tail[:, 1] - t0) / 1e3
        act = tail[:, 3] > 0
        mhz = 1.9e3
        print(label, 'CTAs %d  start us p50 %.1f max %.1f | end us p50 %.1f max %.1f' % (len(tail), np.median(st), st.max(), np.median(en), en.max()))
        print('  field WG0 us/CTA: poll %.0f waitfull %.0f copy %.0f compute %.0f post %.0f  (tiles/CTA both WGs %.1f)' % tuple(list(tail[act, 4:9].mean(0) / mhz) + [tail[act, 3].mean()]))
        print('  producers us/CTA (sum over warps): march %.0f setup %.0f gather %.0f mailwait %.0f fold %.0f slotwait %.0f other %.0f' % tuple(tail[act, 9:16].mean(0) / mhz))
    timers = bool(int(os.environ.get('XRB_FUSED_DBG', '0')) & 8)
    if timers:
        dump_timers(r2._ws_fused, 'first call:')
    only_fused = len(sys.argv) > 2
    globals()['timers'] = timers
    for P in (1, 4):
        tu = 1.0 if only_fused else run(lambda rr, o, d, b: rr.render(o, d, b), P)
        tf = run(lambda rr, o, d, b: rr.render_fused(o, d, b), P)
        print(f'P={P}: unfused {tu:.3f} ms ({N / tu / 1e3:.1f} Mrays/s)   fused {tf:.3f} ms ({N / tf / 1e3:.1f} Mrays/s)', flush=True)


if __name__ == '__main__':
    main()

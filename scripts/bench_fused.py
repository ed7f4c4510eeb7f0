"""Developer timing probe: single-launch fused NGP render vs the 5-launch path (sequential and with batches in flight on several streams)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from xrnerf_b200 import synth, _C
from xrnerf_b200.ngp import NgpField, NgpRenderer


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    nb = 24
    grid = synth.lego_like_density_grid(0)
    bf, _ = synth.bitfield_from_grid_numpy(grid)
    bf = torch.from_numpy(bf).cuda()
    f = NgpField().cuda()
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

    def run(fn, P, steps=96):
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
        return e0.elapsed_time(e1) / steps

    if int(os.environ.get('XRB_FUSED_DBG', '0')) & 8:
        torch.cuda.synchronize()
        tail = r2._ws_fused[-148 * 128:].view(torch.int64).reshape(148, 16).cpu().numpy()
        t0 = tail[:, 0].min()
        st, en = (tail[:, 0] - t0) / 1e3, (tail[:, 1] - t0) / 1e3
        print('CTA start us: min %.1f p50 %.1f max %.1f | end us: min %.1f p50 %.1f max %.1f' % (st.min(), np.median(st), st.max(), en.min(), np.median(en), en.max()))
        print('late starters (>20us):', int((st > 20).sum()), ' distinct SMs:', len(set(tail[:, 2].tolist())), ' tiles/CTA: mean %.1f max %d' % (tail[:, 3].mean(), tail[:, 3].max()))
        act = tail[:, 3] > 0
        mhz = 1.9e3
        print('field WG us/CTA: poll %.0f waitfull %.0f copy %.0f compute %.0f post %.0f  (tiles %.1f)' % tuple(list(tail[act, 4:9].mean(0) / mhz) + [tail[act, 3].mean()]))
        print('producers us/CTA (sum over 10 warps): march %.0f setup %.0f gather %.0f mailwait %.0f fold %.0f slotwait %.0f other %.0f' % tuple(tail[act, 9:16].mean(0) / mhz))
        import collections
        cnt = collections.Counter(tail[:, 2].tolist())
        print('CTAs per SM histogram:', collections.Counter(cnt.values()))
    only_fused = len(sys.argv) > 2
    for P in (1, 4):
        tu = 1.0 if only_fused else run(lambda rr, o, d, b: rr.render(o, d, b), P)
        tf = run(lambda rr, o, d, b: rr.render_fused(o, d, b), P)
        print(f'P={P}: unfused {tu:.3f} ms ({N / tu / 1e3:.1f} Mrays/s)   fused {tf:.3f} ms ({N / tf / 1e3:.1f} Mrays/s)', flush=True)


if __name__ == '__main__':
    main()

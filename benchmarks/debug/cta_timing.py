"""Per-CTA timing of the scatter stage (debug builds only).

Build a library with -DCMR_DBG_TIMING (common.cuh: %globaltimer marks per CTA written to g_dbg), e.g.
    nvcc <flags of cmr_agent_b200/build.py> -DCMR_DBG_TIMING -o /tmp/lib_timing.so cmr_agent_b200/csrc/cmr_b200.cu
and run   CMR_B200_LIB=/tmp/lib_timing.so python benchmarks/debug/cta_timing.py
It prints the span of k_tile_gather, the distribution of CTA durations per role and the slowest CTAs.
Marks (thread 0 of the CTA = light unit of warp 0): 0 start, 1 griddepcontrol.wait returned, 2 count + entries
arrived (bucket CTA: heavy buckets ranked), 3 rows added, 4 stored, 5 all warps of the CTA done.
"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from cmr_agent_b200 import _lib, synth, environment as env
dev = torch.device('cuda:0'); B, N = 32, 40960
HEAVY = 16
cpu = synth.make_batch(B, seed=2023, num_pt=N, img_h=160, img_w=512)
data = dict(cpu)
for k in ("pc", "pc_overlap_pred", "pc_geo_feat", "img_geo_feat"): data[k] = cpu[k].to(dev)
pose, _ = env.init(data)
for _ in range(3): env.observation_from_a_pose(data, pose)
ep = env.episode_state(data); p = _lib.ptr
obs2d = torch.empty(B, 128, 40, 128, device=dev)
obs3d = torch.empty(B, 5, N, device=dev)
lib = _lib.load()
lib.cmr_debug_read.restype = ctypes.c_int
sync_between = len(sys.argv) > 1 and sys.argv[1] == "sync"
for rep in range(3):   # a scatter consumes what ONE project left
    _lib.call("cmr_project", p(ep.pc), p(ep.overlap), p(ep.K), p(pose), p(ep.mean), p(ep.ws), B, N, 64, 40, 128,
              p(obs3d), None, None, p(ep.img_feat), p(obs2d), None, 1, _lib.stream())
    if sync_between:
        torch.cuda.synchronize()
    _lib.call("cmr_tile_scatter", p(ep.img_feat), p(ep.K), p(ep.ws), B, N, 64, 40, 128, 0, p(obs2d), _lib.stream())
torch.cuda.synchronize()
buf = np.zeros(16 * 8192, np.uint64)
rc = lib.cmr_debug_read(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(buf.nbytes)); assert rc == 0, rc
nct = min(8192, (HEAVY + 20) * B)
d = buf.reshape(8192, 16)[:nct].astype(np.int64)
yy = np.arange(nct) // B
role = (yy < 2 * HEAVY) & (yy % 2 == 0)     # bucket CTAs alternate with light CTAs for the first 2 * HEAVY rows
t0 = d[:, 1].min()     # first CTA released by griddepcontrol.wait
print('CTAs seen', nct, ' span from first wait-return to last end us', (d[:, 5].max() - t0) / 1e3)
for name, m in (('light', ~role), ('bucket', role)):
    x = d[m]
    rel, end = (x[:, 1] - t0) / 1e3, (x[:, 5] - t0) / 1e3
    dur = end - rel
    print(f'{name}: n={m.sum()} dur(after wait) mean {dur.mean():.2f} median {np.median(dur):.2f} p90 {np.percentile(dur, 90):.2f} max {dur.max():.2f}; '
          f'wait-return min {rel.min():.1f} median {np.median(rel):.1f} max {rel.max():.1f}; end max {end.max():.1f}; entries mean {x[:,7].mean():.1f} max {x[:,7].max()}')
    if name == 'light':
        ok = x[:, 7] > 0
        ph = lambda a, b_, mm: ((x[mm, a] - x[mm, b_]) / 1e3).mean()
        ok = (x[:, 7] > 0) & (x[:, 7] <= 64)
        print('   warp-0 unit with entries (%d): loads %.2f  order+flags %.2f  add+means %.2f  store %.2f  other warps still busy %.2f' % (ok.sum(), ph(2, 1, ok), ph(8, 2, ok), ph(3, 8, ok), ph(4, 3, ok), ph(5, 4, ok)))
        for lo, hi in ((1, 8), (9, 16), (17, 32), (33, 48), (49, 64)):
            mm = (x[:, 7] >= lo) & (x[:, 7] <= hi)
            if mm.sum(): print('      %d-%d entries (%d units): order %.2f  add %.2f' % (lo, hi, mm.sum(), ph(8, 2, mm), ph(3, 8, mm)))
        em = x[:, 7] == 0
        print('   warp-0 unit empty (%d): loads %.2f  store %.2f  others %.2f' % (em.sum(), ph(2, 1, em), ph(4, 2, em), ph(5, 4, em)))
    else:
        busy = x[:, 6] > HEAVY * 0 + 0
        one = x[:, 6] == 1
        ph = lambda a, b_: ((x[one, a] - x[one, b_]) / 1e3).mean()
        print('   CTAs with one item (%d, %.0f entries): queue read %.2f  entries staged %.2f  bin+sort+flags %.2f  zero+add %.2f  mean+store %.2f' % (one.sum(), x[one, 7].mean(), ph(2, 1), ph(8, 2), ph(9, 8), ph(10, 9), ph(11, 10)))
    order = np.argsort(-dur)[:6]
    for i in order:
        extra = ''
        if name == 'bucket':
            extra = '  phases: queue %.2f stage %.2f sort %.2f add %.2f [rows landed %.2f, warp0 (%d entries) added %.2f, all warps %.2f] mean+store %.2f' % ((x[i, 2] - x[i, 1]) / 1e3, (x[i, 8] - x[i, 2]) / 1e3, (x[i, 9] - x[i, 8]) / 1e3, (x[i, 10] - x[i, 9]) / 1e3, (x[i, 12] - x[i, 9]) / 1e3, x[i, 15], (x[i, 13] - x[i, 12]) / 1e3, (x[i, 14] - x[i, 13]) / 1e3, (x[i, 11] - x[i, 10]) / 1e3)
        print('     cta', i, 'entries', x[i, 7], 'items', x[i, 6], 'released %.1f dur %.1f' % (rel[i], dur[i]), extra)
hist, edges = np.histogram((d[:, 5] - t0) / 1e3, bins=10)
print('end-time histogram us:', hist.tolist(), edges.round(1).tolist())

"""Per-CTA timing of the scatter stage (debug builds only).

Build a library with -DCMR_DBG_TIMING (common.cuh: %globaltimer marks per CTA written to g_dbg), e.g.
    nvcc <flags of cmr_agent_b200/build.py> -DCMR_DBG_TIMING -o /tmp/lib_timing.so cmr_agent_b200/csrc/cmr_b200.cu
and run   CMR_B200_LIB=/tmp/lib_timing.so python benchmarks/debug/cta_timing.py
It prints the kernel span, the distribution of CTA durations and the slowest CTAs of cmr_tile_scatter.
"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from cmr_agent_b200 import _lib, synth, environment as env
dev = torch.device('cuda:0'); B, N = 32, 40960
cpu = synth.make_batch(B, seed=2023, num_pt=N, img_h=160, img_w=512)
data = dict(cpu)
for k in ("pc", "pc_overlap_pred", "pc_geo_feat", "img_geo_feat"): data[k] = cpu[k].to(dev)
pose, _ = env.init(data)
for _ in range(3): env.observation_from_a_pose(data, pose)
ep = data["_cmr_b200_episode"][1]; p = _lib.ptr
obs2d = torch.empty(B, 128, 40, 128, device=dev)
for _ in range(3):
    _lib.call("cmr_tile_scatter", p(ep.img_feat), p(ep.K), p(ep.ws), B, N, 64, 40, 128, 0, p(obs2d), _lib.stream())
torch.cuda.synchronize()
lib = _lib.load()
buf = np.zeros(16 * 8192, np.uint64)
lib.cmr_debug_read.restype = ctypes.c_int
rc = lib.cmr_debug_read(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(buf.nbytes)); assert rc == 0, rc
d = buf.reshape(8192, 16)[:5120].astype(np.int64); d = d[d[:,4] > 0]
t0 = d[:, 0].min()
start, wait, scan, acc_, end = [(d[:, i] - t0) / 1e3 for i in range(5)]
smid, tile, hits = d[:, 5], d[:, 6], d[:, 7]
print('kernel span us', (d[:, 4].max() - t0) / 1e3)
dur = end - start
print('CTA duration us: mean %.2f median %.2f p90 %.2f max %.2f' % (dur.mean(), np.median(dur), np.percentile(dur, 90), dur.max()))
print('phases mean us: zero->wait %.2f  wait->scan_done %.2f  scan_done->acc_done %.2f  acc_done->end %.2f' % ((wait - start).mean(), (scan - wait).mean(), (acc_ - scan).mean(), (end - acc_).mean()))
order = np.argsort(-dur)[:10]
for i in order:
    print('     warp0: staged %.1f  flush_begin %.1f  flush_end %.1f n_own(w0) %d' % ((d[i,8]-t0)/1e3-start[i], (d[i,9]-t0)/1e3-start[i], (d[i,10]-t0)/1e3-start[i], d[i,11]))
    print('  cta', i, 'p0', tile[i], 'hits', hits[i], 'sm', smid[i], 'start %.1f dur %.1f  phases %.1f %.1f %.1f %.1f' % (start[i], dur[i], wait[i]-start[i], scan[i]-wait[i], acc_[i]-scan[i], end[i]-acc_[i]))
print('start time histogram (us):', np.histogram(start, bins=8)[0].tolist(), np.histogram(start, bins=8)[1].round(1).tolist())
print('hits: mean %.1f max %d; corr(dur,hits)=%.2f' % (hits.mean(), hits.max(), np.corrcoef(dur, hits)[0, 1]))
# concurrency per SM
for s in (0, 1):
    m = smid == s
    print('sm', s, 'n ctas', int(m.sum()), 'busy span %.1f..%.1f' % (start[m].min(), end[m].max()))

import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import wordpiece_b200
from wordpiece_b200 import synth
wl = sys.argv[1] if len(sys.argv) > 1 else "en"
g = synth.generator(wl)
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
g.fill(h.numpy(), seed=2, first_block=0, n_threads=16)
d = h.cuda()
v = wordpiece_b200.Vocab(g.spec.vocab, device=0)
ids = torch.empty(n // 2 + 4096, dtype=torch.int32, device="cuda")
L = wordpiece_b200.load_library()
for _ in range(3):
    v.encode_device(d, ids)
out = np.zeros((3, 4096), dtype=np.uint64)
L.wp_debug_k2(out.ctypes.data_as(C.c_void_p))
nw = 444 * 8
t0 = out[0, :nw].astype(np.int64); t1 = out[1, :nw].astype(np.int64); r = out[2, :nw].astype(np.int64)
start = t0.min()
end = (t1 - start) / 1e3
print(wl, "warps", nw, "kernel span us", end.max(), "start spread us", (t0.max() - start) / 1e3)
print("  warp end us: min %.1f p10 %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f mean %.1f" % (end.min(), *np.percentile(end, [10, 50, 90, 99]), end.max(), end.mean()))
print("  rounds: min %d p50 %d p90 %d max %d mean %.1f" % (r.min(), np.percentile(r, 50), np.percentile(r, 90), r.max(), r.mean()))
dur = (t1 - t0) / 1e3
print("  us per round: mean %.2f" % (dur.sum() / r.sum()))

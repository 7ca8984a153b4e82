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
out = (C.c_uint64 * 16)()
for _ in range(3):
    v.encode_device(d, ids)
L.wp_debug_phases(out)
v.encode_device(d, ids)
L.wp_debug_phases(out)
names = ["ticket+init", "load", "classify", "dirty/left", "S1d lists", "S2a probes", "lookback walk", "result out", "memo", "reserve", "emit"]
tiles = out[15]
tot = sum(out[i] for i in range(11))
print(wl, "tiles", tiles, "cycles/tile", tot / tiles)
for i, nm in enumerate(names):
    print(f"  {nm:14s} {out[i] / tiles:9.0f} cyc  {100.0 * out[i] / tot:5.1f} %")

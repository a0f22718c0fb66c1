import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskrcnn_b200 import _lib as L, synth
N = 6000
rng = np.random.default_rng(11)
b = synth.random_rois(N, 11, image=1024.0, min_size=16, max_size=500) * 1024.0
b[N // 2:] = b[:N - N // 2] + rng.uniform(-8, 8, (N - N // 2, 4)).astype(np.float32)
d5 = torch.from_numpy(np.concatenate([b, np.sort(synth.unique_scores(N, 11))[::-1][:, None]], 1).astype(np.float32)).cuda()
keep = torch.empty(N, dtype=torch.int64, device="cuda"); cnt = torch.empty(1, dtype=torch.int32, device="cuda")
ws = torch.empty(L.lib.mrcnn_nms_workspace_bytes(N), dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
raw = ctypes.CDLL(L.LIB_PATH)
for it in range(4):
    L.check(L.lib.mrcnn_nms(d5.data_ptr(), N, 0.7, keep.data_ptr(), cnt.data_ptr(), ws.data_ptr(), ws.numel(), st))
    torch.cuda.synchronize()
    out = (ctypes.c_longlong * 64)()
    raw.mrcnn_debug_nms_trace(out)
    t = np.array(out[:], dtype=np.int64)
    t0 = t[0]
    print("call", it, "stage %d" % (t[1] - t0), end=" | ")
    for p in range(1, 19):
        a, b_, c = t[2 + 3 * p], t[3 + 3 * p], t[4 + 3 * p]
        if a < t0: break
        print("p%d start+%d poll %d comp %d" % (p, a - t0, b_ - a, (c - b_) if c >= b_ else -1), end=" | ")
    print("cta0 exit-loop +%d emit %d" % (t[60] - t0, t[61] - t[60]))
pubt = (ctypes.c_longlong * (16 * 160))(); startt = (ctypes.c_longlong * 160)()
raw.mrcnn_debug_nms_trace2(pubt, startt)
P = np.array(pubt[:], dtype=np.int64).reshape(16, 160)[:, :94]; S0 = np.array(startt[:], dtype=np.int64)[:94]
base = S0.min()
print("CTA start skew: min 0 max %d (cta %d); cta93 at %d" % (S0.max() - base, S0.argmax(), S0[93] - base))
for p in range(1, 8):
    r = P[p] - base
    print("pass %d publish: min %d (cta %d) median %d max %d (cta %d)  cta0 %d cta93 %d" % (p, r.min(), r.argmin(), np.median(r), r.max(), r.argmax(), r[0], r[93]))
    print("   slowest five:", [(int(i), int(r[i])) for i in np.argsort(-r)[:5]])

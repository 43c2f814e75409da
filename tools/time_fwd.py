"""Times the eval forward (predict) of one workload: seq/s and per-launch time of the fused layer-forward kernel."""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import WORKLOADS, synth_batch
from transformergrooveinfilling_b200 import GrooveTransformerEncoder, _lib
name = sys.argv[1] if len(sys.argv) > 1 else "c4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
train = len(sys.argv) > 3 and sys.argv[3] == "train"
w = WORKLOADS[name]
lib = _lib.load()
m = GrooveTransformerEncoder(w["d"], w["E"], 27, w["H"], w["F"], w["p"], w["L"], 32, "cuda")
m.set_precision("bf16").set_seed(1)
m.train() if train else m.eval()
x, y = synth_batch(w, n, 1)
x = x.cuda()
with torch.no_grad():
    for _ in range(3): m(x)
    lib.gt_profile_enable(17, 4096)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): m(x)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
tot, cnt = C.c_double(0), C.c_int64(0)
lib.gt_profile_collect(C.byref(tot), C.byref(cnt))
mac = 32 * (4 * w["d"] ** 2 + 2 * w["d"] * w["F"]) + 2 * 32 * 32 * w["d"]
per = tot.value / cnt.value
print(f"{name} n={n} train={train}: {n/dt:.0f} seq/s fwd; layer fwd kernel {per*1e3:.1f} us/launch = {2*mac*n/(per*1e-3)/1e12:.1f} TFLOP/s")

"""Per-stage device timing of one plan on synthetic reads (no parity check): python tools/time_stages.py [kit] [n_reads]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qcat_b200 import config, engine, scanner, synth
from qcat_b200.tables import Tables

kit = sys.argv[1] if len(sys.argv) > 1 else "PBC096"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
mode = "dual" if kit == "dual" else "epi2me"
sc = (scanner.BarcodeScannerDual() if mode == "dual" else scanner.BarcodeScannerEPI2ME(kit=kit))
d = synth.generate(sc.layouts, n, seed=1)
plan = engine.DevicePlan(Tables(sc.layouts, config.qcatConfig(), mode, sc.min_quality), device=0)
dev = torch.device("cuda", 0)
t = {k: torch.from_numpy(d[k]).to(dev) for k in ("win5", "tail3", "wlen", "read_len")}
out = torch.zeros(n * 32, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
def step():
    plan.detect_device(t["win5"].data_ptr(), t["tail3"].data_ptr(), 160, t["wlen"].data_ptr(), t["read_len"].data_ptr(), n, out.data_ptr(), stream=st)
for _ in range(3): step()
torch.cuda.synchronize()
plan.set_profiling(True); plan.stage_times(reset=True)
for _ in range(3): step()
torch.cuda.synchronize()
s = plan.stage_times()
tot = sum(v[0] for v in s.values()) / 3
print(kit, n, {k: round(v[0] / 3, 3) for k, v in s.items()}, "total_ms %.3f -> %.2f M reads/s" % (tot, n / tot / 1e3))

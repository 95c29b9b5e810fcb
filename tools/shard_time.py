import sys, torch, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import montecarlocuda_b200 as m
from montecarlocuda_b200 import distributed as D
eng = m.Engine(0)
VAN = m.OptionData(100.0, 100.0, 0.05, 0.2, 1.0)
for prec in ("f32", "f64"):
    p = m.plan("vanilla", VAN, 1 << 32, prec)
    for world in (8, 4, 1):
        first, count = m.shard_range(p, 0, world)
        acc = torch.zeros(12, dtype=torch.int64, device="cuda:0")
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            D._launch(eng, "vanilla", p, VAN, 1, first, count, acc, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            D._launch(eng, "vanilla", p, VAN, 1, first, count, acc, st)
        e1.record(); torch.cuda.synchronize()
        print(prec, "shard 0 of", world, "chunks", count, "ms", e0.elapsed_time(e1) / 20)

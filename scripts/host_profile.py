import cProfile, pstats, io, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
torch.cuda.set_device(0)
name = sys.argv[1]
wl = bench.make_workload(name, 0, 1)
for i in range(5): wl.step(i)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for i in range(20): out = wl.step(i)
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18); print(s.getvalue()[:3500])

import cProfile, pstats, sys, io
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import run_compare
pr = cProfile.Profile()
pr.enable()
run_compare.arm(sys.argv[1], 20, 4000, 8000, 0.05)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45)
print(s.getvalue()[:9000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(25)
print(s.getvalue()[:6000])

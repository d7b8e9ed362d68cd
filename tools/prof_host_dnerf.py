import cProfile, pstats, sys, io, runpy
sys.argv = ["ncu_dnerf.py", "500"]
src = open("tools/ncu_dnerf.py").read().replace("for _ in range(3):", "for _ in range(int(__import__('os').environ.get('STEPS','3'))):")
code = compile(src, "tools/ncu_dnerf.py", "exec")
import os
os.environ["STEPS"] = "5"
exec(code, {"__name__": "__main__", "__file__": "tools/ncu_dnerf.py"})
os.environ["STEPS"] = "100"
pr = cProfile.Profile(); pr.enable()
exec(code, {"__name__": "__main__", "__file__": "tools/ncu_dnerf.py"})
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28); print(s.getvalue()[:6000])

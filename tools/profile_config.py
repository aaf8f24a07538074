"""cProfile of one whole plspy_b200.PLS(...) call of a BASELINE config (development aid; tools/time_config.py times it).

    PYTHONPATH=. python tools/profile_config.py --cfg 2 [--analysis device]
"""
import os as _os, sys as _sys, runpy, cProfile, pstats, io
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
argv = _sys.argv[1:]
_sys.argv = ["time_config.py", "--draw", "--reps", "2"] + argv
here = _os.path.dirname(_os.path.abspath(__file__))
runpy.run_path(_os.path.join(here, "time_config.py"), run_name="__main__")         # warm-up (two calls)
import plspy_b200
pr = cProfile.Profile()
_sys.argv = ["time_config.py", "--draw", "--reps", "1"] + argv
pr.enable(); runpy.run_path(_os.path.join(here, "time_config.py"), run_name="__main__"); pr.disable()
buf = io.StringIO(); pstats.Stats(pr, stream=buf).sort_stats("tottime").print_stats(28)
print("\n".join(l[:160] for l in buf.getvalue().splitlines() if l.strip()))

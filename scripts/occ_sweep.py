import os, sys, subprocess
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for B in ("10000", "5000"):
    for lib in ("", "scripts/_variants/lib_rw8.so"):
        env = dict(os.environ, MPCB200_LIB=os.path.join(root, lib) if lib else "")
        out = subprocess.run([sys.executable, os.path.join(root, "scripts/prof_one.py"), B, "kin_cbf"], env=env, capture_output=True, text=True, timeout=120)
        lines = out.stdout.strip().splitlines()
        print(lib or "default(rw12)", lines[-1] if lines else out.stderr[-300:], flush=True)

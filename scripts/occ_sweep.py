import os, sys, subprocess
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for lib in ("", "scripts/_variants/lib_se2.so", "scripts/_variants/lib_se4.so", "scripts/_variants/lib_se1024.so"):
    env = dict(os.environ, MPCB200_LIB=os.path.join(root, lib) if lib else "")
    try:
        out = subprocess.run([sys.executable, os.path.join(root, "scripts/prof_one.py"), "16384"], env=env, capture_output=True, text=True, timeout=120)
        lines = out.stdout.strip().splitlines()
        print(lib or "default(sync every 1)", lines[-1] if lines else out.stderr[-300:], flush=True)
    except subprocess.TimeoutExpired:
        print(lib or "default", "TIMEOUT", flush=True)

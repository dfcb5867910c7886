import os, sys, subprocess
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for lib, pads in (("", [0, 2048, 4096, 8192, 12288]), ("scripts/_variants/lib_mb6.so", [0, 8192, 16384]), ("scripts/_variants/lib_mb8.so", [0, 4096, 8192]), ("scripts/_variants/lib_mb12.so", [0])):
    for pad in pads:
        env = dict(os.environ, MPCB200_LIB=os.path.join(root, lib) if lib else "", MPCB_SMEM_PAD=str(pad))
        out = subprocess.run([sys.executable, os.path.join(root, "scripts/prof_one.py"), "16384"], env=env, capture_output=True, text=True).stdout.strip().splitlines()
        print(lib or "default(mb10)", "pad", pad, out[-1] if out else "?", flush=True)

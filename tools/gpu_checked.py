"""Memory-safety run without compute-sanitizer (closed on this GPU pool): builds libmpcb200 with -DMPCB_CHECKED (device
asserts on every table segment index, work-list slot, class-list slot and problem index) into build/, then drives every
kernel through tools/gpu_sanitize.py's workload with that library.  A failed assert surfaces as a CUDA error.
    python tools/gpu_checked.py build      (here: cross-compiles)
    python tools/gpu_checked.py            (on the GPU box)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "build", "libmpcb200_checked.so")
if len(sys.argv) > 1 and sys.argv[1] == "build":
    import importlib
    b = importlib.import_module("safe_autonomous_driving_mpc_b200._build")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    print(b.build(defines=["MPCB_CHECKED"], out=LIB))
    sys.exit(0)
assert os.path.exists(LIB), "run `python tools/gpu_checked.py build` first"
env = dict(os.environ, MPCB_LIB=LIB)
r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_sanitize.py")], env=env, capture_output=True, text=True)
print(r.stdout[-3000:])
print(r.stderr[-3000:])
print("checked build:", "PASS (no device assert fired, all kernels ran)" if r.returncode == 0 and r.stdout.strip().endswith("ok") else f"FAIL rc={r.returncode}")
sys.exit(r.returncode)

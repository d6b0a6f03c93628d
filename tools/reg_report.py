"""Registers / stack / spills per fused-kernel instantiation, from the -Xptxas -v build logs (csrc/build/*.log)."""
import re, subprocess, sys, glob, os
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tensor-cuda-fft-_b200", "csrc", "build")
pat = sys.argv[1] if len(sys.argv) > 1 else "ext_f32"
for f in sorted(glob.glob(os.path.join(root, f"*{pat}*.log"))):
    txt = open(f).read()
    blocks = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", txt)
    print(os.path.basename(f))
    for name, st, ss, sl, regs in blocks:
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        m = re.search(r"sml_fast_kernel<([^>]*)>", dem)
        if m:
            print("  ", m.group(1), "regs", regs, "stack", st, "spill", ss, sl)

"""Per-kernel SASS opcode histogram of the built library (no GPU needed):
   python scripts/sass_histogram.py [lib.so] > profiles/rNN_sass_histogram.txt
Shows which kernels carry tcgen05 (UTCHMMA / UTCBAR / LDTM / STTM), bulk-async copies (UBLKCP), packed FP32
(FFMA2 / FMUL2 / FADD2) and MUFU work."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "neurallaplace_b200", "lib", "libnl_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "UBLKPF", "SYNCS", "ELECT", "FFMA2", "FMUL2", "FADD2",
        "FFMA", "FMUL", "FADD", "DFMA", "DMMA", "HMMA", "MUFU", "F2FP", "LDS", "STS", "LDG", "STG", "RED", "ATOMG", "BAR", "SHFL"]
print(f"# SASS opcode histogram per kernel of {os.path.relpath(lib, ROOT)} (cuobjdump -sass), sm_100a")
print("# " + " ".join(f"{k:>7s}" for k in ["total"] + KEYS) + "  kernel")
tot = collections.Counter()
for fn in txt.split("Function : ")[1:]:
  name = fn.split("\n")[0].strip()
  try:
    name = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
  except OSError:
    pass
  ops = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", fn, re.M)
  c = collections.Counter(ops)
  tot.update(c)
  short = re.sub(r"\([^()]*\)$", "", name).replace("(bool)", "").replace("(int)", "").replace("nl::", "").replace("(anonymous namespace)::", "")
  print("  " + " ".join(f"{v:7d}" for v in [len(ops)] + [c.get(k, 0) for k in KEYS]) + "  " + short[:120])
print("  " + " ".join(f"{v:7d}" for v in [sum(tot.values())] + [tot.get(k, 0) for k in KEYS]) + "  TOTAL")

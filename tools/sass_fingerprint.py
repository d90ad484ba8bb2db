"""Per-kernel fingerprint of the SASS in libdrake_b200_hydro.so (instruction text without addresses / encodings'
offsets), to show that a source refactor left a GPU-validated kernel's machine code untouched:

  python tools/sass_fingerprint.py [lib.so] > profiles/rN_sass_fingerprints.txt
  diff <(python tools/sass_fingerprint.py old.so) <(python tools/sass_fingerprint.py new.so)
"""
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "drake_b200", "lib", "libdrake_b200_hydro.so")
text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, out = None, {}
for line in text.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        out[name] = [hashlib.md5(), 0]
        continue
    if name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        out[name][0].update(re.sub(r"/\*[0-9a-f]+\*/", "", line).strip().encode())
        out[name][1] += 1
demangle = subprocess.run(["cu++filt"] + list(out), capture_output=True, text=True).stdout.splitlines() if out else []
for (mangled, (h, n)), pretty in sorted(zip(out.items(), demangle), key=lambda t: t[1]):
    print(f"{h.hexdigest()[:16]}  {n:6d} instr  {pretty}")

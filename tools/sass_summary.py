"""SASS mnemonic counts per kernel of the built library (no GPU needed): the proof that the hot kernels are tcgen05 / TMEM /
TMA code.  Writes the text that is committed as profiles/<tag>_sass_summary.txt.
  python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dcvc_b200", "lib", "libdcvc_b200.so")
WANT = re.compile(r"^(UTCHMMA|UTCBAR|UTCATOMSWS|UTMALDG|UTMASTG|UTMAPF|LDTM|STTM|LDGSTS|SYNCS|MUFU\.TANH|UTCCP)")

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n  # noqa: E731
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = re.sub(r"\((int|bool|unsigned int)\)", "", demangle(m.group(1)))
        cur = re.sub(r"\(.*", "", cur)
        per[cur] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]+)", line)
    if m and cur and WANT.match(m.group(1)):
        per[cur][m.group(1)] += 1
print(f"# SASS mnemonic counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, CUDA 12.9, sm_100a; tools/sass_summary.py)")
print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / tcgen05.st, UTMALDG / UTMASTG = TMA load / store,")
print("# UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc / dealloc, LDGSTS = cp.async, SYNCS = mbarrier ops\n")
for k, c in per.items():
    if not c:
        continue
    print(k)
    print("    " + ", ".join(f"{m} x{n}" for m, n in sorted(c.items())))

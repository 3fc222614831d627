"""Per-kernel instruction evidence from the shipped library: `cuobjdump -sass libi2l_b200.so`, one line per kernel
with the counts of the Blackwell mnemonics that prove the tcgen05 / TMEM / TMA / cluster path
(B200_PROFILING.md): UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG (TMA tensor load), UBLKCP (bulk
copy, incl. DSMEM), UTCBAR (tcgen05.commit), SYNCS (mbarrier), MUFU.TANH, REDUX, FFMA.  Runs without a GPU.
Usage: python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hmer-img2latex_b200", "csrc", "libi2l_b200.so")
PATTERNS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "MUFU.TANH", "MUFU.EX2", "REDUX",
            "FFMA", "HMMA", "LDG", "STG", "ATOM"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                           text=True).stdout.splitlines()
    blocks = re.split(r"\n\s*Function : \S+\n", sass)[1:]
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(blocks)} kernels (cuobjdump -sass, sm_100a); columns = instruction counts")
    print("# " + " ".join(f"{p:>9s}" for p in PATTERNS) + "   instrs  kernel")
    for name, body in sorted(zip(names, blocks), key=lambda t: t[0]):
        ops = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body, flags=re.M)
        cnt = collections.Counter()
        for op in ops:
            for p in PATTERNS:
                if op.startswith(p):
                    cnt[p] += 1
        short = re.sub(r"\(anonymous namespace\)::", "", name)
        short = re.sub(r"\(.*$", "", short)
        print("  " + " ".join(f"{cnt[p]:9d}" for p in PATTERNS) + f"  {len(ops):7d}  {short}")


if __name__ == "__main__":
    sys.exit(main())

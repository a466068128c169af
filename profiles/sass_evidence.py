"""Static evidence that the hot kernels are Blackwell-native: per kernel of liblgs_b200.so, counts of the SASS mnemonics
the profiling guide names (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA loads, LDGSTS = cp.async,
HMMA = legacy mma.sync — must be absent) plus 128-bit global loads/stores and atomics.
usage: python profiles/sass_evidence.py > profiles/r1_j_sass_evidence.txt   (needs cuobjdump; no GPU)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "languagegroundedsemseg_b200", "csrc", "liblgs_b200.so")
PAT = [("UTC*MMA", r"\bUTC\w*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
       ("LDGSTS", r"\bLDGSTS"), ("HMMA", r"\bHMMA"), ("LDG.128", r"\bLDG\.E\S*\.128"), ("STG.128", r"\bSTG\.E\S*\.128"),
       ("REDG/ATOMG", r"\b(REDG|ATOMG)\b"), ("SYNCS", r"\bSYNCS")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur:
            kernels[cur]["instructions"] += bool(re.search(r"/\*[0-9a-f]{4}\*/", line))
            for name, pat in PAT:
                if re.search(pat, line):
                    kernels[cur][name] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a) — SASS mnemonic counts per kernel")
    print(f"# {'kernel':70s} " + " ".join(f"{n:>8s}" for n, _ in PAT) + "   instr")
    for (mangled, cnt), name in zip(kernels.items(), demangled):
        short = re.sub(r"\(.*", "", name).replace("void ", "")
        print(f"{short[:72]:72s} " + " ".join(f"{cnt[n]:8d}" for n, _ in PAT) + f" {cnt['instructions']:7d}")
    legacy = sum(c["HMMA"] for c in kernels.values())
    print(f"# legacy mma.sync (HMMA) instructions in the library: {legacy}")


if __name__ == "__main__":
    main()

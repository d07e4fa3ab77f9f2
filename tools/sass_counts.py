"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use, from the in-tree library.
    python tools/sass_counts.py > profiles/sass_counts_r2.txt
UTCHMMA/UTCQMMA... = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA load/store, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops, FFMA2/FADD2 = packed f32x2 arithmetic, MUFU.EX2 = exponentials."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sam_audio_b200", "libsamaudio_b200.so")
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "FFMA2", "FADD2", "MUFU.EX2", "HMMA", "LDGSTS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, counts, size = None, collections.OrderedDict(), {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            size[cur] = 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            size[cur] += 1
            op = m.group(1)
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k == "MUFU.EX2" and op.startswith("MUFU.EX2")):
                    counts[cur][k] += 1
    dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(counts)} kernels; SASS mnemonic counts per kernel")
    print("# " + " ".join(f"{k:>8s}" for k in ["instrs"] + KEYS) + "  kernel")
    tot = collections.Counter()
    for (name, c), d in zip(counts.items(), dem):
        tot.update(c)
        d = re.sub(r"\(CUtensorMap_st.*", "", d).replace("sab::", "")
        print("  " + " ".join(f"{v:8d}" for v in [size[name]] + [c[k] for k in KEYS]) + "  " + d[:110])
    print("# total " + " ".join(f"{k}={tot[k]}" for k in KEYS))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of libpose_b200.so the instruction-mnemonic histogram (cuobjdump -sass) and the
lines that prove which engines a kernel uses (tcgen05: UTCHMMA / UTCBAR / LDTM / STTM; TMA: UTMALDG / UBLKCP; mbarrier: SYNCS).

    python tools/sass_listing.py <kernel-name-regex> <out.txt>
"""
import collections
import re
import subprocess
import sys

LIB = "pytorch-pose-estimation_b200/libpose_b200.so"
KEY = re.compile(r"CREDUX\S*|UTC\w*|LDTM\S*|STTM\S*|UTMALDG\S*|UTMASTG\S*|UBLKCP\S*|SYNCS\S*|UTCATOMSWS\S*|MUFU\S*|LDG\S*|STG\S*|LDS\S*|STS\S*")


def main():
    pat, out = re.compile(sys.argv[1]), sys.argv[2]
    text = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, name = collections.OrderedDict(), None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kernels[name] = []
            continue
        if name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            kernels[name].append(line)
    with open(out, "w") as f:
        f.write(f"# cuobjdump -sass {LIB}, kernels matching /{sys.argv[1]}/ (sm_100a)\n")
        for k, lines in kernels.items():
            if not pat.search(k):
                continue
            hist = collections.Counter()
            for ln in lines:
                m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
                if m:
                    hist[m.group(1)] += 1
            f.write(f"\n== {k}\n   {len(lines)} instructions\n")
            keys = {m: c for m, c in hist.items() if KEY.fullmatch(m)}
            f.write("   engine / memory mnemonics: " + ", ".join(f"{m} x{c}" for m, c in sorted(keys.items())) + "\n")
            f.write("   top mnemonics: " + ", ".join(f"{m} x{c}" for m, c in hist.most_common(14)) + "\n")
            shown = 0
            for ln in lines:
                if re.search(r"UTC\w*MMA|UTCBAR|LDTM|STTM|UTMALDG|UBLKCP", ln) and shown < 24:
                    f.write("   " + re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ln).strip() + "\n")
                    shown += 1
    print(out)


if __name__ == "__main__":
    main()

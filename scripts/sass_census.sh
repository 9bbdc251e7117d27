#!/bin/bash
# opcode census of every kernel in libppn_decode.so: the mnemonics that show which hardware paths are used
# (UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor copies, UBLKCP = 1-D bulk copies, SYNCS = mbarrier,
#  UCGABAR = cluster barrier, UBLKPF = L2 bulk prefetch, HSET2/HMNMX2 = packed 16-bit compares, LDGSTS = cp.async)
SO=pytorch_pose_proposal_network_b200/libppn_decode.so
echo "# cuobjdump -sass $SO ($(date -u +%F)), nvcc $(nvcc --version | grep -o 'release [0-9.]*')"
python3 - "$SO" <<'PY'
import collections, re, subprocess, sys
so = sys.argv[1]
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
keep = re.compile(r"^(UTC\w*|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UBLKCP|UBLKPF|SYNCS|UCGABAR\w*|HSET2|HMNMX2|LDGSTS|ACQBULK|PREEXIT|ATOM\w*|ATOMS|ATOMG|RED|REDG|MUFU|LDS|STS|LDG|STG|BAR|SHFL|VOTE|FSETP|FSEL|FMNMX|NANOSLEEP|ERRBAR|CCTL)$")
fn, total, ops = None, collections.Counter(), collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(1)
        total[fn] += 1
        base = op.split(".")[0]
        if keep.match(base):
            ops[fn][op if base.startswith(("UTC", "UTMA", "UBLK", "SYNCS", "UCGA", "LDTM", "HSET2", "HMNMX2", "MUFU")) else base] += 1
names = subprocess.run(["c++filt"], input="\n".join(total), capture_output=True, text=True).stdout.splitlines()
for mangled, nice in sorted(zip(total, names), key=lambda kv: kv[1]):
    short = re.sub(r"\(.*", "", nice)
    print(f"\n{short}   [{total[mangled]} SASS instructions]")
    row = ops[mangled]
    print("   " + "  ".join(f"{k}:{v}" for k, v in sorted(row.items())))
PY
